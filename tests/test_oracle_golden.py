"""The CPU oracle against the golden vectors produced by the unmodified reference (bit-exact)."""
import numpy as np

import replay


def test_oracle_warp_goldens(oracle):
    replay.run_warper_goldens(replay.OracleWarper)


def test_oracle_blend_goldens(oracle):
    replay.run_blender_goldens(oracle.Blender)


def test_oracle_e2e_goldens(oracle):
    replay.run_e2e_goldens(replay.OracleWarper, oracle.Blender)


def test_oracle_pyramid_goldens(oracle):
    g = replay.load("golden_pyr.npz")
    for i in range(int(g["n"])):
        replay.assert_exact(oracle.pyrdown_s16(g[f"s16_{i}"]), g[f"down_s16_{i}"], f"pyrDown s16 case {i}")
        replay.assert_exact(oracle.pyrup_s16(g[f"s16_{i}"]), g[f"up_s16_{i}"], f"pyrUp s16 case {i}")
        got = oracle.pyrdown_f32(g[f"f32_{i}"])
        assert np.array_equal(got.view(np.uint32), g[f"down_f32_{i}"].view(np.uint32)), f"pyrDown f32 case {i} not bit-exact"
    replay.assert_exact(oracle.convert_scale_abs(g["csa_in"]), g["csa_out"], "convertScaleAbs")
    d = oracle.dist_l1(g["dt_mask"])
    assert np.array_equal(d, g["dt_l1"]), "distanceTransform L1"


def test_oracle_seam_resize_goldens(oracle):
    """SeamFinder.resize of the reference (dilate + cv.resize + AND) == the oracle's restatement."""
    g = replay.load("golden_seam.npz")
    for i in range(int(g["n"])):
        replay.assert_exact(oracle.seam_resize(g[f"seam_{i}"], g[f"mask_{i}"]), g[f"out_{i}"], f"seam resize case {i}")


def test_oracle_image_resize_goldens(oracle):
    """Images.resize_img_by_scaler of the reference (cv.resize INTER_LINEAR_EXACT) == the oracle's restatement."""
    replay.run_resize_goldens(oracle.resize_linear_exact)


def test_oracle_gain_apply_goldens(oracle):
    """ExposureErrorCompensator.apply of the reference (all five compensators) == the oracle's restatement."""
    replay.run_gain_goldens(oracle.gain_apply)


def test_oracle_timelapse_goldens(oracle):
    """Timelapser.process_frame / get_frame of the reference (cv.detail.Timelapser AS_IS / CROP) == the oracle's restatement."""
    replay.run_timelapse_goldens(oracle.Timelapser)
