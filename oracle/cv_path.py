"""The reference's CPU implementation of the hot path, for timing (bench.py) and cross-checks.  TEST/BENCH
INFRASTRUCTURE ONLY -- never imported by the product.

The reference (OpenStitching/stitching v0.7.0) is a thin Python layer over the third-party wheel cv2; the
/root/reference tree does not exist on the GPU box, but cv2 (the code that actually does the arithmetic and
takes the time) ships in the image.  This module issues the same cv2 calls, in the same order, as
  stitching/warper.py:43-52, 58-68, 79-82   (cv.PyRotationWarper.warp / warpRoi)
  stitching/blender.py:23-48                (resultRoi, MultiBand/Feather/NO prepare, feed, blend, convertScaleAbs)
driven like stitcher.py:178-189, 241-259.  bench.py reports it as cpu_baseline.kind = "port" (a restatement
of the 150-line wrapper; the numeric backend is the reference's own).
"""
import time

import numpy as np


def available():
    try:
        import cv2  # noqa: F401

        return True
    except Exception:
        return False


def get_K(cam, aspect=1):
    K = cam.K().astype(np.float32)
    K[0, 0] *= aspect
    K[0, 2] *= aspect
    K[1, 1] *= aspect
    K[1, 2] *= aspect
    return K


def composite(cfg, cams, imgs, threads=None):
    """Returns (pano, mask, seconds per stage dict)."""
    import cv2 as cv

    if threads is not None:
        cv.setNumThreads(int(threads))
    wtype = cfg["warper"]
    scale = float(np.median([c.focal for c in cams]))
    t = {}
    t0 = time.perf_counter()
    warped = []
    for img, cam in zip(imgs, cams):
        w = cv.PyRotationWarper(wtype, scale)
        warped.append(w.warp(img, get_K(cam), cam.R, cv.INTER_LINEAR, cv.BORDER_REFLECT)[1])
    t["warp_images"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    masks = []
    for img, cam in zip(imgs, cams):
        w = cv.PyRotationWarper(wtype, scale)
        m = 255 * np.ones(img.shape[:2], np.uint8)
        masks.append(w.warp(m, get_K(cam), cam.R, cv.INTER_NEAREST, cv.BORDER_CONSTANT)[1])
    t["warp_masks"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    corners, sizes = [], []
    for img, cam in zip(imgs, cams):
        roi = cv.PyRotationWarper(wtype, scale).warpRoi((img.shape[1], img.shape[0]), get_K(cam), cam.R)
        corners.append(roi[0:2])
        sizes.append(roi[2:4])
    dst = cv.detail.resultRoi(corners=corners, sizes=sizes)
    bw = np.sqrt(dst[2] * dst[3]) * cfg["strength"] / 100
    if cfg["blender"] == "no" or bw < 1:
        b = cv.detail.Blender_createDefault(cv.detail.Blender_NO)
    elif cfg["blender"] == "multiband":
        b = cv.detail_MultiBandBlender()
        b.setNumBands(int(np.log(bw) / np.log(2.0) - 1.0))
    else:
        b = cv.detail_FeatherBlender()
        b.setSharpness(1.0 / bw)
    b.prepare(dst)
    t["prepare"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for img, m, c in zip(warped, masks, corners):
        b.feed(cv.UMat(img.astype(np.int16)), m, c)
    t["feed"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    res, res_mask = b.blend(None, None)
    pano = cv.convertScaleAbs(res)
    t["blend"] = time.perf_counter() - t0
    return pano, res_mask, t


def describe():
    import cv2 as cv

    info = cv.getBuildInformation()
    par = [ln.strip() for ln in info.splitlines() if "Parallel framework" in ln]
    return {"cv2": cv.__version__, "threads": cv.getNumThreads(), "parallel": par[0] if par else ""}
