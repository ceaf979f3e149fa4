"""stitching_b200 -- the compositing hot path of OpenStitching/stitching on NVIDIA B200 (sm_100a).

Drop-in replacements for `stitching.warper.Warper`, `stitching.blender.Blender` and `SeamFinder.resize` (same
interface) backed by hand-written CUDA kernels behind a C ABI (include/stitch_b200.h), plus a fused `Compositor`.
`install()` swaps them into an installed `stitching` package so that Stitcher / AffineStitcher / the CLI run
unchanged.
"""
from . import exposure_error_compensator, images, seam_finder  # noqa: F401
from .blender import Blender  # noqa: F401
from .compositor import Compositor  # noqa: F401
from .stitching_error import StitchingError, StitchingWarning  # noqa: F401
from .warper import Warper  # noqa: F401

__version__ = "0.1.0"


def install(stitching_module=None):
    """Route `stitching.Stitcher` (and cropper / seam finder / verbose callers) through the B200 classes.

    The reference modules bind the class names at import time (`from .warper import Warper` in
    stitcher.py, cropper.py, seam_finder.py, verbose.py), so the names are patched in each of them.
    """
    import importlib

    if stitching_module is None:
        stitching_module = importlib.import_module("stitching")
    pkg = stitching_module.__name__
    for mod, names in (
        ("warper", ("Warper",)), ("blender", ("Blender",)), ("stitcher", ("Warper", "Blender")),
        ("cropper", ("Blender",)), ("seam_finder", ("Blender",)), ("verbose", ("Warper", "Blender")),
    ):
        try:
            m = importlib.import_module(f"{pkg}.{mod}")
        except Exception:
            continue
        for name in names:
            if hasattr(m, name):
                setattr(m, name, {"Warper": Warper, "Blender": Blender}[name])
    # the FINAL-resolution step of the seam finder (seam_finder.py:38-43); stitcher.py calls it through the class
    try:
        sf = importlib.import_module(f"{pkg}.seam_finder")
        sf.SeamFinder.resize = staticmethod(seam_finder.resize)
    except Exception:
        pass
    # the resampling to MEDIUM / LOW / FINAL resolution (images.py:120-123)
    try:
        im = importlib.import_module(f"{pkg}.images")
        im.Images.resize_img_by_scaler = staticmethod(images.resize_img_by_scaler)
    except Exception:
        pass
    # the FINAL-resolution step of the exposure compensator (exposure_error_compensator.py:43-45)
    try:
        ec = importlib.import_module(f"{pkg}.exposure_error_compensator")

        def _apply(self, *args):
            return exposure_error_compensator.apply(self.compensator, *args)

        ec.ExposureErrorCompensator.apply = _apply
    except Exception:
        pass
    return stitching_module


def __getattr__(name):
    """`stitching_b200.Stitcher` / `stitching_b200.AffineStitcher` (stitching/__init__.py:1): the reference's pipeline
    classes -- same DEFAULT_SETTINGS, same CLI -- running on the B200 classes.  They live in the reference package
    (registration, seam estimation, cropping ... are its control plane); install() routes their hot path here."""
    if name in ("Stitcher", "AffineStitcher"):
        return getattr(install(), name)
    raise AttributeError(f"module 'stitching_b200' has no attribute {name!r}")
