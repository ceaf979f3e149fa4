"""Where does a drop-in Warper.warp_image call spend its time?  (run on a GPU box: python tests/tools/dropin_probe.py)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stitching_b200 import Warper, _lib, host_pool, rigs  # noqa: E402

cfg = rigs.config("cfg2", 1)
cam = cfg["cameras"][3]
img = rigs.synth_image(cfg["h"], cfg["w"], 3)
L = _lib.lib()
_lib.check(L.sb_init(0), "sb_init")
w = Warper("spherical")
w.set_scale(cfg["cameras"])
wtype, scale, K, R = w._params(cam, 1)
fp = lambda a: a.ctypes.data_as(_lib.c_float_p)  # noqa: E731
rect = (C.c_int * 4)()


def timed(name, fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{name:58s} {np.median(ts):8.2f} ms  (min {min(ts):.2f})", flush=True)


timed("sb_warp_roi", lambda: L.sb_warp_roi(wtype, scale, fp(K), fp(R), cfg["w"], cfg["h"], rect))
W, H = rect[2], rect[3]
print("warped size", W, H)
pageable = np.empty((H, W, 3), np.uint8)
pinned = host_pool.empty((H, W, 3), np.uint8)
pmask = host_pool.empty((H, W), np.uint8)
src_pinned = host_pool.empty(img.shape, np.uint8)
src_pinned[...] = img


def call(src, dst, msk, keep=False):
    k = C.c_void_p()
    rc = L.sb_warp_keep(wtype, scale, fp(K), fp(R), src.ctypes.data_as(C.c_void_p) if src is not None else None, cfg["w"], cfg["h"], cfg["w"] * 3,
                        dst.ctypes.data_as(C.c_void_p) if dst is not None else None, W * 3,
                        msk.ctypes.data_as(C.c_void_p) if msk is not None else None, W, rect, C.byref(k) if keep else None, None)
    _lib.check(rc, "sb_warp_keep")
    if keep and k.value:
        L.sb_devimg_release(k)


timed("np.empty + touch (fresh pageable result array)", lambda: np.empty((H, W, 3), np.uint8).fill(0))
timed("warp: mask only -> pinned", lambda: call(None, None, pmask))
timed("warp: pageable src -> pageable dst (reused array)", lambda: call(img, pageable, None))
timed("warp: pageable src -> fresh np.empty dst", lambda: call(img, np.empty((H, W, 3), np.uint8), None))
timed("warp: pageable src -> pinned dst", lambda: call(img, pinned, None))
timed("warp: pinned src -> pinned dst", lambda: call(src_pinned, pinned, None))
timed("warp: pinned src -> pinned dst, keep twin", lambda: call(src_pinned, pinned, None, True))
timed("Warper.warp_image (drop-in, pageable src)", lambda: w.warp_image(img, cam))
timed("Warper.create_and_warp_mask (drop-in)", lambda: w.create_and_warp_mask((cfg["w"], cfg["h"]), cam))
