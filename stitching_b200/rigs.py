"""Synthetic cameras and images for the BASELINE.json configurations (numpy only, seeded).

Used by bench.py, the tests and the golden-vector generator; not on the compute path.
Rigs follow SURVEY.md section 8(d): a yaw ring for the rotational projections, a 4x4 grid of
near-identity affines for the AffineStitcher configuration.
"""
import numpy as np


class Camera:
    """Duck-typed stand-in for cv.detail.CameraParams (fields the hot path reads: warper.py:36,48,86)."""

    def __init__(self, focal, aspect, ppx, ppy, R, t=None):
        self.focal = float(focal)
        self.aspect = float(aspect)
        self.ppx = float(ppx)
        self.ppy = float(ppy)
        self.R = np.ascontiguousarray(R, dtype=np.float32)
        self.t = np.zeros((3, 1), np.float64) if t is None else t

    def K(self):
        return np.array(
            [[self.focal, 0.0, self.ppx], [0.0, self.focal * self.aspect, self.ppy], [0.0, 0.0, 1.0]], np.float64
        )


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float64)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)


def yaw_ring(n, w, h, focal, step_deg):
    """R_i = Ry(step * (i - (n-1)/2)) * Rx(0.01 * ((i mod 3) - 1)), K = (focal, aspect 1, pp = centre)."""
    cams = []
    for i in range(n):
        R = rot_y(np.deg2rad(step_deg) * (i - (n - 1) / 2.0)) @ rot_x(0.01 * ((i % 3) - 1))
        cams.append(Camera(focal, 1.0, w / 2.0, h / 2.0, R.astype(np.float32)))
    return cams


def affine_grid(n, w, h, cols=4):
    """cfg 5: near-identity affines on a grid with 25% overlap; K = I, scale 1 (AffineStitcher conventions)."""
    cams = []
    for i in range(n):
        r, c = divmod(i, cols)
        th = np.deg2rad(0.5) * ((i % 3) - 1)
        H = np.array(
            [[np.cos(th), -np.sin(th), 0.75 * w * c + 3.3 * r], [np.sin(th), np.cos(th), 0.75 * h * r - 2.7 * c], [0, 0, 1]],
            np.float32,
        )
        cams.append(Camera(1.0, 1.0, 0.0, 0.0, H))
    return cams


def synth_image(h, w, seed, noise=8):
    """Smooth low-frequency colour field (bilinear upsample of a coarse random grid) plus integer noise."""
    rng = np.random.default_rng(seed)
    gh, gw = h // 64 + 2, w // 64 + 2
    grid = rng.integers(0, 256, (gh, gw, 3)).astype(np.float32)
    ys = np.linspace(0, gh - 1, h, dtype=np.float32)
    xs = np.linspace(0, gw - 1, w, dtype=np.float32)
    y0 = np.minimum(ys.astype(np.int32), gh - 2)
    x0 = np.minimum(xs.astype(np.int32), gw - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    top = grid[y0][:, x0] * (1 - fx) + grid[y0][:, x0 + 1] * fx
    bot = grid[y0 + 1][:, x0] * (1 - fx) + grid[y0 + 1][:, x0 + 1] * fx
    img = top * (1 - fy) + bot * fy
    if noise:
        img += rng.integers(-noise, noise + 1, (h, w, 3)).astype(np.float32)
    # C-contiguous like a decoded photograph (the fancy indexing above leaves a transposed layout behind)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def noise_image(h, w, seed):
    """Adversarial parity input: independent uniform bytes."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


# BASELINE.json configs -> (n, w, h, warper_type, rig builder, blender, strength)
def config(name, scale_down=1):
    s = scale_down
    if name == "cfg2":  # 8 x 4000x3000 spherical + multiband
        w, h, n = 4000 // s, 3000 // s, 8
        return dict(n=n, w=w, h=h, warper="spherical", cameras=yaw_ring(n, w, h, 4000 / s, 30), blender="multiband", strength=5)
    if name == "cfg3":  # 32 x 4000x3000 cylindrical + multiband
        w, h, n = 4000 // s, 3000 // s, 32
        return dict(n=n, w=w, h=h, warper="cylindrical", cameras=yaw_ring(n, w, h, 8000 / s, 10), blender="multiband", strength=5)
    if name == "cfg4":  # 8 x 8000x6000 spherical + multiband
        w, h, n = 8000 // s, 6000 // s, 8
        return dict(n=n, w=w, h=h, warper="spherical", cameras=yaw_ring(n, w, h, 8000 / s, 30), blender="multiband", strength=5)
    if name == "cfg5":  # 16 x 2000x1500 affine + feather
        w, h, n = 2000 // s, 1500 // s, 16
        return dict(n=n, w=w, h=h, warper="affine", cameras=affine_grid(n, w, h), blender="feather", strength=5)
    raise KeyError(name)
