"""Synthetic KITTI-shaped input generator (SURVEY.md §8d "Value distribution").

Not white noise: smooth large-scale structure + mid/fine texture + random rectangles confined
to the middle rows (low-texture "sky"/"road" bands exercise the minThFAST fallback of
orb_extractor.cpp:753-757), lightly blurred, plus sensor noise.  Deterministic per seed.
Used by tests/ and bench.py; it is input data, not part of the hot path.
"""
import numpy as np

try:  # cv2 is only a fast Gaussian filter here; scipy is the fallback
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None


def _gauss(field, sigma):
    if _cv2 is not None:
        return _cv2.GaussianBlur(field, (0, 0), sigmaX=sigma, sigmaY=sigma, borderType=_cv2.BORDER_REFLECT_101)
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(field, sigma, mode="mirror")


def texture(h, w, seed):
    """float32 texture of shape (h, w), roughly in [0, 255] before clipping."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 110.0, np.float32)
    for sigma, std in ((32.0, 35.0), (10.0, 18.0), (4.0, 8.0)):
        f = _gauss(rng.standard_normal((h, w)).astype(np.float32), sigma)
        img += f * np.float32(std / max(float(f.std()), 1e-6))
    nrect = (h * w) // 500
    y_lo, y_hi = int(0.22 * h), int(0.82 * h)
    rw = rng.integers(3, 49, nrect)
    rh = rng.integers(3, 37, nrect)
    rx = rng.integers(0, max(w - 3, 1), nrect)
    ry = rng.integers(y_lo, max(y_hi - 3, y_lo + 1), nrect)
    amp = rng.uniform(-60.0, 60.0, nrect).astype(np.float32)
    for i in range(nrect):
        y1 = min(ry[i] + rh[i], y_hi)
        img[ry[i]:y1, rx[i]:rx[i] + rw[i]] += amp[i]
    return _gauss(img, 0.8)


def _finish(img, rng):
    out = img + rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def frame(h=376, w=1241, seed=0):
    """One u8 C-contiguous (h, w) frame (BASELINE config 1)."""
    rng = np.random.default_rng(seed + 1_000_003)
    return np.ascontiguousarray(_finish(texture(h, w, seed), rng))


def stereo_pair(h=376, w=1241, seed=0, max_disp=64, n_bands=8):
    """(left, right) u8 frames cut from one wide texture; right[y, x] = left[y, x + d(y)] with a
    per-row-band disparity d in [4, max_disp], plus independent noise (BASELINE config 2)."""
    rng = np.random.default_rng(seed + 2_000_003)
    tex = texture(h, w + max_disp + 32, seed)
    left = tex[:, :w]
    right = np.empty_like(left)
    edges = np.linspace(0, h, n_bands + 1).astype(int)
    disps = rng.integers(4, max_disp + 1, n_bands)
    for b in range(n_bands):
        d = int(disps[b])
        right[edges[b]:edges[b + 1]] = tex[edges[b]:edges[b + 1], d:d + w]
    return (np.ascontiguousarray(_finish(left, rng)), np.ascontiguousarray(_finish(right, rng)))


def shifted_frame(base_seed, h=376, w=1241, dx=8, dy=4):
    """second monocular frame = first shifted by (dx, dy) + new noise (BASELINE config 4)."""
    rng = np.random.default_rng(base_seed + 3_000_003)
    tex = texture(h + 2 * abs(dy) + 2, w + 2 * abs(dx) + 2, base_seed)
    a = tex[abs(dy):abs(dy) + h, abs(dx):abs(dx) + w]
    b = tex[abs(dy) + dy:abs(dy) + dy + h, abs(dx) + dx:abs(dx) + dx + w]
    return (np.ascontiguousarray(_finish(a, rng)), np.ascontiguousarray(_finish(b, rng)))
