"""Deterministic EuRoC-shaped synthetic inputs ("kernel-level stream").

Images: T(u,v) = 128 + 3-octave integer-hash value noise (cells 6/17/48 px, amplitudes 70/50/35), clamped to u8.
Frame k of stream s samples T at (x,y) + c_k, c_k = (37 sin(0.05k) + 0.9k, 21 sin(0.08k + 1)); the right image is the
left one sampled at x - 12 - 4 sin(0.002 (x+y)) (disparity 8..16 px). Only float64 +,*,floor and uint32 integer
hashing are used, so the bytes are identical on every IEEE-754 host.
"""
import numpy as np

CELLS = (6.0, 17.0, 48.0)
AMPS = (70.0, 50.0, 35.0)


def _hash01(ix, iy, seed):
    h = (ix.astype(np.int64) * 374761393 + iy.astype(np.int64) * 668265263 + seed * 2246822519) & 0xFFFFFFFF
    h = ((h ^ (h >> 13)) * 1274126177) & 0xFFFFFFFF
    h = (h ^ (h >> 16)) & 0xFFFFFFFF
    return h.astype(np.float64) / 4294967296.0


def texture_at(u, v, seed=42):
    """Continuous texture sampled at float64 coordinate arrays u, v -> float64."""
    out = np.full(u.shape, 128.0)
    for o, (cell, amp) in enumerate(zip(CELLS, AMPS)):
        x, y = u / cell, v / cell
        x0, y0 = np.floor(x), np.floor(y)
        fx, fy = x - x0, y - y0
        ix, iy = x0.astype(np.int64), y0.astype(np.int64)
        s = seed + 101 * o
        n00, n01 = _hash01(ix, iy, s), _hash01(ix + 1, iy, s)
        n10, n11 = _hash01(ix, iy + 1, s), _hash01(ix + 1, iy + 1, s)
        n = (n00 * (1 - fx) + n01 * fx) * (1 - fy) + (n10 * (1 - fx) + n11 * fx) * fy
        out += amp * (2.0 * n - 1.0) * 0.5
    return out


def camera_offset(k):
    return 37.0 * np.sin(0.05 * k) + 0.9 * k, 21.0 * np.sin(0.08 * k + 1.0)


def stereo_frame(k, width=752, height=480, seed=42):
    """(left, right) uint8 images of frame k."""
    cx, cy = camera_offset(k)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    left = texture_at(x + cx, y + cy, seed)
    xr = x - 12.0 - 4.0 * np.sin(0.002 * (x + y))
    right = texture_at(xr + cx, y + cy, seed)
    q = lambda a: np.clip(np.floor(a + 0.5), 0, 255).astype(np.uint8)
    return q(left), q(right)


def true_flow(k0, k1):
    """Image motion (dx, dy) of a static texture point between frames k0 and k1."""
    a, b = camera_offset(k0), camera_offset(k1)
    return a[0] - b[0], a[1] - b[1]


def true_disparity(x, y):
    """A left-image point (x,y) appears in the right image at x + d(x,y) (first-order)."""
    return 12.0 + 4.0 * np.sin(0.002 * (x + y))


def feature_points(n, width=752, height=480, seed=7, border=5.0, flat_fraction=0.0, flat_rect=None):
    """n points uniform in [-border, w+border) x [-border, h+border) (exercises padding and OOB exits); a
    `flat_fraction` of them is placed inside `flat_rect` = (x0,y0,x1,y1) (a constant-gray patch: minEig rejection)."""
    rng = np.random.RandomState(seed)   # MT19937
    pts = np.empty((n, 2), np.float64)
    pts[:, 0] = rng.uniform(-border, width + border, n)
    pts[:, 1] = rng.uniform(-border, height + border, n)
    if flat_fraction > 0 and flat_rect is not None:
        m = int(n * flat_fraction)
        x0, y0, x1, y1 = flat_rect
        pts[:m, 0] = rng.uniform(x0, x1, m)
        pts[:m, 1] = rng.uniform(y0, y1, m)
    return pts.astype(np.float32)


def interior_points(n, width=752, height=480, seed=7, margin=40.0):
    """n trackable points well inside the image (the bench's steady-state feature set)."""
    rng = np.random.RandomState(seed)
    pts = np.empty((n, 2), np.float64)
    pts[:, 0] = rng.uniform(margin, width - margin, n)
    pts[:, 1] = rng.uniform(margin, height - margin, n)
    return pts.astype(np.float32)


def stereo_frames_torch(k0, count, width=752, height=480, seed=42, device="cpu"):
    """Frames k0 .. k0+count-1 as a (count, 2, H, W) uint8 torch tensor, same formulas as stereo_frame() evaluated with
    torch float64/int64 ops (bench.py uses it to fill its frame pool quickly on the GPU; tests use the numpy version;
    the two agree except possibly at exact rounding ties)."""
    import torch
    dev = torch.device(device)
    y, x = torch.meshgrid(torch.arange(height, dtype=torch.float64, device=dev), torch.arange(width, dtype=torch.float64, device=dev), indexing="ij")

    def hash01(ix, iy, s):
        h = (ix * 374761393 + iy * 668265263 + s * 2246822519) & 0xFFFFFFFF
        h = ((h ^ (h >> 13)) * 1274126177) & 0xFFFFFFFF
        h = (h ^ (h >> 16)) & 0xFFFFFFFF
        return h.to(torch.float64) / 4294967296.0

    def tex(u, v):
        out = torch.full_like(u, 128.0)
        for o, (cell, amp) in enumerate(zip(CELLS, AMPS)):
            xs, ys = u / cell, v / cell
            x0, y0 = torch.floor(xs), torch.floor(ys)
            fx, fy = xs - x0, ys - y0
            ix, iy = x0.to(torch.int64), y0.to(torch.int64)
            s = seed + 101 * o
            n = (hash01(ix, iy, s) * (1 - fx) + hash01(ix + 1, iy, s) * fx) * (1 - fy) + \
                (hash01(ix, iy + 1, s) * (1 - fx) + hash01(ix + 1, iy + 1, s) * fx) * fy
            out = out + amp * (2.0 * n - 1.0) * 0.5
        return out

    frames = torch.empty((count, 2, height, width), dtype=torch.uint8, device=dev)
    xr = x - 12.0 - 4.0 * torch.sin(0.002 * (x + y))
    for i in range(count):
        cx, cy = camera_offset(k0 + i)
        frames[i, 0] = torch.clamp(torch.floor(tex(x + cx, y + cy) + 0.5), 0, 255).to(torch.uint8)
        frames[i, 1] = torch.clamp(torch.floor(tex(xr + cx, y + cy) + 0.5), 0, 255).to(torch.uint8)
    return frames
