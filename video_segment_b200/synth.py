"""Deterministic synthetic video generator.

``synth(seed, W, H, T)`` yields T uint8 BGR frames (H, W, 3), C-contiguous:
smooth background (per-channel bilinear upsampling of a 9x16 grid of
uniform[40, 215] values), K = 48 moving shapes (half axis-aligned rectangles,
half ellipses; sizes uniform[2 %, 15 %] of min(W, H); colours uniform[0, 255]^3;
constant velocities uniform[-3, 3] px/frame with wrap-around; painted in index
order) plus i.i.d. Gaussian noise sigma = 2.0 (clipped, rounded).  Identical
bytes for the CPU oracle and the GPU path.  ``synth_flow`` rasterises the exact
generating velocities to a backward flow field (float32 x, y; zero on frame 0).
"""
from __future__ import annotations

import numpy as np

K_SHAPES = 48


def _background(rng, W, H):
    grid = rng.uniform(40.0, 215.0, size=(9, 16, 3)).astype(np.float32)
    ys = np.linspace(0.0, 8.0, H, dtype=np.float32)
    xs = np.linspace(0.0, 15.0, W, dtype=np.float32)
    y0 = np.clip(np.floor(ys).astype(np.int32), 0, 7)
    x0 = np.clip(np.floor(xs).astype(np.int32), 0, 14)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    g00 = grid[y0][:, x0]
    g01 = grid[y0][:, x0 + 1]
    g10 = grid[y0 + 1][:, x0]
    g11 = grid[y0 + 1][:, x0 + 1]
    return (g00 * (1 - fy) * (1 - fx) + g01 * (1 - fy) * fx + g10 * fy * (1 - fx) + g11 * fy * fx).astype(np.float32)


class _Scene:
    def __init__(self, seed, W, H):
        rng = np.random.default_rng(seed)
        self.W, self.H = W, H
        self.bg = _background(rng, W, H)
        m = float(min(W, H))
        self.size = rng.uniform(0.02 * m, 0.15 * m, size=(K_SHAPES, 2))      # half-extent-ish (w, h)
        self.color = rng.uniform(0.0, 255.0, size=(K_SHAPES, 3)).astype(np.float32)
        self.pos0 = np.stack([rng.uniform(0, W, K_SHAPES), rng.uniform(0, H, K_SHAPES)], 1)
        self.vel = rng.uniform(-3.0, 3.0, size=(K_SHAPES, 2))
        self.seed = seed

    def _paint(self, img, k, t, vel_img=None):
        W, H = self.W, self.H
        cx = (self.pos0[k, 0] + self.vel[k, 0] * t) % W
        cy = (self.pos0[k, 1] + self.vel[k, 1] * t) % H
        hw, hh = self.size[k, 0] * 0.5, self.size[k, 1] * 0.5
        for ox in (-W, 0, W):
            for oy in (-H, 0, H):
                x0, x1 = int(np.floor(cx + ox - hw)), int(np.ceil(cx + ox + hw))
                y0, y1 = int(np.floor(cy + oy - hh)), int(np.ceil(cy + oy + hh))
                xa, xb = max(x0, 0), min(x1, W)
                ya, yb = max(y0, 0), min(y1, H)
                if xa >= xb or ya >= yb:
                    continue
                if k % 2 == 0:   # rectangle
                    img[ya:yb, xa:xb] = self.color[k]
                    if vel_img is not None:
                        vel_img[ya:yb, xa:xb] = self.vel[k]
                else:            # ellipse
                    yy = (np.arange(ya, yb, dtype=np.float32) - (cy + oy)) / max(hh, 1e-3)
                    xx = (np.arange(xa, xb, dtype=np.float32) - (cx + ox)) / max(hw, 1e-3)
                    mask = (yy[:, None] ** 2 + xx[None, :] ** 2) <= 1.0
                    img[ya:yb, xa:xb][mask] = self.color[k]
                    if vel_img is not None:
                        vel_img[ya:yb, xa:xb][mask] = self.vel[k]

    def frame(self, t, with_flow=False):
        img = self.bg.copy()
        vel_img = np.zeros((self.H, self.W, 2), np.float32) if with_flow else None
        for k in range(K_SHAPES):
            self._paint(img, k, t, vel_img)
        # per-frame noise stream keyed by (seed, t): frames can be generated at any offset, so every
        # rank of a sharded run produces its own segment of the same video
        noise_rng = np.random.default_rng([self.seed, 0x5EED, t])
        img += noise_rng.normal(0.0, 2.0, size=img.shape).astype(np.float32)
        out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        if with_flow:
            # backward flow: where the pixel was in frame t-1 (zero on frame 0)
            flow = -vel_img if t > 0 else np.zeros_like(vel_img)
            return np.ascontiguousarray(out), np.ascontiguousarray(flow)
        return np.ascontiguousarray(out)


def synth(seed: int, W: int, H: int, T: int, start: int = 0):
    """Generator of T uint8 BGR frames, (H, W, 3), frames start .. start + T - 1 of the video."""
    sc = _Scene(seed, W, H)
    for t in range(start, start + T):
        yield sc.frame(t)


def synth_flow(seed: int, W: int, H: int, T: int):
    """Generator of (frame, backward_flow) pairs."""
    sc = _Scene(seed, W, H)
    for t in range(T):
        yield sc.frame(t, with_flow=True)


def synth_clip(seed: int, W: int, H: int, T: int) -> np.ndarray:
    """All frames stacked: uint8 (T, H, W, 3)."""
    return np.stack(list(synth(seed, W, H, T)), 0)
