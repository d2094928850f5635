"""Deterministic synthetic frames for parity tests and bench.py (SURVEY.md 8(d) generator).

Frame t of stream s: multi-octave bilinear value noise + random half-transparent rectangles + N(0,2) pixel
noise, rounded to u8.  Frames t>0 of a stream are the stream's base frame translated by a cumulative integer
offset (reflect fill) with fresh pixel noise, so consecutive frames have ground-truth correspondences.
numpy only (no cv2): the same bytes are produced in the build container and on the GPU box.
"""
import numpy as np


def _bilinear_up(grid, h, w, cell):
    gy = (np.arange(h, dtype=np.float64) + 0.5) / cell + 0.5
    gx = (np.arange(w, dtype=np.float64) + 0.5) / cell + 0.5
    y0 = np.floor(gy).astype(np.int64); x0 = np.floor(gx).astype(np.int64)
    fy = (gy - y0)[:, None]; fx = (gx - x0)[None, :]
    y0 = np.clip(y0, 0, grid.shape[0] - 2); x0 = np.clip(x0, 0, grid.shape[1] - 2)
    g00 = grid[y0][:, x0]; g01 = grid[y0][:, x0 + 1]; g10 = grid[y0 + 1][:, x0]; g11 = grid[y0 + 1][:, x0 + 1]
    return (g00 * (1 - fx) + g01 * fx) * (1 - fy) + (g10 * (1 - fx) + g11 * fx) * fy


def base_frame_f64(w, h, stream):
    rng = np.random.Generator(np.random.PCG64(1234 + 1000 * stream))
    img = np.zeros((h, w), np.float64)
    for cell, amp in ((4, 60.0), (8, 40.0), (16, 25.0), (32, 15.0)):
        grid = rng.random((h // cell + 3, w // cell + 3))
        img += amp * _bilinear_up(grid, h, w, cell)
    nrect = int(round(300.0 * (w * h) / (640.0 * 480.0)))
    for _ in range(nrect):
        rw, rh = rng.integers(6, 60, 2)
        x0 = int(rng.integers(0, max(1, w - rw))); y0 = int(rng.integers(0, max(1, h - rh)))
        g = float(rng.integers(0, 256))
        img[y0:y0 + rh, x0:x0 + rw] = 0.5 * img[y0:y0 + rh, x0:x0 + rw] + 0.5 * g
    return img


def stream_frames(w, h, stream, nframes):
    """Returns (frames u8 [nframes,h,w], offsets int [nframes,2] cumulative (dx,dy) w.r.t. frame 0)."""
    base = base_frame_f64(w, h, stream)
    pad = 8 * max(nframes, 1) + 8
    padded = np.pad(base, pad, mode="reflect")
    out = np.zeros((nframes, h, w), np.uint8)
    offs = np.zeros((nframes, 2), np.int64)
    ox = oy = 0
    for t in range(nframes):
        rng = np.random.Generator(np.random.PCG64(1234 + 1000 * stream + t + 1))
        if t > 0:
            dx, dy = rng.integers(-8, 9, 2)
            ox += int(dx); oy += int(dy)
        offs[t] = (ox, oy)
        view = padded[pad + oy:pad + oy + h, pad + ox:pad + ox + w]
        noisy = view + rng.normal(0.0, 2.0, (h, w))
        out[t] = np.clip(np.rint(noisy), 0, 255).astype(np.uint8)
    return out, offs


def batch(w, h, nframes, nstreams=1):
    """nstreams streams x (nframes // nstreams) consecutive frames, stream-major. u8 [nframes,h,w]."""
    per = nframes // nstreams
    assert per * nstreams == nframes
    return np.concatenate([stream_frames(w, h, s, per)[0] for s in range(nstreams)], axis=0)
