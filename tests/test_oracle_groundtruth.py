"""Functional check of the three extraction oracles on ground truth (CPU only): consecutive synthetic frames of a stream
are integer translations of each other (synth.stream_frames returns the cumulative offsets), so a correct
detector + descriptor must (a) re-detect most keypoints at the shifted position and (b) match them by descriptor.
This is independent of any third-party implementation: it guards the restated algorithms against plausible-looking but
wrong arithmetic (mirrored patterns, wrong rotation sense, mis-indexed histogram bins)."""
import numpy as np
import pytest

from oracle import pyoracle as po


def _matches(d0, d1, metric, ratio):
    if metric == "hamming":
        b0 = np.unpackbits(d0, axis=1).astype(np.int16); b1 = np.unpackbits(d1, axis=1).astype(np.int16)
        dist = (b0[:, None, :] != b1[None, :, :]).sum(-1).astype(np.float32)
    else:
        dist = ((d0[:, None, :] - d1[None, :, :]) ** 2).sum(-1)
    order = np.argsort(dist, axis=1)
    best, second = order[:, 0], order[:, 1]
    r = np.arange(len(d0))
    good = dist[r, best] < ratio * dist[r, second]
    return r[good], best[good]


@pytest.mark.parametrize("feature", ["orb32", "sift128", "akaze61"])
def test_descriptors_match_ground_truth_translation(synth, feature):
    frames, offs = synth.stream_frames(640, 480, 11, 2)
    shift = (offs[1] - offs[0]).astype(np.float64)             # content of frame 0 at (x, y) appears in frame 1 at (x, y) - shift
    if feature == "orb32":
        ex = lambda im: po.orb32_extract(im, 1000)[:2]
        metric, ratio, tol = "hamming", 0.8, 2.5
    elif feature == "sift128":
        ex = lambda im: po.sift128_extract(im, 1000)[:2]
        metric, ratio, tol = "l2", 0.64, 1.5                   # ratio on squared distances (0.8^2)
    else:
        ex = lambda im: po.akaze61_extract(im, 1000)[:2]
        metric, ratio, tol = "hamming", 0.8, 1.5
    k0, d0 = ex(frames[0]); k1, d1 = ex(frames[1])
    i0, i1 = _matches(d0, d1, metric, ratio)
    assert len(i0) >= 150, len(i0)
    dx = k1["x"][i1] - k0["x"][i0] + shift[0]; dy = k1["y"][i1] - k0["y"][i0] + shift[1]
    scale = np.maximum(1.0, k0["size"][i0] / (31.0 if feature == "orb32" else 4.0))
    ok = np.hypot(dx, dy) < tol * scale
    assert ok.mean() > 0.95, (feature, float(ok.mean()), len(i0))
    # orientation is consistent between the two views of the same point (pure translation): within 10 degrees
    a0, a1 = k0["angle"][i0][ok], k1["angle"][i1][ok]
    if feature != "orb32":
        a0, a1 = np.degrees(a0), np.degrees(a1)                # sift128 / akaze61 report radians (reference quirk)
    da = np.abs((a1 - a0 + 180.0) % 360.0 - 180.0)
    assert np.median(da) < 5.0 and (da < 15.0).mean() > 0.9
