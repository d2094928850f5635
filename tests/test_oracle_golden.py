"""CPU oracle (oracle/) against the committed golden vectors made by the pinned reference stack
(cv2 4.13.0 + reference call pattern, tools/make_golden.py) -- bit-exact keypoints and descriptors."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po


def _same_kps(a, b):
    return len(a) == len(b) and all((a[f] == b[f]).all() for f in a.dtype.names)


@pytest.mark.parametrize("name", ["toy0", "toy2"])
@pytest.mark.parametrize("n", [1000, 2000])
def test_toy_frames_bit_exact(golden_dir, name, n):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kps, desc, ksz, ncand = po.orb32_extract(g["gray"], n)
    assert _same_kps(kps, g["kps%d" % n])
    assert (desc == g["desc%d" % n]).all()
    assert n <= len(kps) <= n + 3 * 8


@pytest.mark.parametrize("fn", ["synth_640x480_s0_t0", "synth_640x480_s0_t1", "synth_1280x720_s1_t0"])
def test_synthetic_frames_bit_exact(golden_dir, synth, fn):
    g = np.load(os.path.join(golden_dir, fn + ".npz"))
    fr, _ = synth.stream_frames(int(g["w"]), int(g["h"]), int(g["stream"]), int(g["t"]) + 1)
    img = fr[int(g["t"])]
    assert (int(img.astype(np.uint64).sum()) & 0xffffffff) == int(g["crc"]), "synthetic generator drifted"
    kps, desc, _, _ = po.orb32_extract(img, int(g["nfeatures"]))
    assert _same_kps(kps, g["kps"])
    assert (desc == g["desc"]).all()


def test_detect_stage_matches_cv2_detect(golden_dir):
    """cv::ORB::detect restated (pyramid, FAST+NMS, retainBest x2, Harris, IC angle) == cv2's raw output."""
    g = np.load(os.path.join(golden_dir, "toy0.npz"))
    det = g["det1000"]
    levels, ls = po.pyramid(g["gray"])
    q = po.features_per_level(10000)
    assert q.tolist()[0] == 2172
    total = 0
    for l in range(8):
        ref = det[det["octave"] == l]
        xs, ys, hr, fs = po.detect_level(levels[l], 20, q[l])
        mine = {}
        for x, y, r in zip(xs, ys, hr):
            key = (float(np.float32(x) * ls[l]), float(np.float32(y) * ls[l]))
            mine[key] = (float(r), po.ic_angle(levels[l], x, y), float(np.float32(31) * ls[l]))
        assert len(mine) == len(ref)
        for k in ref:
            assert mine[(float(k["x"]), float(k["y"]))] == (float(k["response"]), float(k["angle"]), float(k["size"]))
        total += len(ref)
    assert total == len(det) == 8539


def test_quota_and_geometry():
    assert po.features_per_level(1000).tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert po.features_per_level(2000).tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    lw, lh, _ = po.level_geometry(640, 480)
    assert lw.tolist() == [640, 533, 444, 370, 309, 257, 214, 179]
    assert lh.tolist() == [480, 400, 333, 278, 231, 193, 161, 134]
    lw, lh, _ = po.level_geometry(1280, 720)
    assert lw.tolist() == [1280, 1067, 889, 741, 617, 514, 429, 357]


def test_octree_edge_cases():
    # empty input, single key, all keys in one cell, N larger than keys
    assert len(po.octree([], [], [], 640, 480, 10)) == 0
    assert po.octree([5.0], [5.0], [1.0], 640, 480, 10).tolist() == [0]
    rng = np.random.default_rng(0)
    x = rng.uniform(3, 636, 50).astype(np.float32); y = rng.uniform(3, 476, 50).astype(np.float32)
    r = rng.uniform(0, 1, 50).astype(np.float32)
    keep = po.octree(x, y, r, 640, 480, 1000)
    assert sorted(keep.tolist()) == list(range(50))          # every key ends alone in its node
    keep = po.octree(x, y, r, 640, 480, 8)
    assert 8 <= len(keep) <= 11 and len(set(keep.tolist())) == len(keep)


def test_gray_conversion_matches_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert (po.gray_from_color(img, rgb=True) == cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)).all()
    assert (po.gray_from_color(img, rgb=False) == cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)).all()
    img4 = rng.integers(0, 256, (33, 47, 4), dtype=np.uint8)
    assert (po.gray_from_color(img4, rgb=True) == cv2.cvtColor(img4, cv2.COLOR_RGBA2GRAY)).all()


def test_gray_conversion_known_answers():
    # committed known answers (cv2 4.13.0): weights 9798/19235/3735, +2^14, >>15
    img = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [10, 200, 77], [128, 128, 128]]], np.uint8)
    assert po.gray_from_color(img, rgb=True).tolist() == [[76, 150, 29, 255, 129, 128]]
    assert po.gray_from_color(img, rgb=False).tolist() == [[29, 150, 76, 255, 142, 128]]
