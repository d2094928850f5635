"""CPU oracle of the FeatureMatcher path: definitional checks (distances, grid, window search) and the
behavioural properties of SearchForInitialization / SearchByBoW on frames with known correspondences."""
import numpy as np
import pytest

from oracle import pyoracle as po


def _popcount(a):
    return int(np.unpackbits(a).sum())


def test_descriptor_distances():
    rng = np.random.default_rng(0)
    for dt, nb in ((po.DESC_ORB, 32), (po.DESC_AKAZE61, 61), (po.DESC_BRISK, 48)):
        a = rng.integers(0, 256, nb, dtype=np.uint8); b = rng.integers(0, 256, nb, dtype=np.uint8)
        assert po.descriptor_distance(dt, a, b) == float(_popcount(a ^ b))
        assert po.descriptor_distance(dt, a, a) == 0.0
    a = rng.normal(size=128).astype(np.float32); b = rng.normal(size=128).astype(np.float32)
    ref = float(np.sum((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    assert abs(po.descriptor_distance(po.DESC_SIFT128, a, b) - ref) <= 1e-5 * ref


def test_reference_bit_hack_equals_popcount():
    # DescriptorDistance_orb32 (reference src/Feature_orb32.cpp:67-83) is the SWAR popcount of 8 int32 words
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32); b = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32)
    dist = 0
    for x, y in zip(a, b):
        v = int(x ^ y)
        v = v - ((v >> 1) & 0x55555555)
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333)
        dist += ((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) & 0xffffffff) >> 24
    assert po.descriptor_distance(po.DESC_ORB, a.view(np.uint8), b.view(np.uint8)) == float(dist)


def _extract_pair(synth, t0=0):
    fr, offs = synth.stream_frames(640, 480, 0, t0 + 2)
    out = [po.orb32_extract(fr[t0 + i], 1000) for i in range(2)]
    d = offs[t0 + 1] - offs[t0]
    return out, (int(d[0]), int(d[1]))


def test_search_for_initialization_recovers_translation(synth):
    (k1, d1, s1, _), (k2, d2, s2, _) = _extract_pair(synth)[0]
    _, (dx, dy) = _extract_pair(synth)
    prev = np.stack([k1["x"], k1["y"]], axis=1)
    n, m12, pm = po.search_for_initialization(0, k1, d1, k2, d2, s2, (0, 0, 640, 480), float(np.float32(1.2) ** 7), prev,
                                              window=100, th_low=75.0, nnratio=0.9, check_ori=True)
    idx = np.nonzero(m12 >= 0)[0]
    assert n == len(idx) and n > 50
    assert (k1["octave"][idx] == 0).all()                       # only octave-0 queries (:491-493)
    # frame t+1 is frame t shifted by (dx,dy): a point at x in frame t+1 shows what was at x+dx in the base
    ex = k2["x"][m12[idx]] - k1["x"][idx]; ey = k2["y"][m12[idx]] - k1["y"][idx]
    good = (np.abs(ex + dx) <= 1.5) & (np.abs(ey + dy) <= 1.5)
    assert good.mean() > 0.8
    assert len(set(m12[idx].tolist())) == n                     # one-to-one after stealing
    assert (pm[idx, 0] == k2["x"][m12[idx]]).all()               # vbPrevMatched update (:552-554)


def test_window_and_bruteforce_consistency(synth):
    (k1, d1, s1, _), (k2, d2, s2, _) = _extract_pair(synth)[0]
    qxy = np.stack([k1["x"], k1["y"]], axis=1)
    r = np.full(len(k1), 1e6, np.float32)
    best, bd, sd, bs, ss = po.match_window(0, d1, qxy, r, np.zeros(len(k1), np.float32), np.full(len(k1), 1e9, np.float32),
                                           k2, d2, s2, (0, 0, 640, 480))
    bb, bbd, bsd = po.match_bruteforce(0, d1, d2)
    # an unbounded window sees every keypoint that landed in the grid: same best distance as brute force
    assert (bd == bbd).all() and (sd == bsd).all()
    # empty inputs
    b0 = po.match_bruteforce(0, d1[:0], d2)
    assert len(b0[0]) == 0


def test_grid_quirk_round_vs_floor():
    # PosInGrid uses round(): x = 639.9 -> cell 64 -> dropped from the grid (reference src/Frame.cc:386-392)
    k = np.zeros(2, po.KP_DTYPE); k["x"] = [639.9, 10.0]; k["y"] = [10.0, 10.0]
    d = np.zeros((2, 32), np.uint8); s = np.ones(2, np.float32)
    best, bd, sd, _, _ = po.match_window(0, d[:1], np.array([[639.0, 10.0]], np.float32), np.array([5.0], np.float32),
                                         np.zeros(1, np.float32), np.full(1, 9.0, np.float32), k, d, s, (0, 0, 640, 480))
    assert best[0] == -1


def test_search_by_bow_buckets(synth):
    (k1, d1, s1, _), (k2, d2, s2, _) = _extract_pair(synth)[0]
    rng = np.random.default_rng(3)
    # synthetic FeatureVectors: bucket = coarse position hash so true matches share a node
    def segs(k, shift):
        node = ((k["x"] + shift[0]) // 80).astype(np.int32) * 10 + ((k["y"] + shift[1]) // 80).astype(np.int32)
        order = np.argsort(node, kind="stable")
        ids, starts = np.unique(node[order], return_index=True)
        return ids.astype(np.int32), np.append(starts, len(k)).astype(np.int32), order.astype(np.int32)
    _, (dx, dy) = _extract_pair(synth)
    n, mf = po.search_by_bow(0, d1, k1, segs(k1, (0, 0)), d2, k2, segs(k2, (dx, dy)), th_low=75.0, nnratio=0.7, check_ori=True)
    idx = np.nonzero(mf >= 0)[0]
    assert n == len(idx) and n > 100
    assert len(set(mf[idx].tolist())) <= n


def test_bow_transform_oracle_properties():
    from tests_bow import make_tree
    rng = np.random.default_rng(7)
    tree = make_tree(rng, k=10, L=3, D=32)
    leaves = np.nonzero(tree["node_word"] >= 0)[0]
    # a feature equal to a leaf's descriptor whose ancestors are also nearest descends to that leaf (distance 0 at the end)
    feats = tree["node_desc"][leaves[::37]]
    wid, w, nid = po.bow_transform(0, feats, tree, levelsup=1)
    assert (wid >= 0).all() and (w > 0).all()
    # node ids at level L - levelsup are inner nodes of depth 2
    assert ((nid >= 11) & (nid <= 110)).all()
    # levelsup >= L -> root
    assert (po.bow_transform(0, feats, tree, levelsup=4)[2] == 0).all()
    ids, starts, order = po.feature_vector_segments(nid)
    assert (np.diff(ids) > 0).all() and starts[-1] == len(nid)


def test_undistort_keypoints_matches_cv2_golden(golden_dir):
    """Frame::UndistortKeyPoints (src/Frame.cc:403-433): the oracle is bit-identical with cv2 4.13.0's undistortPoints."""
    import os
    g = np.load(os.path.join(golden_dir, "undistort_cv2.npz"))
    pts = g["pts"]
    kps = np.zeros(len(pts), po.KP_DTYPE)
    kps["x"] = pts[:, 0]; kps["y"] = pts[:, 1]; kps["size"] = 31.0; kps["angle"] = 12.5; kps["octave"] = 3; kps["class_id"] = -1
    for i in range(3):
        out = po.undistort_keypoints(kps, g["K%d" % i], g["D%d" % i])
        assert (out["x"] == g["und%d" % i][:, 0]).all() and (out["y"] == g["und%d" % i][:, 1]).all()
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert (out[f] == kps[f]).all()
    assert (po.undistort_keypoints(kps, g["K2"], g["D2"])["x"] == kps["x"]).all()       # zero distortion: copy (:405-409)
