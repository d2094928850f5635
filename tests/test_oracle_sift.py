"""CPU tests of the sift128 ORACLE (oracle/afv_oracle_sift.c).  PARITY UNPINNED: SiftGPU is not vendored by the
reference; the checks here are (a) the deterministic elementary functions, (b) structural invariants of the published
algorithm, (c) a family check against cv2.SIFT keypoints stored by tools/make_golden_sift.py, (d) the reference-side
post-processing (octave rule, octree quota, merge order, computeSize)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def frame(synth):
    frames, _ = synth.stream_frames(640, 480, 0, 1)
    return frames[0]


def test_elementary_functions():
    L = po.lib()
    for f in (L.orc_sift_exp, L.orc_sift_exp2):
        f.restype = C.c_float; f.argtypes = [C.c_float]
    L.orc_sift_atan2.restype = C.c_float; L.orc_sift_atan2.argtypes = [C.c_float, C.c_float]
    xs = np.linspace(-20, 0, 4001)
    e = np.array([L.orc_sift_exp(float(x)) for x in xs])
    assert np.max(np.abs(e - np.exp(xs)) / np.exp(xs)) < 5e-6
    th = np.linspace(0, 2 * np.pi, 4001)[:-1]
    a = np.array([L.orc_sift_atan2(float(np.sin(t)), float(np.cos(t))) for t in th])
    assert np.max(np.abs(a - th)) < 2e-6
    s = C.c_float(); c = C.c_float()
    for t in th[::7]:
        L.orc_sift_sincos(C.c_float(t), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(t)) < 2e-6 and abs(c.value - np.cos(t)) < 2e-6


def test_scale_space_structure(frame):
    g0 = po.sift_scale_space(frame, 0, 0, 0)
    g3 = po.sift_scale_space(frame, 0, 0, 3)
    g1_0 = po.sift_scale_space(frame, 0, 1, 0)
    assert g0.shape == (480, 640) and g1_0.shape == (240, 320)
    assert (g1_0 == g3[::2, ::2]).all()                       # next octave = every 2nd pixel of level 2
    d2 = po.sift_scale_space(frame, 1, 0, 2)
    g2 = po.sift_scale_space(frame, 0, 0, 2)
    assert (d2 == g3 - g2).all()
    # level -1 has sigma 1.6: compare with a double-precision Gaussian of sqrt(1.6^2 - 0.5^2)
    import scipy.ndimage as ndi
    ref = ndi.gaussian_filter(frame.astype(np.float64) / 255.0, np.sqrt(1.6 ** 2 - 0.25), mode="nearest", truncate=4.5)
    assert np.abs(ref - g0).max() < 2e-3


def test_detect_invariants(frame):
    xyso, desc = po.sift_detect(frame, 1000)
    assert len(xyso) > 1000                                    # soft limit: the level that crosses it is kept whole
    assert np.allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-5)      # unit L2 (matchingTh 0.5 assumes it)
    assert desc.min() >= 0 and desc.max() < 0.6
    assert (xyso[:, 3] >= 0).all() and (xyso[:, 3] < 2 * np.pi).all()
    assert (xyso[:, 2] > 1.5).all()
    full, _ = po.sift_detect(frame, 10 ** 7, with_desc=False)
    # -tc2 keeps a suffix of the list (coarsest levels), list order preserved
    assert (full[len(full) - len(xyso):] == xyso).all()
    # deterministic
    again, d2 = po.sift_detect(frame, 1000)
    assert (again == xyso).all() and (d2 == desc).all()


def test_family_check_against_cv2(frame, golden_dir):
    g = np.load(os.path.join(golden_dir, "sift_cv2_synth_640x480_s0_t0.npz"))
    full, _ = po.sift_detect(frame, 10 ** 7, with_desc=False)
    from scipy.spatial import cKDTree
    tree = cKDTree(full[:, :2])
    for o in (0, 1, 2):
        m = g["octave"] == o
        d, i = tree.query(g["xys"][m, :2])
        hit = d < 1.0
        assert hit.mean() > 0.8, (o, hit.mean())             # cv2's octave >= 0 extrema are found
        ratio = g["xys"][m, 2][hit] / 2.0 / full[i[hit], 2]
        assert abs(np.median(ratio) - 1.0) < 0.02             # and at the same scale (cv2 size = 2 sigma)


def _match_cv2(full, g, octaves, pos_tol, dlog=0.15):
    """For every cv2 keypoint of the given octaves: index of the nearest oracle keypoint of the same scale (|log ratio| < dlog,
    cv2 size = 2 sigma) and its distance; inf where there is none."""
    idx = np.where(np.isin(g["octave"], octaves))[0]
    best = np.full(len(idx), -1, np.int64); dist = np.full(len(idx), np.inf)
    for n, q in enumerate(idx):
        d = np.hypot(full[:, 0] - g["xys"][q, 0], full[:, 1] - g["xys"][q, 1])
        d = np.where(np.abs(np.log(full[:, 2] / (g["xys"][q, 2] / 2.0))) < dlog, d, np.inf)
        best[n] = int(np.argmin(d)); dist[n] = d[best[n]]
    return idx, best, dist


def test_cv2_default_upscale_bias_explains_the_loose_family_check(frame, golden_dir):
    """Why the check above only holds at 1 px: cv2's default 2x upsampling carries a constant +0.25 px shift in x and y."""
    g = np.load(os.path.join(golden_dir, "sift_cv2_synth_640x480_s0_t0.npz"))
    full, _ = po.sift_detect(frame, 10 ** 7, with_desc=False)
    idx, best, dist = _match_cv2(full, g, (0, 1), 1.0)
    hit = dist < 1.0
    dxy = full[best[hit], :2] - g["xys"][idx[hit], :2]
    assert np.all(np.abs(dxy.mean(axis=0) + 0.24) < 0.03) and np.all(dxy.std(axis=0) < 0.13)


def test_pinned_to_cv2_precise_per_octave(frame, golden_dir):
    """Position, scale and orientation against cv2.SIFT(enable_precise_upscale=True), octave by octave, with the residual
    classified.  SiftGPU is not OpenCV's SIFT (one Newton step instead of up to five, its own first octave, threshold 0.02 / 3
    against 0.04 / 3), so this is agreement between two implementations of one published algorithm, not bit parity:
      * octave 0 / 1: > 85 % / 90 % of cv2's keypoints have an oracle keypoint of the same scale within 0.25 px;
        no systematic offset (|mean| < 0.03 px), 0.08 - 0.10 px standard deviation per axis; scale ratio median within 1 %;
      * orientation of those pairs: no offset (|median| < 0.5 degrees), 2.2 degrees of spread.  (This check is what found the
        oracle -- and the CUDA kernel -- mapping a histogram peak to the lower edge of its 10-degree bin instead of the centre:
        a constant -5.0 degrees against cv2, sift++ and Lowe.  Corrected in both.);
      * what cv2 finds and the oracle does not: two thirds sit on the FINEST layer of octave 0 (sigma ~ 2, where the two ways of
        building the first octave -- blur of the input vs. decimated blur of the upsampled input -- differ most), the rest are
        displaced by 0.5 - 1.5 px at the same scale (refinement) or sit one layer off."""
    g = np.load(os.path.join(golden_dir, "sift_cv2_precise_synth_640x480_s0_t0.npz"))
    full, _ = po.sift_detect(frame, 10 ** 7, with_desc=False)
    for o, need in ((0, 0.85), (1, 0.90)):
        idx, best, dist = _match_cv2(full, g, (o,), 0.25)
        hit = dist < 0.25
        assert hit.mean() > need, (o, hit.mean())
        dxy = full[best[hit], :2] - g["xys"][idx[hit], :2]
        assert np.all(np.abs(dxy.mean(axis=0)) < 0.03) and np.all(dxy.std(axis=0) < 0.12), (o, dxy.mean(axis=0), dxy.std(axis=0))
        ratio = full[best[hit], 2] / (g["xys"][idx[hit], 2] / 2.0)
        assert abs(np.median(ratio) - 1.0) < 0.01
        # orientation: nearest of the (up to two) orientations the oracle emits at that position
        sd = []
        for q, b in zip(idx[hit], best[hit]):
            same = np.where((np.abs(full[:, 0] - full[b, 0]) < 1e-4) & (np.abs(full[:, 1] - full[b, 1]) < 1e-4) & (np.abs(full[:, 2] - full[b, 2]) < 1e-4))[0]
            dd = (np.degrees(full[same, 3]) - g["angle"][q] + 180.0) % 360.0 - 180.0
            sd.append(dd[np.argmin(np.abs(dd))])
        sd = np.array(sd)
        core = sd[np.abs(sd) < 15.0]
        assert len(core) > 0.9 * len(sd)                              # the rest: a second peak only one of the two kept
        assert abs(np.median(core)) < 0.5 and core.std() < 3.0, (o, np.median(core), core.std())
    # residual of octave 0
    idx, best, dist = _match_cv2(full, g, (0,), 0.5)
    miss = idx[dist >= 0.5]
    assert len(miss) < 0.13 * len(idx)
    finest = (g["layer"][miss] == 1).mean()
    assert finest > 0.55, finest
    near_same_scale = (dist[dist >= 0.5] < 1.5).mean()
    assert near_same_scale < 0.25                                      # most misses have no counterpart at all, they are not mislocalised


def test_descriptors_are_cv2_descriptors_up_to_the_orientation_bin_direction(frame, golden_dir):
    """At keypoints both implementations place within 0.25 px, 10 % in scale and 3 degrees, the oracle's 128 floats are cv2's (L2-normalised)
    descriptor with the 8 orientation bins counted in the opposite direction -- bin o <-> (8 - o) mod 8, the sift++ / VLFeat
    layout against Lowe's, same 4 x 4 spatial order: median cosine similarity 0.997.  Without that permutation it is 0.53, so the
    check is sensitive to the layout; it says the histogramming (trilinear weights, Gaussian window, clamp 0.2, renormalise) is SIFT's."""
    g = np.load(os.path.join(golden_dir, "sift_cv2_precise_synth_640x480_s0_t0.npz"))
    full, desc = po.sift_detect(frame, 10 ** 7, with_desc=True)
    a, b = [], []
    for q in np.where((g["octave"] >= 0) & (g["octave"] <= 1))[0]:
        d = np.hypot(full[:, 0] - g["xys"][q, 0], full[:, 1] - g["xys"][q, 1])
        ls = np.abs(np.log(full[:, 2] / (g["xys"][q, 2] / 2.0)))
        da = np.abs((np.degrees(full[:, 3]) - g["angle"][q] + 180.0) % 360.0 - 180.0)
        c = np.where((d < 0.25) & (ls < 0.1) & (da < 3.0))[0]
        if len(c):
            a.append(g["desc"][q].astype(np.float32)); b.append(desc[c[0]])
    a = np.array(a); b = np.array(b)
    assert len(a) > 700
    a /= np.linalg.norm(a, axis=1, keepdims=True); b /= np.linalg.norm(b, axis=1, keepdims=True)
    direct = np.median((a * b).sum(1))
    perm = np.roll(b.reshape(-1, 4, 4, 8)[..., ::-1], 1, axis=3).reshape(-1, 128)
    cos = (a * perm).sum(1)
    assert direct < 0.7 and np.median(cos) > 0.99 and np.percentile(cos, 10) > 0.95, (direct, np.median(cos), np.percentile(cos, 10))


def test_rotation_180(frame):
    """The detector is symmetric under a 180 degree rotation (clamped borders, symmetric taps)."""
    a, _ = po.sift_detect(frame, 10 ** 7, with_desc=False)
    b, _ = po.sift_detect(np.ascontiguousarray(frame[::-1, ::-1]), 10 ** 7, with_desc=False)
    h, w = frame.shape
    pa = {(round(float(x), 2), round(float(y), 2)) for x, y in a[a[:, 2] < 3.2, :2]}
    pb = {(round(float(w - 1 - x), 2), round(float(h - 1 - y), 2)) for x, y in b[b[:, 2] < 3.2, :2]}
    assert len(pa & pb) > 0.98 * len(pa)


def test_extract_postprocessing(frame):
    kps, desc, size, nd = po.sift128_extract(frame, 1000)
    xyso, dfull = po.sift_detect(frame, 1000)
    assert nd == len(xyso)
    q = po.features_per_level(1000, 8, 2.0)
    assert list(q) == [502, 251, 125, 63, 31, 16, 8, 4]
    oc = kps["octave"]
    assert (np.diff(oc) >= 0).all()                            # mergeKeypointLevels: ascending level
    for l in range(8):
        assert (oc == l).sum() <= q[l] + 3
    cid = kps["class_id"]
    assert len(set(cid.tolist())) == len(cid)
    assert (kps["x"] == xyso[cid, 0]).all() and (kps["size"] == xyso[cid, 2]).all() and (kps["angle"] == xyso[cid, 3]).all()
    assert (desc == dfull[cid]).all()                          # computeDescriptors: row gather by class_id
    assert (kps["response"] == 1.0).all()
    expect_oct = np.floor(np.maximum(np.log2(kps["size"].astype(np.float64) / 1.6454), 0)).astype(int)
    assert (oc == expect_oct).all()                            # src/Feature_sift128.cpp:92
    assert np.allclose(size, 2.0 ** oc, rtol=1e-6)
