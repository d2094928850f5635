"""CPU tests of the brisk48 ORACLE (oracle/afv_oracle_brisk.c).  PARITY UNPINNED vs ETH brisk v2 (not vendored by the
reference); the detector, the orientation and the 512-bit descriptor core are pinned to cv2 4.13.0's cv::BRISK (the BRISK
authors' implementation of the published algorithm) through fixtures made by tools/make_golden_brisk.py.  The 48-byte pair
table is a documented stand-in (the 384 shortest of the paper's 512 short pairs)."""
import os
import zlib

import numpy as np
import pytest

from oracle import pyoracle as po

CASES = [("brisk_cv2_synth_640x480_s0_t0.npz", ("synth", 640, 480, 0)), ("brisk_cv2_synth_640x480_s2_t0.npz", ("synth", 640, 480, 2)),
         ("brisk_cv2_toy0.npz", ("toy",))]


def _image(case, synth, golden_dir):
    if case[0] == "synth":
        return synth.stream_frames(case[1], case[2], case[3], 1)[0][0]
    return np.ascontiguousarray(np.load(os.path.join(golden_dir, "toy0.npz"))["gray"])


def _detect_as_kps(img, mode):
    d = po.brisk_detect(img, 34, 4, mode)
    k = np.zeros(len(d), po.KP_DTYPE)
    k["x"] = d[:, 0]; k["y"] = d[:, 1]; k["size"] = d[:, 2]; k["response"] = d[:, 3]; k["octave"] = d[:, 4].astype(np.int32)
    k["angle"] = -1; k["class_id"] = -1
    return k


@pytest.mark.parametrize("name,case", CASES)
def test_pyramid_and_agast_pinned_to_cv2(name, case, synth, golden_dir):
    g = np.load(os.path.join(golden_dir, name))
    img = _image(case, synth, golden_dir)
    for i in range(8):                                   # cv::resize INTER_AREA: 2/3 sample, exact and inexact half samples
        lay = po.brisk_layer(img, 0, i)
        assert zlib.crc32(lay.tobytes()) == int(g["layer_crc"][i]), "layer %d differs from cv2" % i
    assert (po.brisk_layer(img, 0, 5) == g["layer5"]).all() and (po.brisk_layer(img, 0, 7) == g["layer7"]).all()
    for i in (1, 4):                                     # OAST 9-16 detections at threshold 34 and their corner scores
        sc = po.brisk_layer(img, 1, i)
        ag = g["agast%d" % i]
        ys, xs = np.nonzero(sc >= 34)
        assert len(ag) == len(xs)
        assert (ag[:, 0] == xs).all() and (ag[:, 1] == ys).all(), "raster-ordered detections"
        assert (sc[ys, xs] == ag[:, 2]).all(), "corner scores"


@pytest.mark.parametrize("name,case", CASES)
def test_detector_and_descriptor_core_pinned_to_cv2(name, case, synth, golden_dir):
    """Sequential-cache mode: every keypoint field except the last bits of the angle, and the 512-bit descriptors, equal cv2's."""
    g = np.load(os.path.join(golden_dir, name))
    img = _image(case, synth, golden_dir)
    k = _detect_as_kps(img, po.BRISK_SEQUENTIAL)
    mk, md, _ = po.brisk_describe(img, k, 64, libm_angle=0)
    ck = g["kp"]
    assert len(mk) == len(ck), "same keypoints survive the border filter"
    assert (mk["x"] == ck[:, 0]).all() and (mk["y"] == ck[:, 1]).all() and (mk["size"] == ck[:, 2]).all()
    assert (mk["response"] == ck[:, 4]).all() and (mk["octave"] == ck[:, 5].astype(np.int32)).all()
    # orientation: > 90 % of the angles are bit-identical, ~99 % within 3 float ulps (cv2 evaluates atan2 through its libm, the
    # oracle through a double-precision polynomial); a handful of large-scale keypoints differ by up to ~1e-3 degrees (one
    # smoothed sample off by one grey level), which never changes the 1024-step pattern rotation: the descriptors are identical
    da = np.abs(mk["angle"] - ck[:, 3]); da = np.minimum(da, 360.0 - da)
    assert da.max() < 0.02 and np.quantile(da, 0.98) < 1e-4 and (mk["angle"] == ck[:, 3]).mean() > 0.9
    neq = (md != g["desc"]).any(axis=1).sum()
    assert neq <= max(1, len(mk) // 500), "512-bit descriptors: %d of %d rows differ" % (neq, len(mk))


def test_dense_contract_differs_only_in_tie_decisions(synth):
    """ORC_BRISK_DENSE (the order-independent contract of the CUDA path) vs the sequential score cache: same score maps, a few
    percent of the keypoints differ (exact ties between neighbouring maxima on synthetic rectangles); all common keypoints
    have identical fields."""
    img = synth.stream_frames(640, 480, 0, 1)[0][0]
    a = _detect_as_kps(img, po.BRISK_SEQUENTIAL); b = _detect_as_kps(img, po.BRISK_DENSE)
    ka = {(float(p["x"]), float(p["y"]), int(p["octave"])): p for p in a}
    kb = {(float(p["x"]), float(p["y"]), int(p["octave"])): p for p in b}
    common = set(ka) & set(kb)
    assert len(common) > 0.95 * max(len(ka), len(kb))
    for key in common:
        assert ka[key]["size"] == kb[key]["size"] and ka[key]["response"] == kb[key]["response"]


def test_resize_area_shapes():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    half = po.resize_area(img[:96, :130], 65, 48)
    ref = (img[0:96:2, 0:130:2].astype(np.int32) + img[0:96:2, 1:130:2] + img[1:96:2, 0:130:2] + img[1:96:2, 1:130:2] + 2) >> 2
    assert (half == ref).all()
    tt = po.resize_area(img, 2 * (131 // 3), 2 * (97 // 3))
    assert tt.shape == (64, 86)
    assert abs(float(tt.mean()) - float(img.mean())) < 1.0


def test_pattern_and_scale_index():
    sizes = np.zeros(64, np.uint32)
    po.lib().orc_brisk_size_list(po._p(sizes))
    assert sizes[0] == 13 and (np.diff(sizes.astype(np.int64)) >= 0).all()
    assert po.brisk_scale_index(7.2) == 0 and po.brisk_scale_index(1.0) == 0 and po.brisk_scale_index(1e6) == 63
    prev = 0
    for s in np.linspace(7.2, 220.0, 400):
        i = po.brisk_scale_index(float(s)); assert i >= prev; prev = i
    at = po.lib().orc_brisk_atan2
    import ctypes as C
    at.restype = C.c_double; at.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(5)
    for _ in range(2000):
        y, x = (float(v) for v in rng.integers(-3000, 3000, 2))
        if x == 0 and y == 0:
            continue
        assert abs(at(y, x) - np.arctan2(y, x)) < 4e-16 * max(1.0, abs(np.arctan2(y, x)))


def test_brisk48_extract_reference_glue(synth):
    """FeatureExtractor_brisk48::operator(): per-layer octree quota, levels ascending, border keypoints removed by compute,
    48-byte rows = the selected 384 of the 512 short-pair bits, computeSize = 1.5^octave mapped by the settings."""
    img = synth.stream_frames(640, 480, 0, 1)[0][0]
    kps, desc, size, ndet = po.brisk48_extract(img, 1000)
    assert desc.shape[1] == 48 and len(kps) == len(desc) == len(size) and 0 < len(kps) <= 1000 + 24
    assert (np.diff(kps["octave"]) >= 0).all()
    q = po.features_per_level(1000, 8, 1.5)
    for l in range(8):
        assert (kps["octave"] == l).sum() <= q[l] + 3
    assert (kps["angle"] >= 0).all() and (kps["angle"] < 360).all() and (kps["class_id"] == -1).all()
    # the 48-byte rows are a sub-selection of the 64-byte rows of the same keypoints
    k64, d64, _ = po.brisk_describe(img, kps, 64, libm_angle=0)
    assert len(k64) == len(kps)
    b48 = np.unpackbits(desc, axis=1, bitorder="little"); b64 = np.unpackbits(d64, axis=1, bitorder="little")
    assert b48.shape[1] == 384
    # every 48-byte bit column equals some 64-byte column, in increasing order (enumeration order kept)
    j = 0
    for c in range(384):
        while j < 512 and not (b64[:, j] == b48[:, c]).all():
            j += 1
        assert j < 512, "bit %d not found in order" % c
        j += 1
    mx = np.float32(1.2) ** np.float32(7)
    exp = 1.0 + (np.float32(1.5) ** kps["octave"].astype(np.float32) - 1.0) * (mx - 1.0) / (mx - 1.0)
    assert np.allclose(size, exp, rtol=1e-6)
