"""Vanilla ORB-SLAM2 extractor (reference src/ORBextractor.cc:460-676 built with VANILLA_ORB_SLAM2): the oracle
restatement against (a) golden vectors made by the reference's OWN compiled code running on the real cv2 4.13.0 functions
(tools/make_golden_orbslam2.py), (b) the same pipeline run live when oracle/_ref and cv2 are present, (c) cv2 stage by stage."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ("synth_640x480_s0", "synth_640x480_s5", "synth_1280x720_s1", "synth_752x480_s2", "toy0", "lowcontrast")


def golden_frame(name, synth, golden_dir):
    z = np.load(os.path.join(golden_dir, "orbslam2_%s.npz" % name))
    if name == "toy0":
        gray = np.load(os.path.join(golden_dir, "toy0.npz"))["gray"]
    else:
        gray = synth.stream_frames(int(z["w"]), int(z["h"]), int(z["stream"]), 1)[0][0]
        if "lowcontrast" in z.files:
            gray = (gray // 6 + 100).astype(np.uint8)
    return gray, z


def assert_same(k, d, s, rk, rd, rs, tag):
    assert len(k) == len(rk), (tag, len(k), len(rk))
    for f in rk.dtype.names:
        assert (k[f] == rk[f]).all(), (tag, f)
    assert (d == rd).all(), (tag, "descriptors")
    assert (s == rs).all(), (tag, "size")


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_reference_code_on_cv2_golden(name, synth, golden_dir):
    gray, z = golden_frame(name, synth, golden_dir)
    k, d, s = po.orbslam2_extract(gray, int(z["nfeatures"]))
    assert_same(k, d, s, z["kps"], z["desc"], z["size"], name)
    assert (np.diff(k["octave"]) >= 0).all()
    if name == "lowcontrast":                   # the frame only yields its quota through the minThFAST fallback
        assert (k["response"] < 20).mean() > 0.5


def _real_parts():
    cv2 = pytest.importorskip("cv2")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libafv_ref.so")):
        from oracle import build_ref
        if build_ref.build() is None:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    spec = importlib.util.spec_from_file_location("mk_os2", os.path.join(ROOT, "tools", "make_golden_orbslam2.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return cv2, mod


def test_oracle_equals_reference_code_live(synth):
    """A frame that is in no fixture, two quotas, and a 5-level / 1.5 pyramid (other cell grids, other resize ratios)."""
    _, mk = _real_parts()
    gray = synth.stream_frames(640, 480, 11, 1)[0][0]
    for nf, nl, sf in ((500, 8, 1.2), (1500, 5, 1.5)):
        rk, rd, rs = mk.reference_real_parts(gray, nf, nl, sf)
        k, d, s = po.orbslam2_extract(gray, nf, nl, sf)
        assert_same(k, d, s, rk, rd, rs, (nf, nl, sf))


def test_stages_against_cv2(synth):
    cv2, _ = _real_parts()
    gray = synth.stream_frames(752, 480, 4, 1)[0][0]
    lw, lh, _, _ = po.orbslam2_geometry(752, 480)
    prev = gray
    for l in range(1, 8):
        ref = cv2.resize(prev, (int(lw[l]), int(lh[l])), interpolation=cv2.INTER_LINEAR)
        assert (po.resize_linear(prev, int(lw[l]), int(lh[l])) == ref).all(), l
        assert (po.gaussblur7_fixed(ref) == cv2.GaussianBlur(ref, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)).all(), l
        prev = ref
    # odd ratios incl. upscaling-free extremes of the coefficient clamp
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    for dw, dh in ((109, 81), (66, 49), (130, 96), (131, 97)):
        assert (po.resize_linear(img, dw, dh) == cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)).all(), (dw, dh)
    # per-cell FAST with the empty-cell fallback against cv2.FastFeatureDetector on the same sub-images
    lvl = cv2.resize(gray, (int(lw[1]), int(lh[1])), interpolation=cv2.INTER_LINEAR)
    xs, ys, sc = po.orbslam2_detect_level(lvl)
    rows, cols = lvl.shape
    minB, maxBX, maxBY = 16, cols - 16, rows - 16
    ncol, nrow = int((maxBX - minB) / 30), int((maxBY - minB) / 30)
    wc, hc = int(np.ceil((maxBX - minB) / ncol)), int(np.ceil((maxBY - minB) / nrow))
    out = []
    d20 = cv2.FastFeatureDetector_create(20, True); d7 = cv2.FastFeatureDetector_create(7, True)
    for i in range(nrow):
        y0 = minB + i * hc; y1 = min(y0 + hc + 6, maxBY)
        if y0 >= maxBY - 3:
            continue
        for j in range(ncol):
            x0 = minB + j * wc; x1 = min(x0 + wc + 6, maxBX)
            if x0 >= maxBX - 6:
                continue
            sub = np.ascontiguousarray(lvl[y0:y1, x0:x1])
            kp = d20.detect(sub) or d7.detect(sub)
            out += [(int(k.pt[0]) + j * wc, int(k.pt[1]) + i * hc, int(k.response)) for k in kp]
    assert len(out) == len(xs) and out == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
