"""Vanilla ORB-SLAM2 extractor on the GPU (afv_extractor_create(AFV_FEAT_ORB32_VANILLA); reference src/ORBextractor.cc:460-676 built
with VANILLA_ORB_SLAM2) against the oracle, stage by stage and end to end, and against the golden vectors produced by the
reference's own compiled code running on cv2 (tests/golden/orbslam2_*.npz)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from test_oracle_orbslam2 import CASES, golden_frame

pytestmark = pytest.mark.gpu


def same(k, d, s, rk, rd, rs, tag):
    assert len(k) == len(rk), (tag, len(k), len(rk))
    for f in rk.dtype.names:
        assert (k[f] == rk[f]).all(), (tag, f, int((k[f] != rk[f]).sum()))
    assert (d == rd).all(), (tag, "descriptors", int((d != rd).any(axis=1).sum()))
    assert (s == rs).all(), (tag, "size")


def test_stage_taps(pkg, synth):
    gray = synth.stream_frames(640, 480, 0, 1)[0][0]
    ex = pkg.FeatureExtractor("orbslam2", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
    ex(gray)
    pyr = po.orbslam2_pyramid(gray)
    for l in range(8):
        h, w = pyr[l].shape
        if l:
            assert (ex.debug_read(40, 0, l).reshape(h, w) == pyr[l]).all(), ("pyramid", l)
        assert (ex.debug_read(41, 0, l).reshape(h, w) == po.gaussblur7_fixed(pyr[l])).all(), ("blur", l)
        xs, ys, sc = po.orbslam2_detect_level(pyr[l])
        det = ex.debug_read(43, 0, l).view(np.uint32)
        assert len(det) == len(xs), ("detect count", l, len(det), len(xs))
        assert ((det & 0xfff) == xs).all() and (((det >> 12) & 0xfff) == ys).all() and ((det >> 24) == sc).all(), ("detect list", l)


@pytest.mark.parametrize("name", CASES)
def test_golden_reference_code_on_cv2(name, pkg, synth, golden_dir):
    gray, z = golden_frame(name, synth, golden_dir)
    h, w = gray.shape
    ex = pkg.FeatureExtractor("orbslam2", nfeatures=int(z["nfeatures"]), max_batch=1, max_w=w, max_h=h)
    k, d, s = ex(gray)
    same(k, d, s, z["kps"], z["desc"], z["size"], name)


def test_batch_device_api_and_other_pyramids(pkg, synth):
    import torch
    frames = np.stack([synth.stream_frames(640, 480, s, 1)[0][0] for s in (3, 8, 9, 12)])
    frames[3] = frames[3] // 8 + 90                                   # low contrast: minThFAST cells
    for nf, nl, sf in ((1000, 8, 1.2), (1500, 5, 1.5)):
        ex = pkg.FeatureExtractor("orbslam2", nfeatures=nf, max_batch=4, max_w=640, max_h=480, n_octaves=nl, scale_factor=sf)
        out = ex.alloc_device_outputs(4)
        ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
        torch.cuda.synchronize()
        ex.status()
        for b in range(4):
            n = int(out[3][b])
            rk, rd, rs = po.orbslam2_extract(frames[b], nf, nl, sf)
            same(pkg.kps_from_device(out[0][b], n), out[1][b, :n].cpu().numpy(), out[2][b, :n].cpu().numpy(), rk, rd, rs, (nf, nl, b))


def test_odd_sizes_blank_and_too_small(pkg, synth):
    for (w, h) in ((752, 480), (641, 479), (322, 246)):
        gray = synth.stream_frames(w, h, 4, 1)[0][0]
        ex = pkg.FeatureExtractor("orbslam2", nfeatures=700, max_batch=1, max_w=w, max_h=h)
        k, d, s = ex(gray)
        rk, rd, rs = po.orbslam2_extract(gray, 700)
        same(k, d, s, rk, rd, rs, (w, h))
    ex = pkg.FeatureExtractor("orbslam2", nfeatures=700, max_batch=1, max_w=640, max_h=480)
    k, d, s = ex(np.full((480, 640), 77, np.uint8))
    assert len(k) == 0
    with pytest.raises(pkg.AfvError):                                  # top level 54 x 40: no room for one 30-px cell (reference: division by zero)
        ex(np.zeros((144, 192), np.uint8))


def test_matcher_on_vanilla_descriptors(pkg, synth):
    """SearchForInitialization on the vanilla extractor's output (orb32 descriptor type) == oracle."""
    import torch
    frames, _ = synth.stream_frames(640, 480, 6, 2)
    ex = pkg.FeatureExtractor("orbslam2", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(2)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=True, desc_type=0, th_low=75.0)
    pa = torch.tensor([0], dtype=torch.int32, device="cuda"); pb = torch.tensor([1], dtype=torch.int32, device="cuda")
    pm = torch.zeros((1, ex.cap, 2), dtype=torch.float32, device="cuda")
    pm[0] = out[0][0, :, :2]
    max_size = float(np.float32(1.2) ** np.float32(7))
    m12, nm = fm.search_for_initialization(out[0], out[1], out[2], out[3], pa, pb, pm, (0.0, 0.0, 640.0, 480.0), max_size)
    torch.cuda.synchronize()
    r = [po.orbslam2_extract(frames[i], 1000) for i in range(2)]
    prev = np.stack([r[0][0]["x"], r[0][0]["y"]], axis=1).astype(np.float32)
    rn, rm12, _ = po.search_for_initialization(0, r[0][0], r[0][1], r[1][0], r[1][1], r[1][2], (0.0, 0.0, 640.0, 480.0), max_size, prev,
                                               window=100, th_low=75.0, nnratio=0.9, check_ori=True)
    n0 = len(r[0][0])
    assert int(nm[0]) == rn and rn > 50
    assert (m12[0, :n0].cpu().numpy() == rm12).all()
