"""Parity of the CUDA orb32 extraction path (through the C ABI) against the CPU oracle and the committed
golden vectors.  Bit-exact: keypoint fields, 256-bit descriptors, sizes; stage taps localise a failure."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _same_kps(a, b):
    return len(a) == len(b) and all((a[f] == b[f]).all() for f in a.dtype.names)


def _stage_report(ex, frames, nfeatures):
    """Compare every intermediate of the last batch with the oracle; returns a list of mismatch strings."""
    bad = []
    q_orb = po.features_per_level(nfeatures * 10)
    q_ext = po.features_per_level(nfeatures)
    for f, img in enumerate(frames):
        h, w = img.shape
        levels, ls = po.pyramid(img)
        for l in range(8):
            L = levels[l]
            if l > 0:
                g = ex.debug_read(0, f, l).reshape(L.shape)
                if not (g == L).all():
                    bad.append("frame %d level %d: pyramid differs in %d px" % (f, l, int((g != L).sum())))
            xs, ys, sc = po.fast(L, 20)
            ref = set(zip(xs.tolist(), ys.tolist(), sc.tolist()))
            c = ex.debug_read(2, f, l).view(np.uint32)
            got = set(zip((c & 0xfff).tolist(), ((c >> 12) & 0xfff).tolist(), (c >> 24).tolist()))
            if ref != got:
                bad.append("frame %d level %d: FAST set differs (ref %d, gpu %d, common %d)" % (f, l, len(ref), len(got), len(ref & got)))
            dx, dy, hr, fs = po.detect_level(L, 20, q_orb[l])
            ref = set(zip(dx.tolist(), dy.tolist(), hr.view(np.uint32).tolist()))
            d = ex.debug_read(3, f, l).view(np.uint32).reshape(-1, 2)
            got = set(zip((d[:, 0] & 0xfff).tolist(), ((d[:, 0] >> 12) & 0xfff).tolist(), d[:, 1].tolist()))
            if ref != got:
                bad.append("frame %d level %d: detect set differs (ref %d, gpu %d, common %d)" % (f, l, len(ref), len(got), len(ref & got)))
            keep = po.octree(dx.astype(np.float32) * ls[l], dy.astype(np.float32) * ls[l], hr, w, h, q_ext[l])
            ref_list = [(int(dx[i]), int(dy[i])) for i in keep]
            k = ex.debug_read(4, f, l).view(np.uint32).reshape(-1, 2)
            got_list = list(zip((k[:, 0] & 0xfff).tolist(), ((k[:, 0] >> 12) & 0xfff).tolist()))
            if ref_list != got_list:
                bad.append("frame %d level %d: octree list differs (ref %d, gpu %d, same-set %s)" % (
                    f, l, len(ref_list), len(got_list), set(ref_list) == set(got_list)))
            b = ex.debug_read(1, f, l).reshape(L.shape)
            rb = po.blur7(L)
            if not (b == rb).all():
                bad.append("frame %d level %d: blur differs in %d px" % (f, l, int((b != rb).sum())))
    return bad


def _check_batch(pkg, frames, nfeatures, w, h):
    ex = pkg.FeatureExtractor("orb32", nfeatures=nfeatures, max_batch=len(frames), max_w=w, max_h=h)
    kps, desc, size, n = ex.extract_batch(frames)
    problems = []
    for f, img in enumerate(frames):
        rk, rd, rs, _ = po.orb32_extract(img, nfeatures)
        m = int(n[f])
        if not (_same_kps(kps[f, :m], rk) and (desc[f, :m] == rd).all() and (size[f, :m] == rs).all()):
            problems.append("frame %d: final output differs (gpu n=%d, oracle n=%d)" % (f, m, len(rk)))
    if problems:
        problems += _stage_report(ex, frames, nfeatures)
    ex.close()
    assert not problems, "\n".join(problems[:40])


def test_stage_taps_match_oracle(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 0, 2)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    ex.extract_batch(frames)
    bad = _stage_report(ex, frames, 1000)
    ex.close()
    assert not bad, "\n".join(bad[:40])


def test_synthetic_640x480_bit_exact(pkg, synth):
    frames = np.concatenate([synth.stream_frames(640, 480, s, 3)[0] for s in (0, 3)], axis=0)
    _check_batch(pkg, frames, 1000, 640, 480)


def test_synthetic_1280x720_n2000_bit_exact(pkg, synth):
    frames, _ = synth.stream_frames(1280, 720, 1, 2)
    _check_batch(pkg, frames, 2000, 1280, 720)


@pytest.mark.parametrize("name", ["toy0", "toy2"])
def test_golden_toy_frames(pkg, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    for nfeat in (1000, 2000):
        ex = pkg.FeatureExtractor("orb32", nfeatures=nfeat, max_batch=1, max_w=640, max_h=480)
        k, d, s = ex(g["gray"])
        ex.close()
        assert _same_kps(k, g["kps%d" % nfeat]), "keypoints differ from the cv2-generated golden vector"
        assert (d == g["desc%d" % nfeat]).all()


def test_golden_synthetic(pkg, synth, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_640x480_s0_t1.npz"))
    fr, _ = synth.stream_frames(640, 480, 0, 2)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
    k, d, s = ex(fr[1])
    ex.close()
    assert _same_kps(k, g["kps"]) and (d == g["desc"]).all()


def test_device_api_matches_host_api(pkg, synth):
    import torch
    frames, _ = synth.stream_frames(640, 480, 2, 4)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=4, max_w=640, max_h=480)
    hk, hd, hs, hn = ex.extract_batch(frames)
    d_gray = torch.from_numpy(frames).cuda()
    out = ex.alloc_device_outputs(4)
    ex.extract_batch_device(d_gray, out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    assert (n == hn).all()
    for f in range(4):
        m = int(n[f])
        assert _same_kps(pkg.kps_from_device(out[0][f], m), hk[f, :m])
        assert (out[1][f, :m].cpu().numpy() == hd[f, :m]).all()
        assert (out[2][f, :m].cpu().numpy() == hs[f, :m]).all()
    # a strided (non-aliased) device view takes the staging-copy path and must give the same result
    big = torch.zeros((4, 480, 648), dtype=torch.uint8, device="cuda")
    big[:, :, 3:643] = d_gray
    out2 = ex.alloc_device_outputs(4)
    ex.extract_batch_device(big[:, :, 3:643], out2)
    torch.cuda.synchronize()
    assert (out2[3].cpu().numpy() == hn).all() and (out2[1].cpu().numpy() == out[1].cpu().numpy()).all()
    ex.close()


def test_edge_cases(pkg):
    # blank frame -> zero keypoints; frame with features only in one corner; odd size
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
    k, d, s = ex(np.full((480, 640), 128, np.uint8))
    assert len(k) == 0
    rng = np.random.default_rng(5)
    img = np.full((480, 640), 100, np.uint8)
    img[:64, :64] = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    k, d, s = ex(img)
    rk, rd, rs, _ = po.orb32_extract(img, 1000)
    assert _same_kps(k, rk) and (d == rd).all()
    img = rng.integers(0, 256, (301, 413), dtype=np.uint8)         # odd geometry, dense corners everywhere
    k, d, s = ex(img)
    rk, rd, rs, _ = po.orb32_extract(img, 1000)
    assert _same_kps(k, rk) and (d == rd).all()
    ex.close()


def test_gray_conversion_kernel(pkg):
    import torch
    rng = np.random.default_rng(1)
    for ch in (3, 4):
        img = rng.integers(0, 256, (3, 120, 161, ch), dtype=np.uint8)
        for rgb in (True, False):
            g = pkg.gray_from_color(torch.from_numpy(img).cuda(), rgb=rgb)
            torch.cuda.synchronize()
            ref = np.stack([po.gray_from_color(img[i], rgb=rgb) for i in range(3)])
            assert (g.cpu().numpy() == ref).all()


def test_full_size_batch_properties(pkg, synth):
    """BASELINE-size batch (256 frames of 640x480, 1000 kp): size-independent properties + oracle spot checks.
    Frames i and i+128 are the same image at different arena slots -> identical outputs (no cross-frame leakage,
    deterministic regardless of atomics order); every frame has nfeatures..cap keypoints in ascending octaves with
    per-level counts within [quota, quota+3]; three sampled frames are bit-exact against the oracle."""
    import torch
    base = np.concatenate([synth.stream_frames(640, 480, 20 + s, 16)[0] for s in range(8)], axis=0)
    frames = np.concatenate([base, base], axis=0)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=256, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(256)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    kps = out[0].cpu().numpy(); desc = out[1].cpu().numpy()
    assert (n[:128] == n[128:]).all()
    q = po.features_per_level(1000)
    for f in range(128):
        m = int(n[f])
        assert (kps[f, :m].view(np.uint8) == kps[f + 128, :m].view(np.uint8)).all() and (desc[f, :m] == desc[f + 128, :m]).all()
        k = kps[f, :m].view(np.uint8).reshape(m, 28).view(pkg.KP_DTYPE).reshape(m)
        assert 1000 <= m <= ex.cap and (np.diff(k["octave"]) >= 0).all()
        cnt = np.bincount(k["octave"], minlength=8)
        assert ((cnt >= q) & (cnt <= q + 3)).all()
    for f in (0, 77, 200):
        rk, rd, rs, _ = po.orb32_extract(frames[f], 1000)
        m = int(n[f])
        assert _same_kps(pkg.kps_from_device(out[0][f], m), rk) and (desc[f, :m] == rd).all()
    ex.close()
