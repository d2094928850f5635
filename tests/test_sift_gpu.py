"""Parity of the CUDA sift128 extraction path (through the C ABI) against the CPU oracle
(oracle/afv_oracle_sift.c; parity vs SiftGPU itself is UNPINNED, see that file's header).
By construction the two share one arithmetic contract, so everything is compared bit for bit; the
north-star tolerance (descriptors within 1e-5 L2) is asserted as well."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _taps(ex, frames):
    bad = []
    for f, img in enumerate(frames):
        no = po.lib().orc_sift_num_octaves(img.shape[1], img.shape[0])
        for o in range(no):
            for what, n in ((10, 6), (11, 5)):
                for i in range(n):
                    ref = po.sift_scale_space(img, what - 10, o, i)
                    got = ex.debug_read(what, f, o * 8 + i, nbytes_cap=ref.size * 4 + 16).view(np.float32).reshape(ref.shape)
                    if not (got.view(np.uint32) == ref.view(np.uint32)).all():
                        bad.append("frame %d octave %d %s %d: %d px differ (max %.3g)" % (
                            f, o, "G" if what == 10 else "DoG", i, int((got != ref).sum()), float(np.abs(got - ref).max())))
    return bad


def _check(pkg, frames, nfeatures, w, h, taps=True):
    ex = pkg.FeatureExtractor("sift128", nfeatures=nfeatures, max_batch=len(frames), max_w=w, max_h=h)
    kps, desc, size, n = ex.extract_batch(frames)
    desc = desc.view(np.float32).reshape(len(frames), ex.cap, 128)
    problems = []
    for f, img in enumerate(frames):
        rk, rd, rs, nd = po.sift128_extract(img, nfeatures)
        xyso, _ = po.sift_detect(img, nfeatures, with_desc=False)
        lst = ex.debug_read(12, f, 0, nbytes_cap=16 * (len(xyso) + 4096)).view(np.float32).reshape(-1, 4)
        if lst.shape != xyso.shape or not (lst.view(np.uint32) == xyso.view(np.uint32)).all():
            same = lst.shape == xyso.shape
            problems.append("frame %d: detected list differs (ref %d, gpu %d%s)" % (
                f, len(xyso), len(lst), ", %d rows differ" % int((lst != xyso).any(axis=1).sum()) if same else ""))
        m = int(n[f])
        if m != len(rk):
            problems.append("frame %d: %d keypoints, oracle %d" % (f, m, len(rk)))
            continue
        for fld in rk.dtype.names:
            if not (kps[f, :m][fld] == rk[fld]).all():
                problems.append("frame %d: keypoint field %s differs in %d rows" % (f, fld, int((kps[f, :m][fld] != rk[fld]).sum())))
        l2 = np.linalg.norm(desc[f, :m] - rd, axis=1)
        if l2.max() > 1e-5:
            problems.append("frame %d: descriptor L2 error %.3g > 1e-5" % (f, float(l2.max())))
        if not (desc[f, :m].view(np.uint32) == rd.view(np.uint32)).all():
            problems.append("frame %d: %d descriptor rows not bit-identical" % (f, int((desc[f, :m] != rd).any(axis=1).sum())))
        if not (size[f, :m] == rs).all():
            problems.append("frame %d: computeSize differs" % f)
    if problems and taps:
        problems += _taps(ex, frames)
    ex.close()
    assert not problems, "\n".join(problems[:40])


def test_sift_640x480(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 0, 2)
    _check(pkg, frames, 1000, 640, 480)


def test_sift_scale_space_taps(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 3, 1)
    ex = pkg.FeatureExtractor("sift128", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
    ex.extract_batch(frames)
    bad = _taps(ex, frames)
    ex.close()
    assert not bad, "\n".join(bad[:20])


def test_sift_1280x720_c3(pkg, synth):
    """BASELINE configs[2]: sift128 1280x720, 2000 kp/frame."""
    frames, _ = synth.stream_frames(1280, 720, 1, 2)
    _check(pkg, frames, 2000, 1280, 720, taps=False)


def test_sift_odd_size_and_smaller_than_max(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 5, 1)
    img = np.ascontiguousarray(frames[:, 7:7 + 333, 11:11 + 517])
    _check(pkg, img, 500, 640, 480)


def test_sift_blank_and_device_api(pkg, synth):
    import torch
    frames, _ = synth.stream_frames(640, 480, 2, 2)
    frames[1][:] = 37                                          # featureless frame -> 0 keypoints
    ex = pkg.FeatureExtractor("sift128", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    d = torch.from_numpy(frames).cuda()
    out = ex.alloc_device_outputs(2)
    ex.extract_batch_device(d, out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    assert n[1] == 0
    rk, rd, rs, _ = po.sift128_extract(frames[0], 1000)
    assert n[0] == len(rk)
    k = pkg.kps_from_device(out[0][0], int(n[0]))
    assert all((k[f] == rk[f]).all() for f in rk.dtype.names)
    dd = out[1][0, :int(n[0])].cpu().numpy().view(np.float32).reshape(-1, 128)
    assert (dd == rd).all()
    ex.close()


def test_sift_l2_matcher_on_extracted(pkg, synth):
    """C3: L2 matcher (DescriptorDistance_sift128, src/Feature_sift128.cpp:132-134) on real sift128 output:
    brute force best / second between consecutive frames == oracle within 1e-5."""
    import torch
    frames, _ = synth.stream_frames(640, 480, 4, 2)
    ex = pkg.FeatureExtractor("sift128", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    kps, desc, size, n = ex.extract_batch(frames)
    ex.close()
    d0 = desc[0, :n[0]].view(np.float32).reshape(-1, 128); d1 = desc[1, :n[1]].view(np.float32).reshape(-1, 128)
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=False, desc_type=5, th_low=0.5)
    best, bd, sd = fm.match_bruteforce(torch.from_numpy(d0.copy()).cuda(), torch.from_numpy(d1.copy()).cuda())
    torch.cuda.synchronize()
    rb, rbd, rsd = po.match_bruteforce(5, d0, d1)
    assert (best.cpu().numpy() == rb).all()
    assert np.allclose(bd.cpu().numpy(), rbd, rtol=1e-5, atol=1e-7) and np.allclose(sd.cpu().numpy(), rsd, rtol=1e-5, atol=1e-7)
    assert (bd.cpu().numpy() < 0.5).mean() > 0.3               # consecutive synthetic frames do match


def test_sift_full_size_batch_properties(pkg, synth):
    """BASELINE configs[2] size (64 frames of 1280x720, 2000 kp): size-independent properties + one oracle spot check.
    Frames i and i+32 are the same image in different arena slots -> identical outputs (no cross-frame leakage,
    deterministic regardless of atomics order); unit-norm descriptors; keypoints inside the image, ascending octaves,
    per-octave counts within quota+3, octave rule of src/Feature_sift128.cpp:92, unique class_id."""
    import torch
    base = np.concatenate([synth.stream_frames(1280, 720, 40 + s, 8)[0] for s in range(4)], axis=0)
    frames = np.concatenate([base, base], axis=0)
    ex = pkg.FeatureExtractor("sift128", nfeatures=2000, max_batch=64, max_w=1280, max_h=720)
    out = ex.alloc_device_outputs(64)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    kps = out[0].cpu().numpy(); desc = out[1].cpu().numpy().view(np.float32).reshape(64, ex.cap, 128)
    assert (n[:32] == n[32:]).all() and n.min() > 500
    q = po.features_per_level(2000, 8, 2.0)
    for f in range(32):
        m = int(n[f])
        assert (kps[f, :m].view(np.uint8) == kps[f + 32, :m].view(np.uint8)).all()
        assert (desc[f, :m].view(np.uint32) == desc[f + 32, :m].view(np.uint32)).all()
        k = kps[f, :m].view(np.uint8).reshape(m, 28).view(pkg.KP_DTYPE).reshape(m)
        assert m <= ex.cap and (np.diff(k["octave"]) >= 0).all()
        assert (np.bincount(k["octave"], minlength=8) <= q + 3).all()
        assert (k["x"] >= 0).all() and (k["x"] < 1280).all() and (k["y"] >= 0).all() and (k["y"] < 720).all()
        assert (k["octave"] == np.floor(np.maximum(np.log2(k["size"].astype(np.float64) / 1.6454), 0)).astype(int)).all()
        assert len(set(k["class_id"].tolist())) == m and (k["response"] == 1.0).all()
        assert np.allclose(np.linalg.norm(desc[f, :m], axis=1), 1.0, atol=1e-5)
    rk, rd, rs, _ = po.sift128_extract(frames[5], 2000)
    m = int(n[5])
    k = pkg.kps_from_device(out[0][5], m)
    assert m == len(rk) and all((k[fld] == rk[fld]).all() for fld in rk.dtype.names) and (desc[5, :m] == rd).all()
    ex.close()
