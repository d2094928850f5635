"""Parity of the CUDA brisk48 extraction path (through the C ABI) against the CPU oracle (oracle/afv_oracle_brisk.c, score
contract ORC_BRISK_DENSE; parity vs ETH brisk v2 itself is UNPINNED, the oracle's detector / orientation / 512-bit core is
pinned to cv2.BRISK, see that file's header).  Bit-exact by construction: pyramid layers, score images, detect list, final
keypoints, 384-bit descriptors, sizes."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _taps(ex, frames, layers=8):
    bad = []
    for f, img in enumerate(frames):
        for lv in range(layers):
            for what, name in ((0, "image"), (1, "score")):
                ref = po.brisk_layer(img, what, lv, octaves=layers // 2)
                got = ex.debug_read(30 + what, f, lv, nbytes_cap=ref.size + 16).reshape(ref.shape)
                if not (got == ref).all():
                    bad.append("frame %d layer %d %s: %d px differ" % (f, lv, name, int((got != ref).sum())))
    return bad


def _check(pkg, frames, nfeatures, w, h, taps=True):
    ex = pkg.FeatureExtractor("brisk48", nfeatures=nfeatures, max_batch=len(frames), max_w=w, max_h=h)
    kps, desc, size, n = ex.extract_batch(frames)
    problems = []
    for f, img in enumerate(frames):
        rk, rd, rs, nd = po.brisk48_extract(img, nfeatures)
        det = po.brisk_detect(img, 34, 4, po.BRISK_DENSE)
        lst = ex.debug_read(32, f, 0, nbytes_cap=20 * (len(det) + 4096)).view(np.float32).reshape(-1, 5)
        if lst.shape != det.shape or not (lst.view(np.uint32) == det.view(np.uint32)).all():
            same = lst.shape == det.shape
            problems.append("frame %d: detect list differs (ref %d, gpu %d%s)" % (
                f, len(det), len(lst), ", %d rows differ" % int((lst != det).any(axis=1).sum()) if same else ""))
        m = int(n[f])
        if m != len(rk):
            problems.append("frame %d: %d keypoints, oracle %d" % (f, m, len(rk)))
            continue
        for fld in rk.dtype.names:
            if not (kps[f, :m][fld] == rk[fld]).all():
                problems.append("frame %d: keypoint field %s differs in %d rows" % (f, fld, int((kps[f, :m][fld] != rk[fld]).sum())))
        if not (desc[f, :m] == rd).all():
            problems.append("frame %d: %d descriptor rows differ" % (f, int((desc[f, :m] != rd).any(axis=1).sum())))
        if not (size[f, :m] == rs).all():
            problems.append("frame %d: computeSize differs" % f)
    if problems and taps:
        problems += _taps(ex, frames)
    ex.close()
    assert not problems, "\n".join(problems[:40])


def test_brisk_640x480(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 0, 2)
    _check(pkg, frames, 1000, 640, 480)


def test_brisk_layer_and_score_taps(pkg, synth):
    frames, _ = synth.stream_frames(640, 480, 3, 1)
    ex = pkg.FeatureExtractor("brisk48", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
    ex.extract_batch(frames)
    bad = _taps(ex, frames)
    ex.close()
    assert not bad, "\n".join(bad[:20])


def test_brisk_other_sizes(pkg, synth):
    frames, _ = synth.stream_frames(1280, 720, 1, 1)
    _check(pkg, frames, 2000, 1280, 720, taps=False)
    f2, _ = synth.stream_frames(640, 480, 5, 1)
    img = np.ascontiguousarray(f2[:, 6:6 + 335, 10:10 + 517])          # odd sizes: inexact half samples on several layers
    _check(pkg, img, 500, 640, 480)


def test_brisk_toy_frame(pkg, golden_dir):
    import os
    toy = np.ascontiguousarray(np.load(os.path.join(golden_dir, "toy0.npz"))["gray"])
    _check(pkg, toy[None], 1000, 640, 480)


def test_brisk_blank_device_api_and_matcher(pkg, synth):
    """C4: brisk48 descriptors through the 48-byte Hamming matcher (DescriptorDistance_brisk48, src/Feature_brisk48.cpp:62-64)."""
    import torch
    frames, _ = synth.stream_frames(640, 480, 2, 3)
    frames[2][:] = 90
    ex = pkg.FeatureExtractor("brisk48", nfeatures=1000, max_batch=3, max_w=640, max_h=480)
    d = torch.from_numpy(frames).cuda()
    out = ex.alloc_device_outputs(3)
    ex.extract_batch_device(d, out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    assert n[2] == 0 and n[0] > 300 and n[1] > 300
    ref = [po.brisk48_extract(frames[i], 1000) for i in range(2)]
    for i in range(2):
        k = pkg.kps_from_device(out[0][i], int(n[i]))
        assert all((k[fld] == ref[i][0][fld]).all() for fld in k.dtype.names)
        assert (out[1][i, :int(n[i])].cpu().numpy() == ref[i][1]).all()
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=True, desc_type=2, th_low=120.0)
    pa = torch.tensor([0], dtype=torch.int32, device="cuda"); pb = torch.tensor([1], dtype=torch.int32, device="cuda")
    max_size = float(np.float32(1.2) ** np.float32(7))
    m12, nm = fm.search_for_initialization(out[0], out[1], out[2], out[3], pa, pb, None, (0.0, 0.0, 640.0, 480.0), max_size)
    torch.cuda.synchronize()
    rk0, rd0, rs0, _ = ref[0]; rk1, rd1, rs1, _ = ref[1]
    prev = np.stack([rk0["x"], rk0["y"]], axis=1).astype(np.float32)
    rn, rm12, _ = po.search_for_initialization(2, rk0, rd0, rk1, rd1, rs1, (0.0, 0.0, 640.0, 480.0), max_size, prev,
                                               window=100, th_low=120.0, nnratio=0.9, check_ori=True)
    assert int(nm[0]) == rn and rn > 20
    assert (m12[0, :len(rk0)].cpu().numpy() == rm12).all()
    ex.close()


def test_brisk_full_size_batch_properties(pkg, synth):
    """BASELINE configs[3] size (256 frames of 640x480, 1000 kp): frames i and i+128 are the same image in different arena
    slots -> identical outputs; layers ascending, per-layer counts within quota+3, keypoints inside the pattern border, angles in
    [0, 360); two frames bit-exact against the oracle."""
    import torch
    base = np.concatenate([synth.stream_frames(640, 480, 60 + s, 16)[0] for s in range(8)], axis=0)
    frames = np.concatenate([base, base], axis=0)
    ex = pkg.FeatureExtractor("brisk48", nfeatures=1000, max_batch=256, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(256)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    kps = out[0].cpu().numpy(); desc = out[1].cpu().numpy()
    assert (n[:128] == n[128:]).all() and n.min() > 300
    q = po.features_per_level(1000, 8, 1.5)
    for f in range(128):
        m = int(n[f])
        assert (kps[f, :m].view(np.uint8) == kps[f + 128, :m].view(np.uint8)).all() and (desc[f, :m] == desc[f + 128, :m]).all()
        k = kps[f, :m].view(np.uint8).reshape(m, 28).view(pkg.KP_DTYPE).reshape(m)
        assert m <= ex.cap and (np.diff(k["octave"]) >= 0).all() and (k["class_id"] == -1).all()
        assert (np.bincount(k["octave"], minlength=8) <= q + 3).all()
        assert (k["x"] >= 13).all() and (k["x"] < 627).all() and (k["y"] >= 13).all() and (k["y"] < 467).all()
        assert (k["angle"] >= 0).all() and (k["angle"] < 360).all() and (k["response"][k["octave"] < 7] > 34).all()
    for f in (3, 200):
        rk, rd, rs, _ = po.brisk48_extract(frames[f], 1000)
        m = int(n[f])
        k = pkg.kps_from_device(out[0][f], m)
        assert m == len(rk) and all((k[fld] == rk[fld]).all() for fld in rk.dtype.names) and (desc[f, :m] == rd).all()
    ex.close()
