"""Parity of the CUDA FeatureMatcher kernels (through the C ABI) against the CPU oracle: bit-exact match
indices / Hamming distances; L2^2 (sift128 layout) within 1e-5 relative."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
BOUNDS = (0.0, 0.0, 640.0, 480.0)
MAXSZ = float(np.float32(1.2) ** 7)


@pytest.fixture(scope="module")
def extracted(pkg, synth):
    import torch
    frames = np.concatenate([synth.stream_frames(640, 480, s, 4)[0] for s in (0, 5)], axis=0)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=len(frames), max_w=640, max_h=480)
    out = ex.alloc_device_outputs(len(frames))
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    host = []
    for f in range(len(frames)):
        m = int(n[f])
        host.append((pkg.kps_from_device(out[0][f], m), out[1][f, :m].cpu().numpy(), out[2][f, :m].cpu().numpy()))
    yield out, host, ex.cap
    ex.close()


@pytest.mark.parametrize("nnratio,check_ori", [(0.9, True), (0.6, False)])
def test_search_for_initialization_batched(pkg, extracted, nnratio, check_ori):
    import torch
    out, host, cap = extracted
    pairs = [(0, 1), (1, 2), (2, 3), (4, 5), (5, 6), (6, 7), (3, 0), (0, 0)]
    pa = torch.tensor([p[0] for p in pairs], dtype=torch.int32, device="cuda")
    pb = torch.tensor([p[1] for p in pairs], dtype=torch.int32, device="cuda")
    pm = torch.zeros((len(pairs), cap, 2), dtype=torch.float32, device="cuda")
    for i, (a, b) in enumerate(pairs):
        pm[i] = out[0][a, :, :2]
    fm = pkg.FeatureMatcher(nnratio=nnratio, check_ori=check_ori, desc_type=0, th_low=75.0)
    m12, nm = fm.search_for_initialization(out[0], out[1], out[2], out[3], pa, pb, pm, BOUNDS, MAXSZ, window=100)
    torch.cuda.synchronize()
    m12 = m12.cpu().numpy(); nm = nm.cpu().numpy(); pmh = pm.cpu().numpy()
    for i, (a, b) in enumerate(pairs):
        k1, d1, s1 = host[a]; k2, d2, s2 = host[b]
        prev = np.stack([k1["x"], k1["y"]], axis=1)
        rn, rm, rpm = po.search_for_initialization(0, k1, d1, k2, d2, s2, BOUNDS, MAXSZ, prev, window=100, th_low=75.0,
                                                   nnratio=nnratio, check_ori=check_ori)
        assert nm[i] == rn, "pair %s: %d vs %d matches" % ((a, b), nm[i], rn)
        assert (m12[i, :len(k1)] == rm).all()
        assert (pmh[i, :len(k1)] == rpm).all()
    assert nm[:6].min() > 30


def test_second_round_uses_updated_prev_matched(pkg, extracted):
    """vbPrevMatched feeds the next call (reference src/Tracking.cc:473-474): chain two calls."""
    import torch
    out, host, cap = extracted
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=True, desc_type=0, th_low=75.0)
    pa = torch.tensor([0], dtype=torch.int32, device="cuda")
    pm = torch.zeros((1, cap, 2), dtype=torch.float32, device="cuda"); pm[0] = out[0][0, :, :2]
    k1, d1, s1 = host[0]
    prev = np.stack([k1["x"], k1["y"]], axis=1)
    for b in (1, 2):
        pb = torch.tensor([b], dtype=torch.int32, device="cuda")
        m12, nm = fm.search_for_initialization(out[0], out[1], out[2], out[3], pa, pb, pm, BOUNDS, MAXSZ, window=100)
        k2, d2, s2 = host[b]
        rn, rm, prev = po.search_for_initialization(0, k1, d1, k2, d2, s2, BOUNDS, MAXSZ, prev, window=100, th_low=75.0, nnratio=0.9, check_ori=True)
        assert int(nm[0]) == rn and (m12[0, :len(k1)].cpu().numpy() == rm).all()


def test_grid_and_window_search(pkg, extracted):
    import torch
    out, host, cap = extracted
    fm = pkg.FeatureMatcher(desc_type=0)
    cs, ci = fm.grid_build(out[0], out[3], BOUNDS)
    torch.cuda.synchronize()
    rng = np.random.default_rng(4)
    for (qa, tb) in ((0, 1), (5, 4)):
        kq, dq, sq = host[qa]; kt, dt, st = host[tb]
        nq = len(kq)
        qxy = np.stack([kq["x"], kq["y"]], axis=1) + rng.uniform(-6, 6, (nq, 2)).astype(np.float32)
        qr = rng.choice(np.array([5.0, 15.0, 30.0, 100.0, 1000.0], np.float32), nq)
        qmin = (sq / np.float32(1.2)).astype(np.float32); qmax = (sq * np.float32(1.2)).astype(np.float32)
        qmin[::7] = 0.0; qmax[::7] = 1e9
        rb, rbd, rsd, rbs, rss = po.match_window(0, dq, qxy, qr, qmin, qmax, kt, dt, st, BOUNDS)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        b, bd, sd, bs, ss = fm.match_window(out[1][qa, :nq], t(qxy), t(qr), t(qmin), t(qmax), out[0][tb], out[1][tb], out[2][tb],
                                            len(kt), cs[tb], ci[tb], BOUNDS)
        torch.cuda.synchronize()
        assert (b.cpu().numpy() == rb).all()
        assert (bd.cpu().numpy() == rbd).all() and (sd.cpu().numpy() == rsd).all()
        assert (bs.cpu().numpy() == rbs).all() and (ss.cpu().numpy() == rss).all()
        assert (rb >= 0).sum() > nq // 2


@pytest.mark.parametrize("desc_type,D", [(0, 32), (1, 61), (2, 48)])
def test_bruteforce_hamming(pkg, desc_type, D):
    import torch
    rng = np.random.default_rng(10 + desc_type)
    t = rng.integers(0, 256, (777, D), dtype=np.uint8)
    q = t[rng.integers(0, 777, 500)].copy()
    flip = rng.integers(0, 256, q.shape, dtype=np.uint8) & rng.integers(0, 256, q.shape, dtype=np.uint8) & rng.integers(0, 256, q.shape, dtype=np.uint8)
    q ^= flip
    q[:5] = t[:5]; t[100] = t[0]                                     # exact ties: first index must win
    fm = pkg.FeatureMatcher(desc_type=desc_type)
    b, bd, sd = fm.match_bruteforce(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda())
    rb, rbd, rsd = po.match_bruteforce(desc_type, q, t)
    assert (b.cpu().numpy() == rb).all() and (bd.cpu().numpy() == rbd).all() and (sd.cpu().numpy() == rsd).all()
    assert rb[0] == 0 and rbd[0] == 0 and rsd[0] == 0
    d = pkg.FeatureMatcher.descriptor_distance(desc_type, torch.from_numpy(q).cuda(), torch.from_numpy(t[:500].copy()).cuda())
    ref = np.array([po.descriptor_distance(desc_type, q[i], t[i]) for i in range(500)], np.float32)
    assert (d.cpu().numpy() == ref).all()


def test_bruteforce_l2_sift_layout(pkg):
    import torch
    rng = np.random.default_rng(20)
    t = rng.normal(size=(300, 128)).astype(np.float32); t /= np.linalg.norm(t, axis=1, keepdims=True)
    q = t[rng.integers(0, 300, 200)] + rng.normal(scale=0.02, size=(200, 128)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    fm = pkg.FeatureMatcher(desc_type=5)
    b, bd, sd = fm.match_bruteforce(torch.from_numpy(q).cuda().view(torch.uint8), torch.from_numpy(t).cuda().view(torch.uint8))
    rb, rbd, rsd = po.match_bruteforce(5, q, t)
    assert (b.cpu().numpy() == rb).all()
    assert np.allclose(bd.cpu().numpy(), rbd, rtol=1e-5, atol=1e-7) and np.allclose(sd.cpu().numpy(), rsd, rtol=1e-5, atol=1e-7)


def test_search_by_bow(pkg, extracted):
    import torch
    out, host, cap = extracted
    k1, d1, s1 = host[0]; k2, d2, s2 = host[1]

    def segs(k):
        node = (k["x"] // 80).astype(np.int32) * 10 + (k["y"] // 80).astype(np.int32)
        order = np.argsort(node, kind="stable")
        ids, starts = np.unique(node[order], return_index=True)
        return ids.astype(np.int32), np.append(starts, len(k)).astype(np.int32), order.astype(np.int32)
    sa, sb = segs(k1), segs(k2)
    for nnratio, ori in ((0.7, True), (0.75, False)):
        rn, rmf = po.search_by_bow(0, d1, k1, sa, d2, k2, sb, th_low=75.0, nnratio=nnratio, check_ori=ori)
        fm = pkg.FeatureMatcher(nnratio=nnratio, check_ori=ori, desc_type=0, th_low=75.0)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        mf, nm = fm.search_by_bow(out[1][0, :len(k1)], out[0][0], [t(v) for v in sa], out[1][1, :len(k2)], out[0][1], [t(v) for v in sb])
        torch.cuda.synchronize()
        assert int(nm[0]) == rn and (mf.cpu().numpy() == rmf).all() and rn > 100


def test_empty_inputs(pkg):
    import torch
    fm = pkg.FeatureMatcher(desc_type=0)
    q = torch.zeros((0, 32), dtype=torch.uint8, device="cuda"); t = torch.zeros((4, 32), dtype=torch.uint8, device="cuda")
    b, bd, sd = fm.match_bruteforce(q, t)
    assert b.numel() == 0
    b, bd, sd = fm.match_bruteforce(t, q)              # no train descriptors: best -1, distances FLT_MAX
    torch.cuda.synchronize()
    assert (b.cpu().numpy() == -1).all() and (bd.cpu().numpy() == np.finfo(np.float32).max).all()


@pytest.mark.parametrize("feature", ["orb32", "sift128", "akaze61", "brisk48"])
def test_cpp_host_mirror_matches_python_path(pkg, synth, tmp_path, feature):
    """FeatureExtractor_<feat>::operator() + FeatureMatcher::SearchForInitialization through the C++ mirror."""
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "anyfeature-vslam_b200", "host")])
    frames, _ = synth.stream_frames(640, 480, 7, 2)
    paths = []
    for i in range(2):
        p = tmp_path / ("f%d.pgm" % i)
        with open(p, "wb") as f:
            f.write(b"P5\n640 480\n255\n"); f.write(frames[i].tobytes())
        paths.append(str(p))
    st = pkg.FEATURE_SETTINGS[feature]
    yaml = tmp_path / ("%s_settings.yaml" % feature)
    yaml.write_text("%%YAML:1.0\nFeatureExtractor.numOctaves: %d\nFeatureExtractor.scaleFactor: %r\nFeatureExtractor.detectionTh: %r\nFeatureMatcher.matchingTh: %r\n"
                    % (st["n_octaves"], st["scale_factor"], st["detect_th"], st["matching_th"]))
    out = subprocess.check_output([os.path.join(root, "anyfeature-vslam_b200", "host", "host_api_test"), str(yaml)] + paths + [feature], text=True)
    got = dict(kv.split("=") for kv in out.split())
    ex = pkg.FeatureExtractor(feature, nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    k, d, s, n = ex.extract_batch(frames)
    o = ex.alloc_device_outputs(2)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), o)
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=True, desc_type=st["feature_id"], th_low=st["matching_th"])
    pa = torch.tensor([0], dtype=torch.int32, device="cuda"); pb = torch.tensor([1], dtype=torch.int32, device="cuda")
    m12, nm = fm.search_for_initialization(o[0], o[1], o[2], o[3], pa, pb, None, BOUNDS, MAXSZ, window=100)
    torch.cuda.synchronize()
    h = 1469598103934665603
    for i in range(2):
        for b in d[i, :int(n[i])].reshape(-1).tolist():
            h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
    for v in m12[0, :int(n[0])].cpu().numpy().tolist():
        h = ((h ^ ((v + 1) & 0xffffffff)) * 1099511628211) & 0xffffffffffffffff
    assert int(got["n0"]) == int(n[0]) and int(got["n1"]) == int(n[1]) and int(got["matches"]) == int(nm[0])
    assert int(got["levels"]) == 8 and int(got["q0"]) == {"orb32": 217, "sift128": 502, "akaze61": 212, "brisk48": 347}[feature]
    assert int(got["hash"]) == h
    assert int(got["pyr_default"]) == 800             # mvImagePyramid: 8 empty levels in the default build, like the reference
    _check_cpp_mirror_against_oracle(pkg, got, feature, st, frames, k, d, s, n, ex, o)
    ex.close()


def _fnv_ints(v):
    h = 1469598103934665603
    for x in np.asarray(v).tolist():
        h = ((h ^ ((x + 1) & 0xffffffff)) * 1099511628211) & 0xffffffffffffffff
    return h


def _fnv_bytes(a, h=1469598103934665603):
    for b in np.ascontiguousarray(a).view(np.uint8).reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
    return h


def _check_cpp_mirror_against_oracle(pkg, got, feature, st, frames, k, d, s, n, ex, o):
    """Every other method of the C++ mirror (tests/cpp/host_api_test.cpp) against the ORACLE on the same inputs."""
    import torch
    from oracle import pyoracle as po
    dt, th = st["feature_id"], float(st["matching_th"])
    n0, n1 = int(n[0]), int(n[1])
    k0, d0, s0 = k[0, :n0], d[0, :n0], s[0, :n0]
    k1, d1, s1 = k[1, :n1], d[1, :n1], s[1, :n1]
    if dt == 5:
        d0 = d0.view(np.float32); d1 = d1.view(np.float32)
    f32 = np.float32

    def queries(kq, sq, dx, dy):
        xy = np.stack([kq["x"] + f32(dx), kq["y"] + f32(dy)], axis=1).astype(np.float32)
        r = (f32(6.0) * sq).astype(np.float32); r[9::10] = -1.0
        return xy, r, (sq / f32(1.2)).astype(np.float32), (sq * f32(1.2)).astype(np.float32), kq["angle"].astype(np.float32)

    qxy, qr, qmin, qmax, qang = queries(k0, s0, 1.5, -2.0)
    occ = (np.arange(n1) % 7 == 0).astype(np.uint8)
    for name, ratio, ang in (("proj_local", True, False), ("proj_sim3", False, False), ("proj_motion", False, True), ("proj_reloc", False, True)):
        rn, rm = po.search_by_projection_ex(dt, d0, qxy, qr, qmin, qmax, k1, d1, s1, BOUNDS, qangle=qang if ang else None, occupied=occ, claim=True,
                                            th=th, nnratio=0.8, ratio_same_scale=ratio, tol=1.2)
        assert int(got[name]) == rn and int(got[name + "_h"]) == _fnv_ints(rm), name
        assert int(got[name + "_occ"]) == int(occ.sum()) + rn and rn > 50
    inf = (f32(1.0) / (s1 * s1)).astype(np.float32)
    for gate in (1, 0):
        rn, rm = po.search_by_projection_ex(dt, d0, qxy, qr, qmin, qmax, k1, d1, s1, BOUNDS, tinf1d=inf if gate else None, claim=False, th=th,
                                            nnratio=0.8, ratio_same_scale=False, tol=1.2)
        assert int(got["fuse%d" % gate]) == rn and int(got["fuse%d_h" % gate]) == _fnv_ints(rm), gate
    q2 = queries(k1, s1, -1.5, 2.0)
    rn, rm = po.search_by_sim3(dt, k0, d0, s0, qxy, qr, qmin, qmax, k1, d1, s1, q2[0], q2[1], q2[2], q2[3], BOUNDS, th)
    assert int(got["sim3"]) == rn and int(got["sim3_h"]) == _fnv_ints(rm) and rn > 50
    if dt != 5:
        node0 = (d0[:, 0] % 24).astype(np.int32); node1 = (d1[:, 0] % 24).astype(np.int32)
        v0 = (np.arange(n0) % 5 != 0).astype(np.uint8); v1 = (np.arange(n1) % 4 != 0).astype(np.uint8)
        rn, rm = po.bow_match(0, dt, k0, d0, node0, v0, k1, d1, node1, None, th_low=th, nnratio=0.7, check_ori=True)
        assert int(got["bow_kf_f"]) == rn and int(got["bow_kf_f_h"]) == _fnv_ints(rm) and rn > 20
        rn, rm = po.bow_match(1, dt, k0, d0, node0, v0, k1, d1, node1, v1, th_low=th, nnratio=0.7, check_ori=True)
        assert int(got["bow_kf_kf"]) == rn and int(got["bow_kf_kf_h"]) == _fnv_ints(rm) and rn > 20
        h0 = (np.arange(n0) % 3 == 0).astype(np.uint8); h1 = (np.arange(n1) % 3 == 1).astype(np.uint8)
        rn, rm = po.bow_match(2, dt, k0, d0, node0, h0, k1, d1, node1, h1, th_low=th, nnratio=0.6, check_ori=False,
                              F12=[0, 0, 0, 0, 0, -1, 0, 1, 0], epipole=(1.0e6, 240.0), sigma2_2=(s1 * s1).astype(np.float32))
        hp = 1469598103934665603
        for i1 in np.nonzero(rm >= 0)[0].tolist():
            for v in (i1, int(rm[i1])):
                hp = ((hp ^ v) * 1099511628211) & 0xffffffffffffffff
        assert int(got["triang"]) == rn and int(got["triang_h"]) == hp
    un = po.undistort_keypoints(k0, [520.0, 520.0, 320.0, 240.0], [-0.28, 0.07, 0.0002, 0.00002, 0.0])
    assert int(got["undist_h"]) == _fnv_bytes(un)
    cs, ci = pkg.FeatureMatcher.grid_build(o[0], o[3], BOUNDS)
    torch.cuda.synchronize()
    cs1 = cs[1].cpu().numpy(); ci1 = ci[1].cpu().numpy()[:n1]
    assert int(got["grid_items"]) == int(cs1[-1]) and int(got["grid_h"]) == _fnv_bytes(ci1[:int(cs1[-1])], _fnv_bytes(cs1))
    idx = np.arange(n0)
    depth = (f32(2.0) + (idx % 5).astype(np.float32)); depth[::9] *= -1
    Pw = np.stack([(k0["x"] - f32(320)) / f32(520) * depth, (k0["y"] - f32(240)) / f32(520) * depth, depth], axis=1).astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (n0, 1))
    maxd = np.where(idx % 11 == 0, 1.0, 100.0).astype(np.float32)
    pose = np.zeros(16, np.float32); pose[[0, 4, 8]] = 1
    iv, proj, track, fr, fmin, fmax = po.is_in_frustum(Pw, nrm, np.full(n0, 0.1, np.float32), maxd, s0, np.ones(n0, np.float32), np.full(n0, 3.0, np.float32),
                                                       pose, [520, 520, 320, 240, 0], [0, 640, 0, 480], cos_limit=0.5, radius_factor=3.0, size_tol=1.2)
    uv = proj[:, :2].copy()
    assert int(got["frustum_in"]) == int(iv.sum()) and 0 < iv.sum() < n0
    assert int(got["frustum_h"]) == _fnv_bytes(fr, _fnv_bytes(uv, _fnv_bytes(iv)))
    rn, rm = po.search_by_projection_ex(dt, d0, uv, fr, fmin, fmax, k1, d1, s1, BOUNDS, claim=True, th=th, nnratio=0.8, ratio_same_scale=True, tol=1.2)
    assert int(got["local_points"]) == rn and int(got["local_points_h"]) == _fnv_ints(rm)
    if feature == "orb32":
        vk, vd, vs = po.orbslam2_extract(frames[0], 1000)
        assert int(got["vanilla_n"]) == len(vk) and int(got["vanilla_h"]) == _fnv_bytes(vd, _fnv_bytes(vk))
        l3 = po.orbslam2_pyramid(frames[0])[3]
        assert got["pyr3"] == "%dx%d" % (l3.shape[1], l3.shape[0]) and int(got["pyr3_h"]) == _fnv_bytes(l3)
        ex1 = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
        kk, _, _ = ex1(frames[0])
        nd = sum(len(ex1.debug_read(3, 0, l)) // 8 for l in range(8)); nk = sum(len(ex1.debug_read(4, 0, l)) // 8 for l in range(8))
        assert int(got["hook_detect"]) == nd and int(got["hook_kept"]) == nk == len(kk) == int(got["hook_n"]) and nd > 5 * nk
        ex1.close()


@pytest.mark.parametrize("desc_type,D", [(1, 61), (2, 48), (5, 512)])
def test_search_for_initialization_other_descriptor_layouts(pkg, extracted, desc_type, D):
    """akaze61 / brisk48 (Hamming over 61 / 48 bytes, unaligned rows) and sift128 (L2^2 on 128 floats) through the
    same SearchForInitialization path: real orb keypoints, synthetic descriptors with planted correspondences."""
    import torch
    out, host, cap = extracted
    rng = np.random.default_rng(40 + desc_type)
    k1, _, s1 = host[0]; k2, _, s2 = host[1]
    n1, n2 = len(k1), len(k2)
    if desc_type == 5:
        d2 = rng.normal(size=(n2, 128)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        d1 = rng.normal(size=(n1, 128)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
        th = 0.5
    else:
        d2 = rng.integers(0, 256, (n2, D), dtype=np.uint8); d1 = rng.integers(0, 256, (n1, D), dtype=np.uint8)
        th = 128.0 if desc_type == 1 else 120.0
    # plant matches: query i looks like the nearest train keypoint (in position) with a little noise
    for i in range(0, n1, 2):
        j = int(np.argmin((k2["x"] - k1["x"][i]) ** 2 + (k2["y"] - k1["y"][i]) ** 2))
        if desc_type == 5:
            v = d2[j] + rng.normal(scale=0.03, size=128).astype(np.float32); d1[i] = (v / np.linalg.norm(v)).astype(np.float32)
        else:
            d1[i] = d2[j] ^ (rng.integers(0, 256, D, dtype=np.uint8) & rng.integers(0, 256, D, dtype=np.uint8) & rng.integers(0, 256, D, dtype=np.uint8))
    Db = 512 if desc_type == 5 else D
    kps = torch.zeros((2, cap, 7), dtype=torch.float32, device="cuda"); kps[0] = out[0][0]; kps[1] = out[0][1]
    desc = torch.zeros((2, cap, Db), dtype=torch.uint8, device="cuda")
    desc[0, :n1] = torch.from_numpy(d1.view(np.uint8).reshape(n1, Db)).cuda(); desc[1, :n2] = torch.from_numpy(d2.view(np.uint8).reshape(n2, Db)).cuda()
    size = torch.zeros((2, cap), dtype=torch.float32, device="cuda"); size[0] = out[2][0]; size[1] = out[2][1]
    n = torch.tensor([n1, n2], dtype=torch.int32, device="cuda")
    pa = torch.tensor([0, 1], dtype=torch.int32, device="cuda"); pb = torch.tensor([1, 0], dtype=torch.int32, device="cuda")
    pm = torch.zeros((2, cap, 2), dtype=torch.float32, device="cuda"); pm[0] = kps[0, :, :2]; pm[1] = kps[1, :, :2]
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=(desc_type != 5), desc_type=desc_type, th_low=th)
    m12, nm = fm.search_for_initialization(kps, desc, size, n, pa, pb, pm, BOUNDS, MAXSZ, window=100)
    torch.cuda.synchronize()
    for p, (a, b, ka, da, kb, db, sb) in enumerate([(0, 1, k1, d1, k2, d2, s2), (1, 0, k2, d2, k1, d1, s1)]):
        prev = np.stack([ka["x"], ka["y"]], axis=1)
        rn, rm, rpm = po.search_for_initialization(desc_type, ka, da, kb, db, sb, BOUNDS, MAXSZ, prev, window=100, th_low=th,
                                                   nnratio=0.9, check_ori=(desc_type != 5))
        assert int(nm[p]) == rn and (m12[p, :len(ka)].cpu().numpy() == rm).all(), "pair %d" % p
    assert int(nm[0]) > 40


@pytest.mark.parametrize("desc_type,D,fl", [(0, 32, False), (1, 61, False), (2, 48, False), (5, 128, True)])
def test_bow_transform_and_search_by_bow_chain(pkg, extracted, desc_type, D, fl):
    """Vocabulary::transform on the GPU == oracle (word, weight, FeatureVector node), then the FeatureVectors feed
    SearchByBoW (orb only: real descriptors)."""
    import torch
    from tests_bow import make_tree
    rng = np.random.default_rng(60 + desc_type)
    tree = make_tree(rng, k=10, L=3, D=D, float_desc=fl)
    out, host, cap = extracted
    if desc_type == 0:
        feats = host[0][1]
    elif fl:
        feats = rng.normal(size=(500, 128)).astype(np.float32)
    else:
        feats = rng.integers(0, 256, (500, D), dtype=np.uint8)
    voc = pkg.Vocabulary(desc_type, tree)
    dfe = torch.from_numpy(np.ascontiguousarray(feats).view(np.uint8).reshape(len(feats), -1)).cuda()
    for levelsup in (1, 2, 4):
        wid, w, nid = voc.transform(dfe, levelsup=levelsup)
        torch.cuda.synchronize()
        rw, rwt, rn = po.bow_transform(desc_type, feats, tree, levelsup=levelsup)
        assert (wid.cpu().numpy() == rw).all() and (w.cpu().numpy() == rwt).all() and (nid.cpu().numpy() == rn).all()
    if desc_type == 0:
        k1, d1, s1 = host[0]; k2, d2, s2 = host[1]
        segs = []
        for f, d in ((0, d1), (1, d2)):
            _, _, nid = voc.transform(out[1][f, :len(d)], levelsup=1)
            segs.append(pkg.Vocabulary.feature_vector_segments(nid.cpu().numpy()))
        rn, rmf = po.search_by_bow(0, d1, k1, segs[0], d2, k2, segs[1], th_low=75.0, nnratio=0.7, check_ori=True)
        fm = pkg.FeatureMatcher(nnratio=0.7, check_ori=True, desc_type=0, th_low=75.0)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        mf, nm = fm.search_by_bow(out[1][0, :len(k1)], out[0][0], [t(v) for v in segs[0]], out[1][1, :len(k2)], out[0][1], [t(v) for v in segs[1]])
        torch.cuda.synchronize()
        assert int(nm[0]) == rn and (mf.cpu().numpy() == rmf).all()


@pytest.mark.parametrize("ratio_same_scale,nnratio", [(True, 0.8), (False, 0.9)])
def test_search_by_projection_family(pkg, extracted, ratio_same_scale, nnratio):
    """TrackLocalMap-style SearchByProjection (ratio only between same-scale best/second, :139-146) and the best-only
    variant (:381-386), batched over problems, with pre-occupied train keypoints; sequential claiming reproduced."""
    import torch
    out, host, cap = extracted
    rng = np.random.default_rng(90)
    problems = [(0, 1), (2, 3), (5, 4)]                      # (query source frame, train frame)
    qd, qxy, qr, qmin, qmax, qstart, occs = [], [], [], [], [], [0], np.zeros((len(host), cap), np.uint8)
    refs = []
    for (qa, tb) in problems:
        kq, dq, sq = host[qa]; kt, dt, st = host[tb]
        nq = len(kq)
        xy = np.stack([kq["x"], kq["y"]], axis=1) + rng.uniform(-5, 5, (nq, 2)).astype(np.float32)
        r = (rng.choice(np.array([3.0, 8.0, 20.0], np.float32), nq) * sq).astype(np.float32)
        mn = (sq / np.float32(1.2)).astype(np.float32); mx = (sq * np.float32(1.2)).astype(np.float32)
        # duplicate some queries so later ones must skip keypoints claimed by earlier ones
        dup = rng.integers(0, nq, nq // 5)
        dq2 = np.concatenate([dq, dq[dup]]); xy = np.concatenate([xy, xy[dup]]); r = np.concatenate([r, r[dup]])
        mn = np.concatenate([mn, mn[dup]]); mx = np.concatenate([mx, mx[dup]])
        occ = (rng.random(len(kt)) < 0.15).astype(np.uint8)
        occs[tb, :len(kt)] = occ
        qd.append(dq2); qxy.append(xy); qr.append(r); qmin.append(mn); qmax.append(mx); qstart.append(qstart[-1] + len(dq2))
        refs.append(po.search_by_projection(0, dq2, xy, r, mn, mx, kt, dt, st, BOUNDS, occupied=occ, th=75.0, nnratio=nnratio,
                                            ratio_same_scale=ratio_same_scale, tol=1.2))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    fm = pkg.FeatureMatcher(nnratio=nnratio, desc_type=0, th_low=75.0)
    mq, nm = fm.search_by_projection(t(np.concatenate(qd)), t(np.concatenate(qxy)), t(np.concatenate(qr)), t(np.concatenate(qmin)),
                                     t(np.concatenate(qmax)), t(np.array(qstart, np.int32)), out[0], out[1], out[2], out[3],
                                     t(np.array([p[1] for p in problems], np.int32)), BOUNDS, occupied=t(occs),
                                     ratio_same_scale=ratio_same_scale, size_tolerance=1.2)
    torch.cuda.synchronize()
    mq = mq.cpu().numpy(); nm = nm.cpu().numpy()
    for i, (rn, rm) in enumerate(refs):
        assert nm[i] == rn and (mq[qstart[i]:qstart[i + 1]] == rm).all(), "problem %d" % i
        assert rn > 100
        got = rm[rm >= 0]
        assert len(set(got.tolist())) == len(got)             # every train keypoint claimed at most once


@pytest.mark.parametrize("desc_type,D", [(0, 32), (2, 48), (5, 512)])
def test_distinctive_descriptors(pkg, desc_type, D):
    import torch
    rng = np.random.default_rng(70 + desc_type)
    if desc_type == 5:
        desc = rng.normal(size=(600, 128)).astype(np.float32)
    else:
        base = rng.integers(0, 256, (40, D), dtype=np.uint8)
        desc = base[rng.integers(0, 40, 600)] ^ (rng.integers(0, 256, (600, D), dtype=np.uint8) & rng.integers(0, 256, (600, D), dtype=np.uint8) & rng.integers(0, 256, (600, D), dtype=np.uint8))
    lens = [0, 1, 2, 3, 7, 31, 32, 33, 64, 100] + rng.integers(2, 40, 30).tolist()
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    obs = rng.integers(0, 600, seg[-1]).astype(np.int32)
    ref = np.array([po.distinctive_descriptor(desc_type, desc, obs[seg[m]:seg[m + 1]]) for m in range(len(lens))], np.int32)
    d = torch.from_numpy(np.ascontiguousarray(desc).view(np.uint8).reshape(600, -1)).cuda()
    got = pkg.FeatureMatcher.distinctive_descriptors(desc_type, d, torch.from_numpy(obs).cuda(), torch.from_numpy(seg).cuda(), max(lens))
    torch.cuda.synchronize()
    assert (got.cpu().numpy() == ref).all()


def test_undistort_keypoints_and_grid(pkg, extracted, golden_dir):
    """Frame::UndistortKeyPoints on device rows == oracle (pinned to cv2) bit for bit, then AssignFeaturesToGrid on mvKeysUn."""
    import os
    import torch
    out, host, cap = extracted
    kps_d, n_d = out[0], out[3]
    g = np.load(os.path.join(golden_dir, "undistort_cv2.npz"))
    for i in range(3):
        K4, D5 = g["K%d" % i], g["D%d" % i]
        un = pkg.undistort_keypoints(kps_d, n_d, K4, D5)
        torch.cuda.synchronize()
        for b in range(kps_d.shape[0]):
            m = len(host[b][0])
            ref = po.undistort_keypoints(host[b][0], K4, D5)
            got = pkg.kps_from_device(un[b], m)
            assert all((got[f] == ref[f]).all() for f in ref.dtype.names)
    cs, ci = pkg.FeatureMatcher.grid_build(un, n_d, BOUNDS)
    torch.cuda.synchronize()
    assert int(cs[0, -1]) <= int(n_d[0])


def test_pack_results_kernel_matches_layout(pkg):
    """afv_pack_results (one launch) == sharding.pack_layout / unpack_results, the message format of the multi-GPU gather."""
    import torch
    sh = pkg.sharding
    rng = np.random.default_rng(3)
    for B, cap, D in ((3, 40, 32), (2, 36, 48), (4, 64, 61), (1, 16, 512)):
        n = torch.from_numpy(rng.integers(0, cap, B).astype(np.int32)).cuda()
        nm = torch.from_numpy(rng.integers(0, 200, B).astype(np.int32)).cuda()
        m12 = torch.from_numpy(rng.integers(-1, cap, (B, cap)).astype(np.int32)).cuda()
        kps = torch.from_numpy(rng.normal(size=(B, cap, 7)).astype(np.float32)).cuda()
        desc = torch.from_numpy(rng.integers(0, 256, (B, cap, D), dtype=np.uint8)).cuda()
        _, total = sh.pack_layout(B, cap, D)
        assert total % 4 == 0
        pack = torch.zeros(total, dtype=torch.uint8, device="cuda")
        launches = pkg.kernel_launches()
        sh.pack_results(pack, n, nm, m12, kps, desc)
        torch.cuda.synchronize()
        assert pkg.kernel_launches() == launches + 1                      # the library kernel ran, not five torch copies
        u = sh.unpack_results(pack.cpu().numpy(), B, cap, D)
        assert (u["n"] == n.cpu().numpy()).all() and (u["nmatches"] == nm.cpu().numpy()).all()
        assert (u["matches12"] == m12.cpu().numpy()).all()
        assert (u["kps"].view(np.float32).reshape(B, cap, 7) == kps.cpu().numpy()).all()
        assert (u["desc"] == desc.cpu().numpy()).all()
