"""CUDA == oracle for the remaining FeatureMatcher searches, one test per reference method:
SearchByProjection Sim3 (:287-397), relocalisation (:1406-1506) / motion model (:1291-1402), Fuse x2 (:794-1064), SearchBySim3
(:1066-1287), SearchByBoW(KF,F) and (KF,KF) (:186-283, :561-660), SearchForTriangulation (:662-790).  The oracle functions
used here are themselves checked against the reference's own compiled bodies in tests/test_oracle_vs_ref.py."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
BOUNDS = (0.0, 0.0, 640.0, 480.0)


@pytest.fixture(scope="module", params=["orb32", "brisk48"])
def extracted(request, pkg, synth):
    import torch
    feat = request.param
    st = pkg.FEATURE_SETTINGS[feat]
    frames = np.concatenate([synth.stream_frames(640, 480, s, 4)[0] for s in (20, 25)], axis=0)
    ex = pkg.FeatureExtractor(feat, nfeatures=1000, max_batch=len(frames), max_w=640, max_h=480)
    out = ex.alloc_device_outputs(len(frames))
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    host = []
    for f in range(len(frames)):
        m = int(n[f])
        host.append((pkg.kps_from_device(out[0][f], m), out[1][f, :m].cpu().numpy(), out[2][f, :m].cpu().numpy()))
    yield out, host, ex.cap, st["feature_id"], float(st["matching_th"]), np.float32(st["scale_factor"])
    ex.close()


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _queries(host, problems, rng, tol, radius, skip_frac=0.1, dup=True):
    qd, qxy, qr, qmin, qmax, qang, qstart = [], [], [], [], [], [], [0]
    for (qa, tb) in problems:
        kq, dq, sq = host[qa]
        nq = len(kq)
        xy = (np.stack([kq["x"], kq["y"]], axis=1) + rng.uniform(-4, 4, (nq, 2))).astype(np.float32)
        r = (np.float32(radius) * sq).astype(np.float32)
        r[rng.random(nq) < skip_frac] = -1.0
        mn = (sq / tol).astype(np.float32); mx = (sq * tol).astype(np.float32)
        ang = np.mod(kq["angle"] + np.where(rng.random(nq) < 0.3, 95.0, 0.0), 360.0).astype(np.float32)
        d = dq
        if dup:
            ii = rng.integers(0, nq, nq // 5)
            d = np.concatenate([dq, dq[ii]]); xy = np.concatenate([xy, xy[ii]]); r = np.concatenate([r, r[ii]])
            mn = np.concatenate([mn, mn[ii]]); mx = np.concatenate([mx, mx[ii]]); ang = np.concatenate([ang, ang[ii]])
        qd.append(d); qxy.append(xy); qr.append(r); qmin.append(mn); qmax.append(mx); qang.append(ang); qstart.append(qstart[-1] + len(d))
    return qd, qxy, qr, qmin, qmax, qang, qstart


@pytest.mark.parametrize("variant", ["sim3", "reloc", "reloc_noori", "fuse1", "fuse2"])
def test_projection_variants(pkg, extracted, variant):
    out, host, cap, dt, th, tol = extracted
    rng = np.random.default_rng(100)
    problems = [(0, 1), (2, 3), (5, 4), (6, 7)]
    qd, qxy, qr, qmin, qmax, qang, qstart = _queries(host, problems, rng, tol, 6.0)
    nfr = len(host)
    occs = np.zeros((nfr, cap), np.uint8); infs = np.zeros((nfr, cap), np.float32)
    for f in range(nfr):
        occs[f, :len(host[f][0])] = (rng.random(len(host[f][0])) < 0.15)
        infs[f, :len(host[f][0])] = np.float32(1.0) / (host[f][2] ** 2)
    use_occ = variant in ("sim3", "reloc", "reloc_noori"); claim = use_occ
    use_ang = variant == "reloc"; use_inf = variant == "fuse1"
    refs = []
    for i, (qa, tb) in enumerate(problems):
        kt, dtt, stt = host[tb]
        refs.append(po.search_by_projection_ex(dt, qd[i], qxy[i], qr[i], qmin[i], qmax[i], kt, dtt, stt, BOUNDS, qangle=qang[i] if use_ang else None,
                                               tinf1d=infs[tb, :len(kt)] if use_inf else None, occupied=occs[tb, :len(kt)] if use_occ else None,
                                               claim=claim, th=th, ratio_same_scale=False, tol=float(tol)))
    fm = pkg.FeatureMatcher(nnratio=0.8, check_ori=use_ang, desc_type=dt, th_low=th)
    mq, nm = fm.search_by_projection_ex(_t(np.concatenate(qd)), _t(np.concatenate(qxy)), _t(np.concatenate(qr)), _t(np.concatenate(qmin)),
                                        _t(np.concatenate(qmax)), _t(np.array(qstart, np.int32)), out[0], out[1], out[2], out[3],
                                        _t(np.array([p[1] for p in problems], np.int32)), BOUNDS, qangle=_t(np.concatenate(qang)) if use_ang else None,
                                        inf1d=_t(infs) if use_inf else None, occupied=_t(occs) if use_occ else None, claim=claim,
                                        ratio_same_scale=False, size_tolerance=float(tol))
    import torch
    torch.cuda.synchronize()
    mq = mq.cpu().numpy(); nm = nm.cpu().numpy()
    for i, (rn, rm) in enumerate(refs):
        assert nm[i] == rn and (mq[qstart[i]:qstart[i + 1]] == rm).all(), "%s problem %d: %d vs %d" % (variant, i, nm[i], rn)
        assert rn > 100
        if claim:
            got = rm[rm >= 0]; assert len(set(got.tolist())) == len(got)
    if variant in ("fuse1", "fuse2"):                                  # stateless: duplicated queries land on the same keypoint
        assert any(len(set(rm[rm >= 0].tolist())) < (rm >= 0).sum() for _, rm in refs)


def test_search_by_sim3(pkg, extracted):
    out, host, cap, dt, th, tol = extracted
    rng = np.random.default_rng(101)
    pairs = [(0, 1), (2, 3), (5, 4)]
    q1 = _queries(host, pairs, rng, tol, 5.0, skip_frac=0.3, dup=False)
    q2 = _queries(host, [(b, a) for a, b in pairs], rng, tol, 5.0, skip_frac=0.3, dup=False)
    refs = []
    for i, (a, b) in enumerate(pairs):
        ka, da, sa = host[a]; kb, db, sb = host[b]
        refs.append(po.search_by_sim3(dt, ka, da, sa, q1[1][i], q1[2][i], q1[3][i], q1[4][i], kb, db, sb, q2[1][i], q2[2][i], q2[3][i], q2[4][i], BOUNDS, th))
    fm = pkg.FeatureMatcher(nnratio=0.8, desc_type=dt, th_low=th)
    pack = lambda q: (_t(np.concatenate(q[0])), _t(np.concatenate(q[1])), _t(np.concatenate(q[2])), _t(np.concatenate(q[3])), _t(np.concatenate(q[4])),
                      _t(np.array(q[6], np.int32)))
    m12, nf = fm.search_by_sim3(pack(q1), pack(q2), out[0], out[1], out[2], out[3], _t(np.array([p[0] for p in pairs], np.int32)),
                                _t(np.array([p[1] for p in pairs], np.int32)), BOUNDS)
    import torch
    torch.cuda.synchronize()
    m12 = m12.cpu().numpy(); nf = nf.cpu().numpy()
    for i, (rn, rm) in enumerate(refs):
        assert nf[i] == rn and rn > 30 and (m12[q1[6][i]:q1[6][i + 1]] == rm).all(), "pair %d" % i


def _node_ids(host, cap, rng, drop=0.05):
    nfr = len(host)
    nodes = np.full((nfr, cap), -1, np.int32)
    for f in range(nfr):
        k = host[f][0]
        nd = ((k["x"] // 80).astype(np.int32) * 10 + (k["y"] // 80).astype(np.int32)) * 7 + 3
        nd[rng.random(len(k)) < drop] = -1
        nodes[f, :len(k)] = nd
    return nodes


@pytest.mark.parametrize("mode,check_ori", [(0, True), (0, False), (1, True), (1, False)])
def test_search_by_bow_batched(pkg, extracted, mode, check_ori):
    out, host, cap, dt, th, tol = extracted
    rng = np.random.default_rng(102 + mode)
    nodes = _node_ids(host, cap, rng)
    valid = (rng.random((len(host), cap)) < 0.8).astype(np.uint8)
    pairs = [(0, 1), (1, 2), (2, 3), (4, 5), (6, 7), (3, 3)]
    fm = pkg.FeatureMatcher(nnratio=0.75, check_ori=check_ori, desc_type=dt, th_low=th)
    m, nm = fm.bow_match(mode, out[0], out[1], out[3], _t(nodes), _t(valid), _t(np.array([p[0] for p in pairs], np.int32)),
                         _t(np.array([p[1] for p in pairs], np.int32)))
    import torch
    torch.cuda.synchronize()
    m = m.cpu().numpy(); nm = nm.cpu().numpy()
    for i, (a, b) in enumerate(pairs):
        ka, da, _ = host[a]; kb, db, _ = host[b]
        rn, rm = po.bow_match(mode, dt, ka, da, nodes[a, :len(ka)], valid[a, :len(ka)], kb, db, nodes[b, :len(kb)], valid[b, :len(kb)],
                              th_low=th, nnratio=0.75, check_ori=check_ori)
        assert nm[i] == rn and (m[i, :len(rm)] == rm).all() and (m[i, len(rm):] == -1).all(), "mode %d pair %d: %d vs %d" % (mode, i, nm[i], rn)
        assert rn > 30


def test_search_for_triangulation(pkg, extracted):
    out, host, cap, dt, th, tol = extracted
    rng = np.random.default_rng(104)
    nodes = _node_ids(host, cap, rng)
    has_mp = (rng.random((len(host), cap)) < 0.3).astype(np.uint8)
    sigma2 = np.ones((len(host), cap), np.float32)
    for f in range(len(host)):
        sigma2[f, :len(host[f][0])] = host[f][2] ** 2
    pairs = [(0, 1), (1, 2), (4, 5), (6, 7)]
    F12 = np.zeros((len(pairs), 9), np.float32); epi = np.zeros((len(pairs), 2), np.float32)
    for i in range(len(pairs)):
        ang = rng.uniform(0, np.pi)
        tx, ty = np.cos(ang), np.sin(ang)
        F12[i] = np.array([[0, 0, ty], [0, 0, -tx], [-ty, tx, 0]], np.float32).reshape(9) + rng.normal(0, 1e-4, 9).astype(np.float32)
        epi[i] = (rng.uniform(-300, 900), rng.uniform(-200, 700))
    fm = pkg.FeatureMatcher(nnratio=0.6, check_ori=False, desc_type=dt, th_low=th)
    m, nm = fm.bow_match(2, out[0], out[1], out[3], _t(nodes), _t(has_mp), _t(np.array([p[0] for p in pairs], np.int32)),
                         _t(np.array([p[1] for p in pairs], np.int32)), F12=_t(F12), epipole=_t(epi), sigma2=_t(sigma2))
    import torch
    torch.cuda.synchronize()
    m = m.cpu().numpy(); nm = nm.cpu().numpy()
    tot = 0
    for i, (a, b) in enumerate(pairs):
        ka, da, _ = host[a]; kb, db, _ = host[b]
        rn, rm = po.bow_match(2, dt, ka, da, nodes[a, :len(ka)], has_mp[a, :len(ka)], kb, db, nodes[b, :len(kb)], has_mp[b, :len(kb)], th_low=th,
                              F12=F12[i], epipole=tuple(epi[i].tolist()), sigma2_2=sigma2[b, :len(kb)])
        assert nm[i] == rn and (m[i, :len(rm)] == rm).all(), "pair %d: %d vs %d" % (i, nm[i], rn)
        tot += rn
    assert tot > 40


@pytest.mark.parametrize("radius,gated", [(15.0, False), (100.0, False), (30.0, True)])
def test_match_window_pairs(pkg, extracted, radius, gated):
    """Batched windowed matcher (train frame + grid staged in shared memory, warp per 32 queries with a candidate queue) == the oracle's GetFeaturesInArea
    window search, pair by pair: best index (first minimum in reference enumeration order), best and second distance."""
    import torch
    out, host, cap, dt, th, tol = extracted
    rng = np.random.default_rng(105)
    pairs = [(0, 1), (1, 2), (2, 3), (4, 5), (6, 7), (3, 3)]
    fm = pkg.FeatureMatcher(desc_type=dt, th_low=th)
    cs, ci = fm.grid_build(out[0], out[3], BOUNDS)
    P = len(pairs)
    qxy = np.zeros((P, cap, 2), np.float32); qr = np.full((P, cap), -1.0, np.float32)
    qmin = np.zeros((P, cap), np.float32); qmax = np.zeros((P, cap), np.float32)
    for i, (a, b) in enumerate(pairs):
        ka, _, sa = host[a]
        qxy[i, :len(ka), 0] = ka["x"] + rng.uniform(-6, 6, len(ka)); qxy[i, :len(ka), 1] = ka["y"] + rng.uniform(-6, 6, len(ka))
        qr[i, :len(ka)] = np.where(rng.random(len(ka)) < 0.1, -1.0, radius * rng.choice([0.5, 1.0, 2.0], len(ka)))
        qmin[i, :len(ka)] = sa / tol; qmax[i, :len(ka)] = sa * tol
    if gated:
        best, bd, sd = fm.match_window_pairs(out[0], out[1], out[2], out[3], cs, ci, _t(np.array([p[0] for p in pairs], np.int32)),
                                             _t(np.array([p[1] for p in pairs], np.int32)), BOUNDS, qxy=_t(qxy), qr=_t(qr), qmin=_t(qmin), qmax=_t(qmax))
    else:
        best, bd, sd = fm.match_window_pairs(out[0], out[1], out[2], out[3], cs, ci, _t(np.array([p[0] for p in pairs], np.int32)),
                                             _t(np.array([p[1] for p in pairs], np.int32)), BOUNDS, radius=radius)
    torch.cuda.synchronize()
    best = best.cpu().numpy(); bd = bd.cpu().numpy(); sd = sd.cpu().numpy()
    FMAX = np.finfo(np.float32).max
    for i, (a, b) in enumerate(pairs):
        ka, da, sa = host[a]; kb, db, sb = host[b]
        na = len(ka)
        if gated:
            valid = qr[i, :na] >= 0
            rb, rbd, rsd, _, _ = po.match_window(dt, da, qxy[i, :na], np.maximum(qr[i, :na], 0), qmin[i, :na], qmax[i, :na], kb, db, sb, BOUNDS)
            rb = np.where(valid, rb, -1); rbd = np.where(valid, rbd, FMAX); rsd = np.where(valid, rsd, FMAX)
        else:
            xy = np.stack([ka["x"], ka["y"]], axis=1).astype(np.float32)
            rb, rbd, rsd, _, _ = po.match_window(dt, da, xy, np.full(na, radius, np.float32), np.full(na, -FMAX, np.float32), np.full(na, FMAX, np.float32),
                                                 kb, db, sb, BOUNDS)
        assert (best[i, :na] == rb).all() and (bd[i, :na] == rbd).all() and (sd[i, :na] == rsd).all(), "pair %d" % i
        assert (rb >= 0).sum() > 100


@pytest.mark.parametrize("feat,cap,n", [("akaze61", 1200, 1100), ("orb32", 4000, 3500), ("brisk48", 3000, 2000), ("akaze61", 2300, 900)])
def test_match_window_pairs_synthetic(pkg, feat, cap, n):
    """The same kernel on made-up frames: 61-byte descriptor rows (akaze61, byte staging path), capacities where the train frame
    only fits next to 8 warps' queues (256-thread launch), and keypoints piled into a few cells so that one column step overflows
    the per-warp queue (drain in the middle of a step) and many candidates tie on distance (first minimum in enumeration order)."""
    import torch
    st = pkg.FEATURE_SETTINGS[feat]
    dt = st["feature_id"]; D = {"orb32": 32, "brisk48": 48, "akaze61": 61}[feat]
    rng = np.random.default_rng(cap + n)
    B = 3
    kps = np.zeros((B, cap), pkg.KP_DTYPE); desc = np.zeros((B, cap, D), np.uint8); size = np.zeros((B, cap), np.float32)
    nn = np.array([n, n - 37, n // 2], np.int32)
    protos = rng.integers(0, 256, (6, D), dtype=np.uint8)
    for f in range(B):
        m = int(nn[f])
        xy = rng.uniform([2, 2], [638, 478], (m, 2)).astype(np.float32)
        k = m // 3                                              # a third of the points in three 25-px blobs
        xy[:k] = (np.array([[100, 100], [320, 240], [600, 40]], np.float32)[rng.integers(0, 3, k)] + rng.uniform(-12, 12, (k, 2))).astype(np.float32)
        xy[k:k + k // 2] = np.round(xy[k:k + k // 2])          # integer coordinates: points exactly on cell and window borders
        kps[f, :m]["x"] = xy[:, 0]; kps[f, :m]["y"] = xy[:, 1]
        size[f, :m] = rng.choice([31.0, 37.2, 44.64], m).astype(np.float32); kps[f, :m]["size"] = size[f, :m]
        d = protos[rng.integers(0, 6, m)].copy()                # few prototypes with a few flipped bits: plenty of equal distances
        flips = rng.integers(0, D * 8, (m, 3))
        for j in range(3):
            d[np.arange(m), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
        desc[f, :m] = d
    dk = _t(kps.view(np.uint8).reshape(B, cap, 28).view(np.float32).reshape(B, cap, 7)); dd = _t(desc); ds = _t(size); dn = _t(nn)
    fm = pkg.FeatureMatcher(desc_type=dt, th_low=float(st["matching_th"]))
    cs, ci = fm.grid_build(dk, dn, BOUNDS)
    pairs = [(0, 1), (1, 0), (2, 1), (1, 2), (0, 0)]
    pa = _t(np.array([p[0] for p in pairs], np.int32)); pb = _t(np.array([p[1] for p in pairs], np.int32))
    FMAX = np.finfo(np.float32).max
    for radius in (20.0, 90.0):
        best, bd, sd = fm.match_window_pairs(dk, dd, ds, dn, cs, ci, pa, pb, BOUNDS, radius=radius)
        torch.cuda.synchronize()
        best = best.cpu().numpy(); bd = bd.cpu().numpy(); sd = sd.cpu().numpy()
        ties = 0
        for i, (a, b) in enumerate(pairs):
            na, nb = int(nn[a]), int(nn[b])
            xy = np.stack([kps[a, :na]["x"], kps[a, :na]["y"]], axis=1).astype(np.float32)
            rb, rbd, rsd, _, _ = po.match_window(dt, desc[a, :na], xy, np.full(na, radius, np.float32), np.full(na, -FMAX, np.float32),
                                                 np.full(na, FMAX, np.float32), kps[b, :nb], desc[b, :nb], size[b, :nb], BOUNDS)
            assert (best[i, :na] == rb).all() and (bd[i, :na] == rbd).all() and (sd[i, :na] == rsd).all(), "r=%g pair %d" % (radius, i)
            ties += int(((rbd == rsd) & (rb >= 0)).sum())
        assert ties > 100                                          # best == second distance occurs: the enumeration-order rule is exercised


def test_match_window_pairs_full_size_properties(pkg, synth):
    """BASELINE size (bench.py --workload m1: 10 240 frame pairs of 1000-keypoint orb32 frames) through properties that do not need
    the oracle at that size: the result of a pair does not depend on which launch or which CTA computed it (one launch of
    10 240 pairs == launches of 257), two runs are identical, a frame matched against itself returns every keypoint as its own
    best at distance 0 -- or, for keypoints that share position and descriptor with an earlier one, the first of them in
    enumeration order --, best < second wherever both exist; and a random sample of the pairs equals the oracle."""
    import torch
    B, P = 64, 10240
    frames = np.concatenate([synth.stream_frames(640, 480, s, 16)[0] for s in (30, 31, 32, 33)], axis=0)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=B, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(B)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize(); ex.status()
    cap = ex.cap
    n = out[3].cpu().numpy()
    fm = pkg.FeatureMatcher(desc_type=0, th_low=50.0)
    cs, ci = fm.grid_build(out[0], out[3], BOUNDS)
    rng = np.random.default_rng(7)
    pa = rng.integers(0, B, P).astype(np.int32); pb = rng.integers(0, B, P).astype(np.int32)
    pa[:B] = np.arange(B); pb[:B] = np.arange(B)                       # the first B pairs are (f, f)
    d_pa, d_pb = _t(pa), _t(pb)
    def run(a, b):                                                     # rows past a frame's keypoint count are not written: start from zeros
        res = (torch.zeros((a.shape[0], cap), dtype=torch.int32, device="cuda"), torch.zeros((a.shape[0], cap), dtype=torch.float32, device="cuda"),
               torch.zeros((a.shape[0], cap), dtype=torch.float32, device="cuda"))
        return fm.match_window_pairs(out[0], out[1], out[2], out[3], cs, ci, a, b, BOUNDS, radius=15.0, out=res)
    best, bd, sd = run(d_pa, d_pb)
    again = run(d_pa, d_pb)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip((best, bd, sd), again))
    for lo in range(0, P, 257):                                        # partition invariance
        hi = min(lo + 257, P)
        part = run(d_pa[lo:hi].contiguous(), d_pb[lo:hi].contiguous())
        assert torch.equal(part[0], best[lo:hi]) and torch.equal(part[1], bd[lo:hi]) and torch.equal(part[2], sd[lo:hi])
    hb = best.cpu().numpy(); hbd = bd.cpu().numpy(); hsd = sd.cpu().numpy()
    FMAX = np.finfo(np.float32).max
    for f in range(B):                                                 # self pairs
        m = int(n[f])
        k = pkg.kps_from_device(out[0][f], m); d = out[1][f, :m].cpu().numpy()
        # Frame::PosInGrid rounds to the nearest cell and drops keypoints that round to column 64 / row 48 (x >= 635, y >= 475):
        # the reference cannot find those in ANY window, itself included -- they are not candidates here either
        ingrid = (k["x"] < 634.0) & (k["y"] < 474.0)
        assert (hbd[f, :m][ingrid] == 0).all() and ingrid.mean() > 0.98
        other = np.where((hb[f, :m] != np.arange(m)) & ingrid)[0]
        for i in other:                                                # an identical twin earlier in enumeration order
            j = hb[f, i]
            assert (d[i] == d[j]).all() and abs(k["x"][i] - k["x"][j]) < 15 and abs(k["y"][i] - k["y"][j]) < 15
        assert len(other) < 0.05 * m
    for p in range(P):
        m = int(n[pa[p]])
        has2 = hsd[p, :m] < FMAX
        assert (hbd[p, :m][has2] <= hsd[p, :m][has2]).all() and ((hb[p, :m] >= 0) == (hbd[p, :m] < FMAX)).all()
    for p in rng.choice(P, 24, replace=False):                         # oracle on a sample
        a, b = int(pa[p]), int(pb[p]); na, nb = int(n[a]), int(n[b])
        ka = pkg.kps_from_device(out[0][a], na); kb = pkg.kps_from_device(out[0][b], nb)
        xy = np.stack([ka["x"], ka["y"]], axis=1).astype(np.float32)
        rb, rbd, rsd, _, _ = po.match_window(0, out[1][a, :na].cpu().numpy(), xy, np.full(na, 15.0, np.float32), np.full(na, -FMAX, np.float32),
                                             np.full(na, FMAX, np.float32), kb, out[1][b, :nb].cpu().numpy(), out[2][b, :nb].cpu().numpy(), BOUNDS)
        assert (hb[p, :na] == rb).all() and (hbd[p, :na] == rbd).all() and (hsd[p, :na] == rsd).all()
    ex.close()


def test_match_bruteforce_pairs(pkg, extracted):
    import torch
    out, host, cap, dt, th, tol = extracted
    pairs = [(0, 1), (2, 3), (5, 4), (7, 7)]
    fm = pkg.FeatureMatcher(desc_type=dt, th_low=th)
    best, bd, sd = fm.match_bruteforce_pairs(out[1], out[3], _t(np.array([p[0] for p in pairs], np.int32)), _t(np.array([p[1] for p in pairs], np.int32)))
    torch.cuda.synchronize()
    best = best.cpu().numpy(); bd = bd.cpu().numpy(); sd = sd.cpu().numpy()
    for i, (a, b) in enumerate(pairs):
        da = host[a][1]; db = host[b][1]
        rb, rbd, rsd = po.match_bruteforce(dt, da, db)
        assert (best[i, :len(da)] == rb).all() and (bd[i, :len(da)] == rbd).all() and (sd[i, :len(da)] == rsd).all()
