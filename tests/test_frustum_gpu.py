"""Frame::isInFrustum on the GPU (afv_is_in_frustum; reference src/Frame.cc:276-331) == oracle bit for bit, within 1e-5 of a float64
evaluation of the same formulas (the bound that covers the reference's Eigen build), and chained into SearchByProjection
(TrackLocalMap, src/FeatureMatcher.cc:73-154) exactly as Tracking::SearchLocalPoints does."""
import numpy as np
import pytest

from frustum_case import make_case
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_is_in_frustum_equals_oracle(pkg):
    import torch
    for seed in range(4):
        c = make_case(10 + seed, M=6000)
        ref = po.is_in_frustum(cos_limit=0.5, radius_factor=1.5, size_tol=1.5, **c)
        dev = {k: (_t(v) if k in ("Pw", "normal", "min_dist", "max_dist", "ref_size", "ref_sigma", "ref_dist") else v) for k, v in c.items()}
        got = pkg.is_in_frustum(cos_limit=0.5, radius_factor=1.5, size_tol=1.5, **dev)
        torch.cuda.synchronize()
        for g, r, name in zip(got, ref, ("in_view", "proj", "track", "qr", "qmin", "qmax")):
            assert (g.cpu().numpy() == r).all(), (seed, name)
        # float64 evaluation: decisions identical away from the boundaries, values within 1e-5 relative
        R = c["pose16"][:9].reshape(3, 3).astype(np.float64); t = c["pose16"][9:12].astype(np.float64); cw = c["pose16"][12:15].astype(np.float64)
        Pc = c["Pw"].astype(np.float64) @ R.T + t
        v = ref[0].astype(bool)
        u = c["cam5"][0] * Pc[v, 0] / Pc[v, 2] + c["cam5"][2]
        dist = np.linalg.norm(c["Pw"][v].astype(np.float64) - cw, axis=1)
        assert np.allclose(ref[1][v, 0], u, rtol=1e-5, atol=1e-3)
        assert np.allclose(ref[2][v, 0], c["ref_size"][v].astype(np.float64) * c["ref_dist"][v] / dist, rtol=1e-5)
    # empty input is a no-op
    e = torch.zeros((0, 3), device="cuda"); z = torch.zeros(0, device="cuda")
    assert pkg.is_in_frustum(e, e, z, z, z, z, z, c["pose16"], c["cam5"], c["bounds4"])[0].numel() == 0


def test_search_local_points_chain(pkg, synth):
    """Tracking::SearchLocalPoints: isInFrustum over the local map, then SearchByProjection(F, points, th): map points are
    back-projected keypoints of the previous frame (descriptor = theirs), the camera is the identity pose."""
    import torch
    frames, _ = synth.stream_frames(640, 480, 30, 2)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(2)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    n = out[3].cpu().numpy()
    k0 = pkg.kps_from_device(out[0][0], int(n[0])); d0 = out[1][0, :int(n[0])].cpu().numpy(); s0 = out[2][0, :int(n[0])].cpu().numpy()
    k1 = pkg.kps_from_device(out[0][1], int(n[1])); d1 = out[1][1, :int(n[1])].cpu().numpy(); s1 = out[2][1, :int(n[1])].cpu().numpy()
    rng = np.random.default_rng(7)
    fx, fy, cx, cy = 520.0, 520.0, 320.0, 240.0
    M = len(k0)
    depth = rng.uniform(2.0, 6.0, M)
    Pw = np.stack([(k0["x"] - cx) / fx * depth, (k0["y"] - cy) / fy * depth, depth], 1).astype(np.float32)
    Pw[::7, 2] *= -1                                                   # some behind the camera
    nrm = Pw / np.linalg.norm(Pw, axis=1, keepdims=True)
    dist = np.linalg.norm(Pw, axis=1).astype(np.float32)
    pose = np.zeros(16, np.float32); pose[[0, 4, 8]] = 1
    cam = np.array([fx, fy, cx, cy, 0.0], np.float32); bounds = np.array([0, 640, 0, 480], np.float32)
    args = dict(Pw=Pw, normal=nrm.astype(np.float32), min_dist=dist * 0.5, max_dist=dist * 2, ref_size=s0, ref_sigma=s0, ref_dist=dist,
                pose16=pose, cam5=cam, bounds4=bounds)
    tol = 1.2
    r_iv, r_proj, r_track, r_qr, r_qmin, r_qmax = po.is_in_frustum(cos_limit=0.5, radius_factor=3.0, size_tol=tol, **args)
    rn, rm = po.search_by_projection_ex(0, d0, r_proj[:, :2].copy(), r_qr, r_qmin, r_qmax, k1, d1, s1, (0.0, 0.0, 640.0, 480.0), claim=True,
                                        th=100.0, nnratio=0.8, ratio_same_scale=True, tol=tol)
    dev = {k: (_t(v) if k in ("Pw", "normal", "min_dist", "max_dist", "ref_size", "ref_sigma", "ref_dist") else v) for k, v in args.items()}
    iv, proj, track, qr, qmin, qmax = pkg.is_in_frustum(cos_limit=0.5, radius_factor=3.0, size_tol=tol, **dev)
    fm = pkg.FeatureMatcher(nnratio=0.8, desc_type=0, th_low=100.0)
    mq, nm = fm.search_by_projection_ex(out[1][0, :M].contiguous(), proj[:, :2].contiguous(), qr, qmin, qmax, _t(np.array([0, M], np.int32)),
                                        out[0], out[1], out[2], out[3], _t(np.array([1], np.int32)), (0.0, 0.0, 640.0, 480.0), claim=True,
                                        ratio_same_scale=True, size_tolerance=tol)
    torch.cuda.synchronize()
    assert int(nm[0]) == rn and rn > 200
    assert (mq.cpu().numpy() == rm).all()
    assert (mq.cpu().numpy()[::7] == -1).all()
