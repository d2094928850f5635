"""CUDA path vs the reference's OWN code (oracle/_ref/libafv_ref.so, built by oracle/build_ref.py from /root/reference in the
build container; the binary travels to the GPU box).  Direct version of what tests/test_oracle_vs_ref.py + the oracle parity
tests establish transitively: SearchForInitialization and DistributeOctTree (through the extractor's octree tap)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libafv_ref.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libafv_ref.so was not built (needs /root/reference at build time)")
    return C.CDLL(REF_SO)


@pytest.mark.parametrize("feature,dt,dcols,dtype,th", [("orb32", 0, 32, 0, 75.0), ("akaze61", 1, 61, 0, 128.0), ("sift128", 5, 128, 5, 0.5)])
def test_search_for_initialization_gpu_vs_reference_code(pkg, synth, ref, feature, dt, dcols, dtype, th):
    import torch
    frames, _ = synth.stream_frames(640, 480, 21, 2)
    ex = pkg.FeatureExtractor(feature, nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    out = ex.alloc_device_outputs(2)
    ex.extract_batch_device(torch.from_numpy(frames).cuda(), out)
    torch.cuda.synchronize()
    ex.status()
    n = out[3].cpu().numpy()
    fm = pkg.FeatureMatcher(nnratio=0.9, check_ori=True, desc_type=dt, th_low=th)
    pa = torch.tensor([0], dtype=torch.int32, device="cuda"); pb = torch.tensor([1], dtype=torch.int32, device="cuda")
    max_size = float(np.float32(1.2) ** np.float32(7))
    m12, nm = fm.search_for_initialization(out[0], out[1], out[2], out[3], pa, pb, None, (0.0, 0.0, 640.0, 480.0), max_size, window=100)
    torch.cuda.synchronize()
    k = [np.ascontiguousarray(out[0][i, :int(n[i])].cpu().numpy()) for i in range(2)]            # [n,7] float32 rows == cv::KeyPoint
    d = [np.ascontiguousarray(out[1][i, :int(n[i])].cpu().numpy()) for i in range(2)]
    s = [np.ascontiguousarray(out[2][i, :int(n[i])].cpu().numpy()) for i in range(2)]
    prev = np.ascontiguousarray(k[0][:, :2]).copy()
    m_r = np.zeros(int(n[0]), np.int32)
    n_r = ref.ref_search_for_initialization(dt, dcols, dtype, _p(k[0]), _p(d[0]), _p(s[0]), int(n[0]), _p(k[1]), _p(d[1]), _p(s[1]), int(n[1]),
                                            C.c_float(0.0), C.c_float(0.0), C.c_float(640.0), C.c_float(480.0), C.c_float(max_size),
                                            _p(prev), 100, C.c_float(th), C.c_float(0.9), 1, _p(m_r))
    assert int(nm[0]) == n_r and n_r > 20
    assert (m12[0, :int(n[0])].cpu().numpy() == m_r).all()
    ex.close()


def test_octree_gpu_vs_reference_code(pkg, synth, ref):
    """The extractor's octree keep lists (tap 4) == FeatureExtractor::DistributeOctTree of the reference (monotonic heap) run on
    the extractor's own detect lists (tap 3), for every level of two frames."""
    frames, _ = synth.stream_frames(640, 480, 22, 2)
    ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
    ex.extract_batch(frames)
    _, q = ex.levels()
    from oracle import pyoracle as po
    ls = po.level_geometry(640, 480)[2]                                          # cv::ORB level scales (float)
    for f in range(2):
        for l in range(8):
            det = ex.debug_read(3, f, l).view(np.uint32).reshape(-1, 2)
            x = (det[:, 0] & 0xfff).astype(np.int64); y = ((det[:, 0] >> 12) & 0xfff).astype(np.int64)
            order = np.argsort(y * 4096 + x, kind="stable")                      # the list is unordered on the device: raster order
            x, y, resp = x[order], y[order], det[order, 1].copy().view(np.float32)
            px = (x.astype(np.float32) * ls[l]).astype(np.float32); py = (y.astype(np.float32) * ls[l]).astype(np.float32)
            ox = np.zeros(len(px) + 8, np.float32); oy = np.zeros_like(ox); oi = np.zeros_like(ox)
            m = ref.ref_distribute_octree(_p(px), _p(py), _p(np.ascontiguousarray(resp)), len(px), 0, 640, 0, 480, int(q[l]), _p(ox), _p(oy), _p(oi), len(ox))
            keep = ex.debug_read(4, f, l).view(np.uint32).reshape(-1, 2)
            got = list(zip((keep[:, 0] & 0xfff).tolist(), ((keep[:, 0] >> 12) & 0xfff).tolist()))
            want = [(int(x[int(i)]), int(y[int(i)])) for i in oi[:m]]
            assert got == want, (f, l)
    ex.close()
