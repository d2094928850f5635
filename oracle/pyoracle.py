"""ctypes front end of the CPU ORACLE (oracle/libafv_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

DESC_ORB, DESC_AKAZE61, DESC_BRISK, DESC_SIFT128 = 0, 1, 2, 5      # include/Types.h:24-34


def build(force=False):
    so = os.path.join(_DIR, "libafv_oracle.so")
    srcs = [os.path.join(_DIR, f) for f in ("afv_oracle.c", "afv_oracle_match.c", "afv_oracle_sift.c", "afv_oracle_akaze.c", "afv_oracle_brisk.c", "afv_oracle_orbslam2.c", "afv_oracle_batch.c", "afv_oracle.h", "orb_pattern.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _DIR, "libafv_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_orb_pyramid.restype = C.c_long
        _LIB.orc_harris7.restype = C.c_float
        _LIB.orc_ic_angle.restype = C.c_float
        _LIB.orc_fast_atan2.restype = C.c_float
        _LIB.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _LIB.orc_descriptor_distance.restype = C.c_float
        _LIB.orc_orb32_extract_match_batch.restype = C.c_long
        _LIB.orc_sift128_extract_match_batch.restype = C.c_long
        _LIB.orc_orbslam2_extract_match_batch.restype = C.c_long
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(v):
    return C.c_float(float(v))


def level_geometry(w, h, nlevels=8, sf=1.2):
    lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32); ls = np.zeros(nlevels, np.float32)
    lib().orc_orb_level_geometry(w, h, nlevels, _f(sf), _p(lw), _p(lh), _p(ls))
    return lw, lh, ls


def features_per_level(nfeatures, nlevels=8, sf=1.2):
    q = np.zeros(nlevels, np.int32)
    lib().orc_features_per_level(nfeatures, nlevels, _f(sf), _p(q))
    return q


def pyramid(gray, nlevels=8, sf=1.2):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    offs = np.zeros(nlevels, np.int64); lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32)
    ls = np.zeros(nlevels, np.float32)
    total = lib().orc_orb_pyramid(_p(gray), w, h, w, nlevels, _f(sf), None, _p(offs), _p(lw), _p(lh), _p(ls))
    buf = np.zeros(total, np.uint8)
    lib().orc_orb_pyramid(_p(gray), w, h, w, nlevels, _f(sf), _p(buf), _p(offs), _p(lw), _p(lh), _p(ls))
    return [buf[offs[l]:offs[l] + int(lw[l]) * int(lh[l])].reshape(lh[l], lw[l]) for l in range(nlevels)], ls


def fast(img, threshold=20):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h // 4 + 16
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().orc_fast9_16_nms(_p(img), w, h, w, threshold, _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def harris(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    return lib().orc_harris7(_p(img), w, h, w, int(x), int(y))


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    return lib().orc_ic_angle(_p(img), w, h, w, int(x), int(y))


def blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros_like(img)
    lib().orc_blur7_level(_p(img), w, h, w, _p(out), w)
    return out


def detect_level(img, fast_th, quota):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h // 4 + 16
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); fs = np.zeros(cap, np.int32)
    hr = np.zeros(cap, np.float32)
    n = lib().orc_orb_detect_level(_p(img), w, h, w, int(fast_th), int(quota), _p(xs), _p(ys), _p(hr), _p(fs), cap)
    assert n >= 0
    return xs[:n].copy(), ys[:n].copy(), hr[:n].copy(), fs[:n].copy()


def octree(px, py, resp, w, h, N):
    px = np.ascontiguousarray(px, np.float32); py = np.ascontiguousarray(py, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    n = len(px)
    keep = np.zeros(max(n, 1), np.int32)
    m = lib().orc_distribute_octree(_p(px), _p(py), _p(resp), None, n, 0, int(w), 0, int(h), int(N), _p(keep), n)
    assert m >= 0
    return keep[:m].copy()


def orb32_extract(gray, nfeatures=1000, nlevels=8, scale_factor=1.2, detect_th=20.0, cap=None):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = cap or nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); ksz = np.zeros(cap, np.float32)
    n = C.c_int(0); nc = C.c_int(0)
    rc = lib().orc_orb32_extract(_p(gray), w, h, w, nfeatures, nlevels, _f(scale_factor), _f(detect_th),
                                 _p(kps), _p(desc), _p(ksz), cap, C.byref(n), C.byref(nc))
    assert rc == 0, rc
    return kps[:n.value].copy(), desc[:n.value].copy(), ksz[:n.value].copy(), nc.value


def descriptor_distance(desc_type, a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return lib().orc_descriptor_distance(desc_type, _p(a), _p(b))


def search_for_initialization(desc_type, k1, d1, k2, d2, size2, bounds, max_kpt_size, prev_matched,
                              window=100, th_low=75.0, nnratio=0.9, check_ori=True):
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
    d1 = np.ascontiguousarray(d1); d2 = np.ascontiguousarray(d2)
    size2 = np.ascontiguousarray(size2, np.float32)
    pm = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.zeros(max(len(k1), 1), np.int32)
    minX, minY, maxX, maxY = bounds
    n = lib().orc_search_for_initialization(desc_type, _p(k1), _p(d1), len(k1), _p(k2), _p(d2), _p(size2), len(k2),
                                            _f(minX), _f(minY), _f(maxX), _f(maxY), _f(max_kpt_size), _p(pm),
                                            int(window), _f(th_low), _f(nnratio), int(bool(check_ori)), _p(m12))
    return n, m12[:len(k1)].copy(), pm


def match_window(desc_type, q, qxy, qr, qmin, qmax, tk, td, tsize, bounds):
    q = np.ascontiguousarray(q); td = np.ascontiguousarray(td); tk = np.ascontiguousarray(tk)
    qxy = np.ascontiguousarray(qxy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    qmin = np.ascontiguousarray(qmin, np.float32); qmax = np.ascontiguousarray(qmax, np.float32)
    tsize = np.ascontiguousarray(tsize, np.float32)
    nq = len(q)
    best = np.zeros(nq, np.int32); bd = np.zeros(nq, np.float32); sd = np.zeros(nq, np.float32)
    bs = np.zeros(nq, np.float32); ss = np.zeros(nq, np.float32)
    minX, minY, maxX, maxY = bounds
    lib().orc_match_window(desc_type, _p(q), _p(qxy), _p(qr), _p(qmin), _p(qmax), nq, _p(tk), _p(td), _p(tsize),
                           len(tk), _f(minX), _f(minY), _f(maxX), _f(maxY), _p(best), _p(bd), _p(sd), _p(bs), _p(ss))
    return best, bd, sd, bs, ss


def match_bruteforce(desc_type, q, t):
    q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
    nq = len(q)
    best = np.zeros(nq, np.int32); bd = np.zeros(nq, np.float32); sd = np.zeros(nq, np.float32)
    lib().orc_match_bruteforce(desc_type, _p(q), nq, _p(t), len(t), _p(best), _p(bd), _p(sd))
    return best, bd, sd


def search_by_bow(desc_type, dkf, kkf, kf_segs, df, kf_f, f_segs, th_low=75.0, nnratio=0.7, check_ori=True):
    """kf_segs / f_segs: (node_ids, starts(len+1), idx) sorted by node id."""
    dkf = np.ascontiguousarray(dkf); df = np.ascontiguousarray(df)
    kkf = np.ascontiguousarray(kkf); kf_f = np.ascontiguousarray(kf_f)
    a = [np.ascontiguousarray(v, np.int32) for v in kf_segs]
    b = [np.ascontiguousarray(v, np.int32) for v in f_segs]
    nf = len(df)
    out = np.zeros(max(nf, 1), np.int32)
    n = lib().orc_search_by_bow(desc_type, _p(dkf), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(kkf),
                                _p(df), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), _p(kf_f), nf,
                                _f(th_low), _f(nnratio), int(bool(check_ori)), _p(out))
    return n, out[:nf].copy()


def extract_match_batch(frames, pair_a, pair_b, nfeatures=1000, nthreads=1, window=100, th_low=75.0, nnratio=0.9, check_ori=True):
    """One bench step on the CPU, threaded in C (OpenMP). Returns the total number of matches."""
    frames = np.ascontiguousarray(frames, np.uint8)
    B, h, w = frames.shape
    pa = np.ascontiguousarray(pair_a, np.int32); pb = np.ascontiguousarray(pair_b, np.int32)
    r = lib().orc_orb32_extract_match_batch(_p(frames), B, w, h, nfeatures, 8, _f(1.2), _f(20.0), _p(pa), _p(pb), len(pa),
                                            int(window), _f(th_low), _f(nnratio), int(bool(check_ori)), int(nthreads))
    assert r >= 0
    return int(r)


def gray_from_color(img, rgb=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_gray_from_color(_p(img), ch, int(bool(rgb)), w, h, w * ch, _p(out), w)
    return out


def bow_transform(desc_type, desc, tree, levelsup=4):
    """tree = dict(child_off, child_ids, node_desc, node_word, node_weight, L). Returns (word_id, weight, node_id)."""
    desc = np.ascontiguousarray(desc)
    n = len(desc)
    co = np.ascontiguousarray(tree["child_off"], np.int32); ci = np.ascontiguousarray(tree["child_ids"], np.int32)
    nd = np.ascontiguousarray(tree["node_desc"]); nw = np.ascontiguousarray(tree["node_word"], np.int32)
    wt = np.ascontiguousarray(tree["node_weight"], np.float64)
    wid = np.zeros(max(n, 1), np.int32); w = np.zeros(max(n, 1), np.float64); nid = np.zeros(max(n, 1), np.int32)
    lib().orc_bow_transform(desc_type, _p(desc), n, _p(co), _p(ci), _p(nd), _p(nw), _p(wt), int(tree["L"]), int(levelsup),
                            _p(wid), _p(w), _p(nid))
    return wid[:n].copy(), w[:n].copy(), nid[:n].copy()


def feature_vector_segments(node_id):
    """DBoW2::FeatureVector as SearchByBoW consumes it: sorted node ids + CSR of feature indices (ascending)."""
    node_id = np.asarray(node_id)
    order = np.argsort(node_id, kind="stable")
    ids, starts = np.unique(node_id[order], return_index=True)
    return ids.astype(np.int32), np.append(starts, len(node_id)).astype(np.int32), order.astype(np.int32)


def search_by_projection(desc_type, qdesc, qxy, qr, qmin, qmax, tk, td, tsize, bounds, occupied=None, th=75.0, nnratio=0.8,
                         ratio_same_scale=True, tol=1.2):
    qdesc = np.ascontiguousarray(qdesc); td = np.ascontiguousarray(td); tk = np.ascontiguousarray(tk)
    qxy = np.ascontiguousarray(qxy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    qmin = np.ascontiguousarray(qmin, np.float32); qmax = np.ascontiguousarray(qmax, np.float32)
    tsize = np.ascontiguousarray(tsize, np.float32)
    nq = len(qdesc)
    out = np.zeros(max(nq, 1), np.int32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    minX, minY, maxX, maxY = bounds
    n = lib().orc_search_by_projection(desc_type, _p(qdesc), _p(qxy), _p(qr), _p(qmin), _p(qmax), nq, _p(tk), _p(td), _p(tsize), len(tk),
                                       None if occ is None else _p(occ), _f(minX), _f(minY), _f(maxX), _f(maxY), _f(th), _f(nnratio),
                                       int(bool(ratio_same_scale)), _f(tol), _p(out))
    return n, out[:nq].copy()


def distinctive_descriptor(desc_type, desc, obs):
    desc = np.ascontiguousarray(desc); obs = np.ascontiguousarray(obs, np.int32)
    return int(lib().orc_distinctive_descriptor(desc_type, _p(desc), _p(obs), len(obs)))


# ---- sift128 (oracle/afv_oracle_sift.c; PARITY UNPINNED, see its header) --------------------------------------------
def sift_scale_space(gray, what, octave, idx):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    ow = C.c_int(0); oh = C.c_int(0)
    L = lib(); L.orc_sift_scale_space.restype = C.c_long
    n = L.orc_sift_scale_space(_p(gray), w, h, w, int(what), int(octave), int(idx), None, C.byref(ow), C.byref(oh))
    assert n > 0
    out = np.zeros((oh.value, ow.value), np.float32)
    L.orc_sift_scale_space(_p(gray), w, h, w, int(what), int(octave), int(idx), _p(out), C.byref(ow), C.byref(oh))
    return out


def sift_detect(gray, nfeatures, with_desc=True, cap=200000):
    """SiftGPU-equivalent list after the -tc2 limit: (n,4) x,y,s,o and (n,128) descriptors."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    xyso = np.zeros((cap, 4), np.float32)
    desc = np.zeros((cap, 128), np.float32) if with_desc else None
    n = lib().orc_sift_detect(_p(gray), w, h, w, int(nfeatures), _p(xyso), _p(desc) if with_desc else None, cap)
    assert n >= 0, n
    return xyso[:n].copy(), (desc[:n].copy() if with_desc else None)


def sift128_extract(gray, nfeatures, nlevels=8, scale_factor=2.0):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 128), np.float32); size = np.zeros(cap, np.float32)
    n = C.c_int(0); nd = C.c_int(0)
    rc = lib().orc_sift128_extract(_p(gray), w, h, w, int(nfeatures), int(nlevels), _f(scale_factor), _p(kps), _p(desc),
                                   _p(size), cap, C.byref(n), C.byref(nd))
    assert rc == 0, rc
    return kps[:n.value].copy(), desc[:n.value].copy(), size[:n.value].copy(), nd.value


def sift_extract_match_batch(frames, pair_a, pair_b, nfeatures=2000, nthreads=1, window=100, th_low=0.5, nnratio=0.9, check_ori=True):
    """One bench step of the sift128 workload on the CPU (pthreads in C). Returns the total number of matches."""
    frames = np.ascontiguousarray(frames, np.uint8)
    B, h, w = frames.shape
    pa = np.ascontiguousarray(pair_a, np.int32); pb = np.ascontiguousarray(pair_b, np.int32)
    L = lib(); L.orc_sift128_extract_match_batch.restype = C.c_long
    r = L.orc_sift128_extract_match_batch(_p(frames), B, w, h, int(nfeatures), 8, _f(2.0), _p(pa), _p(pb), len(pa),
                                          int(window), _f(th_low), _f(nnratio), int(bool(check_ori)), int(nthreads))
    assert r >= 0
    return int(r)


# ---- akaze61 (oracle/afv_oracle_akaze.c; PARITY UNPINNED, see its header) -------------------------------------------
def akaze_scale_space(gray, what, level, omax=2, nsub=4):
    """what: 0 Lt, 1 Lsmooth, 2 Lx, 3 Ly, 4 Ldet of evolution level `level`; returns (image, kcontrast)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    ow = C.c_int(0); oh = C.c_int(0); kc = C.c_float(0)
    L = lib(); L.orc_akaze_scale_space.restype = C.c_long
    n = L.orc_akaze_scale_space(_p(gray), w, h, w, omax, nsub, int(what), int(level), None, C.byref(ow), C.byref(oh), C.byref(kc))
    assert n > 0
    out = np.zeros((oh.value, ow.value), np.float32)
    L.orc_akaze_scale_space(_p(gray), w, h, w, omax, nsub, int(what), int(level), _p(out), C.byref(ow), C.byref(oh), C.byref(kc))
    return out, kc.value


def akaze_detect(gray, dth=5e-4, omax=2, nsub=4, cap=100000):
    """Feature_Detection list: (n,5) x, y, size, response, class_id."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.zeros((cap, 5), np.float32)
    n = lib().orc_akaze_detect(_p(gray), w, h, w, omax, nsub, _f(dth), _p(out), cap)
    assert n >= 0, n
    return out[:n].copy()


def akaze61_extract(gray, nfeatures, nlevels=8, scale_factor=1.1892, detect_th=5e-4):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 61), np.uint8); size = np.zeros(cap, np.float32)
    n = C.c_int(0); nd = C.c_int(0)
    rc = lib().orc_akaze61_extract(_p(gray), w, h, w, int(nfeatures), int(nlevels), _f(scale_factor), _f(detect_th), _p(kps),
                                   _p(desc), _p(size), cap, C.byref(n), C.byref(nd))
    assert rc == 0, rc
    return kps[:n.value].copy(), desc[:n.value].copy(), size[:n.value].copy(), nd.value


def akaze_extract_match_batch(frames, pair_a, pair_b, nfeatures=1000, nthreads=1, window=100, th_akaze=128.0, th_brisk=120.0,
                              nnratio=0.9, check_ori=True):
    """One bench step of the akaze61 (+ brisk48 layout) workload on the CPU (pthreads in C). Returns the total matches."""
    frames = np.ascontiguousarray(frames, np.uint8)
    B, h, w = frames.shape
    pa = np.ascontiguousarray(pair_a, np.int32); pb = np.ascontiguousarray(pair_b, np.int32)
    L = lib(); L.orc_akaze61_extract_match_batch.restype = C.c_long
    r = L.orc_akaze61_extract_match_batch(_p(frames), B, w, h, int(nfeatures), 8, _f(1.1892), _f(5e-4), _p(pa), _p(pb), len(pa),
                                          int(window), _f(th_akaze), _f(th_brisk), _f(nnratio), int(bool(check_ori)), int(nthreads))
    assert r >= 0
    return int(r)


def undistort_keypoints(kps, K4, dist5):
    """Frame::UndistortKeyPoints on a KP_DTYPE array; K4 = (fx, fy, cx, cy), dist5 = (k1, k2, p1, p2, k3) as float32."""
    kps = np.ascontiguousarray(kps)
    K4 = np.ascontiguousarray(K4, np.float32); dist5 = np.ascontiguousarray(dist5, np.float32)
    out = np.zeros_like(kps)
    lib().orc_undistort_keypoints(_p(kps), len(kps), _p(K4), _p(dist5), _p(out))
    return out


def features_in_area(kps, kpsize, bounds, x, y, r, min_size, max_size):
    """Frame::AssignFeaturesToGrid + GetFeaturesInArea (src/Frame.cc:225-240, :333-382): indices in reference order."""
    kps = np.ascontiguousarray(kps); kpsize = np.ascontiguousarray(kpsize, np.float32)
    n = len(kps)
    minX, minY, maxX, maxY = bounds
    invW = np.float32(64.0) / (np.float32(maxX) - np.float32(minX)); invH = np.float32(48.0) / (np.float32(maxY) - np.float32(minY))
    cs = np.zeros(64 * 48 + 1, np.int32); ci = np.zeros(max(n, 1), np.int32)
    lib().orc_grid_build(_p(kps), n, _f(minX), _f(minY), _f(invW), _f(invH), _p(cs), _p(ci))
    out = np.zeros(max(n, 1), np.int32)
    m = lib().orc_features_in_area(_p(kps), _p(kpsize), _p(cs), _p(ci), _f(minX), _f(minY), _f(invW), _f(invH), _f(x), _f(y), _f(r),
                                   _f(min_size), _f(max_size), _p(out), len(out))
    return out[:m].copy()


# ---- brisk48 (oracle/afv_oracle_brisk.c; PARITY UNPINNED vs ETH brisk v2, semi-pinned to cv2.BRISK, see its header) ----
BRISK_SEQUENTIAL, BRISK_DENSE = 0, 1


def resize_area(img, dw, dh):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_area_u8(_p(img), w, h, w, _p(out), dw, dh, dw)
    return out


def brisk_layer(gray, what, layer, octaves=4):
    """what = 0: pyramid layer image, 1: true AGAST 9-16 score image (0 below 1)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    ow = C.c_int(0); oh = C.c_int(0)
    L = lib(); L.orc_brisk_layer.restype = C.c_long
    n = L.orc_brisk_layer(_p(gray), w, h, w, int(octaves), int(what), int(layer), None, C.byref(ow), C.byref(oh))
    if n < 0:
        raise ValueError("layer %d not built" % layer)
    out = np.zeros((oh.value, ow.value), np.uint8)
    L.orc_brisk_layer(_p(gray), w, h, w, int(octaves), int(what), int(layer), _p(out), C.byref(ow), C.byref(oh))
    return out


def brisk_detect(gray, threshold=34, octaves=4, mode=BRISK_DENSE, cap=200000):
    """BriskFeatureDetector(threshold, octaves).detect: rows (x, y, size, response, layer) in detection order."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.zeros((cap, 5), np.float32)
    n = lib().orc_brisk_detect(_p(gray), w, h, w, int(threshold), int(octaves), int(mode), _p(out), cap)
    if n < 0:
        raise RuntimeError("orc_brisk_detect failed / capacity (%d)" % n)
    return out[:n].copy()


def brisk_describe(gray, kps, nbytes=48, libm_angle=False):
    """BriskDescriptorExtractor.compute on a KP_DTYPE array: returns (kept keypoints with angle, descriptors, kept index)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    k = np.ascontiguousarray(kps, KP_DTYPE).copy()
    desc = np.zeros((max(len(k), 1), nbytes), np.uint8)
    idx = np.zeros(max(len(k), 1), np.int32)
    m = lib().orc_brisk_describe(_p(gray), w, h, w, _p(k), len(k), int(nbytes), int(libm_angle), _p(desc), _p(idx))
    return k[:m], desc[:m], idx[:m]


def brisk_scale_index(size):
    return int(lib().orc_brisk_scale_index(_f(size)))


def brisk48_extract(gray, nfeatures, nlevels=8, scale_factor=1.5, detect_th=34.0, mode=BRISK_DENSE):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 48), np.uint8); size = np.zeros(cap, np.float32)
    n = C.c_int(0); nd = C.c_int(0)
    rc = lib().orc_brisk48_extract(_p(gray), w, h, w, int(nfeatures), int(nlevels), _f(scale_factor), _f(detect_th), int(mode),
                                   _p(kps), _p(desc), _p(size), cap, C.byref(n), C.byref(nd))
    if rc:
        raise RuntimeError("orc_brisk48_extract failed (%d)" % rc)
    return kps[:n.value], desc[:n.value], size[:n.value], nd.value


# ---- the remaining FeatureMatcher searches (rows a19 / a20) on arrays --------------------------------------------------
def search_by_projection_ex(desc_type, qdesc, qxy, qr, qmin, qmax, tk, td, tsize, bounds, qangle=None, tinf1d=None, occupied=None,
                            claim=True, th=75.0, nnratio=0.8, ratio_same_scale=False, tol=1.2):
    """All projection-type searches after their projection prologue (see orc_search_by_projection_ex)."""
    qdesc = np.ascontiguousarray(qdesc); td = np.ascontiguousarray(td); tk = np.ascontiguousarray(tk)
    qxy = np.ascontiguousarray(qxy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    qmin = np.ascontiguousarray(qmin, np.float32); qmax = np.ascontiguousarray(qmax, np.float32)
    tsize = np.ascontiguousarray(tsize, np.float32)
    qa = None if qangle is None else np.ascontiguousarray(qangle, np.float32)
    ti = None if tinf1d is None else np.ascontiguousarray(tinf1d, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    nq = len(qdesc)
    out = np.zeros(max(nq, 1), np.int32)
    minX, minY, maxX, maxY = bounds
    n = lib().orc_search_by_projection_ex(desc_type, _p(qdesc), _p(qxy), _p(qr), _p(qmin), _p(qmax), None if qa is None else _p(qa), nq,
                                          _p(tk), _p(td), _p(tsize), None if ti is None else _p(ti), len(tk), None if occ is None else _p(occ),
                                          int(bool(claim)), _f(minX), _f(minY), _f(maxX), _f(maxY), _f(th), _f(nnratio),
                                          int(bool(ratio_same_scale)), _f(tol), _p(out))
    return n, out[:nq].copy()


def search_by_sim3(desc_type, k1, d1, size1, q1xy, q1r, q1min, q1max, k2, d2, size2, q2xy, q2r, q2min, q2max, bounds, th_high):
    """SearchBySim3: direction 1 -> 2 queries are KF1's keypoints (descriptor d1[i], projected to q1xy[i] in KF2), and vice versa."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2); d1 = np.ascontiguousarray(d1); d2 = np.ascontiguousarray(d2)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    size1, q1xy, q1r, q1min, q1max, size2, q2xy, q2r, q2min, q2max = map(f, (size1, q1xy, q1r, q1min, q1max, size2, q2xy, q2r, q2min, q2max))
    out = np.zeros(max(len(k1), 1), np.int32)
    minX, minY, maxX, maxY = bounds
    n = lib().orc_search_by_sim3(desc_type, _p(d1), _p(q1xy), _p(q1r), _p(q1min), _p(q1max), len(k1), _p(d2), _p(q2xy), _p(q2r), _p(q2min), _p(q2max),
                                 len(k2), _p(k1), _p(d1), _p(size1), _p(k2), _p(d2), _p(size2), _f(minX), _f(minY), _f(maxX), _f(maxY), _f(th_high), _p(out))
    return n, out[:len(k1)].copy()


def bow_match(mode, desc_type, k1, d1, node1, valid1, k2, d2, node2, valid2, th_low=75.0, nnratio=0.7, check_ori=True, F12=None, epipole=(0.0, 0.0),
              sigma2_2=None):
    """mode 0 SearchByBoW(KF,F) -> out[i2] = i1; 1 SearchByBoW(KF,KF) -> out[i1] = i2; 2 SearchForTriangulation -> out[i1] = i2."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2); d1 = np.ascontiguousarray(d1); d2 = np.ascontiguousarray(d2)
    node1 = np.ascontiguousarray(node1, np.int32); node2 = np.ascontiguousarray(node2, np.int32)
    v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8); v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
    F = np.zeros(9, np.float32) if F12 is None else np.ascontiguousarray(F12, np.float32).reshape(9)
    s2 = np.ones(len(k2), np.float32) if sigma2_2 is None else np.ascontiguousarray(sigma2_2, np.float32)
    nout = len(k2) if mode == 0 else len(k1)
    out = np.zeros(max(nout, 1), np.int32)
    n = lib().orc_bow_match(int(mode), desc_type, _p(k1), _p(d1), _p(node1), None if v1 is None else _p(v1), len(k1), _p(k2), _p(d2), _p(node2),
                            None if v2 is None else _p(v2), len(k2), _f(th_low), _f(nnratio), int(bool(check_ori)), _p(F), _f(epipole[0]),
                            _f(epipole[1]), _p(s2), _p(out))
    return n, out[:nout].copy()


# ---- vanilla ORB-SLAM2 extractor (src/ORBextractor.cc:460-676 built with VANILLA_ORB_SLAM2) --------------------------------
def resize_linear(img, dw, dh):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(img), w, h, w, _p(out), dw, dh, dw)
    return out


def gaussblur7_fixed(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros_like(img)
    lib().orc_gaussblur7_fixed_u8(_p(img), w, h, w, _p(out), w)
    return out


def orbslam2_geometry(w, h, nlevels=8, sf=1.2):
    lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32)
    s = np.zeros(nlevels, np.float32); i = np.zeros(nlevels, np.float32)
    assert lib().orc_orbslam2_geometry(w, h, nlevels, _f(sf), _p(lw), _p(lh), _p(s), _p(i)) == 0
    return lw, lh, s, i


def orbslam2_pyramid(gray, nlevels=8, sf=1.2):
    lw, lh, _, _ = orbslam2_geometry(gray.shape[1], gray.shape[0], nlevels, sf)
    pyr = [np.ascontiguousarray(gray, np.uint8)]
    for l in range(1, nlevels):
        pyr.append(resize_linear(pyr[-1], int(lw[l]), int(lh[l])))
    return pyr


def orbslam2_detect_level(img, ini_th=20, min_th=7):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h // 4 + 16
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().orc_orbslam2_detect_level(_p(img), w, h, w, int(ini_th), int(min_th), _p(xs), _p(ys), _p(sc), cap)
    assert 0 <= n <= cap, n
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def orbslam2_extract(gray, nfeatures=1000, nlevels=8, scale_factor=1.2, ini_th=20, min_th=7, cap=None):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = cap or nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); ksz = np.zeros(cap, np.float32)
    n = C.c_int(0)
    rc = lib().orc_orbslam2_extract(_p(gray), w, h, w, nfeatures, nlevels, _f(scale_factor), int(ini_th), int(min_th),
                                    _p(kps), _p(desc), _p(ksz), cap, C.byref(n))
    assert rc == 0, rc
    return kps[:n.value].copy(), desc[:n.value].copy(), ksz[:n.value].copy()


def is_in_frustum(Pw, normal, min_dist, max_dist, ref_size, ref_sigma, ref_dist, pose16, cam5, bounds4, cos_limit=0.5,
                  radius_factor=1.0, size_tol=1.5):
    """Frame::isInFrustum (src/Frame.cc:276-331) + SearchByProjection's window prologue; returns in_view, proj3, track3, qr, qmin, qmax."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    Pw, normal = f32(Pw), f32(normal)
    M = len(Pw)
    iv = np.zeros(M, np.uint8); proj = np.zeros((M, 3), np.float32); track = np.zeros((M, 3), np.float32)
    qr = np.zeros(M, np.float32); qmin = np.zeros(M, np.float32); qmax = np.zeros(M, np.float32)
    lib().orc_is_in_frustum(_p(Pw), _p(normal), _p(f32(min_dist)), _p(f32(max_dist)), _p(f32(ref_size)), _p(f32(ref_sigma)), _p(f32(ref_dist)),
                            M, _p(f32(pose16)), _p(f32(cam5)), _p(f32(bounds4)), _f(cos_limit), _f(radius_factor), _f(size_tol),
                            _p(iv), _p(proj), _p(track), _p(qr), _p(qmin), _p(qmax))
    return iv, proj, track, qr, qmin, qmax


def orbslam2_extract_match_batch(frames, pair_a, pair_b, nfeatures=1000, nthreads=1, window=100, th_low=75.0, nnratio=0.9, check_ori=True):
    """One bench step of the vanilla ORB-SLAM2 workload on the CPU (pthreads in C).  Returns the total number of matches."""
    frames = np.ascontiguousarray(frames, np.uint8)
    B, h, w = frames.shape
    pa = np.ascontiguousarray(pair_a, np.int32); pb = np.ascontiguousarray(pair_b, np.int32)
    r = lib().orc_orbslam2_extract_match_batch(_p(frames), B, w, h, nfeatures, 8, _f(1.2), _p(pa), _p(pb), len(pa), int(window), _f(th_low),
                                               _f(nnratio), int(bool(check_ori)), int(nthreads))
    assert r >= 0
    return int(r)
