"""FeatureMatcher mirror (reference include/FeatureMatcher.h:36-118) above the C ABI, on device arrays.

The reference methods take Frame / KeyFrame / MapPoint graphs; here the same searches run on the flat
outputs of FeatureExtractor.extract_batch_device (torch cuda tensors).  Thresholds follow
FeatureMatcher::setDescriptorDistanceThresholds (src/FeatureMatcher.cc:1533-1545): TH_LOW = TH_HIGH =
FeatureMatcher.matchingTh of settings/<feat>_settings.yaml.
"""
import ctypes as C

import numpy as np


def _afv():
    from . import lib, _check, _vp, _stream_ptr
    return lib(), _check, _vp, _stream_ptr


class FeatureMatcher:
    HISTO_LENGTH = 30                       # src/FeatureMatcher.cc:64

    def __init__(self, nnratio=0.6, check_ori=True, desc_type=0, th_low=75.0):
        self.nnratio = float(nnratio)
        self.check_ori = bool(check_ori)
        self.desc_type = int(desc_type)
        self.th_low = float(th_low)         # TH_LOW == TH_HIGH == reloc thresholds

    # -- SearchForInitialization (src/FeatureMatcher.cc:399-557), batched over frame pairs ------------
    def sfi_workspace(self, P, cap, device):
        """Caller-owned scratch for search_for_initialization(workspace=...): nothing is allocated on the call path."""
        import torch
        lib, _, _, _ = _afv()
        lib.afv_search_for_initialization_workspace_bytes.restype = C.c_size_t
        nbytes = int(lib.afv_search_for_initialization_workspace_bytes(self.desc_type, int(P), int(cap)))
        return torch.empty(nbytes, dtype=torch.uint8, device=device)

    def search_for_initialization(self, kps, desc, kpsize, n, pair_a, pair_b, prev_matched, bounds,
                                  max_kpt_size, window=100, matches12=None, nmatches=None, stream=None, workspace=None):
        """kps [B,cap,7] f32, desc [B,cap,D] u8, kpsize [B,cap] f32, n [B] i32, pair_a/pair_b [P] i32,
        prev_matched [P,cap,2] f32 (in/out).  Returns (matches12 [P,cap] i32, nmatches [P] i32)."""
        import torch
        lib, _check, _vp, _sp = _afv()
        B, cap = kps.shape[0], kps.shape[1]
        P = pair_a.shape[0]
        if matches12 is None:
            matches12 = torch.empty((P, cap), dtype=torch.int32, device=kps.device)
        if nmatches is None:
            nmatches = torch.empty((P,), dtype=torch.int32, device=kps.device)
        minx, miny, maxx, maxy = bounds
        ws_ptr, ws_bytes = (None, 0) if workspace is None else (_vp(workspace), workspace.numel() * workspace.element_size())
        _check(lib.afv_search_for_initialization_ws(
            self.desc_type, _vp(kps), _vp(desc), _vp(kpsize), _vp(n), B, cap, _vp(pair_a), _vp(pair_b), P,
            C.c_float(minx), C.c_float(miny), C.c_float(maxx), C.c_float(maxy), C.c_float(max_kpt_size),
            _vp(prev_matched), int(window), C.c_float(self.th_low), C.c_float(self.nnratio), int(self.check_ori),
            _vp(matches12), _vp(nmatches), ws_ptr, C.c_size_t(ws_bytes), _sp(stream)))
        return matches12, nmatches

    # -- Frame grid + stateless window search (core of SearchByProjection) ------------------------------
    @staticmethod
    def grid_build(kps, n, bounds, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        B, cap = kps.shape[0], kps.shape[1]
        cs = torch.empty((B, 64 * 48 + 1), dtype=torch.int32, device=kps.device)
        ci = torch.empty((B, cap), dtype=torch.int32, device=kps.device)
        minx, miny, maxx, maxy = bounds
        _check(lib.afv_grid_build(_vp(kps), _vp(n), B, cap, C.c_float(minx), C.c_float(miny), C.c_float(maxx),
                                  C.c_float(maxy), _vp(cs), _vp(ci), _sp(stream)))
        return cs, ci

    def match_window(self, q, qxy, qr, qmin, qmax, tk, td, tsize, nt, cell_start, cell_items, bounds, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        nq = q.shape[0]
        dev = q.device
        best = torch.empty(nq, dtype=torch.int32, device=dev)
        bd = torch.empty(nq, dtype=torch.float32, device=dev); sd = torch.empty_like(bd)
        bs = torch.empty_like(bd); ss = torch.empty_like(bd)
        minx, miny, maxx, maxy = bounds
        _check(lib.afv_match_window(self.desc_type, _vp(q), _vp(qxy), _vp(qr), _vp(qmin), _vp(qmax), nq,
                                    _vp(tk), _vp(td), _vp(tsize), int(nt), _vp(cell_start), _vp(cell_items),
                                    C.c_float(minx), C.c_float(miny), C.c_float(maxx), C.c_float(maxy),
                                    _vp(best), _vp(bd), _vp(sd), _vp(bs), _vp(ss), _sp(stream)))
        return best, bd, sd, bs, ss

    # -- SearchByProjection family (src/FeatureMatcher.cc:73-154, :287-397), batched over problems ------------------
    def search_by_projection(self, qdesc, qxy, qr, qmin, qmax, q_start, kps, desc, kpsize, n, frame, bounds, occupied=None,
                             ratio_same_scale=True, size_tolerance=1.2, stream=None):
        """qdesc [Q,D] u8, qxy [Q,2], qr/qmin/qmax [Q] f32, q_start [P+1] i32, frame [P] i32 (train frame of problem p in the
        B x cap arrays), occupied [B,cap] u8 or None.  Returns (match_q [Q] i32 train index or -1, nmatches [P] i32)."""
        import torch
        lib, _check, _vp, _sp = _afv()
        P = frame.shape[0]
        Q = qdesc.shape[0]
        B, cap = kps.shape[0], kps.shape[1]
        match_q = torch.empty(max(Q, 1), dtype=torch.int32, device=kps.device)
        nm = torch.empty(max(P, 1), dtype=torch.int32, device=kps.device)
        minx, miny, maxx, maxy = bounds
        _check(lib.afv_search_by_projection(self.desc_type, _vp(qdesc), _vp(qxy), _vp(qr), _vp(qmin), _vp(qmax), _vp(q_start), P,
                                            _vp(kps), _vp(desc), _vp(kpsize), _vp(n), B, cap, _vp(frame), _vp(occupied),
                                            C.c_float(minx), C.c_float(miny), C.c_float(maxx), C.c_float(maxy), C.c_float(self.th_low),
                                            C.c_float(self.nnratio), int(bool(ratio_same_scale)), C.c_float(size_tolerance),
                                            _vp(match_q), _vp(nm), _sp(stream)))
        return match_q[:Q], nm[:P]

    def search_by_projection_ex(self, qdesc, qxy, qr, qmin, qmax, q_start, kps, desc, kpsize, n, frame, bounds, qangle=None, inf1d=None,
                                occupied=None, claim=True, ratio_same_scale=False, size_tolerance=1.2, th=None, nq_total=None,
                                workspace=None, stream=None):
        """Every projection-type search of the reference after its projection prologue (include/afv.h lists the option set of each
        reference method).  qr < 0 skips a query; qangle [Q] adds the orientation histogram; inf1d [B,cap] adds Fuse's reprojection
        gate.  Returns (match_q [Q] i32 train index or -1, nmatches [P] i32)."""
        import torch
        lib, _check, _vp, _sp = _afv()
        P = frame.shape[0]
        Q = qdesc.shape[0]
        B, cap = kps.shape[0], kps.shape[1]
        match_q = torch.empty(max(Q, 1), dtype=torch.int32, device=kps.device)
        nm = torch.empty(max(P, 1), dtype=torch.int32, device=kps.device)
        minx, miny, maxx, maxy = bounds
        ws_ptr, ws_bytes = (None, 0) if workspace is None else (_vp(workspace), workspace.numel() * workspace.element_size())
        lib.afv_search_by_projection_workspace_bytes.restype = C.c_size_t
        _check(lib.afv_search_by_projection_ex(self.desc_type, _vp(qdesc), _vp(qxy), _vp(qr), _vp(qmin), _vp(qmax), _vp(qangle), _vp(q_start), P,
                                               int(Q if nq_total is None else nq_total), _vp(kps), _vp(desc), _vp(kpsize), _vp(inf1d), _vp(n), B, cap,
                                               _vp(frame), _vp(occupied), int(bool(claim)), C.c_float(minx), C.c_float(miny), C.c_float(maxx),
                                               C.c_float(maxy), C.c_float(self.th_low if th is None else th), C.c_float(self.nnratio),
                                               int(bool(ratio_same_scale)), C.c_float(size_tolerance), _vp(match_q), _vp(nm), ws_ptr,
                                               C.c_size_t(ws_bytes), _sp(stream)))
        return match_q[:Q], nm[:P]

    # the reference's method names on top of the generic search (arrays after the projection prologue)
    def fuse(self, qdesc, qxy, qr, qmin, qmax, q_start, kps, desc, kpsize, n, frame, bounds, inf1d=None, size_tolerance=1.2, stream=None):
        """Fuse (src/FeatureMatcher.cc:794-942 with inf1d = GetKeyPt1DInf, :944-1064 without): keypoint each map point lands on."""
        return self.search_by_projection_ex(qdesc, qxy, qr, qmin, qmax, q_start, kps, desc, kpsize, n, frame, bounds, inf1d=inf1d, occupied=None,
                                            claim=False, ratio_same_scale=False, size_tolerance=size_tolerance, stream=stream)

    def search_by_projection_reloc(self, qdesc, qxy, qr, qmin, qmax, qangle, q_start, kps, desc, kpsize, n, frame, bounds, occupied=None,
                                   size_tolerance=1.2, stream=None):
        """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, useHigh) (:1406-1506) / (CurrentFrame, LastFrame) (:1291-1402)."""
        return self.search_by_projection_ex(qdesc, qxy, qr, qmin, qmax, q_start, kps, desc, kpsize, n, frame, bounds,
                                            qangle=qangle if self.check_ori else None, occupied=occupied, claim=True, ratio_same_scale=False,
                                            size_tolerance=size_tolerance, stream=stream)

    def search_by_sim3(self, q1, q2, kps, desc, kpsize, n, frame1, frame2, bounds, stream=None):
        """SearchBySim3 (:1066-1287).  q1 / q2 = (qdesc, qxy, qr, qmin, qmax, q_start) of the two directions; query j of problem p is
        keypoint j of frame1[p] (resp. frame2[p]).  Returns (match12 [Q1] i32, nfound [P] i32)."""
        import torch
        lib, _check, _vp, _sp = _afv()
        P = frame1.shape[0]
        B, cap = kps.shape[0], kps.shape[1]
        Q1, Q2 = q1[0].shape[0], q2[0].shape[0]
        m12 = torch.empty(max(Q1, 1), dtype=torch.int32, device=kps.device)
        nf = torch.empty(max(P, 1), dtype=torch.int32, device=kps.device)
        minx, miny, maxx, maxy = bounds
        _check(lib.afv_search_by_sim3(self.desc_type, _vp(q1[0]), _vp(q1[1]), _vp(q1[2]), _vp(q1[3]), _vp(q1[4]), _vp(q1[5]), Q1,
                                      _vp(q2[0]), _vp(q2[1]), _vp(q2[2]), _vp(q2[3]), _vp(q2[4]), _vp(q2[5]), Q2, P, _vp(kps), _vp(desc), _vp(kpsize),
                                      _vp(n), B, cap, _vp(frame1), _vp(frame2), C.c_float(minx), C.c_float(miny), C.c_float(maxx), C.c_float(maxy),
                                      C.c_float(self.th_low), _vp(m12), _vp(nf), _sp(stream)))
        return m12[:Q1], nf[:P]

    def bow_match(self, mode, kps, desc, n, node_id, valid, pair_a, pair_b, F12=None, epipole=None, sigma2=None, stream=None):
        """mode 0 SearchByBoW(KF,F) (:186-283, match[p][i2] = i1), 1 SearchByBoW(KF,KF) (:561-660, match[p][i1] = i2),
        2 SearchForTriangulation (:662-790, match[p][i1] = i2).  node_id [B,cap] i32 (afv_bow_transform's node ids, < 0 = none),
        valid [B,cap] u8 or None.  Returns (match [P,cap] i32, nmatches [P] i32)."""
        import torch
        lib, _check, _vp, _sp = _afv()
        B, cap = kps.shape[0], kps.shape[1]
        P = pair_a.shape[0]
        match = torch.empty((max(P, 1), cap), dtype=torch.int32, device=kps.device)
        nm = torch.empty(max(P, 1), dtype=torch.int32, device=kps.device)
        _check(lib.afv_bow_match(int(mode), self.desc_type, _vp(kps), _vp(desc), _vp(n), B, cap, _vp(node_id), _vp(valid), _vp(pair_a), _vp(pair_b), P,
                                 C.c_float(self.th_low), C.c_float(self.nnratio), int(self.check_ori), _vp(F12), _vp(epipole), _vp(sigma2),
                                 _vp(match), _vp(nm), _sp(stream)))
        return match[:P], nm[:P]

    def match_window_pairs(self, kps, desc, kpsize, n, cell_start, cell_items, pair_a, pair_b, bounds, radius=15.0, qxy=None, qr=None,
                           qmin=None, qmax=None, out=None, stream=None):
        """Windowed best / second search for P frame pairs (see afv_match_window_pairs).  Returns (best, bestd, secondd) [P,cap]."""
        import torch
        lib, _check, _vp, _sp = _afv()
        B, cap = kps.shape[0], kps.shape[1]
        P = pair_a.shape[0]
        if out is None:
            out = (torch.empty((P, cap), dtype=torch.int32, device=kps.device), torch.empty((P, cap), dtype=torch.float32, device=kps.device),
                   torch.empty((P, cap), dtype=torch.float32, device=kps.device))
        minx, miny, maxx, maxy = bounds
        _check(lib.afv_match_window_pairs(self.desc_type, _vp(kps), _vp(desc), _vp(kpsize), _vp(n), B, cap, _vp(cell_start), _vp(cell_items),
                                          _vp(pair_a), _vp(pair_b), P, _vp(qxy), _vp(qr), C.c_float(radius), _vp(qmin), _vp(qmax),
                                          C.c_float(minx), C.c_float(miny), C.c_float(maxx), C.c_float(maxy), _vp(out[0]), _vp(out[1]), _vp(out[2]),
                                          _sp(stream)))
        return out

    def match_bruteforce_pairs(self, desc, n, pair_a, pair_b, out=None, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        B, cap = desc.shape[0], desc.shape[1]
        P = pair_a.shape[0]
        if out is None:
            out = (torch.empty((P, cap), dtype=torch.int32, device=desc.device), torch.empty((P, cap), dtype=torch.float32, device=desc.device),
                   torch.empty((P, cap), dtype=torch.float32, device=desc.device))
        _check(lib.afv_match_bruteforce_pairs(self.desc_type, _vp(desc), _vp(n), B, cap, _vp(pair_a), _vp(pair_b), P, _vp(out[0]), _vp(out[1]),
                                              _vp(out[2]), _sp(stream)))
        return out

    def match_bruteforce(self, q, t, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        nq, nt = q.shape[0], t.shape[0]
        best = torch.empty(nq, dtype=torch.int32, device=q.device)
        bd = torch.empty(nq, dtype=torch.float32, device=q.device); sd = torch.empty_like(bd)
        _check(lib.afv_match_bruteforce(self.desc_type, _vp(q), nq, _vp(t), nt, _vp(best), _vp(bd), _vp(sd), _sp(stream)))
        return best, bd, sd

    # -- SearchByBoW(KF, F) (src/FeatureMatcher.cc:186-283) ---------------------------------------------
    def search_by_bow(self, dkf, kkf, kf_segs, df, kf_f, f_segs, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        nf = df.shape[0]
        match_f = torch.empty(max(nf, 1), dtype=torch.int32, device=df.device)
        nm = torch.zeros(1, dtype=torch.int32, device=df.device)
        _check(lib.afv_search_by_bow(self.desc_type, _vp(dkf), _vp(kkf), _vp(kf_segs[0]), _vp(kf_segs[1]), _vp(kf_segs[2]),
                                     kf_segs[0].shape[0], _vp(df), _vp(kf_f), nf, _vp(f_segs[0]), _vp(f_segs[1]),
                                     _vp(f_segs[2]), f_segs[0].shape[0], C.c_float(self.th_low), C.c_float(self.nnratio),
                                     int(self.check_ori), _vp(match_f), _vp(nm), _sp(stream)))
        return match_f[:nf], nm

    # -- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:279-348), batched over map points ----------------
    @staticmethod
    def distinctive_descriptors(desc_type, desc, obs, seg_start, max_seg, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        M = seg_start.shape[0] - 1
        best = torch.empty(max(M, 1), dtype=torch.int32, device=desc.device)
        _check(lib.afv_distinctive_descriptors(int(desc_type), _vp(desc), _vp(obs), _vp(seg_start), M, int(max_seg), _vp(best), _sp(stream)))
        return best[:M]

    # -- static DescriptorDistance (src/FeatureMatcher.cc:1508-1531) -------------------------------------
    @staticmethod
    def descriptor_distance(desc_type, a, b, stream=None):
        import torch
        lib, _check, _vp, _sp = _afv()
        n = a.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=a.device)
        _check(lib.afv_descriptor_distance(int(desc_type), _vp(a), _vp(b), n, _vp(out), _sp(stream)))
        return out


class Vocabulary:
    """Mirror of the reference's Vocabulary::transform (src/Vocabulary.cpp:156-207) on a flat k-ary tree held on the
    device (DBoW2 TemplatedVocabulary semantics, levelsup = 4).  `tree`: dict of numpy arrays child_off [N+1], child_ids,
    node_desc [N,D], node_word [N] (-1 for inner nodes), node_weight [N] float64, and depth L."""

    def __init__(self, desc_type, tree, device=0):
        import torch
        dev = torch.device("cuda", device)
        self.desc_type = int(desc_type)
        self.L = int(tree["L"])
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
        self.child_off = t(tree["child_off"], np.int32); self.child_ids = t(tree["child_ids"], np.int32)
        nd = np.ascontiguousarray(tree["node_desc"])
        self.node_desc = torch.from_numpy(nd.view(np.uint8).reshape(nd.shape[0], -1)).to(dev)
        self.node_word = t(tree["node_word"], np.int32); self.node_weight = t(tree["node_weight"], np.float64)
        self.n_nodes = int(self.node_word.shape[0])

    def transform(self, desc, levelsup=4, stream=None):
        """desc: uint8 cuda tensor [n, D bytes]. Returns (word_id i32 [n], weight f64 [n], node_id i32 [n])."""
        import torch
        lib, _check, _vp, _sp = _afv()
        n = desc.shape[0]
        wid = torch.empty(max(n, 1), dtype=torch.int32, device=desc.device)
        w = torch.empty(max(n, 1), dtype=torch.float64, device=desc.device)
        nid = torch.empty(max(n, 1), dtype=torch.int32, device=desc.device)
        _check(lib.afv_bow_transform(self.desc_type, _vp(desc), n, _vp(self.child_off), _vp(self.child_ids), _vp(self.node_desc),
                                     _vp(self.node_word), _vp(self.node_weight), self.n_nodes, self.L, int(levelsup),
                                     _vp(wid), _vp(w), _vp(nid), _sp(stream)))
        return wid[:n], w[:n], nid[:n]

    @staticmethod
    def feature_vector_segments(node_id):
        """FeatureVector (sorted node ids + CSR of ascending feature indices) from per-feature node ids (numpy)."""
        node_id = np.asarray(node_id)
        order = np.argsort(node_id, kind="stable")
        ids, starts = np.unique(node_id[order], return_index=True)
        return ids.astype(np.int32), np.append(starts, len(node_id)).astype(np.int32), order.astype(np.int32)
