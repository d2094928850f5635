"""CPU tests of the akaze61 ORACLE (oracle/afv_oracle_akaze.c).  PARITY UNPINNED vs libAKAZE (not vendored by the
reference); cv2.AKAZE (OpenCV's port of libAKAZE) stored by tools/make_golden_akaze.py PINS the pipeline: with OpenCV's
variant of the duplicate filter selected the keypoint lists are identical, and MLDB agrees bit for bit where the orientation
searches agree.  Plus the reference-side post-processing."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def frame(synth):
    frames, _ = synth.stream_frames(640, 480, 0, 1)
    return frames[0]


def test_schedule():
    lw = np.zeros(16, np.int32); lh = np.zeros(16, np.int32); oc = np.zeros(16, np.int32); ss = np.zeros(16, np.int32)
    es = np.zeros(16, np.float32); ns = np.zeros(16, np.int32); tau = np.zeros((16, 64), np.float32)
    n = po.lib().orc_akaze_schedule(640, 480, 2, 4, po._p(lw), po._p(lh), po._p(oc), po._p(ss), po._p(es), po._p(ns), po._p(tau))
    assert n == 8
    assert list(lw[:8]) == [640] * 4 + [320] * 4 and list(oc[:8]) == [0] * 4 + [1] * 4
    assert np.allclose(es[:8], 1.6 * 2.0 ** (np.arange(8) / 4.0), rtol=1e-6)
    assert list(ns[:8]) == [0, 3, 3, 4, 4, 5, 6, 7]                      # FED cycle lengths
    for i in range(1, 8):                                                 # a FED cycle integrates exactly the level's time step
        T = 0.5 * (es[i] ** 2 - es[i - 1] ** 2)
        assert abs(tau[i, :ns[i]].sum() - T) < 1e-4 * T


def test_scale_space_structure(frame):
    lt0, kc = po.akaze_scale_space(frame, 0, 0)
    import scipy.ndimage as ndi
    ref = ndi.gaussian_filter(frame.astype(np.float64) / 255.0, 1.6, mode="nearest", truncate=2.5)
    assert np.abs(ref - lt0).max() < 3e-3
    assert 0.0 < kc < 1.0
    lt3, _ = po.akaze_scale_space(frame, 0, 3)
    lt4, _ = po.akaze_scale_space(frame, 0, 4)
    assert lt4.shape == (240, 320)
    # diffusion keeps the mean (up to rounding) and only smooths
    assert abs(float(lt3.mean()) - float(lt0.mean())) < 1e-4
    assert lt3.var() < lt0.var()


def test_pinned_to_cv2_with_opencv_duplicate_filter(frame, golden_dir):
    """With OpenCV's variant of the duplicate filter selected, the oracle reproduces cv2.AKAZE's keypoint list exactly (count, order,
    level, position, size, response): this pins the FED scale space, the Hessian, the maxima, the border rule (10 sqrt 2 * sigma_size)
    and the sub-pixel refinement to the cv2 4.13.0 binary.  MLDB is pinned where the two orientation searches agree, and -- with OpenCV's
    orientation search selected too -- on every keypoint."""
    g = np.load(os.path.join(golden_dir, "akaze_cv2_synth_640x480_s0_t0.npz"))
    kp = g["kp"]
    po.lib().orc_akaze_set_cv2_filter(1)
    try:
        det = po.akaze_detect(frame)
        kps, desc, size, nd = po.akaze61_extract(frame, 20000)
    finally:
        po.lib().orc_akaze_set_cv2_filter(0)
    assert len(det) == len(kp) == 1751
    assert (det[:, 4] == kp[:, 6]).all()                                   # same evolution level, same ORDER
    assert np.abs(det[:, :2] - kp[:, :2]).max() < 5e-4                     # sub-pixel positions
    assert (det[:, 2] == kp[:, 2]).all()                                   # size
    assert (np.abs(det[:, 3] - kp[:, 4]) / kp[:, 4]).max() < 1e-4          # Hessian response
    from scipy.spatial import cKDTree
    d2, i2 = cKDTree(kp[:, :2]).query(np.stack([kps["x"], kps["y"]], 1))
    m = (d2 < 0.05) & (kp[i2, 6] == kps["class_id"])
    assert m.sum() == len(kp)                                               # no octree at this quota: every keypoint is described
    da = np.abs(((np.degrees(kps["angle"][m]) - kp[i2[m], 3]) + 180) % 360 - 180)
    hd = np.unpackbits(desc[m] ^ g["desc"][i2[m]], axis=1).sum(1)
    close = da < 1e-3                                                       # the two orientation searches agree
    assert close.mean() > 0.4 and (hd[close] == 0).mean() > 0.98 and hd[close].max() <= 2
    assert (hd[da < 1e-2] <= 4).all()
    assert (da > 1.0).mean() < 0.15                                         # OpenCV's 42-slice search vs libAKAZE's exact-angle windows
    # ... and with OpenCV's orientation search selected as well (mode bits 0 + 1) the sampling and MLDB code is pinned on EVERY keypoint:
    po.lib().orc_akaze_set_cv2_filter(3)
    try:
        kps, desc, size, nd = po.akaze61_extract(frame, 20000)
    finally:
        po.lib().orc_akaze_set_cv2_filter(0)
    d2, i2 = cKDTree(kp[:, :2]).query(np.stack([kps["x"], kps["y"]], 1))
    m = (d2 < 0.05) & (kp[i2, 6] == kps["class_id"])
    assert m.sum() == len(kp)
    da = np.abs(((np.degrees(kps["angle"][m]) - kp[i2[m], 3]) + 180) % 360 - 180)
    hd = np.unpackbits(desc[m] ^ g["desc"][i2[m]], axis=1).sum(1)
    assert (da < 1e-2).mean() > 0.995 and da.max() < 1.0                    # all 1751 orientations (the derivative arrays differ in the last bits)
    assert (hd == 0).mean() > 0.94 and (hd <= 2).mean() > 0.995             # 95.7 % of the 486-bit descriptors identical, 99.8 % within 2 bits


def test_family_check_against_cv2(frame, golden_dir):
    g = np.load(os.path.join(golden_dir, "akaze_cv2_synth_640x480_s0_t0.npz"))
    det = po.akaze_detect(frame)
    from scipy.spatial import cKDTree
    tree = cKDTree(det[:, :2])
    d, i = tree.query(g["kp"][:, :2])
    same = (d < 0.05) & (det[i, 4] == g["kp"][:, 6])
    assert same.mean() > 0.995                                 # libAKAZE's filter keeps (at least) cv2's keypoints: same position, same level
    assert np.allclose(det[i[same], 2], g["kp"][same, 2], rtol=1e-6)                    # size
    rel = np.abs(det[i[same], 3] - g["kp"][same, 4]) / g["kp"][same, 4]
    assert np.median(rel) < 1e-5 and np.percentile(rel, 99) < 1e-3                     # Hessian response
    # descriptors at the keypoints both detectors keep (no octree: quota large)
    kps, desc, size, nd = po.akaze61_extract(frame, 20000)
    tr2 = cKDTree(g["kp"][:, :2])
    d2, i2 = tr2.query(np.stack([kps["x"], kps["y"]], 1))
    m = (d2 < 0.05) & (g["kp"][i2, 6] == kps["class_id"])
    assert m.sum() > 1000
    da = np.abs(((np.degrees(kps["angle"][m]) - g["kp"][i2[m], 3]) + 180) % 360 - 180)
    assert np.median(da) < 0.05
    hd = np.unpackbits(desc[m] ^ g["desc"][i2[m]], axis=1).sum(1)
    assert np.median(hd) <= 2 and (hd == 0).mean() > 0.5        # MLDB-486 bits


def test_extract_postprocessing(frame):
    kps, desc, size, nd = po.akaze61_extract(frame, 1000)
    det = po.akaze_detect(frame)
    assert nd == len(det)
    q = po.features_per_level(1000, 8, 1.1892)
    cl = kps["class_id"]
    assert (np.diff(cl) >= 0).all()                             # levels merged in ascending order (std::map)
    for l in range(8):
        assert (cl == l).sum() <= q[l] + 3
        assert (cl == l).sum() <= (det[:, 4] == l).sum()
    assert (kps["octave"] == cl // 4).all()
    assert np.allclose(size, np.float32(1.1892) ** cl, rtol=1e-5)
    assert desc.shape[1] == 61 and (desc[:, 60] >> 6 == 0).all()     # 486 bits
    # every kept keypoint is one of the detected ones
    s = {(float(a), float(b)) for a, b in det[:, :2]}
    assert all((float(a), float(b)) in s for a, b in zip(kps["x"], kps["y"]))
    again = po.akaze61_extract(frame, 1000)
    assert (again[1] == desc).all()
