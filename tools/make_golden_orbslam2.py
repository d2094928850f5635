#!/usr/bin/env python3
"""Golden vectors for the vanilla ORB-SLAM2 extractor, produced by the REFERENCE'S OWN CODE + the cv2 4.13.0 binary.

oracle/_ref/libafv_ref.so holds FeatureExtractor::operator()(..., vanillaOrbslam), ComputePyramid, ComputeKeyPointsOctTree,
DistributeOctTree, IC_Angle, computeOrbDescriptor, computeSize / computeSigma cut from /root/reference (src/ORBextractor.cc:79-177,
:460-676) and compiled unmodified; the OpenCV functions they call are plugged in here as callbacks that run the real
cv2.FastFeatureDetector / cv2.resize(INTER_LINEAR) / cv2.GaussianBlur / cv2.fastAtan2.  Runs only in the build container
(needs cv2 and /root/reference); writes tests/golden/orbslam2_*.npz, which tests/test_oracle_orbslam2.py compares with the oracle.
"""
import ctypes as C
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po

cv2.setNumThreads(1)
OUT = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libafv_ref.so")

FAST_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int)
RESIZE_CB = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int)
BLUR_CB = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int)
ATAN2_CB = C.CFUNCTYPE(C.c_float, C.c_float, C.c_float)


def _view(ptr, rows, cols, step):
    buf = np.ctypeslib.as_array(ptr, shape=(rows * step,))
    return np.lib.stride_tricks.as_strided(buf, shape=(rows, cols), strides=(step, 1))


_detectors = {}


def _fast(ptr, rows, cols, step, th, out, cap):
    img = np.ascontiguousarray(_view(ptr, rows, cols, step))
    det = _detectors.setdefault(th, cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                                   type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16))
    kps = det.detect(img)
    assert len(kps) <= cap
    for i, k in enumerate(kps):
        out[3 * i] = k.pt[0]; out[3 * i + 1] = k.pt[1]; out[3 * i + 2] = k.response
    return len(kps)


def _resize(sp, sr, sc, ss, dp, dr, dc, ds):
    src = np.ascontiguousarray(_view(sp, sr, sc, ss))
    _view(dp, dr, dc, ds)[:, :] = cv2.resize(src, (dc, dr), interpolation=cv2.INTER_LINEAR)


def _blur(p, r, c, s):
    v = _view(p, r, c, s)
    v[:, :] = cv2.GaussianBlur(np.ascontiguousarray(v), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)


def _atan2(y, x):
    return float(cv2.fastAtan2(y, x))


_CBS = (FAST_CB(_fast), RESIZE_CB(_resize), BLUR_CB(_blur), ATAN2_CB(_atan2))


def reference_real_parts(gray, nfeatures, nlevels=8, scale_factor=1.2, ini_th=20, min_th=7):
    """The reference's vanilla operator() (compiled from /root/reference) on top of the real cv2 functions."""
    lib = C.CDLL(REF_SO)
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, po.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); ksz = np.zeros(cap, np.float32)
    n = lib.ref_orbslam2_extract(gray.ctypes.data_as(C.c_void_p), w, h, nfeatures, nlevels, C.c_float(scale_factor), ini_th, min_th,
                                 _CBS[0], _CBS[1], _CBS[2], _CBS[3], kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p),
                                 ksz.ctypes.data_as(C.c_void_p), cap)
    assert 0 <= n <= cap, n
    return kps[:n].copy(), desc[:n].copy(), ksz[:n].copy()


def main():
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "anyfeature-vslam_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    cases = []
    for (w, h, s, nf) in ((640, 480, 0, 1000), (640, 480, 5, 2000), (1280, 720, 1, 2000), (752, 480, 2, 1000)):
        cases.append(("synth_%dx%d_s%d" % (w, h, s), synth.stream_frames(w, h, s, 1)[0][0], nf, dict(w=w, h=h, stream=s)))
    toy = np.load(os.path.join(OUT, "toy0.npz"))["gray"]
    cases.append(("toy0", toy, 1000, {}))
    low = (synth.stream_frames(640, 480, 7, 1)[0][0] // 6 + 100).astype(np.uint8)         # low contrast: the minThFAST = 7 fallback carries the frame
    cases.append(("lowcontrast", low, 1000, dict(w=640, h=480, stream=7, lowcontrast=1)))
    for name, gray, nf, meta in cases:
        k, d, sz = reference_real_parts(gray, nf)
        np.savez_compressed(os.path.join(OUT, "orbslam2_%s.npz" % name), kps=k, desc=d, size=sz, nfeatures=nf,
                            **({"gray": gray} if name == "toy0_unused" else {}), **meta)
        print(name, len(k), "keypoints; per level", np.bincount(k["octave"], minlength=8))


if __name__ == "__main__":
    main()
