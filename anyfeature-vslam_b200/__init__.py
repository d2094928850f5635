"""anyfeature-vslam_b200 -- B200-native feature front end (orb32 / sift128 / akaze61 extract + FeatureMatcher kernels).

Python host side above the C ABI (include/afv.h), used by tests/ and bench.py.  It mirrors the reference's
FeatureExtractor / FeatureMatcher call surface on arrays (numpy for host buffers, torch tensors for device
buffers: torch is only plumbing for device memory, streams and torch.distributed).  There is NO CPU fallback:
importing works without a GPU (so the symbol/ABI tests can run), but every compute entry point raises
AfvError when the CUDA library or device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafv_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
FEAT_ORB32, FEAT_AKAZE61, FEAT_BRISK48, FEAT_SIFT128 = 0, 1, 2, 5       # reference include/Types.h:35-45
FEAT_ORB32_VANILLA = 100                                                 # extractor built with VANILLA_ORB_SLAM2 (include/Definitions.h:8)
DESC_BYTES = {0: 32, 1: 61, 2: 48, 5: 512, 100: 32}

# settings/<feat>_settings.yaml of the reference (numOctaves, scaleFactor, detectionTh, matchingTh)
FEATURE_SETTINGS = {
    "orb32": dict(feature_id=0, n_octaves=8, scale_factor=1.2, detect_th=20.0, matching_th=75.0),
    "akaze61": dict(feature_id=1, n_octaves=8, scale_factor=1.1892, detect_th=0.0005, matching_th=128.0),
    "brisk48": dict(feature_id=2, n_octaves=8, scale_factor=1.5, detect_th=34.0, matching_th=120.0),
    "sift128": dict(feature_id=5, n_octaves=8, scale_factor=2.0, detect_th=10.0, matching_th=0.5),
    # vanilla ORB-SLAM2 build: the settings constructor pins scaleFactor 1.2, 8 octaves, iniThFAST 20 (src/FeatureExtractor.cpp:40-46)
    "orbslam2": dict(feature_id=100, n_octaves=8, scale_factor=1.2, detect_th=20.0, matching_th=75.0),
}


class AfvError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; raises loudly if it was not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfvError("CUDA extension %s is missing: run `python __graft_entry__.py` / build.py first; "
                           "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.afv_last_error.restype = C.c_char_p
        _lib.afv_version.restype = C.c_char_p
        _lib.afv_kernel_launches.restype = C.c_longlong
    return _lib


def _check(rc):
    if rc != 0:
        raise AfvError("afv error %d: %s" % (rc, lib().afv_last_error().decode()))


def kernel_launches():
    return int(lib().afv_kernel_launches())


def _vp(x):
    """void* of a numpy array or a torch tensor (device or host)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    return C.c_void_p(x.data_ptr())


def _stream_ptr(stream):
    """cudaStream_t for the C ABI.  torch's default stream has handle 0, which the extractor entry points read as "use the
    extractor's own non-blocking stream" -- unordered against torch work on the default stream that produced the inputs.  The
    legacy default stream is therefore passed explicitly as cudaStreamLegacy (0x1), which keeps the call ordered after it."""
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    h = stream.cuda_stream
    return C.c_void_p(h if h else 1)


class FeatureExtractor:
    """Mirror of the reference's FeatureExtractor_<feat> (include/FeatureExtractor.h:68-161).

    __call__(gray) == operator()(Image, keypoints, descriptors, ..., size) for one host frame; extract_batch
    is the batched extension.  Construction == the reference factory (src/Tracking.cc:1505-1553) with the
    values of settings/<feat>_settings.yaml.
    """

    def __init__(self, feature="orb32", nfeatures=1000, device=0, max_batch=1, max_w=640, max_h=480,
                 n_octaves=None, scale_factor=None, detect_th=None):
        s = FEATURE_SETTINGS[feature]
        self.feature = feature
        self.nfeatures = nfeatures
        self.n_octaves = n_octaves or s["n_octaves"]
        self.scale_factor = scale_factor or s["scale_factor"]
        self.detect_th = detect_th if detect_th is not None else s["detect_th"]
        self.desc_bytes = DESC_BYTES[s["feature_id"]]
        self.max_batch = max_batch
        self._h = C.c_void_p()
        _check(lib().afv_extractor_create(C.byref(self._h), s["feature_id"], nfeatures, self.n_octaves,
                                          C.c_float(self.scale_factor), C.c_float(self.detect_th), device,
                                          max_batch, max_w, max_h))
        self.cap = lib().afv_extractor_output_cap(self._h)
        self.device = device

    def close(self):
        if self._h:
            lib().afv_extractor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def levels(self):
        sf = np.zeros(self.n_octaves, np.float32); q = np.zeros(self.n_octaves, np.int32)
        lib().afv_extractor_levels(self._h, _vp(sf), _vp(q))
        return sf, q

    # ---- host buffers (numpy) -------------------------------------------------------------------------
    def extract_batch(self, gray, out=None):
        """gray: uint8 [B,H,W] numpy (C-contiguous).  Returns (kps [B,cap] KP_DTYPE, desc [B,cap,D], size [B,cap], n [B])."""
        gray = np.ascontiguousarray(gray, np.uint8)
        B, h, w = gray.shape
        if out is None:
            out = (np.zeros((B, self.cap), KP_DTYPE), np.zeros((B, self.cap, self.desc_bytes), np.uint8),
                   np.zeros((B, self.cap), np.float32), np.zeros(B, np.int32))
        kps, desc, size, n = out
        _check(lib().afv_extract_batch(self._h, _vp(gray), B, w, h, w, C.c_long(w * h), _vp(kps), _vp(desc), _vp(size),
                                       self.cap, _vp(n)))
        return kps, desc, size, n

    def __call__(self, gray):
        kps, desc, size, n = self.extract_batch(np.asarray(gray)[None])
        m = int(n[0])
        return kps[0, :m], desc[0, :m], size[0, :m]

    # ---- device buffers (torch) -----------------------------------------------------------------------
    def alloc_device_outputs(self, B):
        import torch
        dev = torch.device("cuda", self.device)
        return (torch.zeros((B, self.cap, 7), dtype=torch.float32, device=dev),       # afv_keypoint rows (28 B)
                torch.zeros((B, self.cap, self.desc_bytes), dtype=torch.uint8, device=dev),
                torch.zeros((B, self.cap), dtype=torch.float32, device=dev),
                torch.zeros((B,), dtype=torch.int32, device=dev))

    def extract_batch_device(self, d_gray, out, stream=None):
        """d_gray: uint8 cuda tensor [B,H,W]; out from alloc_device_outputs. Asynchronous on `stream`."""
        B, h, w = d_gray.shape
        kps, desc, size, n = out
        _check(lib().afv_extract_batch_device(self._h, _vp(d_gray), B, w, h, d_gray.stride(1), C.c_long(d_gray.stride(0)),
                                              _vp(kps), _vp(desc), _vp(size), self.cap, _vp(n), _stream_ptr(stream)))
        return out

    def status(self):
        _check(lib().afv_extractor_status(self._h))

    def debug_read(self, what, frame, level, nbytes_cap=1 << 24):
        buf = np.zeros(nbytes_cap, np.uint8)
        n = C.c_long(0)
        _check(lib().afv_debug_read(self._h, what, frame, level, _vp(buf), C.c_long(nbytes_cap), C.byref(n)))
        return buf[:n.value].copy()


def gray_from_color(d_img, rgb=True, out=None, stream=None):
    """Image::GetGrayImage on device: uint8 cuda tensor [B,H,W,C] (C = 3 or 4) -> [B,H,W].  rgb=True reproduces the
    reference (CV_RGB2GRAY applied to imread's BGR data)."""
    import torch
    B, h, w, ch = d_img.shape
    if out is None:
        out = torch.empty((B, h, w), dtype=torch.uint8, device=d_img.device)
    _check(lib().afv_gray_from_color(_vp(d_img), ch, int(bool(rgb)), B, w, h, d_img.stride(1), C.c_long(d_img.stride(0)), _vp(out),
                                     out.stride(1), C.c_long(out.stride(0)), _stream_ptr(stream)))
    return out


def undistort_keypoints(kps, n, K4, dist5, out=None, stream=None):
    """Frame::UndistortKeyPoints (src/Frame.cc:403-433) on device keypoint rows [B,cap,7]; K4 = (fx, fy, cx, cy) and
    dist5 = (k1, k2, p1, p2, k3) are host float32 values.  Returns the undistorted rows (mvKeysUn)."""
    import torch
    B, cap = kps.shape[0], kps.shape[1]
    if out is None:
        out = torch.zeros_like(kps)
    K4 = np.ascontiguousarray(K4, np.float32); dist5 = np.ascontiguousarray(dist5, np.float32)
    _check(lib().afv_undistort_keypoints(_vp(kps), _vp(n), B, cap, _vp(K4), _vp(dist5), _vp(out), _stream_ptr(stream)))
    return out


def is_in_frustum(Pw, normal, min_dist, max_dist, ref_size, ref_sigma, ref_dist, pose16, cam5, bounds4, cos_limit=0.5,
                  radius_factor=1.0, size_tol=1.5, stream=None):
    """Frame::isInFrustum (src/Frame.cc:276-331) for M map points (float32 cuda tensors [M,3] / [M]) against one frame, fused with
    SearchByProjection's window prologue.  pose16 / cam5 / bounds4 are host values.  Returns device tensors
    (in_view [M] uint8, proj [M,3], track [M,3], qr [M], qmin [M], qmax [M])."""
    import torch
    M = Pw.shape[0]
    dev = Pw.device
    iv = torch.zeros(M, dtype=torch.uint8, device=dev)
    proj = torch.zeros((M, 3), dtype=torch.float32, device=dev); track = torch.zeros((M, 3), dtype=torch.float32, device=dev)
    qr = torch.zeros(M, dtype=torch.float32, device=dev); qmin = torch.zeros_like(qr); qmax = torch.zeros_like(qr)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    _check(lib().afv_is_in_frustum(_vp(Pw), _vp(normal), _vp(min_dist), _vp(max_dist), _vp(ref_size), _vp(ref_sigma), _vp(ref_dist), M,
                                   _vp(f32(pose16)), _vp(f32(cam5)), _vp(f32(bounds4)), C.c_float(cos_limit), C.c_float(radius_factor),
                                   C.c_float(size_tol), _vp(iv), _vp(proj), _vp(track), _vp(qr), _vp(qmin), _vp(qmax), _stream_ptr(stream)))
    return iv, proj, track, qr, qmin, qmax


def kps_from_device(t, n):
    """[cap,7] float32 device rows -> numpy structured afv_keypoint array of length n."""
    a = t[:n].contiguous().cpu().numpy()
    return a.view(np.uint8).reshape(n, 28).view(KP_DTYPE).reshape(n)


from . import matcher as matcher  # noqa: E402  (FeatureMatcher mirror)
from . import synth as synth  # noqa: E402  (synthetic frame generator)
from . import sharding as sharding  # noqa: E402  (multi-GPU host logic)
from .matcher import FeatureMatcher, Vocabulary  # noqa: E402,F401
