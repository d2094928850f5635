"""Family-check fixture for the brisk48 oracle (run in the build container, where cv2 4.13.0 is importable).

The reference's BRISK is ETH brisk v2 (not vendored, parity unpinned).  cv2.BRISK is the BRISK authors' own implementation
of the published algorithm as contributed to OpenCV: same AGAST scale-space detector (BriskFeatureDetector(34, 4 octaves) ==
BRISK_create(thresh=34, octaves=4)), same 60-point pattern, long-pair orientation and short-pair tests; it emits the paper's
64-byte (512 short pairs) descriptor, NOT the 48-byte briskV2 table.  This script stores, per test frame: cv2's keypoints
(after its border filter, with orientation) and 64-byte descriptors, the cv2 INTER_AREA pyramid layers (small ones verbatim,
CRC32 of all), and the AGAST 9-16 detections of two layers.  tests/test_oracle_brisk.py holds the oracle against them.
"""
import importlib.util
import os
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("afv_synth", os.path.join(ROOT, "anyfeature-vslam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)


def pyramid(img, octaves=4):
    lay = [img, cv2.resize(img, (2 * (img.shape[1] // 3), 2 * (img.shape[0] // 3)), interpolation=cv2.INTER_AREA)]
    for i in range(2, 2 * octaves):
        p = lay[i - 2]
        lay.append(cv2.resize(p, (p.shape[1] // 2, p.shape[0] // 2), interpolation=cv2.INTER_AREA))
    return lay


def make(img, name):
    b = cv2.BRISK_create(thresh=34, octaves=4)
    k, d = b.detectAndCompute(img, None)
    kp = np.array([[p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave, p.class_id] for p in k], np.float32)
    lay = pyramid(img)
    crc = np.array([zlib.crc32(l.tobytes()) for l in lay], np.uint32)
    ag = {}
    for i in (1, 4):
        det = cv2.AgastFeatureDetector_create(threshold=34, nonmaxSuppression=False, type=cv2.AgastFeatureDetector_OAST_9_16)
        pts = det.detect(lay[i])
        ag["agast%d" % i] = np.array([[p.pt[0], p.pt[1], p.response] for p in pts], np.int32)
    out = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(out, kp=kp, desc=d, layer_crc=crc, layer5=lay[5], layer7=lay[7], cv2_version=cv2.__version__, **ag)
    print(out, kp.shape, d.shape)


if __name__ == "__main__":
    make(synth.stream_frames(640, 480, 0, 1)[0][0], "brisk_cv2_synth_640x480_s0_t0.npz")
    make(synth.stream_frames(640, 480, 2, 1)[0][0], "brisk_cv2_synth_640x480_s2_t0.npz")
    toy = np.load(os.path.join(ROOT, "tests", "golden", "toy0.npz"))["gray"]
    make(np.ascontiguousarray(toy), "brisk_cv2_toy0.npz")
