"""Family-check fixture for the sift128 oracle (run in the build container, where cv2 4.13.0 is importable).

The reference's SIFT is SiftGPU, which cannot run here, so there is no golden vector for it (parity unpinned).  What
CAN be checked is that the oracle is a DoG-SIFT detector: cv2.SIFT (an independent implementation of the same published
algorithm) must find the same scale-space extrema on octaves >= 0.  This script stores cv2's keypoints for a seeded
synthetic frame; tests/test_oracle_sift.py measures the oracle's recall against them.
"""
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("afv_synth", os.path.join(ROOT, "anyfeature-vslam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

frames, _ = synth.stream_frames(640, 480, 0, 1)
sift = cv2.SIFT_create(nfeatures=0, nOctaveLayers=3, contrastThreshold=0.04, edgeThreshold=10, sigma=1.6)
k = sift.detect(frames[0], None)
oc = np.array([(p.octave & 255) if (p.octave & 255) < 128 else (p.octave & 255) - 256 for p in k], np.int32)
xys = np.array([[p.pt[0], p.pt[1], p.size] for p in k], np.float32)
out = os.path.join(ROOT, "tests", "golden", "sift_cv2_synth_640x480_s0_t0.npz")
np.savez_compressed(out, xys=xys, octave=oc, cv2_version=cv2.__version__)
print(out, len(k), np.bincount(oc + 1))

# Second fixture: the same detector with enable_precise_upscale=True.  cv2's default 2x upsampling (INTER_LINEAR, no half-pixel
# compensation) shifts every keypoint by +0.25 px in x and y; the precise variant removes that bias, and only against it can
# positions be compared below half a pixel.  Stored with layer, angle and response so that the test can classify the residual.
sift = cv2.SIFT_create(nfeatures=0, nOctaveLayers=3, contrastThreshold=0.04, edgeThreshold=10, sigma=1.6, enable_precise_upscale=True)
k, dsc = sift.detectAndCompute(frames[0], None)
assert (dsc == np.round(dsc)).all() and dsc.min() >= 0 and dsc.max() <= 255      # cv2 quantises to integers in [0, 255]
oc = np.array([(p.octave & 255) if (p.octave & 255) < 128 else (p.octave & 255) - 256 for p in k], np.int32)
layer = np.array([(p.octave >> 8) & 255 for p in k], np.int32)
xys = np.array([[p.pt[0], p.pt[1], p.size] for p in k], np.float32)
out = os.path.join(ROOT, "tests", "golden", "sift_cv2_precise_synth_640x480_s0_t0.npz")
np.savez_compressed(out, xys=xys, octave=oc, layer=layer, angle=np.array([p.angle for p in k], np.float32),
                    response=np.array([p.response for p in k], np.float32), desc=dsc.astype(np.uint8), cv2_version=cv2.__version__)
print(out, len(k), np.bincount(oc + 1))
