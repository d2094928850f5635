"""Family-check fixture for the akaze61 oracle (run in the build container, where cv2 4.13.0 is importable).

The reference's AKAZE is libAKAZE (not vendored, parity unpinned).  cv2.AKAZE is OpenCV's port of the same library by the
same author: with the reference's options (threshold 5e-4, 2 octaves x 4 sublevels, PM_G2, MLDB-486) it must find the
same Hessian maxima with the same responses and nearly the same descriptors.  This script stores cv2's keypoints and
descriptors for a seeded synthetic frame; tests/test_oracle_akaze.py measures the oracle against them.
"""
import importlib.util
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("afv_synth", os.path.join(ROOT, "anyfeature-vslam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

frames, _ = synth.stream_frames(640, 480, 0, 1)
ak = cv2.AKAZE_create(descriptor_type=cv2.AKAZE_DESCRIPTOR_MLDB, descriptor_size=0, descriptor_channels=3, threshold=5e-4,
                      nOctaves=2, nOctaveLayers=4, diffusivity=cv2.KAZE_DIFF_PM_G2)
k, d = ak.detectAndCompute(frames[0], None)
kp = np.array([[p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave, p.class_id] for p in k], np.float32)
out = os.path.join(ROOT, "tests", "golden", "akaze_cv2_synth_640x480_s0_t0.npz")
np.savez_compressed(out, kp=kp, desc=d, cv2_version=cv2.__version__)
print(out, kp.shape, d.shape)
