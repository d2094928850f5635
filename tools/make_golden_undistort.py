"""Golden vectors for Frame::UndistortKeyPoints (reference src/Frame.cc:403-433 -> cv::undistortPoints), generated with
cv2 4.13.0 in the build container: random float pixel positions, three camera models (TUM fr1-like strong radial, mild
radial + tangential, zero distortion)."""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(7)
cams = [
    ((517.3, 516.5, 318.6, 255.3), (0.2624, -0.9531, -0.0054, 0.0026, 1.1633)),
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),
    ((535.4, 539.2, 320.1, 247.6), (0.0, 0.0, 0.0, 0.0, 0.0)),
]
pts = rng.uniform([-20, -20], [760, 500], (3000, 2)).astype(np.float32)
out = {"pts": pts}
for i, (k, d) in enumerate(cams):
    K = np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float32)
    D = np.array(d, np.float32)
    out["K%d" % i] = np.array(k, np.float32); out["D%d" % i] = D
    out["und%d" % i] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
path = os.path.join(ROOT, "tests", "golden", "undistort_cv2.npz")
np.savez_compressed(path, cv2_version=cv2.__version__, **out)
print(path)
