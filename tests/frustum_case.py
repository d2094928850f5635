"""Shared generator of isInFrustum test cases: a camera pose and a cloud of map points around its viewing cone, with every
rejection reason of src/Frame.cc:276-331 represented (behind the camera, outside the image, outside the distance range, viewing
angle) and margins so that no decision sits on a float rounding boundary."""
import numpy as np


def make_case(seed, M=4000):
    rng = np.random.default_rng(seed)
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(-0.6, 0.6)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = rng.uniform(-1, 1, 3)
    R32 = R.astype(np.float32); t32 = t.astype(np.float32)
    twc = (-(R32.T.astype(np.float64) @ t32.astype(np.float64))).astype(np.float32)
    pose = np.zeros(16, np.float32); pose[:9] = R32.reshape(-1); pose[9:12] = t32; pose[12:15] = twc
    cam = np.array([517.3, 516.5, 318.6, 255.3, 40.0], np.float32)
    bounds = np.array([0.0, 640.0, 0.0, 480.0], np.float32)
    # points in camera coordinates, then to the world
    z = rng.uniform(-2.0, 12.0, M); z[np.abs(z) < 0.05] = 0.5
    x = rng.uniform(-1.2, 1.2, M) * np.abs(z); y = rng.uniform(-0.9, 0.9, M) * np.abs(z)
    Pc = np.stack([x, y, z], 1)
    Pw = ((Pc - t) @ R).astype(np.float32)                         # R^T (Pc - t)
    PO = Pw.astype(np.float64) - twc
    dist = np.linalg.norm(PO, axis=1)
    nrm = PO / dist[:, None] + rng.normal(scale=0.6, size=(M, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    mind = (dist * rng.choice([0.3, 0.8, 1.3], M, p=[0.5, 0.35, 0.15])).astype(np.float32)
    maxd = (dist * rng.choice([3.0, 1.4, 0.9], M, p=[0.5, 0.35, 0.15])).astype(np.float32)
    rsz = rng.uniform(1.0, 4.0, M).astype(np.float32); rsg = rng.uniform(0.5, 2.0, M).astype(np.float32)
    rds = rng.uniform(0.5, 8.0, M).astype(np.float32)
    return dict(Pw=Pw, normal=nrm.astype(np.float32), min_dist=mind, max_dist=maxd, ref_size=rsz, ref_sigma=rsg, ref_dist=rds,
                pose16=pose, cam5=cam, bounds4=bounds)
