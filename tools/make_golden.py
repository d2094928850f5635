#!/usr/bin/env python3
"""Generate tests/golden/*.npz -- golden vectors for the orb32 path, produced by the PINNED reference stack.

Runs only in the build container (needs cv2 4.13.0 and /root/reference for the toy frames).  For each case
the reference's call pattern (src/Feature_orb32.cpp:11-53) is executed with cv2 for the OpenCV stages:
  cv2.ORB_create(); setMaxFeatures(10*n); setEdgeThreshold(0); setFastThreshold(20); setNLevels(8)
  detect -> bucket by octave -> DistributeOctTree (oracle C restatement; the only in-repo stage, no third-party
  numerics) -> one cv2 ORB.compute call PER LEVEL -> merge ascending.
The per-level candidate order handed to the octree is canonical raster order (y, x) -- cv::ORB's own order
after retainBest is std::nth_element's and only matters for exact float response ties inside one octree node.
Also stores cv2's raw detect() output for one frame (stage-level golden) and, for the toy frames, the gray
image itself (COLOR_RGB2GRAY applied to imread's BGR: quirk a1 in SURVEY.md 8(a)).
"""
import os, sys, glob, importlib.util
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "anyfeature-vslam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
cv2.setNumThreads(1)
OUT = os.path.join(ROOT, "tests", "golden")


def mk_orb(nfeatures):
    orb = cv2.ORB_create()
    orb.setMaxFeatures(nfeatures * 10); orb.setEdgeThreshold(0)
    orb.setFastThreshold(20); orb.setNLevels(8)
    return orb


def kp_array(kps):
    a = np.zeros(len(kps), po.KP_DTYPE)
    for i, k in enumerate(kps):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return a


def reference_flow(gray, nfeatures):
    h, w = gray.shape
    orb = mk_orb(nfeatures)
    det = orb.detect(gray)
    q = po.features_per_level(nfeatures)
    ls = po.level_geometry(w, h)[2]
    out_k, out_d = [], []
    for l in range(8):
        kl = [k for k in det if k.octave == l]
        if not kl:
            continue
        inv = np.float32(1) / ls[l]
        key = [(int(np.rint(np.float32(k.pt[1]) * inv)), int(np.rint(np.float32(k.pt[0]) * inv))) for k in kl]
        kl = [kl[i] for i in sorted(range(len(kl)), key=lambda i: key[i])]
        keep = po.octree([k.pt[0] for k in kl], [k.pt[1] for k in kl], [k.response for k in kl], w, h, q[l])
        sel, d = orb.compute(gray, [kl[i] for i in keep])
        assert len(sel) == len(keep)
        out_k += list(sel); out_d.append(d)
    return kp_array(out_k), np.vstack(out_d), kp_array(det)


def main():
    os.makedirs(OUT, exist_ok=True)
    toy = sorted(glob.glob("/root/reference/docs/toy_sequence/rgb/*.png"))
    for i in (0, 2):
        gray = cv2.cvtColor(cv2.imread(toy[i], cv2.IMREAD_UNCHANGED), cv2.COLOR_RGB2GRAY)
        k1, d1, det1 = reference_flow(gray, 1000)
        k2, d2, _ = reference_flow(gray, 2000)
        extra = {"det1000": det1} if i == 0 else {}
        np.savez_compressed(os.path.join(OUT, "toy%d.npz" % i), gray=gray, kps1000=k1, desc1000=d1,
                            kps2000=k2, desc2000=d2, **extra)
        print("toy", i, len(k1), len(k2))
    for (w, h, n, stream, t) in ((640, 480, 1000, 0, 0), (640, 480, 1000, 0, 1), (1280, 720, 2000, 1, 0)):
        fr, _ = synth.stream_frames(w, h, stream, t + 1)
        k, d, _ = reference_flow(fr[t], n)
        np.savez_compressed(os.path.join(OUT, "synth_%dx%d_s%d_t%d.npz" % (w, h, stream, t)),
                            kps=k, desc=d, nfeatures=n, w=w, h=h, stream=stream, t=t,
                            crc=np.uint32(int(fr[t].astype(np.uint64).sum()) & 0xffffffff))
        print("synth", w, h, len(k))


if __name__ == "__main__":
    main()
