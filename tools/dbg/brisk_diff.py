"""Debug helper: CUDA brisk48 detect list vs the oracle's, first differing rows."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import pyoracle as po
pkg = ge._load_pkg()
frames, _ = pkg.synth.stream_frames(640, 480, 0, 2)
ex = pkg.FeatureExtractor("brisk48", nfeatures=1000, max_batch=2, max_w=640, max_h=480)
kps, desc, size, n = ex.extract_batch(frames)
for f in range(2):
    det = po.brisk_detect(frames[f], 34, 4, po.BRISK_DENSE)
    lst = ex.debug_read(32, f, 0, nbytes_cap=20 * (len(det) + 4096)).view(np.float32).reshape(-1, 5)
    print("frame", f, "oracle", len(det), "gpu", len(lst))
    for l in range(8):
        a = det[det[:, 4] == l]; b = lst[lst[:, 4] == l]
        if len(a) != len(b) or (a != b).any():
            print(" layer", l, len(a), len(b))
            sa = {tuple(r) for r in a.tolist()}; sb = {tuple(r) for r in b.tolist()}
            oa = sorted(sa - sb)[:6]; ob = sorted(sb - sa)[:6]
            print("  oracle only", len(sa - sb), oa)
            print("  gpu only   ", len(sb - sa), ob)
            # order check
            common = [r for r in a.tolist() if tuple(r) in sb]
            commonb = [r for r in b.tolist() if tuple(r) in sa]
            print("  common same order:", common == commonb)
