import sys, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g._load_pkg()
from oracle import pyoracle as po
frames, _ = pkg.synth.stream_frames(640, 480, 0, 1)
ex = pkg.FeatureExtractor("orb32", nfeatures=1000, max_batch=1, max_w=640, max_h=480)
try:
    ex.extract_batch(frames)
except Exception as e:
    print("extract raised", e)
L = frames[0]
xs, ys, sc = po.fast(L, 20)
ref = {(x, y): s for x, y, s in zip(xs.tolist(), ys.tolist(), sc.tolist())}
c = ex.debug_read(2, 0, 0).view(np.uint32)
got = {(int(v & 0xfff), int((v >> 12) & 0xfff)): int(v >> 24) for v in c}
print("ref", len(ref), "gpu", len(got), "same pos", len(set(ref) & set(got)))
both = sorted(set(ref) & set(got))
diffs = [(p, ref[p], got[p]) for p in both if ref[p] != got[p]]
print("same pos, score differs:", len(diffs), diffs[:20])
only_ref = sorted(set(ref) - set(got))[:20]; only_gpu = sorted(set(got) - set(ref))[:20]
print("only ref", [(p, ref[p]) for p in only_ref])
print("only gpu", [(p, got[p]) for p in only_gpu])
