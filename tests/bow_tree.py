"""Synthetic DBoW2-style vocabulary tree for tests (the reference's vocabularies are HF downloads, not in the repo)."""
import numpy as np


def make_tree(rng, k=10, L=3, D=32, float_desc=False, flip=0.12):
    """k-ary tree of depth L in BFS order; children descriptors = parent's with random perturbation so descent is
    meaningful; leaves get consecutive word ids and idf-like weights. Returns the dict the oracle / Vocabulary take."""
    descs = [np.zeros(D, np.uint8) if not float_desc else np.zeros(128, np.float32)]
    child_off, child_ids, word, weight = [0], [], [], []
    level_nodes = [0]
    nxt = 1
    children_of = {}
    for lvl in range(L):
        new_level = []
        for p in level_nodes:
            ids = []
            for _ in range(k):
                if float_desc:
                    d = descs[p] + rng.normal(scale=0.5 / (lvl + 1), size=128).astype(np.float32)
                else:
                    base = descs[p] if lvl > 0 else rng.integers(0, 256, D, dtype=np.uint8)
                    mask = (rng.random((D, 8)) < flip).astype(np.uint8)
                    d = base ^ np.packbits(mask, axis=1).ravel()
                descs.append(d); ids.append(nxt); new_level.append(nxt); nxt += 1
            children_of[p] = ids
        level_nodes = new_level
    n = nxt
    off = 0
    wcount = 0
    for i in range(n):
        ids = children_of.get(i, [])
        child_ids += ids
        off += len(ids); child_off.append(off)
        if ids:
            word.append(-1); weight.append(0.0)
        else:
            word.append(wcount); weight.append(float(rng.uniform(0.5, 3.0))); wcount += 1
    # one duplicated child descriptor to exercise the first-minimum-wins rule
    descs[children_of[0][3]] = descs[children_of[0][1]].copy()
    return dict(child_off=np.array(child_off, np.int32), child_ids=np.array(child_ids, np.int32), node_desc=np.stack(descs),
                node_word=np.array(word, np.int32), node_weight=np.array(weight, np.float64), L=L)
