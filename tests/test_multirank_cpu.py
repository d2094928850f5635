"""Host logic of the N>1 path on CPU (gloo, world_size 2): stream sharding and the packed-result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_pkg():
    import __graft_entry__ as g
    return g._load_pkg()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, cap, q, D=32):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_pkg()
    sh = pkg.sharding
    rng = np.random.default_rng(100 + rank)
    n = torch.from_numpy(rng.integers(0, cap, B).astype(np.int32))
    nm = torch.from_numpy(rng.integers(0, 200, B).astype(np.int32))
    m12 = torch.from_numpy(rng.integers(-1, cap, (B, cap)).astype(np.int32))
    kps = torch.from_numpy(rng.normal(size=(B, cap, 7)).astype(np.float32))
    desc = torch.from_numpy(rng.integers(0, 256, (B, cap, D), dtype=np.uint8))
    _, total = sh.pack_layout(B, cap, D)
    pack = torch.empty(total, dtype=torch.uint8)
    sh.pack_results(pack, n, nm, m12, kps, desc)
    gathered = [torch.empty(total, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    sh.gather_to_root(pack, gathered, world, rank)
    if rank == 0:
        ok = True
        for r in range(world):
            u = sh.unpack_results(gathered[r].numpy(), B, cap, D)
            rr = np.random.default_rng(100 + r)
            ok &= (u["n"] == rr.integers(0, cap, B).astype(np.int32)).all()
            ok &= (u["nmatches"] == rr.integers(0, 200, B).astype(np.int32)).all()
            ok &= (u["matches12"] == rr.integers(-1, cap, (B, cap)).astype(np.int32)).all()
            ok &= (u["kps"].view(np.float32).reshape(B, cap, 7) == rr.normal(size=(B, cap, 7)).astype(np.float32)).all()
            ok &= (u["desc"] == rr.integers(0, 256, (B, cap, D), dtype=np.uint8)).all()
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("D", [32, 61, 512])          # orb32, akaze61 (unaligned rows), sift128 (128 floats)
def test_gather_packed_results_world2(D):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, 40, q, D)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_stream_sharding_is_a_partition():
    sh = load_pkg().sharding
    for world in (1, 2, 4, 8):
        parts = [sh.streams_of_rank(8, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(8))
        assert all(len(p) == 8 // world for p in parts)


def test_bench_frames_differ_across_ranks_and_pairs_stay_in_stream():
    import importlib.util
    spec = importlib.util.spec_from_file_location("afv_bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    pkg = load_pkg()
    f0, pa, pb = b.make_frames(pkg, 32, 0, unique_streams=2, frames_per_stream=16)
    f1, _, _ = b.make_frames(pkg, 32, 1, unique_streams=2, frames_per_stream=16)
    assert f0.shape == (32, 480, 640) and (f0 != f1).any()
    assert (pa == np.arange(32)).all() and (pb // 16 == pa // 16).all() and ((pb - pa) % 16 == 1).all()


def test_c5_strong_partition_is_fixed_work():
    """bench.py's c5 data plan: 8 fixed streams, stream i on rank i % world, total frames independent of world, pairs
    stay inside their stream, and the union over ranks is the same set of streams for every world size."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("afv_bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    pkg = load_pkg()
    ref = None
    for world in (1, 2, 4, 8):
        per_rank = [b.c5_rank_frames(pkg, world, r, frames_per_stream=8, unique=4, w=96, h=64) for r in range(world)]
        assert sum(len(f) for f, _, _, _ in per_rank) == 8 * 8
        assert sorted(sum((m for _, _, _, m in per_rank), [])) == list(range(8))
        crc = 0
        for f, pa, pb, mine in per_rank:
            assert len(f) == 8 * len(mine) and (pa == np.arange(len(f))).all()
            assert (pb // 4 == pa // 4).all() and (pb // 8 == pa // 8).all()          # next frame of the same 4-cycle of the same stream
            crc += int(f.astype(np.int64).sum())
        ref = crc if ref is None else ref
        assert crc == ref
