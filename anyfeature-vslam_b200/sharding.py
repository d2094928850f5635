"""Multi-GPU host logic (one process per GPU): frames / streams are independent, so ranks take disjoint streams
(weak scaling) and the only exchange is a gather of fixed-capacity packed results to rank 0 (BASELINE config C5:
"NCCL over NVLink only to gather results").  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def streams_of_rank(nstreams_total, world, rank):
    """Stream i lives on rank i % world."""
    return [s for s in range(nstreams_total) if s % world == rank]


def pack_layout(B, cap, desc_bytes=32):
    """Byte layout of one rank's packed results: n[B] i32 | nmatches[B] i32 | matches12[B,cap] i32 | kps[B,cap,28] | desc[B,cap,D]."""
    sizes = [("n", 4 * B), ("nmatches", 4 * B), ("matches12", 4 * B * cap), ("kps", 28 * B * cap), ("desc", desc_bytes * B * cap)]
    off, lay = 0, {}
    for name, sz in sizes:
        lay[name] = (off, sz)
        off += sz
    return lay, off


def pack_results(pack, n, nmatches, matches12, kps, desc):
    """Copy the five device (or host) tensors into the flat uint8 tensor `pack` (same device)."""
    import torch
    lay, total = pack_layout(n.shape[0], matches12.shape[1], desc.shape[2])
    assert pack.numel() == total
    if pack.is_cuda and total % 4 == 0:
        # one launch of the library's pack kernel (afv_pack_results) on the current stream instead of five torch copies
        import ctypes as C
        from . import lib, _check, _vp, _stream_ptr
        _check(lib().afv_pack_results(_vp(n), _vp(nmatches), _vp(matches12), _vp(kps), _vp(desc), n.shape[0], matches12.shape[1], desc.shape[2],
                                      _vp(pack), _stream_ptr(None)))
        return pack
    for name, t in (("n", n), ("nmatches", nmatches), ("matches12", matches12), ("kps", kps), ("desc", desc)):
        o, sz = lay[name]
        pack[o:o + sz].copy_(t.contiguous().view(torch.uint8).view(-1), non_blocking=True)
    return pack


def gather_to_root(pack, gathered, world, rank):
    """torch.distributed.gather of the packed buffers (gathered: list of `world` tensors on rank 0, else None)."""
    import torch.distributed as dist
    if world > 1:
        dist.gather(pack, gathered, dst=0)
    return gathered


def unpack_results(buf, B, cap, desc_bytes=32):
    """numpy view of one rank's packed buffer (host uint8 array) -> dict of arrays."""
    lay, total = pack_layout(B, cap, desc_bytes)
    buf = np.asarray(buf, np.uint8)
    assert buf.size == total
    out = {}
    for name, (o, sz) in lay.items():
        out[name] = buf[o:o + sz]
    return {"n": out["n"].view(np.int32), "nmatches": out["nmatches"].view(np.int32),
            "matches12": out["matches12"].view(np.int32).reshape(B, cap), "kps": out["kps"].reshape(B, cap, 28),
            "desc": out["desc"].reshape(B, cap, desc_bytes)}
