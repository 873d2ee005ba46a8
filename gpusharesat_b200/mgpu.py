"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The library does the work; this module only moves
the per-batch payload (run parameters + assignment deltas, produced by rank 0) to every rank and
the per-rank hit lists back to rank 0.  See include/gpushare_b200.h (gss_mgpu_*)."""
import ctypes as C

import numpy as np
import torch

from .api import RAW_HIT_DTYPE

VARUPDATE_BYTES = 12


def _host_view(ptr, nbytes):
    """numpy view of `nbytes` bytes of (pinned) host memory owned by the library"""
    if nbytes == 0:
        return np.zeros(0, dtype=np.uint8)
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr))


def broadcast_batch(dist, rank, device, collected):
    """rank 0 passes the result of GpuClauseSharer.mgpuCollect() (or None); every rank gets
    (status, params_tensor, updates_tensor, n_updates) with the tensors on `device`.
    status: -1 nothing to run, 0 normal batch, 1 batch that rebuilds the tables."""
    header = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == 0:
        if collected is None:
            header[0] = -1
        else:
            rebuild, _, pbytes, _, nupd = collected
            header[0], header[1], header[2] = rebuild, pbytes, nupd
    dist.broadcast(header, 0)
    status, pbytes, nupd = (int(x) for x in header.tolist())
    if status < 0:
        return status, None, None, 0
    params = torch.empty(pbytes, dtype=torch.uint8, device=device)
    updates = torch.empty(max(nupd, 1) * VARUPDATE_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        _, pptr, _, uptr, _ = collected
        params.copy_(torch.from_numpy(_host_view(pptr, pbytes)))
        if nupd:
            updates[: nupd * VARUPDATE_BYTES].copy_(torch.from_numpy(_host_view(uptr, nupd * VARUPDATE_BYTES)))
    dist.broadcast(params, 0)
    if nupd:
        dist.broadcast(updates, 0)
    return status, params, updates, nupd


def gather_hits(dist, rank, world, device, my_hits):
    """every rank passes its hits (numpy RAW_HIT_DTYPE); rank 0 gets the concatenation (others None)"""
    n = torch.tensor([len(my_hits)], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c) for c in counts]
    m = max(counts)
    if m == 0:
        return np.zeros(0, dtype=RAW_HIT_DTYPE) if rank == 0 else None
    buf = torch.zeros(m * RAW_HIT_DTYPE.itemsize, dtype=torch.uint8, device=device)
    if len(my_hits):
        raw = torch.from_numpy(np.ascontiguousarray(my_hits).view(np.uint8).reshape(-1))
        buf[: raw.numel()].copy_(raw)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    parts = [b[: c * RAW_HIT_DTYPE.itemsize].cpu().numpy().view(RAW_HIT_DTYPE) for b, c in zip(bufs, counts) if c]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=RAW_HIT_DTYPE)


def run_batch(sh, dist, rank, world, device):
    """One batch through the sharded path.  Rank 0 returns the number of hits handed over
    (None when nothing ran); other ranks return their local hit count."""
    collected = sh.mgpuCollect() if rank == 0 else None
    status, params, updates, nupd = broadcast_batch(dist, rank, device, collected)
    if status < 0:
        return None
    if device.type == "cuda":
        torch.cuda.synchronize(device)  # the library launches on its own stream
    sh.mgpuRun(params.data_ptr(), params.numel(), updates.data_ptr(), nupd, status)
    mine = sh.mgpuWait()
    allhits = gather_hits(dist, rank, world, device, mine)
    if rank == 0:
        sh.mgpuImport(allhits)
        return len(allhits)
    return len(mine)
