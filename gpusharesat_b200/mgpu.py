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


HEADER_BYTES = 64


def _round_up(n, q):
    return (n + q - 1) // q * q


def _next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


class ShardedRunner:
    """Fast path: the library runs on torch's current stream, the batch travels as ONE packed
    payload ([header][params][deltas], gss_mgpu_collect_to) and is broadcast with a single
    collective whose size every rank predicts from the previous batch (a second collective only
    when a batch outgrows the prediction); the per-rank hits are all-gathered from device memory.

    Collectives per batch: one broadcast (payload) + one all_gather ([count][hits] per rank)."""

    def __init__(self, sh, dist, rank, world, device, payload_cap=64 << 20, min_bcast=1 << 16):
        self.sh, self.dist, self.rank, self.world, self.device = sh, dist, rank, world, device
        self.payload = torch.empty(payload_cap, dtype=torch.uint8, device=device)
        self.pred = min_bcast
        self.min_bcast = min_bcast
        self.hit_buf = torch.empty(1 << 16, dtype=torch.uint8, device=device)
        self.counts = torch.zeros(world, dtype=torch.int64, device=device)
        if device.type == "cuda":
            sh.setStream(torch.cuda.current_stream(device).cuda_stream)
        self.ev0 = self.ev1 = None
        self.hit_pred = 4096

    def device_us(self):
        """device time of the last step from the start of the payload broadcast to the gathered hits"""
        if self.ev0 is None or self.ev1 is None:
            return 0.0
        self.ev1.synchronize()
        return self.ev0.elapsed_time(self.ev1) * 1e3

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _gather(self, enqueue=True):
        """all-gather of [64 B header][hits x hit_pred] from every rank; returns (host headers
        [world, 8] int64, gathered device tensor, slot bytes)"""
        rec = RAW_HIT_DTYPE.itemsize
        slot = HEADER_BYTES + self.hit_pred * rec
        if self.hit_buf.numel() < slot:
            self.hit_buf = torch.empty(_next_pow2(slot), dtype=torch.uint8, device=self.device)
        self.sh.mgpuEnqueueResult(self.hit_buf.data_ptr(), self.hit_pred)
        gathered = torch.empty(self.world * slot, dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(gathered, self.hit_buf[:slot])
        return gathered, slot

    def step(self):
        """One batch.  Between the payload broadcast and the gathered hits nothing synchronises with
        the host: receivers enqueue the batch blindly (gss_mgpu_enqueue_payload), every rank
        enqueues its result block, one all-gather, ONE synchronisation at the end.  Mis-predicted
        sizes (payload or hit block) are detected from the headers afterwards and repaired with a
        second round.  Returns the total number of hits (rank 0) / 0 (others), None if nothing ran."""
        sh, dist, rank, world = self.sh, self.dist, self.rank, self.world
        cap = self.payload.numel()
        cuda = self.device.type == "cuda"
        rec = RAW_HIT_DTYPE.itemsize
        total = sh.mgpuCollectTo(self.payload.data_ptr(), cap) if rank == 0 else 0
        if cuda:  # device-side time of the exchange + kernels: from the broadcast to the gathered hits
            self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.ev0.record()
        n = min(self.pred, cap)
        dist.broadcast(self.payload[:n], 0)
        if rank == 0:
            status = sh.mgpuRunPayload(self.payload.data_ptr(), cap)
        else:
            status = sh.mgpuEnqueuePayload(self.payload.data_ptr(), n)
        if status < 0:  # no clause anywhere yet (every rank holds the same clause stream)
            self._sync()
            return None
        gathered, slot = self._gather()
        if cuda:
            self.ev1.record()
        # ---- the only synchronisation of the step ----
        heads = gathered.view(world, slot)[:, :HEADER_BYTES].contiguous().cpu().numpy().view(np.int64)
        if rank != 0:
            total = int(self.payload[:HEADER_BYTES].cpu().numpy().view(np.int64)[3])
        sh.mgpuFinish()  # a rank whose header says "overflow" grows its buffers and runs again in here
        counts, flags = heads[:, 0].tolist(), heads[:, 1].tolist()
        redo_in_flight = False
        if total > n:  # the batch outgrew the predicted broadcast: ship the rest, receivers run again
            dist.broadcast(self.payload[n:min(cap, _round_up(total, 1 << 16))], 0)
            if rank != 0:
                sh.mgpuRedoPayload(self.payload.data_ptr(), total)
                redo_in_flight = True
            need = True
        else:
            need = any(flags) or max(counts) > self.hit_pred
        while need:  # second round(s): every rank re-contributes its (now complete) result block
            self.hit_pred = max(self.hit_pred, _next_pow2(max(counts) + 1))
            gathered, slot = self._gather()
            heads = gathered.view(world, slot)[:, :HEADER_BYTES].contiguous().cpu().numpy().view(np.int64)
            if redo_in_flight:
                sh.mgpuFinish()
                redo_in_flight = False
            counts, flags = heads[:, 0].tolist(), heads[:, 1].tolist()
            need = any(flags) or max(counts) > self.hit_pred
        self.pred = max(self.min_bcast, _round_up(total + total // 4, 1 << 16))
        m = max(counts)
        self.hit_pred = max(1024, _round_up(m + m // 2, 1024))
        if rank != 0:
            return 0
        # the union stays on the device: the library concatenates, sorts and resolves it there
        sh.mgpuImportGathered(gathered.data_ptr(), world, slot, counts)
        return int(sum(counts))


class PeerRunner:
    """Data path over peer memory instead of collectives (csrc/peer.cu): every rank exports a
    window of its device memory with CUDA IPC, workers read the batch out of rank 0's memory and
    append their hits into rank 0's memory over NVLink, streams wait on flags.  torch.distributed
    only carries the 128-byte window handles once, at start-up.

    SPMD: every rank calls step() for every batch (like the clause stream, which every rank sees)."""

    def __init__(self, sh, dist, rank, world, payload_cap=64 << 20, slot_hits=1 << 20, records=None):
        self.sh, self.rank, self.world = sh, rank, world
        if records is not None:  # (None: the library's default / GPUSHARE_PEER_RECORDS)
            sh.debugSetPeerRecords(records)
        blob = sh.peerInit(rank, world, payload_cap, slot_hits)
        blobs = [None] * world
        if world > 1:
            dist.all_gather_object(blobs, blob)
        else:
            blobs[0] = blob
        sh.peerConnect(blobs)
        if world > 1:
            dist.barrier()

    def set_records(self, on):
        """workers export their sorted record keys + masks (rank 0: activity bumps for every rank's hits, debugLastHits
        over every rank) or every rank bumps its own hits; call on every rank between two batches"""
        self.sh.debugSetPeerRecords(on)

    def step(self):
        """one batch; rank 0 returns the number of hits handed over, workers their own count, None
        when there is no clause yet"""
        if self.sh.peerEnqueue() < 0:
            return None
        return self.sh.peerFinish()

    def device_us(self):
        """device time of the last batch on this rank from "batch resident in rank 0's HBM" to the end
        of the step (rank 0: every rank's hits have arrived in its memory)"""
        t = self.sh.debugLastRunTimes()
        return (t[3] - t[0]) if t else 0.0
