"""Multi-GPU plumbing for the one exchange step of the path: one process per GPU, the
batch is sharded by rank, every rank runs forward + decode + NMS on its slice independently, and the
fixed-size padded detections are all-gathered (NCCL over NVLink; gloo in CPU tests) so that rank 0 can
run the unchanged host-side mAP accumulation in rank-then-image order.

ONE collective per step: detections and their per-image counts travel in one packed fp32 buffer
`[B*max_det*6 | B counts (int32 bits)]` — the layout `ops.nms_batched` already writes (`Detections.packed`), so
nothing is copied before the exchange.  Payload: B_local*(max_det*6+1)*4 B (461 KB at 64x300) — latency-bound,
nothing to fuse it with; `all_gather_packed_async` returns at once (NCCL works on its own stream) so the exchange
of step i overlaps the forward of step i+1, and the ranks are not locked to the slowest GPU inside every step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def pack_detections(padded: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """(padded [B, max_det, 6] fp32, counts [B] int32) -> one flat fp32 buffer [B*max_det*6 + B]."""
    b = padded.shape[0]
    buf = torch.empty(padded.numel() + b, dtype=torch.float32, device=padded.device)
    buf[:padded.numel()] = padded.reshape(-1)
    buf[padded.numel():] = counts.to(torch.int32).view(torch.float32)
    return buf


def unpack_detections(buf: torch.Tensor, b: int, max_det: int):
    """Inverse of pack_detections on the last dimension of `buf` ([..., B*max_det*6 + B]); leading dims are ranks."""
    n = b * max_det * 6
    lead = tuple(buf.shape[:-1])
    padded = buf[..., :n].reshape(lead + (b, max_det, 6))
    counts = buf[..., n:].contiguous().view(torch.int32).reshape(lead + (b,))
    return padded, counts


def pad_detections(dets, max_det: int, device=None):
    """list of (n_i, 6) -> (padded [B, max_det, 6] fp32, counts [B] int32)."""
    device = device if device is not None else (dets[0].device if dets else 'cpu')
    out = torch.zeros((len(dets), max_det, 6), dtype=torch.float32, device=device)
    cnt = torch.zeros(len(dets), dtype=torch.int32, device=device)
    for i, d in enumerate(dets):
        n = min(int(d.shape[0]), max_det)
        out[i, :n] = d[:n]
        cnt[i] = n
    return out, cnt


class GatherHandle:
    """An all-gather in flight.  `result()` makes the current stream wait for it (no host block on CUDA) and returns
    (all_padded [W*B, max_det, 6], all_counts [W*B]) in rank-major then image order."""

    def __init__(self, work, recv, keep, b, max_det, world):
        self.work, self.recv, self.keep, self.b, self.max_det, self.world = work, recv, keep, b, max_det, world

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self

    def result(self):
        self.wait()
        padded, counts = unpack_detections(self.recv, self.b, self.max_det)
        return padded.reshape(self.world * self.b, self.max_det, 6), counts.reshape(self.world * self.b)


def all_gather_packed_async(packed: torch.Tensor, b: int, max_det: int, group=None, recv: torch.Tensor | None = None) -> GatherHandle:
    """One collective for detections + counts; returns immediately.  `recv` ([W, len(packed)]) may be passed to reuse
    a buffer (the caller must not reuse it before the previous handle on it was waited for)."""
    w = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if w == 1:
        return GatherHandle(None, packed.unsqueeze(0), packed, b, max_det, 1)
    if recv is None:
        recv = torch.empty((w, packed.numel()), dtype=packed.dtype, device=packed.device)
    if packed.is_cuda:
        work = dist.all_gather_into_tensor(recv, packed, group=group, async_op=True)
    else:  # gloo has no all_gather_into_tensor on every build: the list form, into the rows of `recv`
        work = dist.all_gather(list(recv.unbind(0)), packed, group=group, async_op=True)
    return GatherHandle(work, recv, packed, b, max_det, w)


def all_gather_detections(padded: torch.Tensor, counts: torch.Tensor, group=None, packed: torch.Tensor | None = None):
    """-> (all_padded [W*B, max_det, 6], all_counts [W*B]) on every rank, rank-major then image order (blocking form).
    `packed`: the buffer `padded` / `counts` are views of (Detections.packed) — saves the packing copy."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return padded, counts
    b, max_det = padded.shape[0], padded.shape[1]
    if packed is None:
        packed = pack_detections(padded, counts)
    return all_gather_packed_async(packed, b, max_det, group).result()


def unpad(all_padded: torch.Tensor, all_counts: torch.Tensor):
    c = all_counts.tolist()
    return [all_padded[i, :n] for i, n in enumerate(c)]


def shard_batch(n_total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a global batch for `rank` (images are independent: no halo, no exchange)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)
