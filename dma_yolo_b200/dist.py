"""Multi-GPU plumbing for the one exchange step of the path: one process per GPU, the
batch is sharded by rank, every rank runs forward + decode + NMS on its slice independently, and the
fixed-size padded detections are all-gathered (NCCL over NVLink; gloo in CPU tests) so that rank 0 can
run the unchanged host-side mAP accumulation in rank-then-image order.  Payload: B_local*max_det*6*4 B
(461 KB at 64x300) + counts — latency-bound, nothing to fuse it with.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


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


def all_gather_detections(padded: torch.Tensor, counts: torch.Tensor, group=None):
    """-> (all_padded [W*B, max_det, 6], all_counts [W*B]) on every rank, rank-major then image order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return padded, counts
    w = dist.get_world_size(group)
    allp = torch.empty((w * padded.shape[0],) + tuple(padded.shape[1:]), dtype=padded.dtype, device=padded.device)
    allc = torch.empty((w * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    if padded.is_cuda:
        dist.all_gather_into_tensor(allp, padded.contiguous(), group=group)
        dist.all_gather_into_tensor(allc, counts.contiguous(), group=group)
    else:  # gloo has no all_gather_into_tensor on every build: use the list form
        lp = [torch.empty_like(padded) for _ in range(w)]
        lc = [torch.empty_like(counts) for _ in range(w)]
        dist.all_gather(lp, padded.contiguous(), group=group)
        dist.all_gather(lc, counts.contiguous(), group=group)
        allp, allc = torch.cat(lp, 0), torch.cat(lc, 0)
    return allp, allc


def unpad(all_padded: torch.Tensor, all_counts: torch.Tensor):
    c = all_counts.tolist()
    return [all_padded[i, :n] for i, n in enumerate(c)]


def shard_batch(n_total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a global batch for `rank` (images are independent: no halo, no exchange)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)
