"""world_size-2 gloo test of the N>1 host path: batch sharding + detection all-gather order."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dma_yolo_b200.dist import (all_gather_detections, all_gather_packed_async, pack_detections, pad_detections,
                                    shard_batch, unpad)
    g = torch.Generator().manual_seed(0)
    n_total = 6
    all_dets = [torch.rand(int(torch.randint(0, 9, (1,), generator=g)), 6, generator=g) for _ in range(n_total)]
    lo, hi = shard_batch(n_total, rank, world)
    padded, counts = pad_detections(all_dets[lo:hi], 8)
    ap, ac = all_gather_detections(padded, counts)
    got = unpad(ap, ac)
    ok = len(got) == n_total and all(torch.equal(a, b[:8]) for a, b in zip(got, all_dets))
    # the asynchronous one-collective form the GPU path uses: two exchanges in flight one after the other
    h1 = all_gather_packed_async(pack_detections(padded, counts), hi - lo, 8)
    h2 = all_gather_packed_async(pack_detections(padded * 2, counts), hi - lo, 8)
    p1, c1 = h1.result()
    p2, c2 = h2.result()
    ok = ok and torch.equal(p1, ap) and torch.equal(c1, ac) and torch.equal(p2, ap * 2) and torch.equal(c2, ac)
    ok = ok and c1.dtype == torch.int32 and tuple(p1.shape) == (n_total, 8, 6)
    torch.save(ok, os.path.join(tmp, f'ok{rank}.pt'))
    dist.destroy_process_group()


def test_all_gather_detections_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(torch.load(tmp_path / f'ok{r}.pt') for r in range(2))


def test_shard_batch_covers_everything():
    from dma_yolo_b200.dist import shard_batch
    for n in (1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
