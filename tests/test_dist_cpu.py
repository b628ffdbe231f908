"""CPU, world_size 2 over gloo: sample sharding + the final counter all-reduce give the single-process answer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eventclip_b200.dist import AccuracyMeter, gather_predictions, shard_range


def _fake_outputs(n, n_cls, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, n_cls, generator=g)
    probs = torch.softmax(torch.randn(n, n_cls, generator=g), -1)
    valid = torch.rand(n, 3, generator=g) > 0.3
    labels = torch.randint(0, n_cls, (n,), generator=g)
    return logits, probs, valid, labels


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    logits, probs, valid, labels = _fake_outputs(n, 13, 5)
    lo, hi = shard_range(n)
    meter = AccuracyMeter()
    for a in range(lo, hi, 4):                      # a few "batches" of the local shard
        b = min(a + 4, hi)
        meter.update(dict(logits=logits[a:b], probs=probs[a:b], valid_masks=valid[a:b]), labels[a:b])
    meter.all_reduce()
    preds = gather_predictions(logits[lo:hi].argmax(-1))
    q.put((rank, meter.counters.tolist(), preds.tolist(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 100, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_reduction_matches_single_process():
    n = 37
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    logits, probs, valid, labels = _fake_outputs(n, 13, 5)
    single = AccuracyMeter()
    single.update(dict(logits=logits, probs=probs, valid_masks=valid), labels)
    for rank, counters, preds, span in res:
        assert counters == single.counters.tolist()
        assert preds == logits.argmax(-1).tolist()
    assert sorted(r[3] for r in res) == [(0, 19), (19, 37)]
    r = single.result()
    assert r["n"] == n and 0 <= r["logits_acc"] <= r["logits_acc5"] <= 1
