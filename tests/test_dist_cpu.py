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


# ---- fine-tune step: one flat gradient buffer, one averaging all-reduce (eventclip_b200.dist.FlatParams) ----------------
def _flat_worker(rank, world, port, q):
    from eventclip_b200.dist import FlatParams
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)                   # same parameters on every rank
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((11, 64), (128, 4), (4, 128), (7,))]
    before = [p.detach().clone() for p in params]
    flat = FlatParams([params[:1], params[1:]])
    for p, b in zip(params, before):                       # re-homing keeps the values and makes the parameters views
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() >= flat.flat_p.data_ptr() and p.data_ptr() < flat.flat_p.data_ptr() + 4 * flat.numel
    gr = torch.Generator().manual_seed(100 + rank)         # different gradients per rank
    local = [torch.randn(p.shape, generator=gr) for p in params]
    for p, t in zip(params, local):
        flat.grad_view(p).copy_(t)
    flat.average_gradients()
    flat.flat_p.add_(flat.flat_g, alpha=-0.1)              # an SGD step on the flat buffer is visible through the parameters
    q.put((rank, flat.spans, [flat.grad_view(p).clone() for p in params], [p.detach().clone() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes = ((11, 64), (128, 4), (4, 128), (7,))
    g = torch.Generator().manual_seed(3)
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = []
    for rank in range(2):
        gr = torch.Generator().manual_seed(100 + rank)
        grads.append([torch.randn(s, generator=gr) for s in shapes])
    for rank, spans, got_g, got_p in res:
        assert spans == [(0, 704), (704, 704 + 512 + 512 + 7)]
        for i in range(4):
            mean = (grads[0][i] + grads[1][i]) / 2
            assert torch.allclose(got_g[i], mean, atol=1e-7)
            assert torch.allclose(got_p[i], params[i] - 0.1 * mean, atol=1e-6)
    for a, b in zip(res[0][3], res[1][3]):                 # both ranks end with identical parameters
        assert torch.equal(a, b)


def _bcast_worker(rank, world, port, q):
    import torch.distributed as dist
    from eventclip_b200.dist import FlatParams
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(50 + rank)           # DIFFERENT initial parameters per rank (LoRA lora_down is drawn
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((5, 8), (3,))]   # from the global RNG at injection)
    flat = FlatParams([params[:1], params[1:]])
    flat.broadcast()                                       # what DDP does at construction; train.FineTuner calls it
    q.put((rank, [p.detach().clone() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_params_broadcast_rank0_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(50)
    want = [torch.randn(s, generator=g) for s in ((5, 8), (3,))]
    for rank, got in res:
        for a, b in zip(got, want):
            assert torch.equal(a, b), rank                 # every rank holds rank 0's draw
