"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding, the global-batch reduction
hooks and the flat gradient bucket reproduce a single-process run at the global batch."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loss_sums(logits, target, K, ignore=255):
    """torch restatement of the [2+3K] partial sums of essb_task_loss_fwd (host-logic test only)."""
    mask = target != ignore
    lp = torch.log_softmax(logits.double(), 1)
    p = lp.exp() * mask.unsqueeze(1)
    t = torch.nn.functional.one_hot((target * mask).long(), K).permute(0, 3, 1, 2).double() * mask.unsqueeze(1)
    ce = -(lp * t).sum()
    return torch.cat([ce.view(1), mask.sum().double().view(1), (p * t).sum((0, 2, 3)), (p * p).sum((0, 2, 3)),
                      t.sum((0, 2, 3))])


def _loss_from_sums(s, K):
    dice = sum(1 - (2 * s[2 + k] + 1) / (s[2 + K + k] + s[2 + 2 * K + k] + 1) for k in range(K)) / K
    return dice + s[0] / s[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ess_b200 import dp
    r, w, _ = dp.init_from_env('gloo')
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(0)
    B, K = 4, 5
    logits = torch.randn(B, K, 6, 7, generator=g)
    target = torch.randint(0, K, (B, 6, 7), generator=g)
    target[:, 0] = 255
    events = torch.randn(B, 6, 6, 7, generator=g) * (torch.rand(B, 6, 6, 7, generator=g) < 0.3)
    lo, hi = dp.shard_batch(B, rank, world)
    # (1) event statistics: [T,3] sums all-reduced == statistics of the whole batch tensor
    T, C = 2, 3
    ev = events[lo:hi].view(hi - lo, T, -1).double()
    stats = torch.stack([ev.sum((0, 2)), (ev * ev).sum((0, 2)), (ev != 0).sum((0, 2)).double()], 1)
    dp.allreduce_sum_(stats)
    full = events.view(B, T, -1).double()
    ref = torch.stack([full.sum((0, 2)), (full * full).sum((0, 2)), (full != 0).sum((0, 2)).double()], 1)
    assert torch.allclose(stats, ref, rtol=1e-12, atol=1e-12)
    # (2) loss partial sums -> the GLOBAL-batch loss on every rank
    sums = dp.allreduce_sum_(_loss_sums(logits[lo:hi], target[lo:hi], K))
    assert torch.allclose(_loss_from_sums(sums, K), _loss_from_sums(_loss_sums(logits, target, K), K), rtol=1e-12)
    # (3) flat gradient bucket: SUM (no 1/N) of per-rank gradients of the global loss
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(0.5)
        lin.bias.fill_(0.1)
    bucket = dp.GradBucket(lin.parameters())
    bucket.zero_()
    x = torch.arange(12.).view(4, 3)
    (lin(x[lo:hi]).sum()).backward()
    bucket.allreduce_()
    lin2 = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin2.weight.fill_(0.5)
        lin2.bias.fill_(0.1)
    lin2(x).sum().backward()
    assert torch.allclose(lin.weight.grad, lin2.weight.grad) and torch.allclose(lin.bias.grad, lin2.bias.grad)
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    # (4) overlapped exchange: gradients are handed to the bucket's sink in COMPLETION order (reverse parameter order,
    # as _DecoderFn.backward does); each stage's slice is all-reduced as soon as it is complete, the rest at the end
    net = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    with torch.no_grad():
        for i, p_ in enumerate(net.parameters()):
            p_.copy_(torch.linspace(-1, 1, p_.numel()).view_as(p_) * (i + 1) * 0.1)
    ob = dp.GradBucket(net.parameters(), module=net, stage_floats=10)
    assert net._grad_sink == ob._sink and len(ob.stages) >= 2
    assert sorted(i for lo_, hi_ in ob.stages for i in range(lo_, hi_)) == list(range(6))
    for rep in range(2):
        ob.zero_()
        loss = net(x[lo:hi]).pow(2).sum()
        gs = torch.autograd.grad(loss, list(net.parameters()))
        names = [n for n, _ in net.named_parameters()]
        sent_before_flush = 0
        for n, g_ in reversed(list(zip(names, gs))):
            assert ob._sink(n, g_) is True
        sent_before_flush = sum(ob.sent)
        assert sent_before_flush == len(ob.stages)          # every stage left as soon as its last gradient arrived
        assert ob._sink('not_a_parameter', gs[0]) is False
        ob.allreduce_()
        ref_net = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
        ref_net.load_state_dict(net.state_dict())
        ref_g = torch.autograd.grad(ref_net(x).pow(2).sum(), list(ref_net.parameters()))
        for p_, rg in zip(net.parameters(), ref_g):
            assert torch.allclose(p_.grad, rg, rtol=1e-6, atol=1e-6)
    # a gradient delivered twice in one step (UDA: two backward passes) accumulates
    ob.zero_()
    ob._sink(names[-1], gs[-1])
    ob._sink(names[-1], gs[-1])
    assert torch.allclose(list(net.parameters())[-1].grad, 2 * gs[-1])
    torch.distributed.destroy_process_group()
    q.put((rank, 'ok'))


def test_dp_host_logic_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, 'ok'), (1, 'ok')]
