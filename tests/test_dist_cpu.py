"""world_size-2 gloo test of the data-parallel exchange (host-side logic, CPU)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from deepspeech_pytorch_b200 import dist as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1000, generator=g)
    local = grad.clone()
    D.allreduce_flat_grad(grad)
    # mean over ranks == DDP semantics once grad_scale = 1/world is applied
    ref = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + k)) for k in range(world))
    assert torch.allclose(grad, ref, atol=1e-6)
    assert not torch.equal(grad, local)
    m = torch.nn.BatchNorm1d(4)
    m.running_mean.fill_(float(rank + 1))
    D.broadcast_buffers(m, src=0)
    assert float(m.running_mean[0]) == 1.0
    assert D.shard_bins(7, rank, world) == list(range(7))[rank::world]
    # split exchange (conv head / recurrent+fc tail) == one all-reduce of the whole buffer
    from deepspeech_pytorch_b200.optim import FlatParams

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Linear(5, 7)
            self.rest = torch.nn.Linear(7, 3)
    tiny = Tiny()
    flat = FlatParams(tiny)
    ex = D.OverlappedGradAllReduce(flat, tiny)
    assert 0 < ex.split < flat.n
    for fire_hook in (True, False):
        flat.grad.copy_(torch.arange(flat.n, dtype=torch.float32) * (rank + 1))
        if fire_hook:
            tiny.front_end_grad_hook()          # what DeepSpeech.forward's tensor hook does during backward
        ex.finish()
        assert torch.allclose(flat.grad, torch.arange(flat.n, dtype=torch.float32) * sum(range(1, world + 1)))
    # per-block exchange (gradient sinks on): head first, then one recurrent layer at a time, conv in finish();
    # the union of the pieces must be the whole buffer, each element reduced exactly once
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Linear(5, 7)
            self.rnns = torch.nn.Sequential(torch.nn.Linear(7, 9), torch.nn.Linear(9, 9), torch.nn.Linear(9, 4))
            self.lookahead = torch.nn.Linear(4, 4)
            self.fc = torch.nn.Linear(4, 3)
    net = Net()
    flat = FlatParams(net, direct_grads=True)
    ex = D.OverlappedGradAllReduce(flat, net)
    assert ex.per_block and net.block_grad_hook is not None
    assert 0 < ex.split == ex.block_start["rnn0"] < ex.block_start["rnn1"] < ex.block_start["rnn2"] < ex.block_start["head"]
    for order in (["head", "rnn2", "rnn1", "rnn0"], ["head", "rnn1"], []):
        flat.grad.copy_(torch.arange(flat.n, dtype=torch.float32) * (rank + 1))
        for name in order:
            net.block_grad_hook(name)               # what DeepSpeech.forward's tensor hooks do during backward
        if order:
            net.front_end_grad_hook()
        ex.finish()
        assert torch.allclose(flat.grad, torch.arange(flat.n, dtype=torch.float32) * sum(range(1, world + 1))), order
    dist.destroy_process_group()
    out.put(rank)


def test_flat_gradient_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
