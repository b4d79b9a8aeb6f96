"""Utterance-level data parallelism: one process per GPU, every rank runs the whole
model on its own minibatch, and the only exchange is ONE all-reduce(sum) of the flat fp32 gradient
buffer over NCCL (NVLink 5 / NVSwitch) after backward.  DDP's division by world size is folded into
the fused optimizer's `grad_scale`.  BatchNorm running statistics are not synchronised
(sync_batchnorm=False in the reference, lightning_config.py:53); rank 0's are broadcast on demand
(`broadcast_buffers`) before evaluation / checkpointing, which is what DDP's per-forward buffer
broadcast amounts to for a training run.

Works with any torch.distributed backend (`gloo` in the CPU tests, `nccl` on the GPU box)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style bootstrap; returns (rank, world, local_rank)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_flat_grad(flat_grad: torch.Tensor, async_op: bool = False):
    """sum of per-rank gradients, in place on the flat buffer; returns the work handle when async"""
    if world_size() == 1:
        return None
    return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)


class OverlappedGradAllReduce:
    """The same single-buffer exchange, split in two so that most of it hides behind the conv backward.

    Parameters are laid out in `model.parameters()` order: the conv front-end first (0.3 % of the buffer), then
    the recurrent stack and the fc head.  The front-end's backward runs LAST, and by the time autograd reaches it
    every gradient of the tail [split:] is final (AccumulateGrad nodes run before the next function node).
    `DeepSpeech.front_end_grad_hook` fires exactly there: the tail's all-reduce is started asynchronously (NCCL
    orders it after the work already queued on the compute stream) and overlaps the ~5 ms conv backward;
    `finish()` after `backward()` reduces the small head and waits for the tail."""

    def __init__(self, flat, model):
        self.flat = flat
        conv_ids = {id(p) for p in model.conv.parameters()}
        split = 0
        for p, o in zip(flat.params, flat.offsets):
            if id(p) in conv_ids:
                split = max(split, o + (p.numel() + 63) // 64 * 64)
        # the conv parameters must form a prefix of the buffer, otherwise fall back to one exchange in finish()
        prefix_ok = all((id(p) in conv_ids) == (o < split) for p, o in zip(flat.params, flat.offsets))
        self.split = split if prefix_ok else flat.n
        self.work = None
        model.front_end_grad_hook = self._tail_ready

    def _tail_ready(self):
        if world_size() > 1 and self.split < self.flat.n:
            self.work = dist.all_reduce(self.flat.grad[self.split:], op=dist.ReduceOp.SUM, async_op=True)

    def finish(self):
        if world_size() == 1:
            return
        if self.split > 0:
            dist.all_reduce(self.flat.grad[:self.split], op=dist.ReduceOp.SUM)
        if self.work is not None:
            self.work.wait()
            self.work = None
        elif self.split < self.flat.n:       # hook did not fire (e.g. frozen front-end): exchange the tail now
            dist.all_reduce(self.flat.grad[self.split:], op=dist.ReduceOp.SUM)


def broadcast_buffers(model, src: int = 0):
    if world_size() == 1:
        return
    for b in model.buffers():
        dist.broadcast(b, src)


def shard_bins(n_bins: int, rank: int, world: int):
    """the reference's DSElasticDistributedSampler dealing of batch bins (data_loader.py:326-357):
    bins[rank::world]"""
    return list(range(n_bins))[rank::world]
