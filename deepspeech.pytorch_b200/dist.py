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


def broadcast_buffers(model, src: int = 0):
    if world_size() == 1:
        return
    for b in model.buffers():
        dist.broadcast(b, src)


def shard_bins(n_bins: int, rank: int, world: int):
    """the reference's DSElasticDistributedSampler dealing of batch bins (data_loader.py:326-357):
    bins[rank::world]"""
    return list(range(n_bins))[rank::world]
