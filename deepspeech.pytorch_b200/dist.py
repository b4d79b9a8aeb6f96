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
    """The single-buffer exchange, issued in pieces as the gradients become final so that it hides behind the rest of
    the backward pass ("all-reduce RNN-layer grads as soon as each layer's backward finishes").

    Parameters are laid out in `model.parameters()` order: conv front-end (0.3 % of the buffer), recurrent layers
    0..L-1, [lookahead], fc head.  Backward visits them in reverse.  With `FlatParams(direct_grads=True)` the backward
    kernels write straight into the flat buffer, so a block's slice is final the moment its backward returns:
    `DeepSpeech.forward` registers a tensor hook on every block's input and calls `block_grad_ready(name)` from it,
    which starts the asynchronous all-reduce (NCCL orders it after the kernels already queued on the compute stream) of
    that block's contiguous slice — the fc head + lookahead first, then one recurrent layer at a time while the next
    layer's sweep runs.  Without gradient sinks only the conv / rest split of round 1 is safe (AccumulateGrad nodes
    may run after the hook): the whole tail is exchanged when autograd reaches the front-end.  `finish()` after
    `backward()` reduces what is left (the conv slice, or everything if no hook fired) and waits for all pieces."""

    def __init__(self, flat, model):
        self.flat = flat
        self.works = []
        self.done_upto = flat.n                    # everything at offsets >= done_upto has been handed to NCCL
        conv_ids = {id(p) for p in model.conv.parameters()}
        split = 0
        for p, o in zip(flat.params, flat.offsets):
            if id(p) in conv_ids:
                split = max(split, o + (p.numel() + 63) // 64 * 64)
        # the conv parameters must form a prefix of the buffer, otherwise fall back to one exchange in finish()
        prefix_ok = all((id(p) in conv_ids) == (o < split) for p, o in zip(flat.params, flat.offsets))
        self.split = split if prefix_ok else flat.n
        # start offset of every block's slice (blocks are contiguous and in forward order)
        self.block_start = {}
        if prefix_ok:
            offs = {id(p): o for p, o in zip(flat.params, flat.offsets)}
            for i, layer in enumerate(getattr(model, "rnns", None) or []):
                ps = [offs[id(p)] for p in layer.parameters() if id(p) in offs]
                if ps:
                    self.block_start[f"rnn{i}"] = min(ps)
            tail_mods = [m for m in (getattr(model, "lookahead", None), getattr(model, "fc", None)) if m is not None]
            tail = [offs[id(p)] for m in tail_mods for p in m.parameters() if id(p) in offs]
            if tail:
                self.block_start["head"] = min(tail)
        self.per_block = bool(getattr(flat, "direct_grads", False)) and bool(self.block_start)
        model.front_end_grad_hook = self._tail_ready
        model.block_grad_hook = self.block_grad_ready if self.per_block else None

    def _reduce_from(self, start):
        """hand [start, done_upto) to NCCL"""
        if world_size() > 1 and start < self.done_upto:
            from . import ops
            piece = self.flat.grad[start:self.done_upto]
            side = ops.side_stream()
            if side is not None and piece.is_cuda:
                # Deferred weight-gradient GEMMs (side stream) write into this slice.  Issue the collective FROM the
                # side stream, after it has also picked up what the compute stream produced so far (biases, BatchNorm
                # gradients): NCCL then waits for both without stalling the compute stream, whose next kernel is the
                # next layer's sweep.
                side.wait_stream(torch.cuda.current_stream(piece.device))
                with torch.cuda.stream(side):
                    self.works.append(dist.all_reduce(piece, op=dist.ReduceOp.SUM, async_op=True))
            else:
                self.works.append(dist.all_reduce(piece, op=dist.ReduceOp.SUM, async_op=True))
        self.done_upto = min(self.done_upto, start)

    def block_grad_ready(self, name):
        """the gradients of block `name` and of every block after it are final"""
        if name in self.block_start:
            self._reduce_from(self.block_start[name])

    def _tail_ready(self):
        self._reduce_from(self.split)

    def finish(self):
        if world_size() > 1:
            self._reduce_from(0)
            for w in self.works:
                w.wait()
        self.works = []
        self.done_upto = self.flat.n


def broadcast_buffers(model, src: int = 0):
    if world_size() == 1:
        return
    for b in model.buffers():
        dist.broadcast(b, src)


def shard_bins(n_bins: int, rank: int, world: int):
    """the reference's DSElasticDistributedSampler dealing of batch bins (data_loader.py:326-357):
    bins[rank::world]"""
    return list(range(n_bins))[rank::world]
