"""Not a test: target of the compute-sanitizer runs committed under profiles/ (memcheck / racecheck / synccheck).

One train step (forward + CTC + backward + fused AdamW) of a small bi-LSTM model in the benchmarked tensor-core
mode — H=128 is eligible for every tcgen05 kernel the bench uses (split-K cluster sweeps, TF32 GEMMs, conv2
forward / data / weight gradient) — plus a uni-GRU + lookahead model for the remaining kernels.

    compute-sanitizer --tool memcheck  --kernel-regex kns=ds2 python tests/gpu_sanitize_small.py
    compute-sanitizer --tool racecheck --kernel-regex kns=ds2 python tests/gpu_sanitize_small.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import deepspeech_pytorch_b200 as ds  # noqa: E402
from deepspeech_pytorch_b200.optim import FlatParams, FusedOptimizer  # noqa: E402
from gpu_helpers import make_model  # noqa: E402
from oracle import ds2_oracle as O  # noqa: E402


def one_step(rnn, bidir, H, layers, T, prec, direct, deferred):
    ds.set_precision(prec)
    torch.manual_seed(1)
    model = make_model(rnn, bidir, H, layers, ctx=5).train()
    flat = FlatParams(model, direct_grads=direct)
    if deferred:
        ds.ops.enable_deferred_weight_grads(torch.device("cuda", 0))
    opt = FusedOptimizer(flat, model.optim_cfg)
    x, targets, pct, tsz = O.synth_batch(4, T, seed=2, lmin=3, lmax=8)
    loss = model.training_step((x.cuda(), targets, pct.clone(), tsz), 0)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    print(f"{rnn} bidir={bidir} H={H} {prec} direct={direct} deferred={deferred}: loss {float(loss):.4f} "
          f"fallbacks {ds.get_lib().ds2_fallback_count(0)}", flush=True)


def main():
    # the tensor-core modes of the bench on shapes every tcgen05 kernel accepts (split-K cluster sweeps, TF32 / fp16
    # GEMMs, conv2 forward / data / weight gradient), then the bench's own configuration of the plumbing: fp16-operand
    # GEMMs, gradients written into the flat buffer, weight-gradient GEMMs on the side stream
    one_step("lstm", True, 128, 2, 64, "tf32", False, False)
    one_step("gru", False, 128, 2, 64, "tf32", False, False)
    one_step("lstm", True, 128, 2, 64, "fp16", True, True)
    if os.environ.get("DS2_SANITIZE_BIG") == "1":
        # H=1024 selects the compile-time-unrolled sweep instantiations (NKR_T) the benchmark runs
        one_step("lstm", True, 1024, 1, 48, "fp16", True, True)
    print("sanitize target done", flush=True)


if __name__ == "__main__":
    main()
