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


def main():
    ds.set_precision("tf32")
    for rnn, bidir in (("lstm", True), ("gru", False)):
        torch.manual_seed(1)
        model = make_model(rnn, bidir, 128, 2, ctx=5).train()
        flat = FlatParams(model)
        opt = FusedOptimizer(flat, model.optim_cfg)
        x, targets, pct, tsz = O.synth_batch(4, 64, seed=2, lmin=3, lmax=8)
        loss = model.training_step((x.cuda(), targets, pct.clone(), tsz), 0)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        print(f"{rnn} bidir={bidir}: loss {float(loss):.4f} fallbacks {ds.get_lib().ds2_fallback_count(0)}", flush=True)
    print("sanitize target done", flush=True)


if __name__ == "__main__":
    main()
