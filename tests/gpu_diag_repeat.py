"""Not a test: which outputs of one tensor-core-mode layer differ between identical runs, and by how much."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepspeech_pytorch_b200 as ds  # noqa: E402
from test_gpu_fullsize import _layer_case, _run_b200  # noqa: E402


def main():
    for prec in ("tf32", "fp16"):
        ds.set_precision(prec)
        rnn, bidir, T, B, In, H = "lstm", True, 150, 32, 256, 1024
        x, lens, P, dy = _layer_case(rnn, bidir, T, B, In, H, seed=29)
        first = _run_b200(rnn, bidir, x, lens, P, dy)
        worst = {}
        for it in range(10):
            again = _run_b200(rnn, bidir, x, lens, P, dy)
            for n in ("y", "hn", "cn", "dx"):
                d = float((first[n] - again[n]).abs().max())
                worst[n] = max(worst.get(n, 0.0), d)
            for k in first["grads"]:
                a, b = first["grads"][k], again["grads"][k]
                d = float((a - b).abs().max())
                if d > worst.get(k, (0.0,))[0] if isinstance(worst.get(k), tuple) else True:
                    idx = int((a - b).abs().argmax())
                    worst[k] = (max(d, worst.get(k, (0.0, 0, 0))[0]), idx, int(((a - b) != 0).sum()))
        print(prec, {k: v for k, v in worst.items()}, flush=True)


if __name__ == "__main__":
    main()
