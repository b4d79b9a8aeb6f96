"""Not a test: one cfg-L recurrent layer (bi-LSTM 1024, B=32, T'=500), forward + backward once in a tensor-core mode
(argv: rnn type, precision mode — default the benchmarked precision-16 mode).
Target of the Nsight Compute captures (`ncu -k regex:rnn_ ... python tests/gpu_one_layer.py lstm fp16`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_diag_rnn as dr  # noqa: E402


def main():
    rnn = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    x, lens, ws, bnp = dr.make(rnn, True, 500, 32, 1024, 1024, True, ragged=False)
    dy = torch.randn(500, 32, 1024)
    out = dr.run(prec, rnn, True, x, lens, ws, bnp, dy)
    print("ok", float(out[0].abs().mean()))


if __name__ == "__main__":
    main()
