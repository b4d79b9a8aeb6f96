import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_diag_rnn as D
args = ("lstm", True, 37, 32, 96, 128, True)
x, lens, ws, bnp = D.make(*args)
dy = torch.randn(37, 32, 128)
out = D.run("tf32", args[0], args[1], x, lens, ws, bnp, dy)
torch.cuda.synchronize()
print("ok", float(out[0].abs().sum()))
