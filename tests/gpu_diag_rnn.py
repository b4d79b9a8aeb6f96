"""Not a test: recurrent-layer bring-up diagnostics. fp32 FFMA step kernels vs tcgen05 persistent sweep.
(In-kernel clock64 traces: tests/gpu_diag_sweep.py.)"""
import ctypes
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepspeech_pytorch_b200 as ds  # noqa: E402
from deepspeech_pytorch_b200 import _lib  # noqa: E402
from gpu_helpers import rel  # noqa: E402

CODE = {"lstm": _lib.RNN_LSTM, "gru": _lib.RNN_GRU, "rnn": _lib.RNN_TANH}
G = {"lstm": 4, "gru": 3, "rnn": 1}


def make(rnn, bidir, T, B, In, H, bn, seed=0, ragged=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, B, In, generator=g)
    lens = [max(1, T - (3 * i if ragged else 0)) for i in range(B)]
    for b, l in enumerate(lens):
        x[l:, b] = 0
    k = 1.0 / H ** 0.5
    ws = []
    for _ in range(2 if bidir else 1):
        ws += [(torch.rand(G[rnn] * H, In, generator=g) * 2 - 1) * k, (torch.rand(G[rnn] * H, H, generator=g) * 2 - 1) * k,
               (torch.rand(G[rnn] * H, generator=g) * 2 - 1) * k, (torch.rand(G[rnn] * H, generator=g) * 2 - 1) * k]
    bnp = None
    if bn:
        bnp = [1 + 0.1 * torch.randn(In, generator=g), 0.1 * torch.randn(In, generator=g), torch.zeros(In), torch.ones(In)]
    return x, torch.tensor(lens, dtype=torch.int32), ws, bnp


def run(prec, rnn, bidir, x, lens, ws, bnp, dy, training=True):
    ds.set_precision(prec)
    xc = x.cuda().requires_grad_(True)
    wc = [w.cuda().requires_grad_(True) for w in ws]
    bn = [t.cuda() for t in bnp] if bnp else [None] * 4
    if bnp:
        bn[0].requires_grad_(True); bn[1].requires_grad_(True)
    y, hn, cn = ds.ops.RnnLayer.apply(xc, lens.cuda(), CODE[rnn], bidir, training, 0.1, 1e-5, bn[0], bn[1], bn[2], bn[3],
                                      None, None, *wc)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    return y.detach(), hn.detach(), (cn.detach() if cn is not None else None), xc.grad, [w.grad for w in wc]


def compare(rnn, bidir, T, B, In, H, bn):
    x, lens, ws, bnp = make(rnn, bidir, T, B, In, H, bn)
    dy = torch.randn(T, B, H, generator=torch.Generator().manual_seed(9))
    for b in range(B):
        dy[int(lens[b]):, b] = 0
    a = run("fp32", rnn, bidir, x, lens, ws, bnp, dy)
    c = run("tf32", rnn, bidir, x, lens, ws, bnp, dy)
    msg = f"{rnn} bidir={bidir} T={T} B={B} In={In} H={H} bn={bn}: y {rel(c[0], a[0]):.2e} hn {rel(c[1], a[1]):.2e}"
    if a[2] is not None:
        msg += f" cn {rel(c[2], a[2]):.2e}"
    msg += f" dx {rel(c[3], a[3]):.2e} dW " + " ".join(f"{rel(gc, ga):.1e}" for gc, ga in zip(c[4], a[4]))
    print(msg, flush=True)


def timing(rnn, bidir, T, B, In, H, bn, iters=3):
    x, lens, ws, bnp = make(rnn, bidir, T, B, In, H, bn, ragged=False)
    dy = torch.randn(T, B, H)
    lib = ds.get_lib()
    for prec in ("tf32",):
        run(prec, rnn, bidir, x, lens, ws, bnp, dy)
        lib.ds2_prof_enable(1)
        for _ in range(iters):
            run(prec, rnn, bidir, x, lens, ws, bnp, dy)
        buf = (ctypes.c_char * 4096)()
        lib.ds2_prof_report(buf, 4096)
        lib.ds2_prof_enable(0)
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                tag, ms, cnt = item.split(":")
                out[tag] = round(float(ms) / iters, 3)
        print(f"timing {prec} {rnn} bidir={bidir} T={T} B={B} In={In} H={H}: {out}", flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    for args in [("lstm", True, 20, 5, 64, 64, False), ("lstm", True, 37, 32, 96, 128, True), ("gru", True, 25, 7, 64, 64, True),
                 ("gru", False, 30, 33, 32, 96, False), ("rnn", True, 19, 4, 32, 32, False), ("lstm", False, 50, 40, 128, 256, True)]:
        try:
            compare(*args)
        except Exception:
            print("[EXC]", args, traceback.format_exc(), flush=True)
    try:
        timing("lstm", True, 500, 32, 1024, 1024, True)
    except Exception:
        print("[EXC] timing", traceback.format_exc(), flush=True)


if __name__ == "__main__":
    main()
