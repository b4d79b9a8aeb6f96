"""Not a test: tcgen05 GEMM bring-up diagnostics (errors vs fp64, vs cuBLAS TF32 as a yardstick, timing)."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepspeech_pytorch_b200 as ds  # noqa: E402


def err(c, ref):
    d = (c.double() - ref)
    return (float(d.abs().max() / ref.abs().max()), float(d.mean() / ref.abs().mean()),
            float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))


def one(tA, tB, M, N, K, seed=0, positive=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn((K, M) if tA else (M, K), generator=g, device="cuda")
    b = torch.randn((N, K) if tB else (K, N), generator=g, device="cuda")
    if positive:
        a, b = a.abs(), b.abs()
    opa, opb = (a.t() if tA else a), (b.t() if tB else b)
    ref = opa.double() @ opb.double()
    ds.set_precision("tf32")
    c = ds.ops.gemm(a, b, bool(tA), bool(tB))
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = True
    cb = opa @ opb
    torch.backends.cuda.matmul.allow_tf32 = False
    e, eb = err(c, ref), err(cb, ref)
    print(f"tA={tA} tB={tB} M={M} N={N} K={K} pos={positive}: ds2 max={e[0]:.2e} bias={e[1]:+.2e} rms={e[2]:.2e} | "
          f"cublas-tf32 max={eb[0]:.2e} bias={eb[1]:+.2e} rms={eb[2]:.2e}", flush=True)


def timing(M, N, K, iters=10):
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for prec in ("tf32", "fp32"):
        ds.set_precision(prec)
        ds.ops.gemm(a, b, False, True, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            ds.ops.gemm(a, b, False, True, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"gemm {prec} {M}x{N}x{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.matmul(a, b.t(), out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    torch.backends.cuda.matmul.allow_tf32 = False
    print(f"gemm cublas-tf32 {M}x{N}x{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


def main():
    print("DS2_TMAP_TF32 =", os.environ.get("DS2_TMAP_TF32"))
    for args in [(0, 1, 128, 256, 32), (0, 1, 128, 256, 64), (0, 1, 256, 512, 256), (0, 1, 300, 200, 132),
                 (0, 1, 1000, 4096, 1312), (1, 0, 512, 320, 2000), (0, 0, 640, 1312, 512), (1, 1, 160, 96, 96)]:
        try:
            one(*args)
        except Exception:
            print("[EXC]", args, traceback.format_exc(), flush=True)
    one(0, 1, 512, 512, 1024, positive=True)
    try:
        timing(16000, 8192, 1312)
        timing(4096, 1312, 16000)
    except Exception:
        print("[EXC] timing", traceback.format_exc(), flush=True)


if __name__ == "__main__":
    main()
