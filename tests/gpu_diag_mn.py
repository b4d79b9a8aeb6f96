"""Not a test: bring-up of MN-major TF32 operands in the tcgen05 GEMM (DS2_GEMM_MN_MAJOR / DS2_GEMM_MN_CFG)."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepspeech_pytorch_b200 as ds  # noqa: E402
from gpu_diag_tc import one  # noqa: E402


def timing(tA, tB, M, N, K, iters=10):
    a = torch.randn((K, M) if tA else (M, K), device="cuda")
    b = torch.randn((N, K) if tB else (K, N), device="cuda")
    ds.set_precision("tf32")
    ds.ops.gemm(a, b, bool(tA), bool(tB))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ds.ops.gemm(a, b, bool(tA), bool(tB))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"gemm tA={tA} tB={tB} {M}x{N}x{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s (incl. transposes if any)",
          flush=True)


def pair_kernel():
    """CTA-pair kernel (B tile multicast inside 2-CTA clusters), DS2_GEMM_CFG=4"""
    os.environ["DS2_GEMM_MN_MAJOR"] = "1"
    os.environ["DS2_GEMM_CFG"] = "4"
    print("=== DS2_GEMM_CFG 4 (256x256 tiles, 2-CTA clusters sharing the B tile)", flush=True)
    for args in [(0, 1, 512, 512, 256), (0, 1, 600, 520, 300), (1, 0, 512, 320, 2000), (0, 0, 640, 1312, 512),
                 (1, 1, 4096, 1024, 4100)]:
        try:
            one(*args)
        except Exception:
            print("[EXC]", args, traceback.format_exc(), flush=True)
    timing(1, 0, 4096, 1024, 16000)
    timing(0, 0, 16000, 1024, 4096)
    timing(0, 1, 16000, 4096, 1024)
    timing(0, 1, 16000, 8192, 1312)
    del os.environ["DS2_GEMM_CFG"]


def main():
    if os.environ.get("DS2_DIAG_PAIR") == "only":
        return pair_kernel()
    # tma_swizzle (3 = 128B, 4 = 128B_ATOM_32B), descriptor layout type, LBO, SBO
    for cfg in ("4,1,4096,512",):
        os.environ["DS2_GEMM_MN_MAJOR"] = "1"
        os.environ["DS2_GEMM_MN_CFG"] = cfg
        print("=== DS2_GEMM_MN_CFG", cfg, flush=True)
        for args in [(1, 1, 128, 256, 32), (0, 0, 128, 256, 32), (1, 0, 512, 320, 2000), (0, 0, 640, 1312, 512),
                     (1, 1, 160, 96, 96)]:
            try:
                one(*args)
            except Exception:
                print("[EXC]", args, traceback.format_exc(), flush=True)
    del os.environ["DS2_GEMM_MN_CFG"]
    os.environ["DS2_GEMM_MN_MAJOR"] = "1"
    for args in [(1, 0, 4096, 1024, 16000), (0, 0, 16000, 1024, 4096), (0, 1, 16000, 4096, 1024), (1, 0, 4064, 1000, 15990),
                 (0, 1, 1000, 520, 300), (1, 1, 4096, 1024, 4100)]:
        try:
            one(*args)
        except Exception:
            print("[EXC]", args, traceback.format_exc(), flush=True)
    for cfg in ("1", "2", "3", "0"):
        os.environ["DS2_GEMM_CFG"] = cfg
        print("=== timing DS2_GEMM_CFG", cfg, "(1: 128x256x4st, 2: 256x256x3st, 3: 128x256x2st 2 CTA/SM, 0: auto)", flush=True)
        timing(1, 0, 4096, 1024, 16000)    # dW = dG^T . X
        timing(0, 0, 16000, 1024, 4096)    # dX = dG . W
        timing(0, 1, 16000, 4096, 1024)    # projection
        timing(0, 1, 16000, 8192, 1312)
    del os.environ["DS2_GEMM_CFG"]
    if os.environ.get("DS2_DIAG_PAIR"):
        pair_kernel()

if __name__ == "__main__":
    main()
