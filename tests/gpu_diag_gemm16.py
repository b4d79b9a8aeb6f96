"""Not a test: times the fp16-operand tcgen05 GEMM (ds2_gemm_f16) on the recurrent stack's shapes for every tile
configuration (DS2_GEMM16_CFG = 1: 128x256x4 stages, 2: 256x256x3 stages, 3: 128x256x2 stages with 2 CTAs per SM)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepspeech_pytorch_b200 as ds  # noqa: E402


def gemm16(a16, b16, out, alpha=1.0, beta=0.0):
    lib = ds.get_lib()
    M, K = a16.shape
    N = b16.shape[0]
    rc = lib.ds2_gemm_f16(M, N, K, float(alpha), C.c_void_p(a16.data_ptr()), a16.stride(0), C.c_void_p(b16.data_ptr()),
                          b16.stride(0), float(beta), C.c_void_p(out.data_ptr()), out.stride(0),
                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(lib.ds2_last_error().decode())
    return out


def main():
    shapes = [("projection", 16000, 8192, 1024), ("projection layer 0", 16000, 8192, 1312), ("dX", 16000, 1024, 8192),
              ("dW_ih / dW_hh", 4096, 1024, 16000), ("uni-GRU projection", 16000, 3072, 1024)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, M, N, K in shapes:
        a = (torch.randn(M, K, generator=g, device="cuda") * 0.5).half()
        b = (torch.randn(N, K, generator=g, device="cuda") * 0.5).half()
        out = torch.empty(M, N, device="cuda")
        ref = None
        for cfg in ("0", "1", "2", "3"):
            if cfg == "0":
                os.environ.pop("DS2_GEMM16_CFG", None)
            else:
                os.environ["DS2_GEMM16_CFG"] = cfg
            for _ in range(3):
                gemm16(a, b, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                gemm16(a, b, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            if ref is None:
                ref = (a[:512].double() @ b.double().t())
            err = float((out[:512].double() - ref).norm() / ref.norm())
            print(f"{name:22s} {M}x{N}x{K} cfg {cfg}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  rel-L2 {err:.1e}", flush=True)
    os.environ.pop("DS2_GEMM16_CFG", None)


if __name__ == "__main__":
    main()
