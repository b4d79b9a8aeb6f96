// Micro-benchmark + layout check: tcgen05.mma with the A operand in TENSOR MEMORY (TS form) vs in shared memory (SS form)
// for the shapes of the recurrent sweeps (kind::f16, N = 32, M = 128 / 64).  The recurrent weights are constant over
// a sweep: if they can live in TMEM the per-step MMA chain no longer re-reads 4 KB of A from shared memory per
// instruction.  Prints cycles per MMA for both forms and the maximum difference of the two accumulators (and against
// a host reference), for the guessed TMEM layout: row m -> lane m (M = 128) or lane 32*(m/16) + m%16 (M = 64),
// K element k -> 32-bit column k/2, half k%2.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_ts_microbench tools/mma_ts_microbench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../deepspeech.pytorch_b200/csrc/tc_common.cuh"

using namespace ds2::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// warp-converged issue (CUTLASS style): every lane executes the sequence, elect.sync picks the one that issues.  With
// warp-uniform operands the descriptors stay in uniform registers: no R2UR / ELECT waterfall per instruction.
__device__ __forceinline__ void mma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

constexpr int N = 32, KCH = 8;            // K = 8 chunks x 64 = 512
constexpr int K = KCH * 64;
constexpr int D_COL = 0, A_COL = 64;      // TMEM columns: accumulator at 0 (32 cols), A operand from column 64 (K/2 = 256 cols)

// A: [M][K] halfs, B: [N][K] halfs (row-major, K contiguous).  out: [2 modes][M][N] floats, cyc: [2 modes][2]
template <int M>
__global__ void __launch_bounds__(192, 1) bench(const __half* __restrict__ A, const __half* __restrict__ B, float* out,
                                                long long* cyc, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  __shared__ long long tstart[2];
  constexpr int A_BYTES = M * 128, B_BYTES = N * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + KCH * A_BYTES;
  // K-major SW128 tiles written by threads (generic proxy) + proxy fence
  for (int i = threadIdx.x; i < M * K / 8; i += blockDim.x) {
    const int row = i / (K / 8), p = i % (K / 8), c = p / 8, j = p % 8;
    *reinterpret_cast<uint4*>(sa + c * A_BYTES + row * 128 + ((j ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(A + (size_t)row * K + p * 8);
  }
  for (int i = threadIdx.x; i < N * K / 8; i += blockDim.x) {
    const int row = i / (K / 8), p = i % (K / 8), c = p / 8, j = p % 8;
    *reinterpret_cast<uint4*>(sb + c * B_BYTES + row * 128 + ((j ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(B + (size_t)row * K + p * 8);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 2); fence_barrier_init(); }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 1) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  // A into TMEM: warp (2..5) owns TMEM lanes 32q..32q+31, q = warp % 4
  if (warp >= 2) {
    const int q = warp % 4;
    int row;                                     // matrix row held by TMEM lane 32q + lane
    bool used = true;
    if (M == 128) row = 32 * q + lane;
    else { row = 16 * q + lane; used = lane < 16; }
    for (int c8 = 0; c8 < K / 2; c8 += 8) {      // 8 columns = 16 k values per instruction
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        r[j] = used ? *reinterpret_cast<const uint32_t*>(A + (size_t)row * K + 2 * (c8 + j)) : 0u;
      tmem_st8(tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_COL + c8), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = instr_desc(FMT_F16, M, N);
  const uint64_t a_base = smem_desc_sw128(smem_u32(sa)), b_base = smem_desc_sw128(smem_u32(sb));
  uint32_t phase = 0;
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 2) {
      if (warp == 1) {
        long long t0 = 0, t1 = 0, t2 = 0;
        for (int rep = 0; rep < reps; ++rep) {
          t0 = clock64();
#pragma unroll
          for (int c = 0; c < KCH; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_elect(0u + D_COL, a_base + (uint64_t)c * (A_BYTES >> 4) + 2 * k,
                            b_base + (uint64_t)c * (B_BYTES >> 4) + 2 * k, idesc, (c | k) != 0);
          }
          t1 = clock64();
          mma_commit_elect(&bar);
          mbar_wait(&bar, phase);
          phase ^= 1;
          t2 = clock64();
        }
        if (lane == 0) { cyc[4] = t1 - t0; cyc[5] = t2 - t0; }
      }
    } else
    if (threadIdx.x == 32) {
      long long t0 = 0, t1 = 0, t2 = 0;
      for (int rep = 0; rep < reps; ++rep) {
        // fully unrolled issue sequences (what the sweeps do): the single issuing thread must not be the limit
        if (mode == 0) {
          t0 = clock64();
#pragma unroll
          for (int c = 0; c < KCH; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16(0u + D_COL, a_base + (uint64_t)c * (A_BYTES >> 4) + 2 * k, b_base + (uint64_t)c * (B_BYTES >> 4) + 2 * k,
                      idesc, (c | k) != 0);
          }
          t1 = clock64();
        } else {
          t0 = clock64();
#pragma unroll
          for (int c = 0; c < KCH; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ts(0u + D_COL, (uint32_t)(A_COL + c * 32 + k * 8), b_base + (uint64_t)c * (B_BYTES >> 4) + 2 * k, idesc,
                         (c | k) != 0);
          }
          t1 = clock64();
        }
        mma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        t2 = clock64();
      }
      cyc[mode * 2] = t1 - t0;
      cyc[mode * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 2) {
      const int q = warp % 4;
      float v[32];
      tmem_ld32(tb + ((uint32_t)(q * 32) << 16) + D_COL, v);
      int row = M == 128 ? 32 * q + lane : 16 * q + lane;
      if (M == 128 || lane < 16)
        for (int n = 0; n < N; ++n) out[((size_t)mode * M + row) * N + n] = v[n];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // mode 3: TWO warps issue concurrently, each half of the K chunks into its own accumulator (columns 0 / 32)
  {
    uint32_t ph2 = 0;
    for (int rep = 0; rep < reps; ++rep) {
      __syncthreads();
      if (warp < 2) {
        const long long t0 = clock64();
#pragma unroll
        for (int c = 0; c < KCH / 2; ++c) {
          const int cc = warp * (KCH / 2) + c;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_elect((uint32_t)(warp * 32), a_base + (uint64_t)cc * (A_BYTES >> 4) + 2 * k,
                          b_base + (uint64_t)cc * (B_BYTES >> 4) + 2 * k, idesc, (c | k) != 0);
        }
        const long long t1 = clock64();
        mma_commit_elect(&bar2);
        mbar_wait(&bar2, ph2);
        ph2 ^= 1;
        const long long t2 = clock64();
        if (lane == 0 && warp == 1) { cyc[6] = t1 - t0; cyc[7] = t2 - t0; }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 2) {
      const int q = warp % 4;
      float v0[32], v1[32];
      tmem_ld32(tb + ((uint32_t)(q * 32) << 16) + 0, v0);
      tmem_ld32(tb + ((uint32_t)(q * 32) << 16) + 32, v1);
      int row = M == 128 ? 32 * q + lane : 16 * q + lane;
      if (M == 128 || lane < 16)
        for (int n = 0; n < N; ++n) out[((size_t)3 * M + row) * N + n] = v0[n] + v1[n];
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 1) tmem_dealloc<512>(tb);
}

template <int M>
static void run() {
  std::vector<__half> hA((size_t)M * K), hB((size_t)N * K);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) hA[(size_t)m * K + k] = __float2half((float)(((m * 7 + k * 3) % 13) - 6) * 0.125f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hB[(size_t)n * K + k] = __float2half((float)(((n * 5 + k) % 11) - 5) * 0.25f);
  std::vector<float> ref((size_t)M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)__half2float(hA[(size_t)m * K + k]) * (double)__half2float(hB[(size_t)n * K + k]);
      ref[(size_t)m * N + n] = (float)s;
    }
  __half *dA, *dB;
  float* dout;
  long long* dcyc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dout, 4 * M * N * 4); cudaMalloc(&dcyc, 8 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 4 * M * N * 4);
  const int smem = 1024 + KCH * (M * 128 + N * 128);
  cudaFuncSetAttribute(bench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<M><<<1, 192, smem>>>(dA, dB, dout, dcyc, 4);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d: %s\n", M, cudaGetErrorString(e)); exit(1); }
  std::vector<float> out((size_t)4 * M * N);
  long long cyc[8];
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dcyc, 64, cudaMemcpyDeviceToHost);
  double e_ss = 0, e_ts = 0, e_x = 0, e_el = 0;
  for (size_t i = 0; i < (size_t)M * N; ++i) e_el = fmax(e_el, fabs(out[(size_t)2 * M * N + i] - ref[i]));
  for (size_t i = 0; i < (size_t)M * N; ++i) {
    e_ss = fmax(e_ss, fabs(out[i] - ref[i]));
    e_ts = fmax(e_ts, fabs(out[(size_t)M * N + i] - ref[i]));
    e_x = fmax(e_x, fabs(out[i] - out[(size_t)M * N + i]));
  }
  const int nm = KCH * 4;
  printf("M=%3d N=%d K=%d (%d MMAs): SS issue %5.1f complete %5.1f cyc/mma | TS issue %5.1f complete %5.1f cyc/mma | "
         "max|SS-ref| %.3g  max|TS-ref| %.3g  max|SS-TS| %.3g\n", M, N, K, nm, (double)cyc[0] / nm, (double)cyc[1] / nm,
         (double)cyc[2] / nm, (double)cyc[3] / nm, e_ss, e_ts, e_x);
  printf("        warp-converged elect issue (SS): issue %5.1f complete %5.1f cyc/mma, max|ref diff| %.3g\n", (double)cyc[4] / nm,
         (double)cyc[5] / nm, e_el);
  double e_2 = 0;
  for (size_t i = 0; i < (size_t)M * N; ++i) e_2 = fmax(e_2, fabs(out[(size_t)3 * M * N + i] - ref[i]));
  printf("        two issuing warps, two accumulators: issue %6.0f complete %6.0f cycles for all %d MMAs (one warp: %6.0f), max|ref diff| %.3g\n",
         (double)cyc[6], (double)cyc[7], nm, (double)cyc[5], e_2);
}

int main() {
  run<128>();
  run<64>();
  return 0;
}
