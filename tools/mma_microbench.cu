// Micro-benchmark: cycles per tcgen05.mma for small shapes (what bounds the recurrent step).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench mma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../deepspeech.pytorch_b200/csrc/tc_common.cuh"

using namespace ds2::tc;

// smem: A tile 128 rows x 128 B (16 KB) x KCH chunks, B tile 256 rows x 128 B x KCH
template <int KIND>  // 0 tf32, 1 f16
__global__ void bench(int M, int N, int nmma, int nacc, int acc_stride, int kch, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 * kch / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb = slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = instr_desc(KIND == 0 ? FMT_TF32 : FMT_F16, M, N);
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      int ch = (i / 4) % kch;
      uint64_t ad = smem_desc_sw128(smem_u32(smem + ch * 49152)) + (uint64_t)((i % 4) * 2);
      uint64_t bd = smem_desc_sw128(smem_u32(smem + ch * 49152 + 16384)) + (uint64_t)((i % 4) * 2);
      uint32_t d = tb + (uint32_t)((i % nacc) * acc_stride);
      if (KIND == 0) mma_tf32(d, ad, bd, idesc, i >= nacc);
      else mma_f16(d, ad, bd, idesc, i >= nacc);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

// tight issue loop: descriptors precomputed, 4 MMAs unrolled per "chunk", optional mbarrier wait+commit per chunk
template <int KIND>
__global__ void bench_tight(int M, int N, int nchunks, int with_commit, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, ebar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 * 4 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&ebar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = instr_desc(KIND == 0 ? FMT_TF32 : FMT_F16, M, N);
    const uint64_t a0 = smem_desc_sw128(smem_u32(smem)), b0 = smem_desc_sw128(smem_u32(smem + 16384));
    const uint64_t stage_step = 49152 >> 4;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 3;
      const uint64_t ad = a0 + (uint64_t)s * stage_step, bd = b0 + (uint64_t)s * stage_step;
      if (KIND == 0) {
        mma_tf32(tb, ad, bd, idesc, acc); acc = 1;
        mma_tf32(tb, ad + 2, bd + 2, idesc, 1);
        mma_tf32(tb, ad + 4, bd + 4, idesc, 1);
        mma_tf32(tb, ad + 6, bd + 6, idesc, 1);
      } else {
        mma_f16(tb, ad, bd, idesc, acc); acc = 1;
        mma_f16(tb, ad + 2, bd + 2, idesc, 1);
        mma_f16(tb, ad + 4, bd + 4, idesc, 1);
        mma_f16(tb, ad + 6, bd + 6, idesc, 1);
      }
      if (with_commit) mma_commit(&ebar[s]);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

int main() {
  {
    long long* d;
    cudaMalloc(&d, 16);
    int smem = 1024 + 49152 * 4;
    cudaFuncSetAttribute(bench_tight<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench_tight<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int kind = 0; kind < 2; ++kind)
      for (int M : {64, 128})
        for (int N : {16, 32, 64, 128, 256})
          for (int wc : {0, 1}) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
              if (kind == 0) bench_tight<0><<<1, 128, smem>>>(M, N, 64, wc, d);
              else bench_tight<1><<<1, 128, smem>>>(M, N, 64, wc, d);
              if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("TIGHT %s M=%3d N=%3d commit/chunk=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma\n",
                   kind == 0 ? "tf32" : "f16 ", M, N, wc, (double)h[0] / 256, (double)h[1] / 256);
          }
  }
  return 0;
}
int main_old() {
  long long* d;
  cudaMalloc(&d, 16);
  int smem = 1024 + 49152 * 4;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nmma = 256;
  int Ms[] = {64, 128};
  int Ns[] = {8, 16, 32, 64, 128, 256};
  for (int kind = 0; kind < 2; ++kind)
    for (int M : Ms)
      for (int N : Ns) {
        if (M == 128 && N % 16) continue;
        for (int nacc : {1, 2, 4}) {
          if (nacc * (N < 32 ? 32 : N) > 512) continue;
          for (int kch : {1, 4}) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
              if (kind == 0) bench<0><<<1, 128, smem>>>(M, N, nmma, nacc, N < 32 ? 32 : N, kch, d);
              else bench<1><<<1, 128, smem>>>(M, N, nmma, nacc, N < 32 ? 32 : N, kch, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s M=%3d N=%3d nacc=%d kchunks=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma\n",
                   kind == 0 ? "tf32" : "f16 ", M, N, nacc, kch, (double)h[0] / nmma, (double)h[1] / nmma);
          }
        }
      }
  return 0;
}
