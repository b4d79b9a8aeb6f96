// Micro-benchmark of the flag-in-data ("LL") exchange used by the recurrent sweeps: 128 persistent CTAs (one per SM),
// each step every CTA (a) stores its 16 units x 32 batch rows of fp16 "h" (64 packets of 16 bytes), (b) fetches the
// 512-unit K half it multiplies (32 rows x 1 KB = 2048 packets, produced by 32 other CTAs) into shared memory, polling
// every 2-byte element against the 0xFFFF sentinel, (c) burns `work` cycles (stand-in for MMA + gates).
// Prints the median step period for several ways of issuing the loads.  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ll_microbench tools/ll_microbench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int H = 1024, B = 32, UT = 16, NCTA_DIR = 64, THREADS = 192;

enum { M_RELAXED = 0, M_VOLATILE = 1, M_CG = 2, M_CV = 3, M_CA = 4, M_CPASYNC = 5, M_RELAXED_2X = 6, M_NC = 7,
       M_BULK = 8, M_BARRIER_BULK = 9 };

template <int MODE>
__device__ __forceinline__ uint4 ld16(const void* p) {
  uint4 v;
  if (MODE == M_RELAXED || MODE == M_RELAXED_2X)
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else if (MODE == M_VOLATILE)
    asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else if (MODE == M_CG)
    asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else if (MODE == M_CV)
    asm volatile("ld.global.cv.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else if (MODE == M_NC)
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else
    asm volatile("ld.global.ca.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool ready(const uint4& v) {
  return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, unsigned a, unsigned b) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// h16: [2 dirs][T][B][H] halfs, pre-filled with 0xFFFF.  stats: per CTA per step clock of "all fetched".
template <int MODE, int NLOAD>   // NLOAD: loader threads (128 = the 4 epilogue warps, 64, ...)
__global__ void __launch_bounds__(THREADS, 1) pingpong(__half* h16, int T, int work, long long* stamps, int* err,
                                                       unsigned* ctr, long long* retries) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  unsigned par[4] = {0u, 0u, 0u, 0u};
  long long retries_total = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int c = blockIdx.x, d = c / NCTA_DIR, i = c % NCTA_DIR, rank = i & 1;
  const int u0 = (i / 2) * 32 + rank * UT;          // the 16 units this CTA produces
  const int e = threadIdx.x - 64;                   // epilogue / loader threads 0..127
  __half* base = h16 + (size_t)d * T * B * H;
  for (int step = 0; step < T; ++step) {
    if (e >= 0) {
      // (a) produce: thread = 4 units x one row
      const int uq = 4 * (e & 3), b = e >> 2;
      const __half2 v = __floats2half2_rn(0.001f * (step + 1), 0.5f);
      const unsigned bits = *reinterpret_cast<const unsigned*>(&v);
      st_relaxed_v2(base + ((size_t)step * B + b) * H + u0 + uq, bits, bits);
      // (b) fetch the K half `rank` of h[step]: 32 rows x 64 packets
      if (e < NLOAD) {
        const __half* src = base + (size_t)step * B * H + rank * 512;
        constexpr int PER = 2048 / NLOAD;
        if (MODE == M_BULK || MODE == M_BARRIER_BULK) {
          // transport = bulk async copies (TMA engine, 1 KB per batch row), issued by 8 lanes of each of the 4 warps
          // into that warp's quarter of the tile; M_BULK: no barrier, every 2-byte element checked in shared
          // memory, the warp re-issues its quarter until it is complete.  M_BARRIER_BULK: round-1 protocol
          // (bar.sync -> proxy fence -> red.release -> poll -> acquire -> one issue, no check).
          const int q = e >> 5, lane = e & 31;
          uint64_t* mb = bars + q;
          if (MODE == M_BARRIER_BULK) {
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (e == 0) {
              asm volatile("fence.proxy.async.global;" ::: "memory");
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + d) : "memory");
            }
            if (lane == 0) {
              const unsigned target = (unsigned)NCTA_DIR * (step + 1);
              unsigned v = 0;
              int guard = 0;
              do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + d) : "memory"); } while (v < target && ++guard < 10000000);
              asm volatile("fence.acquire.gpu;" ::: "memory");
              asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            __syncwarp();
          }
          bool ok = false;
          int guard = 0;
          while (!ok && ++guard < 100000) {
            if (lane == 0) mbar_expect(mb, 8 * 1024);
            __syncwarp();
            if (lane < 8) bulk_load(smem + (8 * q + lane) * 1024, src + (size_t)(8 * q + lane) * H, 1024, mb);
            mbar_wait(mb, par[q] & 1u);
            par[q] ^= 1u;
            bool mine = true;
            if (MODE == M_BULK) {
#pragma unroll
              for (int k = 0; k < 16; ++k) mine = mine && ready(*reinterpret_cast<const uint4*>(smem + q * 8192 + (k * 32 + lane) * 16));
            }
            ok = __all_sync(0xffffffffu, mine);
          }
          if (!ok) *err = 1;
          if (lane == 0 && e == 0) retries_total += guard - 1;
        } else if (MODE == M_CPASYNC) {
          bool pending = true;
          int guard = 0;
          unsigned need[(PER + 31) / 32];
          for (int w = 0; w < (PER + 31) / 32; ++w) need[w] = 0xFFFFFFFFu;
          while (pending && ++guard < 100000) {
            for (int k = 0; k < PER; ++k) {
              if (!((need[k / 32] >> (k % 32)) & 1u)) continue;
              const int idx = k * NLOAD + e, row = idx >> 6, pc = idx & 63;
              const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (pc >> 3) * 4096 + row * 128 + (((pc & 7) ^ (row & 7)) << 4));
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)row * H + pc * 8) : "memory");
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            pending = false;
            for (int k = 0; k < PER; ++k) {
              if (!((need[k / 32] >> (k % 32)) & 1u)) continue;
              const int idx = k * NLOAD + e, row = idx >> 6, pc = idx & 63;
              const uint4 v = *reinterpret_cast<const uint4*>(smem + (pc >> 3) * 4096 + row * 128 + (((pc & 7) ^ (row & 7)) << 4));
              if (ready(v)) need[k / 32] &= ~(1u << (k % 32)); else pending = true;
            }
          }
          if (pending) *err = 1;
        } else {
          constexpr int CH = (MODE == M_RELAXED_2X) ? 16 : 8;      // loads in flight per thread
          for (int k0 = 0; k0 < PER; k0 += CH) {
            uint4 v[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
              const int idx = (k0 + k) * NLOAD + e, row = idx >> 6, pc = idx & 63;
              v[k] = ld16<MODE>(src + (size_t)row * H + pc * 8);
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
              const int idx = (k0 + k) * NLOAD + e, row = idx >> 6, pc = idx & 63;
              int guard = 0;
              while (!ready(v[k]) && ++guard < 100000) v[k] = ld16<MODE>(src + (size_t)row * H + pc * 8);
              if (guard >= 100000) *err = 1;
              *reinterpret_cast<uint4*>(smem + (pc >> 3) * 4096 + row * 128 + (((pc & 7) ^ (row & 7)) << 4)) = v[k];
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (e == 0) stamps[(size_t)c * T + step] = clock64();
      // (c) the rest of the step
      const long long t0 = clock64();
      while (clock64() - t0 < work) {
      }
    }
  }
  if (threadIdx.x == 64) retries[blockIdx.x] = retries_total;
}

template <int MODE, int NLOAD>
static void run(const char* name, int T, int work) {
  __half* h16;
  const size_t n = (size_t)2 * T * B * H;
  CK(cudaMalloc(&h16, n * 2));
  CK(cudaMemset(h16, 0xFF, n * 2));
  long long* stamps;
  int* err;
  CK(cudaMalloc(&stamps, sizeof(long long) * 128 * T));
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  unsigned* ctr;
  long long* retries;
  CK(cudaMalloc(&ctr, 64));
  CK(cudaMemset(ctr, 0, 64));
  CK(cudaMalloc(&retries, 8 * 128));
  CK(cudaMemset(retries, 0, 8 * 128));
  auto kern = pingpong<MODE, NLOAD>;
  const int smem = 120 * 1024;   // one CTA per SM
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* args[] = {&h16, &T, &work, &stamps, &err, &ctr, &retries};
  CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(128), dim3(THREADS), args, smem, 0));
  CK(cudaDeviceSynchronize());
  std::vector<long long> s((size_t)128 * T);
  int herr = 0;
  CK(cudaMemcpy(s.data(), stamps, s.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  std::vector<long long> per;
  for (int c = 0; c < 128; c += 7)
    for (int t = 20; t + 1 < T; ++t) per.push_back(s[(size_t)c * T + t + 1] - s[(size_t)c * T + t]);
  std::sort(per.begin(), per.end());
  std::vector<long long> rt(128);
  CK(cudaMemcpy(rt.data(), retries, 8 * 128, cudaMemcpyDeviceToHost));
  long long rsum = 0;
  for (auto v : rt) rsum += v;
  printf("%-34s loaders %3d  work %5d : step period median %6lld  p10 %6lld  p90 %6lld  -> exchange ~%lld cycles  (warp-0 re-issues per step %.2f)%s\n",
         name, NLOAD, work, per[per.size() / 2], per[per.size() / 10], per[per.size() * 9 / 10], per[per.size() / 2] - work,
         (double)rsum / 128.0 / T, herr ? "  [TIMEOUT]" : "");
  CK(cudaFree(ctr)); CK(cudaFree(retries));
  CK(cudaFree(h16)); CK(cudaFree(stamps)); CK(cudaFree(err));
}

int main() {
  const int T = 300;
  for (int work : {0, 2000, 4000, 6000}) {
    run<M_BARRIER_BULK, 128>("barrier + bulk copy (round 1)", T, work);
    run<M_BULK, 128>("bulk copy + smem sentinel check", T, work);
    run<M_CPASYNC, 128>("cp.async.cg 16B -> smem, poll smem", T, work);
    if (work == 4000) run<M_RELAXED, 128>("ld.relaxed.gpu.v4 (8 in flight)", T, work);
  }
  return 0;
}
