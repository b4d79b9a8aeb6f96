// Persistent, warp-specialised recurrent sweeps on tcgen05.
//
// One launch per layer covers ALL time steps and both directions.  A CTA owns 16 hidden units of
// one direction: the G*16 rows of W_hh for those units form the M=64 operand of
//     acc[(g,u), b] = sum_k W_hh[g*H+u, k] * h_{t-1}[b, k]        (tcgen05.mma kind::tf32, N = batch)
// with fp32 accumulators in TMEM.  Per step:
//   warp 0   TMA producer: weight chunks (3-D box over [gate][unit][k], prefetched ahead of the
//            barrier because they do not depend on it) and h_{t-1} chunks (after the grid barrier)
//            into an 8-stage 128B-swizzled shared-memory ring
//   warp 1   single-thread MMA issue, tcgen05.commit -> mbarriers
//   warps 2-5 epilogue: TMEM -> registers, + input projection + biases, gate non-linearities spread
//            over 64 lanes, exchange through shared memory, cell update (c / h state stays in shared
//            memory for the whole sweep), masked stores of h_t (and the saved tensors for backward)
// Steps are separated by a per-direction grid barrier (monotonic counter in global memory, release /
// acquire, bounded spin so that a fault cannot hang the GPU).  All CTAs must be co-resident: the host
// checks occupancy and launches cooperatively; shapes that do not fit return 1 (FFMA step kernels).
//
// Backward sweep: the same skeleton on W_hh^T with split-K over the gates inside a 4-CTA cluster
// (partial sums reduced through distributed shared memory) — see rnn_bwd_persist_kernel.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"
#include "rnn_cells.cuh"
#include "rnn_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace ds2 {

namespace rp {
constexpr int UT = 16;        // hidden units per CTA
constexpr int MM = 64;        // MMA M (G*UT rows, padded for GRU / tanh)
constexpr int BK = 32;        // fp32 elements per 128-byte swizzle row
constexpr int STAGES = 8;
constexpr int THREADS = 192;  // producer, mma, 4 epilogue warps
constexpr int A_BYTES = MM * 128;
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks
}  // namespace rp

struct PersistParams {
  CUtensorMap tmW[2];   // 3-D (k, unit, gate) over W_hh[d]
  CUtensorMap tmV[2];   // 2-D (k, row) over the per-direction vector sequence (hseq[d] as [T*B, H])
  CUtensorMap tmV2[2];  // bwd GRU: the n-gate part of dGh lives in the aux buffer
  CUtensorMap tmV3[2];  // resident: 3-D view (k in chunk, row, chunk) of the streamed fp16 operand, box = 4 chunks
  int box3;             // 1: tmV3 is valid (number of chunks divisible by 4): one TMA instruction per group
  int T, B, NB, H, D, NT, G, training;
  const float* dy;      // bwd: (T,B,H)
  __half* h16;          // resident fwd: fp16 copy of hseq (D,T,B,H), the MMA operand of the next step
  __half* dg16;         // resident bwd: scaled fp16 copy of dGh (T*B, D*G*H)
  unsigned int* gmax;   // resident bwd: [D][T+1] float bits of max|dGh| per processed step (slot 0: bound from dY)
  unsigned int* dymax;  // resident bwd: [T] float bits of max_b,u |dY[t]| (the scale of a step also covers its own dY)
  __half* dgn16;        // resident bwd, optional: fp16 copy of the gate gradients x nscale, (T*B, D*G*H) row-major
  __half* dgn16T;       //   ... transposed (D*G*H, T*B)
  __half* auxn16T;      //   ... GRU h-side n-gate gradient, transposed (D*H, T*B)
  const float* nscale;  //   device: power-of-two scale of those copies
  unsigned int* gmeta;  // LL bwd: [D][T][NT*CL*4] float bits of max|dGh| per (step, epilogue warp), 0xFFFFFFFF = not yet
  long long* trace;     // optional: clock64 stamps of CTA 0, 4 per step
  const int32_t* len;
  float* gates;
  float* hseq;
  float* aux;
  const float* b_ih[2];
  const float* b_hh[2];
  unsigned int* bar;    // [2][NT] per-CTA step flags (zeroed by the host): flag = number of finished steps
  int nacc, acc_cols;   // independent TMEM accumulator chains (K is dealt round-robin over them)
  int defer;            // 1: stores that only later kernels read are issued after the barrier arrival
  int d0;               // first direction handled by this launch (directions can be launched one at a time
                        // when both together would not be co-resident, e.g. H = 1536)
  float* dbias[2];      // split-K bwd: bias-gradient accumulators (G*H per direction, zeroed by the host) or null
  float* dbias_hn[2];   // split-K bwd, GRU: sum of dGh_n (H per direction)
  int* err;             // set to 1 if a barrier wait timed out
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Cross-proxy ordering for data that one CTA writes with ordinary stores and another CTA reads with TMA.  The
// unqualified fence.proxy.async compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC (measured ~950 cycles per call inside
// the sweeps, three of them on the per-step critical path); the .global form is the proxy fence alone, and the
// gpu-scope ordering comes from the release / acquire pair on the barrier counter, which is needed anyway.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Grid barrier of one direction: every CTA publishes the number of steps it has finished in its own
// 4-byte flag (plain release store, no atomic serialisation); a whole warp polls all NT flags
// (coalesced acquire loads) until each is >= target.  Bounded spin: a fault cannot hang the GPU.
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_wait_flags(const unsigned int* flags, int nt, unsigned int target, int* err) {
  const int lane = threadIdx.x % 32;
  const long long t0 = clock64();
  unsigned int it = 0;
  for (;;) {
    // up to 4 independent loads in flight per lane (nt <= 128 flags per direction), relaxed polling
    unsigned int v0 = lane < nt ? ld_relaxed(flags + lane) : target;
    unsigned int v1 = lane + 32 < nt ? ld_relaxed(flags + lane + 32) : target;
    unsigned int v2 = lane + 64 < nt ? ld_relaxed(flags + lane + 64) : target;
    unsigned int v3 = lane + 96 < nt ? ld_relaxed(flags + lane + 96) : target;
    bool ok = v0 >= target && v1 >= target && v2 >= target && v3 >= target;
    for (int i = lane + 128; i < nt; i += 32) ok = ok && (ld_relaxed(flags + i) >= target);
    if (__all_sync(0xffffffffu, ok)) break;
    if ((++it & 63u) == 0) {
      if (*(volatile int*)err) return;
      if (clock64() - t0 > rp::SPIN_LIMIT) {
        *(volatile int*)err = 1;
        if (lane == 0) printf("ds2: recurrent sweep barrier timeout (block %d, target %u)\n", blockIdx.x, target);
        return;
      }
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire side of the flag protocol
}
// single-thread variant on a monotonically increasing arrival counter
// (relaxed polling loads, then one acquire fence: an acquire load would invalidate L1 on every poll)
__device__ __forceinline__ void grid_wait_counter(const unsigned int* ctr, unsigned int target, int* err) {
  if (ld_relaxed(ctr) < target) {
    const long long t0 = clock64();
    unsigned int it = 0;
    while (ld_relaxed(ctr) < target) {
      if ((++it & 255u) == 0) {
        if (*(volatile int*)err) break;
        if (clock64() - t0 > rp::SPIN_LIMIT) {
          *(volatile int*)err = 1;
          printf("ds2: recurrent sweep barrier timeout (block %d, target %u, have %u)\n", blockIdx.x, target,
                 ld_relaxed(ctr));
          break;
        }
      }
    }
  }
  asm volatile("fence.acquire.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// MUFU.EX2 / MUFU.RCP without the denormal-range fix-up code of the CUDA fast-math intrinsics (that code
// serialises independent chains through one predicate register); inputs are gate pre-activations.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.f + ex2_ftz(-LOG2E * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, rcp_ftz(1.f + ex2_ftz(-2.f * LOG2E * x)), -1.f); }

// debug trace: 16 clock64 slots per (CTA, step); slot 12 holds %globaltimer (ns) instead
constexpr int TRACE_SLOTS = 16;
__device__ __forceinline__ void trace_stamp(long long* trace, int T, int step, int slot) {
  if (trace) trace[((size_t)blockIdx.x * T + step) * TRACE_SLOTS + slot] = clock64();
}
__device__ __forceinline__ void trace_stamp_ns(long long* trace, int T, int step, int slot) {
  if (trace) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    trace[((size_t)blockIdx.x * T + step) * TRACE_SLOTS + slot] = (long long)ns;
  }
}

// Resident variants: the streamed operand arrives in groups of 4 K chunks (256 k), one mbarrier per group.
// (A single-chunk first group was measured: the MMA chain starts ~250 cycles earlier but the extra
// producer instructions delay the later groups by as much; what limits the start of a step is the ~110
// cycles the single producer thread needs per TMA instruction, hence one 3-D box per group where possible.)
__device__ __forceinline__ int grp_begin(int g) { return 4 * g; }
__device__ __forceinline__ int grp_count(int nkr) { return (nkr + 3) / 4; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- flag-in-data exchange ("LL": the data is its own ready flag) ---------------------------------------------
// The per-step grid barrier costs store -> MEMBAR.ALL.GPU -> RED -> poll -> acquire fence -> TMA round trip, ~3.6k of
// the 8.8k cycles of a forward step.  In the LL variants the streamed fp16 operand buffer is pre-filled with the
// bit pattern 0xFFFF (an fp16 NaN that neither h in (-1,1) nor a saturated gate gradient can produce); producers
// store their values with relaxed gpu-scope stores and NOTHING else, consumers poll the 16-byte packets they need
// with relaxed gpu-scope loads until no 2-byte element equals the sentinel (every element is its own flag, so
// torn 16-byte packets are harmless) and copy them into the 128B-swizzled K-major tile the MMA reads.  No fence, no
// atomic, no barrier counter on the critical path; one L2 round trip from "stored" to "in shared memory".
constexpr unsigned int LL_SENTINEL = 0xFFFFFFFFu;
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, unsigned int a, unsigned int b) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ bool ll_ready(const uint4& v) {
  return (__vcmpeq2(v.x, LL_SENTINEL) | __vcmpeq2(v.y, LL_SENTINEL) | __vcmpeq2(v.z, LL_SENTINEL) |
          __vcmpeq2(v.w, LL_SENTINEL)) == 0u;
}
// poll one packet until it is complete; bounded (a protocol fault sets *err and lets the kernel run to its end)
__device__ __forceinline__ void ll_wait(uint4& v, const void* src, int* err) {
  if (ll_ready(v)) return;
  const long long t0 = clock64();
  unsigned int it = 0;
  do {
    v = ld_relaxed_v4(src);
    if ((++it & 63u) == 0) {
      if (*(volatile int*)err) return;
      if (clock64() - t0 > rp::SPIN_LIMIT) {
        *(volatile int*)err = 1;
        printf("ds2: recurrent sweep data-flag timeout (block %d)\n", blockIdx.x);
        return;
      }
    }
  } while (!ll_ready(v));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void st_relaxed_u32(void* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// poll one 32-bit word until it differs from the sentinel (bounded like ll_wait)
__device__ __forceinline__ unsigned int ll_wait_u32(const unsigned int* src, int* err) {
  unsigned int v = ld_relaxed(src);
  if (v != LL_SENTINEL) return v;
  const long long t0 = clock64();
  unsigned int it = 0;
  do {
    v = ld_relaxed(src);
    if ((++it & 63u) == 0) {
      if (*(volatile int*)err) return 0u;
      if (clock64() - t0 > rp::SPIN_LIMIT) { *(volatile int*)err = 1; return 0u; }
    }
  } while (v == LL_SENTINEL);
  return v;
}


// RES = true: "resident" variant.  W_hh is converted to fp16 once per call and this CTA's slice
// (G*16 rows x H, 128 KB at H=1024) stays in shared memory for the whole sweep; per step only the
// fp16 copy of h_{t-1} (B x H, 64 KB) is streamed by TMA, and the MMAs are kind::f16 (K=16, half the
// instruction count).  fp16 has the same 10-bit mantissa as TF32 and |h| < 1, |w| << 65504, so this
// is the same arithmetic class as the TF32 path (fp32 accumulation in TMEM either way).
template <int RNN, bool RES>
__global__ void __launch_bounds__(rp::THREADS, 1) rnn_fwd_persist_kernel(const __grid_constant__ PersistParams p) {
  using namespace rp;
  using namespace tc;
  constexpr int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still __shared__
  const int NB = p.NB, B = p.B, T = p.T, H = p.H, D = p.D;
  const int B_BYTES = NB * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const int NBp = NB + 1;
  const int NKR = H / 64;                                  // resident: 64 fp16 (128 B) of K per chunk
  const int NG = grp_count(NKR);
  // streaming layout: STAGES x (A | B) ; resident layout: NKR x A (weights) then NKR x B (h chunks)
  const int ring_bytes = RES ? NKR * STAGE_BYTES : STAGES * STAGE_BYTES;
  float* ex = reinterpret_cast<float*>(smem + ring_bytes);            // [4][UT][NBp]
  float* cst = ex + 4 * UT * NBp;                                       // [UT][NBp] cell (LSTM) / hidden (GRU) state
  int* lens_s = reinterpret_cast<int*>(cst + UT * NBp);                 // [NB] (padded to even)
  uint64_t* full = reinterpret_cast<uint64_t*>(lens_s + ((NB + 1) & ~1));   // resident: one per h chunk (<= 32)
  uint64_t* empty = full + (RES ? 32 : STAGES);                             // resident: [0] = weights landed
  uint64_t* accum_bar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int d = p.d0 + blockIdx.x / p.NT, tile = blockIdx.x % p.NT, u0 = tile * UT;
  const int NK = H / BK;
  const int GH = G * H;
  unsigned int* ctr = p.bar + 32 * d;   // one 128-byte line per direction

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[d]);
    tma_prefetch_desc(&p.tmV[d]);
    for (int i = 0; i < (RES ? 32 : STAGES); ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < STAGES; ++i) mbar_init(&empty[i], 1);
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < UT * NBp; i += THREADS) cst[i] = 0.f;
  for (int i = threadIdx.x; i < NB; i += THREADS) lens_s[i] = i < B ? p.len[i] : 0;
  if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The CTA owns all 512 TMEM columns, so the allocation starts at address 0.  Using the literal keeps the
  // accumulator operand of tcgen05.mma in a uniform register without the ELECT / R2UR.BROADCAST waterfall
  // loop nvcc otherwise emits per MMA for a value it cannot prove warp-uniform.
  if (tmem_base != 0 && threadIdx.x == 0) {
    *(volatile int*)p.err = 2;
    printf("ds2: unexpected TMEM base %u (block %d)\n", tmem_base, blockIdx.x);
  }
  const uint32_t tx_bytes = (uint32_t)(G * UT * 128 + B * 128);

  if (warp == 0) {
    if (RES) {
      if (lane == 0) {
        // weights once: NKR chunks of (G*16 rows x 64 fp16) into the resident region
        mbar_arrive_expect_tx(&empty[0], (uint32_t)(NKR * G * UT * 128));
        for (int c = 0; c < NKR; ++c) tma_load_3d(smem + c * A_BYTES, &p.tmW[d], &empty[0], c * 64, u0, 0);
        uint8_t* hbuf = smem + NKR * A_BYTES;
        for (int step = 1; step < T; ++step) {
          const int t = d == 0 ? step : T - 1 - step;
          const int tp = d == 0 ? t - 1 : t + 1;
          // the previous step's MMAs have finished reading hbuf: this CTA's epilogue (which waited for them)
          // arrived at the barrier we are about to pass
          grid_wait_counter(ctr, (unsigned int)p.NT * (unsigned int)step, p.err);
          fence_proxy_async_global();
          trace_stamp(p.trace, p.T, step, 0);
          for (int g = 0; g < NG; ++g) {  // one mbarrier per group of chunks
            uint64_t* fb = full + g;
            const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
            if (p.box3) {
              mbar_arrive_expect_tx(fb, (uint32_t)(4 * B_BYTES));
              tma_load_3d(hbuf + c0 * B_BYTES, &p.tmV3[d], fb, 0, tp * B, c0);
            } else {
              mbar_arrive_expect_tx(fb, (uint32_t)((c1 - c0) * B * 128));
              for (int c = c0; c < c1; ++c) tma_load_2d(hbuf + c * B_BYTES, &p.tmV[d], fb, c * 64, tp * B);
            }
          }
          trace_stamp(p.trace, p.T, step, 1);
        }
      }
    } else
    if (lane == 0) {
      // strict chunk order: slot free -> weight chunk -> (first chunk of a step: grid barrier) -> h chunk.
      // The weight chunk of the next step's first slot is therefore in flight while the barrier is awaited.
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        const int t = d == 0 ? step : T - 1 - step;
        const int tp = d == 0 ? t - 1 : t + 1;
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], tx_bytes);
          tma_load_3d(smem + s * STAGE_BYTES, &p.tmW[d], &full[s], c * BK, u0, 0);
          if (c == 0) {
            grid_wait_counter(ctr, (unsigned int)p.NT * (unsigned int)step, p.err);
            fence_proxy_async_global();
            trace_stamp(p.trace, p.T, step, 0);
          }
          tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &p.tmV[d], &full[s], c * BK, tp * B);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (RES) {
      {   // the whole warp runs the issue loop converged (mma_f16_w: elect.sync inside), see tc_common.cuh
        const uint32_t idesc = instr_desc(FMT_F16, MM, NB);
        const uint64_t a_base = warp_uniform(smem_desc_sw128(smem_u32(smem)));
        const uint64_t b_base = warp_uniform(smem_desc_sw128(smem_u32(smem + NKR * A_BYTES)));
        const uint64_t a_step = (uint64_t)(A_BYTES >> 4), b_step = (uint64_t)(B_BYTES >> 4);
        mbar_wait(&empty[0], 0);                 // weights resident
        uint32_t ph = 0;
        for (int step = 1; step < T; ++step) {
          for (int g = 0; g < NG; ++g) {
            mbar_wait(full + g, ph);
            tc_fence_after();
            if (g == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
            const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
            if (c1 == NKR && lane == 0) trace_stamp(p.trace, p.T, step, 3);
            for (int c = c0; c < c1; ++c) {
              const uint64_t ad = a_base + (uint64_t)c * a_step, bd = b_base + (uint64_t)c * b_step;
              mma_f16_w(0u, ad, bd, idesc, c > 0);
              mma_f16_w(0u, ad + 2, bd + 2, idesc, 1);
              mma_f16_w(0u, ad + 4, bd + 4, idesc, 1);
              mma_f16_w(0u, ad + 6, bd + 6, idesc, 1);
            }
          }
          mma_commit_w(accum_bar);
          if (lane == 0) trace_stamp(p.trace, p.T, step, 4);
          ph ^= 1;
        }
      }
    } else
    {
      // The issue chain is the critical path of a step: the whole warp runs it converged (mma_tf32_w: elect.sync
      // inside; a single divergent lane costs ~70 cycles per tcgen05.mma, converged ~30-40) and the loop is free of
      // div/mod and descriptor construction.
      const uint32_t idesc = instr_desc(FMT_TF32, MM, NB);
      const uint64_t a_base = smem_desc_sw128(smem_u32(smem));
      const uint64_t b_base = smem_desc_sw128(smem_u32(smem + A_BYTES));
      const uint64_t stage_step = (uint64_t)(STAGE_BYTES >> 4);
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&full[s], ph);
          if (c == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 1);
          if (c == NK - 1 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
          tc_fence_after();
          const uint64_t ad = a_base + (uint64_t)s * stage_step, bd = b_base + (uint64_t)s * stage_step;
          mma_tf32_w(0u, ad, bd, idesc, c > 0);
          mma_tf32_w(0u, ad + 2, bd + 2, idesc, 1);
          mma_tf32_w(0u, ad + 4, bd + 4, idesc, 1);
          mma_tf32_w(0u, ad + 6, bd + 6, idesc, 1);
          mma_commit_w(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_w(accum_bar);
      }
    }
  } else {
    // ---------------- epilogue warps: quarter q of TMEM == gate q (rows q*16 .. q*16+15 in lanes 0..15)
    // All 32 lanes work: lane L<16 owns row (q, L); its 32 batch columns are split with lane L+16
    // (columns 16..31 travel by shuffle), so every lane activates 16 values per 32-column chunk.
    const int q = warp % 4;
    const int e = threadIdx.x - 64;          // 0..127
    const int ul = lane & 15, half = lane >> 4;
    const int u = u0 + ul;
    const bool has_row = (q < G) || (RNN == DS2_RNN_GRU && q == 3);
    const int gsel = (RNN == DS2_RNN_GRU && q == 3) ? 2 : q;   // GRU: warp 3 carries x_n + b_in
    float bias_x = 0.f, bias_h = 0.f;
    if (has_row) bias_x = p.b_ih[d][gsel * H + u];
    if (q < G) bias_h = p.b_hh[d][q * H + u];
    uint32_t acc_phase = 0;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? step : T - 1 - step;
      for (int cb = 0; cb < NB; cb += 32) {
        // input-projection values (independent of the recurrence: issued before waiting for the MMAs)
        float gx[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int b = cb + half * 16 + j;
          gx[j] = (has_row && b < B) ? p.gates[(((size_t)t * B + b) * D + d) * GH + (size_t)gsel * H + u] : 0.f;
        }
        float a16[16];
        if (step > 0 && q < G) {
          if (cb == 0) { mbar_wait(accum_bar, acc_phase); tc_fence_after(); if (e == 0) trace_stamp(p.trace, p.T, step, 5); }
          float acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, acc);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float hi = __shfl_sync(0xffffffffu, acc[16 + j], ul);   // row owner's upper columns
            a16[j] = half ? hi : acc[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) a16[j] = 0.f;
        }
        // The gate of a warp is uniform: v = sc * sigmoid(sc * pre) + (1 - sc) is the sigmoid for sc = 1 and tanh
        // for sc = 2, i.e. one EX2 + one RCP per value.  The 16 values of a lane are processed stage by stage so
        // that the special-function latencies overlap (the straightforward per-value form was compiled into two
        // serial chains and cost ~2200 cycles per step on the critical path).
        float v[16];
        const bool act = RNN == DS2_RNN_LSTM || (RNN == DS2_RNN_GRU && q < 2) || (RNN == DS2_RNN_TANH && q == 0);
        if (act) {
          const float sc = ((RNN == DS2_RNN_LSTM && q == 2) || RNN == DS2_RNN_TANH) ? 2.f : 1.f;
          const float bsum = bias_x + bias_h, nsc = -sc * LOG2E, off = 1.f - sc;
          float ev[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) ev[j] = ex2_ftz(nsc * ((gx[j] + a16[j]) + bsum));
#pragma unroll
          for (int j = 0; j < 16; ++j) ev[j] = rcp_ftz(1.f + ev[j]);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(sc, ev[j], off);
        } else if (RNN == DS2_RNN_GRU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = q == 2 ? a16[j] + bias_h      // W_hn h + b_hn
                                                     : gx[j] + bias_x;      // x_n + b_in
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int b = cb + half * 16 + j;
          if (b < B) ex[(q * UT + ul) * NBp + b] = v[j];
        }
      }
      if (step > 0) acc_phase ^= 1;
      tc_fence_before();
      if (e == 0) trace_stamp(p.trace, p.T, step, 6);
      named_bar_sync(1, 128);
      if (e == 0) trace_stamp(p.trace, p.T, step, 7);
      // ---------------- combine: thread e owns the 4 consecutive units 4*(e&3)..+3 of batch column (e>>2) [+32
      // per block], so every global access is one 8- or 16-byte vector, and the four cells are processed stage by
      // stage (independent special-function chains).  Only the fp16 copy of h_t (resident variant) is read by
      // the next step; the fp32 outputs (sequence output, saved activations / cell state for the backward) are
      // returned in o[] so that the caller can store them after the barrier arrival.
      constexpr int NDF = 4;
      const int uq = 4 * (e & 3);                  // first of this thread's units inside the CTA's 16
      auto combine4 = [&](int b, float (&o)[6][NDF]) {
        float e0[NDF], e1[NDF], e2[NDF], e3[NDF], cp[NDF], hval[NDF];
        const bool valid = t < lens_s[b];
#pragma unroll
        for (int j = 0; j < NDF; ++j) {
          const int ui = uq + j;
          e0[j] = ex[(0 * UT + ui) * NBp + b]; e1[j] = ex[(1 * UT + ui) * NBp + b];
          e2[j] = ex[(2 * UT + ui) * NBp + b]; e3[j] = ex[(3 * UT + ui) * NBp + b];
          cp[j] = cst[ui * NBp + b];
        }
        if (RNN == DS2_RNN_LSTM) {
          float cval[NDF], th[NDF];
#pragma unroll
          for (int j = 0; j < NDF; ++j) cval[j] = valid ? fmaf(e1[j], cp[j], e0[j] * e2[j]) : 0.f;
#pragma unroll
          for (int j = 0; j < NDF; ++j) th[j] = ex2_ftz(-2.f * LOG2E * cval[j]);
#pragma unroll
          for (int j = 0; j < NDF; ++j) th[j] = rcp_ftz(1.f + th[j]);
#pragma unroll
          for (int j = 0; j < NDF; ++j) {
            hval[j] = valid ? e3[j] * fmaf(2.f, th[j], -1.f) : 0.f;
            if (valid) cst[(uq + j) * NBp + b] = cval[j];
            o[0][j] = valid ? e0[j] : 0.f; o[1][j] = valid ? e1[j] : 0.f;
            o[2][j] = valid ? e2[j] : 0.f; o[3][j] = valid ? e3[j] : 0.f;
            o[4][j] = cval[j];
          }
        } else if (RNN == DS2_RNN_GRU) {
          float th[NDF];
#pragma unroll
          for (int j = 0; j < NDF; ++j) th[j] = ex2_ftz(-2.f * LOG2E * fmaf(e0[j], e2[j], e3[j]));
#pragma unroll
          for (int j = 0; j < NDF; ++j) th[j] = rcp_ftz(1.f + th[j]);
#pragma unroll
          for (int j = 0; j < NDF; ++j) {
            const float nval = valid ? fmaf(2.f, th[j], -1.f) : 0.f;
            hval[j] = valid ? fmaf(e1[j], cp[j] - nval, nval) : 0.f;
            if (valid) cst[(uq + j) * NBp + b] = hval[j];
            o[0][j] = valid ? e0[j] : 0.f; o[1][j] = valid ? e1[j] : 0.f; o[2][j] = nval; o[3][j] = 0.f;
            o[4][j] = valid ? e2[j] : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < NDF; ++j) {
            hval[j] = valid ? e0[j] : 0.f;
            o[0][j] = hval[j]; o[1][j] = o[2][j] = o[3][j] = o[4][j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < NDF; ++j) o[5][j] = hval[j];
        if (RES) {
          const __half2 lo = __floats2half2_rn(hval[0], hval[1]), hi = __floats2half2_rn(hval[2], hval[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned int*>(&lo);
          pk.y = *reinterpret_cast<const unsigned int*>(&hi);
          *reinterpret_cast<uint2*>(p.h16 + (((size_t)d * T + t) * B + b) * H + u0 + uq) = pk;
        }
      };
      auto st4 = [](float* dst, const float (&v)[NDF]) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      };
      auto store4 = [&](int b, const float (&o)[6][NDF]) {
        const size_t so = (((size_t)d * T + t) * B + b) * H + u0 + uq;
        float* gp = p.gates + (((size_t)t * B + b) * D + d) * GH + u0 + uq;
        if (RNN != DS2_RNN_TANH) st4(p.aux + so, o[4]);
        if (p.training) {
#pragma unroll
          for (int g = 0; g < G; ++g) st4(gp + g * H, o[g]);
        }
        st4(p.hseq + so, o[5]);
      };
      const bool defer = RES && p.defer && B <= 32;
      const int b_own = e >> 2;
      float sv[6][NDF];
      if (defer) {
        if (b_own < B) combine4(b_own, sv);
      } else {
        for (int b = b_own; b < B; b += 32) {
          combine4(b, sv);
          store4(b, sv);
        }
      }
      if (e == 0) trace_stamp(p.trace, p.T, step, 8);
      named_bar_sync(1, 128);          // CTA-scope: every epilogue thread's stores happen-before thread 0's release
      if (e == 0) {
        trace_stamp(p.trace, p.T, step, 9);
        fence_proxy_async_global();
        trace_stamp(p.trace, p.T, step, 10);
        // release is cumulative over the stores ordered by the named barrier
        red_release(ctr, 1u);
        trace_stamp(p.trace, p.T, step, 11);
        trace_stamp_ns(p.trace, p.T, step, 12);
      }
      if (defer) {
        if (b_own < B) store4(b_own, sv);
        if (e == 0) trace_stamp(p.trace, p.T, step, 13);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc<512>(tmem_base);
  }
}

// A sweep whose barrier wait timed out (or that found an unexpected TMEM base) leaves garbage behind: fail loudly.
// Runs right after every sweep launch; the trap makes the next CUDA call of the process return an error.
__global__ void sweep_check_kernel(const int* err) {
  if (*err != 0) {
    printf("ds2: recurrent sweep failed (code %d): results are invalid\n", *err);
    __trap();
  }
}

// ------------------------------------------------------------------------------------------------
// Every persistent CTA allocates all 512 TMEM columns, so two of them must never share an SM (the second
// tcgen05.alloc would block while the first waits for it at the grid barrier): ask for more than half of
// the 227 KB of shared memory even when the tiles are small.
static size_t one_cta_per_sm(size_t smem) { return smem < 116 * 1024 ? 116 * 1024 : smem; }

static long long* trace_ptr_from_env(const char* name) {
  const char* e = getenv(name);   // debug: device address of an int64 buffer of 6*T entries
  return e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 10)) : nullptr;
}

// the epilogues use 8- / 16-byte vector accesses on these buffers
static bool vec_ok(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}
// Nsight Compute cannot replay a cooperative launch of a kernel with a cluster dimension (it aborts the target:
// `ncu_rc=9` in the round-1 driver record).  Under the profiler's injection — or with DS2_SPLITK_NONCOOP=1 — the
// cluster sweeps are launched without the cooperative attribute; co-residency is still verified with
// cudaOccupancyMaxActiveClusters, and a lone kernel of <= 148 one-per-SM CTAs on an otherwise idle stream becomes
// resident as a whole either way.
static bool noncoop_cluster_launch() {
  const char* e = getenv("DS2_SPLITK_NONCOOP");
  if (e) return atoi(e) != 0;
  static const bool under_ncu = [] {
    if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") ||
        getenv("NV_TPS_LAUNCH_TOKEN"))
      return true;
    const char* inj = getenv("CUDA_INJECTION64_PATH");
    if (inj && (strstr(inj, "nsight-compute") || strstr(inj, "cuda-injection"))) return true;
    void* h = dlopen("libcuda-injection.so", RTLD_NOLOAD | RTLD_LAZY);   // already mapped by the profiler?
    if (h) { dlclose(h); return true; }
    return false;
  }();
  return under_ncu;
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// DS2_SWEEP_DEFER=0: store everything before the barrier arrival (see PersistParams::defer)
static int sweep_defer_default() { return env_flag("DS2_SWEEP_DEFER", 1); }

static size_t fwd_smem_bytes(int NB) {
  using namespace rp;
  size_t NBp = NB + 1;
  return 1024 + (size_t)STAGES * (A_BYTES + (size_t)NB * 128) + (5 * UT * NBp + NB + 4) * sizeof(float) +
         (2 * STAGES + 2) * sizeof(uint64_t) + 64;
}

size_t rnn_sweep_tc_workspace_bytes(int rnn, int T, int B, int H, int D) {
  const int G = rnn == DS2_RNN_LSTM ? 4 : (rnn == DS2_RNN_GRU ? 3 : 1);
  const size_t fwd = 4096 + align_up((size_t)D * G * H * H * 2, 256) + align_up((size_t)D * T * B * H * 2, 256);
  const size_t GH = (size_t)G * H;
  const size_t bwd = 4096 + align_up((size_t)D * (T + 1) * 4, 256) + align_up((size_t)T * 4, 256) +
                     align_up((size_t)D * T * (H / 16 + 1) * 4 * 4, 256) + align_up((size_t)D * H * GH * 2, 256) +
                     align_up((size_t)T * B * D * GH * 2, 256);   // >= splitk_res_ws_bytes()
  return (fwd > bwd ? fwd : bwd) + 256;
}

static void set_acc_layout(PersistParams& p) {
  // accumulator width = power of two >= max(32, NB) columns; as many chains as fit in the 512 TMEM columns
  int cols = 32;
  while (cols < p.NB) cols *= 2;
  p.acc_cols = cols;
  p.nacc = 1;   // one chain: switching accumulators between MMAs was measured slower, not faster
}

static bool fwd_eligible(const SeqArgs& a) {
  if (a.h0 || a.c0) return false;                 // initial states -> generic step kernels
  if (a.H % 32 != 0 || a.B > 256 || a.T < 2) return false;
  return vec_ok(a.gates, a.hseq, a.aux);
}

__global__ void f32_to_f16_kernel(size_t n, const float* __restrict__ in, __half* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2half_rn(in[i]);
}

static size_t res_smem_bytes(int NB, int H) {
  using namespace rp;
  size_t NBp = NB + 1;
  return 1024 + (size_t)(H / 64) * (A_BYTES + (size_t)NB * 128) + (5 * UT * NBp + NB + 4) * sizeof(float) +
         (32 + STAGES + 2) * sizeof(uint64_t) + 64;
}

// workspace of the resident forward: [4 KB control][W16: D*G*H*H halfs][h16: D*T*B*H halfs]
static size_t res_ws_bytes(int G, int T, int B, int H, int D) {
  return 4096 + align_up((size_t)D * G * H * H * 2, 256) + align_up((size_t)D * T * B * H * 2, 256);
}

template <int RNN>
static int launch_fwd_resident(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  const int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  if (a.H % 64 != 0 || a.H / 64 > 32) return 1;
  if (!vec_ok(a.gates, a.hseq, a.aux)) return 1;
  if (ws_bytes < res_ws_bytes(G, a.T, a.B, a.H, a.D)) return 1;
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 7) / 8 * 8; p.H = a.H; p.D = a.D; p.NT = a.H / UT; p.G = G;
  p.training = a.training;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux;
  p.trace = trace_ptr_from_env("DS2_TRACE_FWD");
  p.defer = sweep_defer_default();
  set_acc_layout(p);
  p.err = static_cast<int*>(ws);
  p.bar = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 128);
  __half* w16 = reinterpret_cast<__half*>(static_cast<char*>(ws) + 4096);
  p.h16 = reinterpret_cast<__half*>(static_cast<char*>(ws) + 4096 + align_up((size_t)a.D * G * a.H * a.H * 2, 256));
  const size_t smem = one_cta_per_sm(res_smem_bytes(p.NB, a.H));
  if (smem > 227 * 1024) return 1;
  auto kern = rnn_fwd_persist_kernel<RNN, true>;
  static DeviceOnce attr_once;
  const int num_sms = device_sm_count();
  if (attr_once.first()) {
    DS2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  int max_blocks_per_sm = 0;
  DS2_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, kern, THREADS, smem));
  int grid = a.D * p.NT, launches = 1;
  if (max_blocks_per_sm < 1) return 1;
  if (grid > max_blocks_per_sm * num_sms) {           // both directions do not fit: one launch per direction
    if (p.NT > max_blocks_per_sm * num_sms) return 1;
    grid = p.NT;
    launches = a.D;
  }
  const size_t wn = (size_t)G * a.H * a.H;
  for (int d = 0; d < a.D; ++d) {
    p.b_ih[d] = a.b_ih[d];
    p.b_hh[d] = a.b_hh[d];
    DS2_LAUNCH(f32_to_f16_kernel, 148 * 4, 256, 0, st, wn, a.w_hh[d], w16 + (size_t)d * wn);
    int rc = make_tmap_f16(&p.tmW[d], w16 + (size_t)d * wn, 3, a.H, a.H, G, (size_t)a.H, (size_t)a.H * a.H, 64, UT, G);
    if (rc) return rc;
    rc = make_tmap_f16(&p.tmV[d], p.h16 + (size_t)d * a.T * a.B * a.H, 2, a.H, a.T * a.B, 1, (size_t)a.H, 0, 64, a.B, 1);
    if (rc) return rc;
    p.box3 = (a.H / 64) % 4 == 0;
    if (p.box3) {   // rows b >= B of a box belong to the next time step (or are zero-filled): those N columns are discarded
      rc = make_tmap_f16(&p.tmV3[d], p.h16 + (size_t)d * a.T * a.B * a.H, 3, 64, a.T * a.B, a.H / 64, (size_t)a.H, 64, 64,
                         p.NB, 4);
      if (rc) return rc;
    }
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096, st));
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    void* args[] = {&p};
    DS2_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(THREADS), args, smem, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  return DS2_OK;
}

template <int RNN, bool LL>
static int launch_fwd_splitk(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);

template <int RNN>
static int launch_fwd(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  const int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  if (!getenv("DS2_NO_RESIDENT")) {
    if (RNN != DS2_RNN_TANH && env_flag("DS2_FWD_SPLITK", 1)) {   // 2-CTA clusters, half the MMA chain per step
      constexpr int R = RNN == DS2_RNN_TANH ? DS2_RNN_LSTM : RNN;
      // DS2_FWD_LL=1: flag-in-data exchange instead of grid barrier + TMA of h_{t-1}.  Measured SLOWER (13.9k vs 8.7k
      // cycles per step, profiles/r02_ll_exchange.md): strong per-thread loads do not pipeline (~600 cycles each),
      // and even with the TMA engine as transport the exchange costs what the barrier costs — the floor is the
      // store -> L2 -> load visibility latency, not the barrier.  Kept as a tested, selectable variant.
      int rc = env_flag("DS2_FWD_LL", 0) ? launch_fwd_splitk<R, true>(a, ws, ws_bytes, st)
                                         : launch_fwd_splitk<R, false>(a, ws, ws_bytes, st);
      if (rc != 1) return rc;
    }
    int rc = launch_fwd_resident<RNN>(a, ws, ws_bytes, st);
    if (rc != 1) return rc;
  }
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 7) / 8 * 8; p.H = a.H; p.D = a.D; p.NT = a.H / UT; p.G = G;
  p.training = a.training;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux;
  p.trace = trace_ptr_from_env("DS2_TRACE_FWD");
  if (ws_bytes < 4096) return 1;
  set_acc_layout(p);
  p.err = static_cast<int*>(ws);
  p.bar = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 128);
  const size_t smem = one_cta_per_sm(fwd_smem_bytes(p.NB));
  auto kern = rnn_fwd_persist_kernel<RNN, false>;
  static DeviceOnce attr_once;
  int max_blocks_per_sm = 0;
  const int num_sms = device_sm_count();
  if (attr_once.first()) {
    DS2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  if (smem > 227 * 1024) return 1;
  DS2_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, kern, THREADS, smem));
  int grid = a.D * p.NT, launches = 1;
  if (max_blocks_per_sm < 1) return 1;
  if (grid > max_blocks_per_sm * num_sms) {           // both directions do not fit: one launch per direction
    if (p.NT > max_blocks_per_sm * num_sms) return 1;
    grid = p.NT;
    launches = a.D;
  }   // cannot be co-resident
  for (int d = 0; d < a.D; ++d) {
    p.b_ih[d] = a.b_ih[d];
    p.b_hh[d] = a.b_hh[d];
    int rc = make_tmap_3d(&p.tmW[d], a.w_hh[d], a.H, a.H, G, (size_t)a.H, (size_t)a.H * a.H, BK, UT, G);
    if (rc) return rc;
    rc = make_tmap_2d(&p.tmV[d], a.hseq + (size_t)d * a.T * a.B * a.H, a.T * a.B, a.H, a.H, a.B, BK);
    if (rc) return rc;
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096, st));
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    void* args[] = {&p};
    DS2_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(THREADS), args, smem, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  return DS2_OK;
}

int rnn_sweep_fwd_tc(int rnn, const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!fwd_eligible(a)) return 1;
  if (rnn == DS2_RNN_LSTM) return launch_fwd<DS2_RNN_LSTM>(a, ws, ws_bytes, st);
  if (rnn == DS2_RNN_GRU) return launch_fwd<DS2_RNN_GRU>(a, ws, ws_bytes, st);
  return launch_fwd<DS2_RNN_TANH>(a, ws, ws_bytes, st);
}

// ------------------------------------------------------------------------------------------------
// Backward sweep.  A CTA owns 16 hidden units of one direction: rows u0..u0+15 of W_hh^T (H x G*H)
// are the (16 valid of 64) M rows, the gate-gradient vector of the previously processed step
// dGh[t_next] (B x G*H) is the N operand, K = G*H:
//     dh_rec[u, b] = sum_k W_hh^T[u, k] * dGh[t_next][b, k]
// Epilogue: dh = dY[t] + dh_rec (+ carried terms), gate backward from the saved activations, the
// gate gradients overwrite the activations in place (and are the next step's N operand).  The
// cell-state / hidden-state carry of the owned units stays in shared memory for the whole sweep.
template <int RNN>
__global__ void __launch_bounds__(rp::THREADS, 1) rnn_bwd_persist_kernel(const __grid_constant__ PersistParams p) {
  using namespace rp;
  using namespace tc;
  constexpr int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still __shared__
  const int NB = p.NB, B = p.B, T = p.T, H = p.H, D = p.D;
  const int B_BYTES = NB * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const int NBp = NB + 1;
  float* ex = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);   // [UT][NBp] dh_rec
  float* cst = ex + 4 * UT * NBp;                                       // [UT][NBp] carried dc (LSTM) / dh (GRU)
  uint64_t* full = reinterpret_cast<uint64_t*>(cst + UT * NBp);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_bar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int d = p.d0 + blockIdx.x / p.NT, tile = blockIdx.x % p.NT, u0 = tile * UT;
  const int GH = G * H;
  const int NK = GH / BK;
  unsigned int* ctr = p.bar + 32 * d;   // one 128-byte line per direction

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[d]);
    tma_prefetch_desc(&p.tmV[d]);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < UT * NBp; i += THREADS) { cst[i] = 0.f; ex[i] = 0.f; }
  if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The CTA owns all 512 TMEM columns, so the allocation starts at address 0.  Using the literal keeps the
  // accumulator operand of tcgen05.mma in a uniform register without the ELECT / R2UR.BROADCAST waterfall
  // loop nvcc otherwise emits per MMA for a value it cannot prove warp-uniform.
  if (tmem_base != 0 && threadIdx.x == 0) {
    *(volatile int*)p.err = 2;
    printf("ds2: unexpected TMEM base %u (block %d)\n", tmem_base, blockIdx.x);
  }
  const uint32_t tx_bytes = (uint32_t)(UT * 128 + B * 128);

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        const int t = d == 0 ? T - 1 - step : step;
        const int tn = d == 0 ? t + 1 : t - 1;
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], tx_bytes);
          tma_load_2d(smem + s * STAGE_BYTES, &p.tmW[d], &full[s], c * BK, u0);
          if (c == 0) {
            grid_wait_counter(ctr, (unsigned int)p.NT * (unsigned int)step, p.err);
            fence_proxy_async_global();
            trace_stamp(p.trace, p.T, step, 0);
          }
          const int k0 = c * BK;
          if (RNN == DS2_RNN_GRU && k0 >= 2 * H)
            tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &p.tmV2[d], &full[s], k0 - 2 * H, tn * B);
          else
            tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &p.tmV[d], &full[s], d * GH + k0, tn * B);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // The issue chain is the critical path of a step: the whole warp runs it converged (mma_tf32_w: elect.sync
      // inside; a single divergent lane costs ~70 cycles per tcgen05.mma, converged ~30-40) and the loop is free of
      // div/mod and descriptor construction.
      const uint32_t idesc = instr_desc(FMT_TF32, MM, NB);
      const uint64_t a_base = smem_desc_sw128(smem_u32(smem));
      const uint64_t b_base = smem_desc_sw128(smem_u32(smem + A_BYTES));
      const uint64_t stage_step = (uint64_t)(STAGE_BYTES >> 4);
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&full[s], ph);
          if (c == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 1);
          if (c == NK - 1 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
          tc_fence_after();
          const uint64_t ad = a_base + (uint64_t)s * stage_step, bd = b_base + (uint64_t)s * stage_step;
          mma_tf32_w(0u, ad, bd, idesc, c > 0);
          mma_tf32_w(0u, ad + 2, bd + 2, idesc, 1);
          mma_tf32_w(0u, ad + 4, bd + 4, idesc, 1);
          mma_tf32_w(0u, ad + 6, bd + 6, idesc, 1);
          mma_commit_w(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_w(accum_bar);
      }
    }
  } else {
    const int q = warp % 4;
    const int e = threadIdx.x - 64;
    uint32_t acc_phase = 0;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? T - 1 - step : step;
      const int tp = d == 0 ? t - 1 : t + 1;
      const bool tp_in = tp >= 0 && tp < T;
      if (q == 0 && step > 0) {          // rows 0..15 of the accumulator live in lanes 0..15 of quarter 0
        mbar_wait(accum_bar, acc_phase);
        tc_fence_after();
        if (lane == 0) trace_stamp(p.trace, p.T, step, 3);
        for (int cb = 0; cb < NB; cb += 32) {
          float acc[32];
          const int nsum = min(p.nacc, NK * (BK / 8));
          tmem_ld32(tmem_base + (uint32_t)cb, acc);
          for (int a2 = 1; a2 < nsum; ++a2) {
            float part[32];
            tmem_ld32(tmem_base + (uint32_t)(a2 * p.acc_cols + cb), part);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += part[j];
          }
          if (lane < UT) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cb + j < B) ex[lane * NBp + cb + j] = acc[j];
          }
        }
      }
      if (step > 0) acc_phase ^= 1;
      tc_fence_before();
      named_bar_sync(1, 128);
      if (e == 0) trace_stamp(p.trace, p.T, step, 4);
      for (int pi = e; pi < UT * B; pi += 128) {
        const int ui = pi % UT, b = pi / UT;
        const bool valid = t < p.len[b];
        const bool pin = tp_in && (d == 0 || tp < p.len[b]);
        const size_t si = (((size_t)d * T + t) * B + b) * H + u0 + ui;
        const size_t sp = (((size_t)d * T + (tp_in ? tp : 0)) * B + b) * H + u0 + ui;
        float* gp = p.gates + (((size_t)t * B + b) * D + d) * GH + u0 + ui;
        if (!valid) {
#pragma unroll
          for (int g = 0; g < G; ++g) gp[g * H] = 0.f;
          if (RNN == DS2_RNN_GRU) p.aux[si] = 0.f;
        } else {
          float dh = p.dy[((size_t)t * B + b) * H + u0 + ui] + ex[ui * NBp + b];
          if (RNN == DS2_RNN_LSTM) {
            const float c_prev = pin ? p.aux[sp] : 0.f;
            LstmBwd r = lstm_cell_bwd(gp[0], gp[H], gp[2 * H], gp[3 * H], p.aux[si], c_prev, dh, cst[ui * NBp + b]);
            gp[0] = r.di; gp[H] = r.df; gp[2 * H] = r.dg; gp[3 * H] = r.d_o;
            cst[ui * NBp + b] = r.dc_prev;
          } else if (RNN == DS2_RNN_GRU) {
            const float h_prev = pin ? p.hseq[sp] : 0.f;
            dh += cst[ui * NBp + b];
            GruBwd r = gru_cell_bwd(gp[0], gp[H], gp[2 * H], p.aux[si], h_prev, dh);
            gp[0] = r.dr; gp[H] = r.dz; gp[2 * H] = r.dxn;
            p.aux[si] = r.dhn;
            cst[ui * NBp + b] = r.dh_prev;
          } else {
            const float h = p.hseq[si];
            gp[0] = dh * (1.f - h * h);
          }
        }
      }
      named_bar_sync(1, 128);          // CTA-scope: every epilogue thread's stores happen-before thread 0's release
      if (e == 0) {
        fence_proxy_async_global();
        red_release(ctr, 1u);            // release is cumulative over the stores ordered by the named barrier
        trace_stamp(p.trace, p.T, step, 5);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward sweep, split-K variant (the fast path): a 4-CTA cluster owns 64 hidden units; CTA `ks` of
// the cluster reduces over a quarter of K = G*H, so every CTA issues the same number of MMAs per
// step as the forward sweep with all 64 M rows useful.  The partial sums (64 units x B) are exchanged
// through distributed shared memory in push form: TMEM quarter q of every CTA holds the rows that CTA q
// of the cluster finishes, so epilogue warp q sends them straight into CTA q's shared memory with
// st.async (16-byte stores that complete_tx on the destination's mbarrier); each CTA then reduces the four
// slices it received from its own shared memory and finishes its 16 units (gate backward, carried dc / dh).
// (The earlier pull form — publish, cluster barrier, 4-byte ld.shared::cluster reads with a 132-byte lane
// stride — cost ~5.9k of the 16k cycles per step.)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err) {
  uint32_t ok = 0, it = 0;
  long long t0 = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tc::smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && (++it & 1023u) == 0) {            // bounded: a protocol fault must not hang the GPU
      if (t0 == 0) t0 = clock64();
      if (*(volatile int*)err) return;
      if (clock64() - t0 > rp::SPIN_LIMIT) { *(volatile int*)err = 1; return; }
    }
  }
}
// Layout of the partial tiles exchanged inside a cluster: [source CTA][group of 4 columns][row position][4 columns].
// A lane owns an accumulator row, so for one group of 4 columns the 32 lanes of a warp store to (nearly)
// consecutive 16-byte slots (with a [row][column] layout every lane hit a different 144-byte row: ~170 cycles
// per st.async instruction).  Rows are skewed by one slot every 8 rows so that the reader's rows r, r+4, r+8,
// r+12 fall into different banks, and a column group is padded by one slot.
__host__ __device__ constexpr int xt_cgs(int rows) { return (rows + rows / 8) * 4 + 4; }                   // floats / group
__host__ __device__ constexpr int xt_slice(int rows, int nb) { return (nb / 4) * xt_cgs(rows); }           // floats / source
__device__ __forceinline__ int xt_off(int rows, int row, int col) {
  return (col >> 2) * xt_cgs(rows) + (row + (row >> 3)) * 4 + (col & 3);
}
// 16-byte store into a cluster peer's shared memory that also counts 16 bytes on the peer's mbarrier
// (st.async: data and completion travel together, no fence / separate arrive on the critical path)
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   cluster_addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// RES = true: the CTA's slice of W_hh^T (64 units x K/4, fp16, 128 KB at H=1024) is resident in shared memory and
// the streamed operand is a SCALED fp16 copy of the gate gradients: dg16 = fp16(dGh * S_s), S_s a power of two
// chosen from the maximum |dGh| of the previously processed step (global atomicMax, final at the grid barrier;
// step 0 uses the bound max|dY|), so that the largest value sits near 2^8: 256x of fp16 headroom above, 2^-22 of
// the maximum still representable below.  fp16 carries the same 10-bit mantissa as TF32.
__device__ __forceinline__ __half to_half_sat(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65000.f), 65000.f));
}
// scale exponent for a step whose maximum magnitude is m: m * 2^sx lands in [2^7, 2^8).  Exponent arithmetic on
// the float bits (m = f * 2^(E-126), f in [0.5, 1)) instead of frexpf / ldexpf / a division on the critical path.
__device__ __forceinline__ int pow2_exp_for(unsigned int m_bits, int sx_prev) {
  if (m_bits == 0u || m_bits >= 0x7f800000u) return sx_prev;     // 0, inf, nan: keep the previous scale
  const int E = (int)((m_bits >> 23) & 0xffu);
  return max(-100, min(100, 134 - E));
}
__device__ __forceinline__ float pow2f(int ex) { return __int_as_float((ex + 127) << 23); }   // |ex| <= 100

// CL = CTAs per cluster = K split: 4 (64 units per cluster, MMA M = 64) or 8 (128 units, M = 128: the same
// number of CTAs, but each reduces only K/8, i.e. half as many MMA instructions on the per-step critical path).
// LL (RES only): no grid barrier; the scaled fp16 gate gradients are their own ready flags (see the LL helpers) and
// the per-step maxima travel as one word per (CTA, epilogue warp) in `gmeta`.
template <int RNN, bool RES, int CL, bool LL = false, int NKR_T = 0>   // NKR_T: see rnn_fwd_splitk_kernel
__global__ void __launch_bounds__(rp::THREADS, 1) rnn_bwd_splitk_kernel(const __grid_constant__ PersistParams p) {
  using namespace rp;
  using namespace tc;
  static_assert(!LL || RES, "the flag-in-data exchange streams the fp16 copy of the resident variant");
  constexpr int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  constexpr int UM = UT * CL;                  // units per cluster (all M rows valid): 64 or 128 = MMA M
  constexpr int A_BYTES = UM * 128;            // one K chunk of the weight tile (shadows rp::A_BYTES)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still __shared__
  const int NB = p.NB, B = p.B, T = p.T, H = p.H, D = p.D;
  const int B_BYTES = NB * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const int NBp = NB + 1;
  const int NKR = NKR_T ? NKR_T : (G * H / CL) / 64;                     // resident: 64 fp16 of K per chunk
  const int NG = grp_count(NKR);
  const int ring_bytes = RES ? NKR * STAGE_BYTES : STAGES * STAGE_BYTES;
  float* part = reinterpret_cast<float*>(smem + ring_bytes);             // [CL sources] partial dh_rec tiles (xt_* layout)
  float* cst = part + CL * xt_slice(UT, NB);                             // [16][NBp] carried dc / dh
  int* lens_s = reinterpret_cast<int*>(cst + UT * NBp);
  unsigned int* cta_max = reinterpret_cast<unsigned int*>(lens_s + ((NB + 1) & ~1));   // [2] (8 bytes)
  unsigned int* wmax = cta_max + 2;                                      // LL: [2 step parities][4 epilogue warps]
  uint64_t* full = reinterpret_cast<uint64_t*>(cta_max + 10);            // resident: one per group of 4 chunks
  uint64_t* empty = full + (RES ? 32 : STAGES);                          // resident: [0] = weights landed
  uint64_t* accum_bar = empty + STAGES;
  uint64_t* part_bar = accum_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int ks = blockIdx.x % CL;                         // rank in the cluster == K split
  const int cl = blockIdx.x / CL;
  const int NTc = H / UM;                                 // clusters per direction
  const int d = p.d0 + cl / NTc, ut = cl % NTc;
  const int GH = G * H;
  const int Kc = GH / CL, NK = Kc / BK;
  const int kbase = ks * Kc;
  const int u0 = ut * UM + ks * UT;                       // the 16 units this CTA finishes
  unsigned int* ctr = p.bar + 32 * d;
  const unsigned int n_arrive = (unsigned int)(NTc * CL);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[d]);
    if (!LL) tma_prefetch_desc(&p.tmV[d]);
    for (int i = 0; i < (RES ? 32 : STAGES); ++i) mbar_init(&full[i], LL ? 4 : 1);   // LL: one arrival per loader warp
    for (int i = 0; i < STAGES; ++i) mbar_init(&empty[i], 1);
    mbar_init(accum_bar, 1);
    mbar_init(part_bar, 1);
    fence_barrier_init();
    cta_max[0] = 0u;
    for (int i = 0; i < 8; ++i) wmax[i] = 0u;
  }
  for (int i = threadIdx.x; i < UT * NBp; i += THREADS) cst[i] = 0.f;
  for (int i = threadIdx.x; i < NB; i += THREADS) lens_s[i] = i < B ? p.len[i] : 0;
  if (LL) {   // batch-padding rows of the streamed tile are never loaded: keep them finite (their columns are unused)
    uint8_t* vb = smem + NKR * A_BYTES;
    for (int i = threadIdx.x; i < NKR * (NB - B) * 8; i += THREADS) {
      const int c = i / ((NB - B) * 8), r = B + (i / 8) % (NB - B), j = i % 8;
      *reinterpret_cast<uint4*>(vb + c * B_BYTES + r * 128 + j * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                                     // peers' mbarriers are initialised
  const uint32_t tmem_base = *tmem_slot;
  // The CTA owns all 512 TMEM columns, so the allocation starts at address 0.  Using the literal keeps the
  // accumulator operand of tcgen05.mma in a uniform register without the ELECT / R2UR.BROADCAST waterfall
  // loop nvcc otherwise emits per MMA for a value it cannot prove warp-uniform.
  if (tmem_base != 0 && threadIdx.x == 0) {
    *(volatile int*)p.err = 2;
    printf("ds2: unexpected TMEM base %u (block %d)\n", tmem_base, blockIdx.x);
  }
  const uint32_t tx_bytes = (uint32_t)(UM * 128 + B * 128);

  if (warp == 0) {
    if (RES) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&empty[0], (uint32_t)(NKR * UM * 128));
        for (int c = 0; c < NKR; ++c) tma_load_2d(smem + c * A_BYTES, &p.tmW[d], &empty[0], kbase + c * 64, ut * UM);
        uint8_t* vbuf = smem + NKR * A_BYTES;
        for (int step = 1; !LL && step < T; ++step) {     // LL: the epilogue warps fetch dGh[t_next] themselves
          const int t = d == 0 ? T - 1 - step : step;
          const int tn = d == 0 ? t + 1 : t - 1;
          grid_wait_counter(ctr, n_arrive * (unsigned int)step, p.err);
          fence_proxy_async_global();
          trace_stamp(p.trace, p.T, step, 0);
          // every CTA's atomicMax of the previous step precedes its arrival: the maximum is final.  It is handed
          // to the epilogue warps through shared memory (published before the last group is armed, so the
          // full-barrier -> MMA -> accumulator-barrier chain orders it) instead of a ~300-cycle global load
          // on their critical path.
          const unsigned int gm = ld_relaxed(p.gmax + (size_t)d * (T + 1) + step);
          for (int g = 0; g < NG; ++g) {
            uint64_t* fb = full + g;
            const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
            if (c1 == NKR) *(volatile unsigned int*)(cta_max + 1) = gm;
            if (p.box3) {
              mbar_arrive_expect_tx(fb, (uint32_t)(4 * B_BYTES));
              tma_load_3d(vbuf + c0 * B_BYTES, &p.tmV3[d], fb, 0, tn * B, (d * GH + kbase) / 64 + c0);
            } else {
              mbar_arrive_expect_tx(fb, (uint32_t)((c1 - c0) * B * 128));
              for (int c = c0; c < c1; ++c)
                tma_load_2d(vbuf + c * B_BYTES, &p.tmV[d], fb, d * GH + kbase + c * 64, tn * B);
            }
          }
          trace_stamp(p.trace, p.T, step, 1);
        }
      }
    } else
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        const int t = d == 0 ? T - 1 - step : step;
        const int tn = d == 0 ? t + 1 : t - 1;
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], tx_bytes);
          const int k0 = kbase + c * BK;
          tma_load_2d(smem + s * STAGE_BYTES, &p.tmW[d], &full[s], k0, ut * UM);
          if (c == 0) {
            grid_wait_counter(ctr, n_arrive * (unsigned int)step, p.err);
            fence_proxy_async_global();
            trace_stamp(p.trace, p.T, step, 0);
          }
          if (RNN == DS2_RNN_GRU && k0 >= 2 * H)
            tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &p.tmV2[d], &full[s], k0 - 2 * H, tn * B);
          else
            tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &p.tmV[d], &full[s], d * GH + k0, tn * B);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (RES) {
      {   // the whole warp runs the issue loop converged (mma_f16_w: elect.sync inside): 64 MMAs per step in ~1.9k
          // cycles instead of ~4.9k from a single divergent lane
        const uint32_t idesc = instr_desc(FMT_F16, UM, NB);
        const uint64_t a_base = warp_uniform(smem_desc_sw128(smem_u32(smem)));
        const uint64_t b_base = warp_uniform(smem_desc_sw128(smem_u32(smem + NKR * A_BYTES)));
        const uint64_t a_step = (uint64_t)(A_BYTES >> 4), b_step = (uint64_t)(B_BYTES >> 4);
        mbar_wait(&empty[0], 0);
        uint32_t ph = 0;
        for (int step = 1; step < T; ++step) {
          if constexpr (NKR_T > 0) {
            constexpr int NGT = (NKR_T + 3) / 4;
#pragma unroll
            for (int g = 0; g < NGT; ++g) {
              mbar_wait(full + g, ph);
              tc_fence_after();
              if (g == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
              if (g == NGT - 1 && lane == 0) trace_stamp(p.trace, p.T, step, 3);
#pragma unroll
              for (int c = 4 * g; c < (4 * g + 4 < NKR_T ? 4 * g + 4 : NKR_T); ++c) {
                const uint64_t ad = a_base + (uint64_t)c * a_step, bd = b_base + (uint64_t)c * b_step;
                mma_f16_w(0u, ad, bd, idesc, c > 0);
                mma_f16_w(0u, ad + 2, bd + 2, idesc, 1);
                mma_f16_w(0u, ad + 4, bd + 4, idesc, 1);
                mma_f16_w(0u, ad + 6, bd + 6, idesc, 1);
              }
            }
          } else {
            for (int g = 0; g < NG; ++g) {
              mbar_wait(full + g, ph);
              tc_fence_after();
              if (g == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
              const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
              if (c1 == NKR && lane == 0) trace_stamp(p.trace, p.T, step, 3);
              for (int c = c0; c < c1; ++c) {
                const uint64_t ad = a_base + (uint64_t)c * a_step, bd = b_base + (uint64_t)c * b_step;
                mma_f16_w(0u, ad, bd, idesc, c > 0);
                mma_f16_w(0u, ad + 2, bd + 2, idesc, 1);
                mma_f16_w(0u, ad + 4, bd + 4, idesc, 1);
                mma_f16_w(0u, ad + 6, bd + 6, idesc, 1);
              }
            }
          }
          mma_commit_w(accum_bar);
          if (lane == 0) trace_stamp(p.trace, p.T, step, 4);
          ph ^= 1;
        }
      }
    } else
    {   // converged issue, see mma_tf32_w
      const uint32_t idesc = instr_desc(FMT_TF32, UM, NB);
      const uint64_t a_base = smem_desc_sw128(smem_u32(smem));
      const uint64_t b_base = smem_desc_sw128(smem_u32(smem + A_BYTES));
      const uint64_t stage_step = (uint64_t)(STAGE_BYTES >> 4);
      int s = 0;
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        for (int c = 0; c < NK; ++c) {
          mbar_wait(&full[s], ph);
          if (c == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 1);
          if (c == NK - 1 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
          tc_fence_after();
          const uint64_t ad = a_base + (uint64_t)s * stage_step, bd = b_base + (uint64_t)s * stage_step;
          mma_tf32_w(0u, ad, bd, idesc, c > 0);
          mma_tf32_w(0u, ad + 2, bd + 2, idesc, 1);
          mma_tf32_w(0u, ad + 4, bd + 4, idesc, 1);
          mma_tf32_w(0u, ad + 6, bd + 6, idesc, 1);
          mma_commit_w(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_w(accum_bar);
      }
    }
  } else {
    const int q = warp % 4;
    const int e = threadIdx.x - 64;
    const int ul = lane & 15, half = lane >> 4;
    uint32_t acc_phase = 0, part_phase = 0;
    // Destination of this lane's accumulator row: the CTA that finishes the row's unit, slice ks (= this CTA's rank)
    // of its tile.  M = 64 (CL 4): TMEM quarter q holds rows 16q..16q+15 in lanes 0..15 -> CTA q, the 32 columns are
    // split with lane+16 by shuffle.  M = 128 (CL 8): lane l of quarter q holds row 32q+l -> CTA 2q + l/16, all 32
    // columns of the row are sent by that lane.
    const int dst_cta = CL == 4 ? q : 2 * q + half;
    const uint32_t dst_row = mapa_u32(smem_u32(part), (uint32_t)dst_cta) + (uint32_t)(ks * xt_slice(UT, NB) * 4);
    const uint32_t dst_bar = mapa_u32(smem_u32(part_bar), (uint32_t)dst_cta);
    const uint32_t part_tx = (uint32_t)(CL * UT * NB * 4);               // bytes this CTA receives per step
    // resident: s_cur scales what this step writes, s_prev un-scales what this step's MMAs consumed
    const unsigned int* gmax_d = RES ? p.gmax + (size_t)d * (T + 1) : nullptr;
    // bias gradients: this thread's 4 units x (gate) sums over all steps of its batch column(s)
    float bsum[5][4];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) bsum[i][j] = 0.f;
    int sx_prev = 0, sx_cur = 0;
    // step-0 scale: |dGh| <= |dh| = |dY[t_first]| for every cell type
    if (RES) sx_cur = pow2_exp_for(__ldg(p.dymax + (d == 0 ? T - 1 : 0)), 0);
    float s_cur = pow2f(sx_cur), inv_prev = 1.f;
    const int cta_in_dir = ut * CL + ks, nmeta = NTc * CL * 4;
    (void)gmax_d;
    const float nscale = (RES && p.dgn16) ? __ldg(p.nscale) : 1.f;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? T - 1 - step : step;
      const int tp = d == 0 ? t - 1 : t + 1;
      const bool tp_in = tp >= 0 && tp < T;
      float lmax = 0.f;
      // the scale of this step's outputs must also cover this step's own upstream gradient: a frame of dY far above
      // its neighbours would otherwise saturate the fp16 copy (the previous step's maximum knows nothing of it)
      const unsigned int dym = RES ? __ldg(p.dymax + t) : 0u;
      // Saved activations / states / dY of this step do not depend on the recurrence, and every gate gradient is
      // linear in dh (LSTM: in dh and dc): fetch them and reduce them to per-pair coefficients BEFORE waiting for
      // the MMAs, so that global-load and special-function latencies hide behind the tensor-core phase and only
      // a handful of multiplies remain on the critical path.
      //   LSTM  k = {o(1-tc^2), tc o(1-o), g i(1-i), c_prev f(1-f), i(1-g^2), f}   tc = tanh(c_t)
      //   GRU   k = {cn hn r(1-r), (h_prev-n) z(1-z), cn, cn r, z}                  cn = (1-z)(1-n^2)
      //   tanh  k = {1-h^2}
      // Thread e owns the 4 consecutive units 4*(e&3)..+3 of batch column (e>>2) [+32 per block]: every global
      // access is one 8- or 16-byte vector.
      constexpr int NPF = 4;
      const bool single = B <= 32;
      const int uq = 4 * (e & 3), b_own = e >> 2;
      auto ld4 = [](const float* src, float (&v)[NPF]) {
        const float4 x = *reinterpret_cast<const float4*>(src);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
      };
      auto load_coefs = [&](int b, float (&k)[6][NPF], float (&dyv)[NPF], bool& valid) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < NPF; ++j) k[i][j] = 0.f;
#pragma unroll
        for (int j = 0; j < NPF; ++j) dyv[j] = 0.f;
        valid = b < B && t < lens_s[b < B ? b : 0];
        if (!valid) return;
        const bool pin = tp_in && (d == 0 || tp < lens_s[b]);
        const size_t si = (((size_t)d * T + t) * B + b) * H + u0 + uq;
        const size_t sp = (((size_t)d * T + (tp_in ? tp : 0)) * B + b) * H + u0 + uq;
        const float* gp = p.gates + (((size_t)t * B + b) * D + d) * GH + u0 + uq;
        ld4(p.dy + ((size_t)t * B + b) * H + u0 + uq, dyv);
        if (RNN == DS2_RNN_LSTM) {
          float gi[NPF], gf[NPF], gg[NPF], go[NPF], c[NPF], c_prev[NPF] = {0.f, 0.f, 0.f, 0.f}, tc[NPF];
          ld4(gp, gi); ld4(gp + H, gf); ld4(gp + 2 * H, gg); ld4(gp + 3 * H, go);
          ld4(p.aux + si, c);
          if (pin) ld4(p.aux + sp, c_prev);
#pragma unroll
          for (int j = 0; j < NPF; ++j) tc[j] = ex2_ftz(-2.f * LOG2E * c[j]);
#pragma unroll
          for (int j = 0; j < NPF; ++j) tc[j] = fmaf(2.f, rcp_ftz(1.f + tc[j]), -1.f);   // the forward used this tanh
#pragma unroll
          for (int j = 0; j < NPF; ++j) {
            k[0][j] = go[j] * (1.f - tc[j] * tc[j]); k[1][j] = tc[j] * go[j] * (1.f - go[j]);
            k[2][j] = gg[j] * gi[j] * (1.f - gi[j]); k[3][j] = c_prev[j] * gf[j] * (1.f - gf[j]);
            k[4][j] = gi[j] * (1.f - gg[j] * gg[j]); k[5][j] = gf[j];
          }
        } else if (RNN == DS2_RNN_GRU) {
          float r[NPF], z[NPF], n[NPF], hn[NPF], h_prev[NPF] = {0.f, 0.f, 0.f, 0.f};
          ld4(gp, r); ld4(gp + H, z); ld4(gp + 2 * H, n);
          ld4(p.aux + si, hn);
          if (pin) ld4(p.hseq + sp, h_prev);
#pragma unroll
          for (int j = 0; j < NPF; ++j) {
            const float cn = (1.f - z[j]) * (1.f - n[j] * n[j]);
            k[0][j] = cn * hn[j] * r[j] * (1.f - r[j]); k[1][j] = (h_prev[j] - n[j]) * z[j] * (1.f - z[j]);
            k[2][j] = cn; k[3][j] = cn * r[j]; k[4][j] = z[j];
          }
        } else {
          float hv[NPF];
          ld4(p.hseq + si, hv);
#pragma unroll
          for (int j = 0; j < NPF; ++j) k[0][j] = 1.f - hv[j] * hv[j];
        }
      };
      float kc[6][NPF], pdy[NPF];
      bool pvalid = false;
      if (single) load_coefs(b_own, kc, pdy, pvalid);
      if (step > 0) {
        if (e == 0) mbar_arrive_expect_tx(part_bar, part_tx);   // arm this step's phase (peers may already have sent)
        // send this CTA's partial tile: row (16q + ul), 32 columns split over the two half-warps
        mbar_wait(accum_bar, acc_phase);
        tc_fence_after();
        if (RES) {
          // non-LL: this step's MMAs ran, so the grid barrier was passed: the maximum of step-1 the producer forwarded
          // is final.  LL: the four loader warps left the maximum over all (CTA, warp) words of step-1 in wmax[parity]
          // before their last full-barrier arrival (ordered by the full -> MMA -> accumulator barrier chain).
          unsigned int gm;
          if (LL) {
            const volatile unsigned int* wm = wmax + (step & 1) * 4;
            gm = max(max(wm[0], wm[1]), max(wm[2], wm[3]));
          } else {
            gm = *(volatile unsigned int*)(cta_max + 1);
          }
          sx_prev = sx_cur;
          sx_cur = pow2_exp_for(max(gm, dym), sx_prev);      // non-negative floats order like their bit patterns
          s_cur = pow2f(sx_cur);
          inv_prev = pow2f(-sx_prev);
        }
        if (e == 0) trace_stamp(p.trace, p.T, step, 5);
        for (int cb = 0; cb < NB; cb += 32) {
          float acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, acc);
          if (CL == 4) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float hi = __shfl_sync(0xffffffffu, acc[16 + j], ul);
              v[j] = half ? hi : acc[j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c0 = cb + half * 16 + 4 * i;      // NB is a multiple of 8: a group of 4 columns is in or out
              if (c0 < NB)
                st_async_v4(dst_row + (uint32_t)(xt_off(UT, ul, c0) * 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], dst_bar);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c0 = cb + 4 * i;
              if (c0 < NB)
                st_async_v4(dst_row + (uint32_t)(xt_off(UT, ul, c0) * 4), acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3], dst_bar);
            }
          }
        }
        acc_phase ^= 1;
        tc_fence_before();
        if (e == 0) trace_stamp(p.trace, p.T, step, 6);
        mbar_wait_cluster(part_bar, part_phase, p.err);   // all CL slices of this CTA's units have landed
        part_phase ^= 1;
      }
      if (e == 0) trace_stamp(p.trace, p.T, step, 7);
      // Four cells: recurrent term = sum of the four received slices, gate backward from the coefficients.  The
      // scaled fp16 copy (resident variant: the next step's MMA operand) is stored here; the fp32 gate gradients,
      // which only the weight-gradient GEMMs after the sweep read, are returned in o[] (o[4]: GRU dGh_n).
      auto st_h4 = [&](__half* dst, const float (&v)[NPF]) {     // 4 scaled, saturated halves = one 8-byte store
        const __half2 lo = __halves2half2(to_half_sat(v[0] * s_cur), to_half_sat(v[1] * s_cur));
        const __half2 hi = __halves2half2(to_half_sat(v[2] * s_cur), to_half_sat(v[3] * s_cur));
        uint2 pk;
        pk.x = *reinterpret_cast<const unsigned int*>(&lo);
        pk.y = *reinterpret_cast<const unsigned int*>(&hi);
        if (LL) st_relaxed_v2(dst, pk.x, pk.y);            // saturated at +-65000: never the 0xFFFF sentinel
        else *reinterpret_cast<uint2*>(dst) = pk;
      };
      auto finish4 = [&](int b, const float (&k)[6][NPF], const float (&dyv)[NPF], bool valid, float (&o)[5][NPF]) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < NPF; ++j) o[i][j] = 0.f;
        if (b >= B) return;
        __half* hp16 = RES ? p.dg16 + (((size_t)t * B + b) * D + d) * GH + u0 + uq : nullptr;
        if (!valid) {
          if (RES) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (LL) st_relaxed_v2(hp16 + g * H, 0u, 0u);
              else *reinterpret_cast<uint2*>(hp16 + g * H) = make_uint2(0u, 0u);
            }
          }
          return;
        }
        float dh[NPF];
#pragma unroll
        for (int j = 0; j < NPF; ++j) {
          float rec = 0.f;
          if (step > 0) {
            const float* pr = part + xt_off(UT, uq + j, b);
            const int ss = xt_slice(UT, NB);
            rec = (pr[0] + pr[ss]) + (pr[2 * ss] + pr[3 * ss]);
            if (CL == 8) rec += (pr[4 * ss] + pr[5 * ss]) + (pr[6 * ss] + pr[7 * ss]);
          }
          dh[j] = dyv[j] + (RES ? rec * inv_prev : rec);
        }
        if (RNN == DS2_RNN_LSTM) {
#pragma unroll
          for (int j = 0; j < NPF; ++j) {
            const int ci = (uq + j) * NBp + b;
            const float dc = fmaf(dh[j], k[0][j], cst[ci]);
            o[0][j] = dc * k[2][j]; o[1][j] = dc * k[3][j]; o[2][j] = dc * k[4][j]; o[3][j] = dh[j] * k[1][j];
            cst[ci] = dc * k[5][j];
            lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(o[0][j]), fabsf(o[1][j])), fmaxf(fabsf(o[2][j]), fabsf(o[3][j]))));
          }
          if (RES) {
#pragma unroll
            for (int g = 0; g < 4; ++g) st_h4(hp16 + g * H, o[g]);
          }
        } else if (RNN == DS2_RNN_GRU) {
#pragma unroll
          for (int j = 0; j < NPF; ++j) {
            const int ci = (uq + j) * NBp + b;
            const float dht = dh[j] + cst[ci];
            o[0][j] = dht * k[0][j]; o[1][j] = dht * k[1][j]; o[2][j] = dht * k[2][j]; o[4][j] = dht * k[3][j];
            cst[ci] = dht * k[4][j];
            lmax = fmaxf(lmax, fmaxf(fabsf(o[0][j]), fmaxf(fabsf(o[1][j]), fabsf(o[4][j]))));
          }
          if (RES) {
            st_h4(hp16, o[0]); st_h4(hp16 + H, o[1]);
            st_h4(hp16 + 2 * H, o[4]);                        // the h-side n-gate gradient (dGh_n)
          }
        } else {
#pragma unroll
          for (int j = 0; j < NPF; ++j) {
            o[0][j] = dh[j] * k[0][j];
            lmax = fmaxf(lmax, fabsf(o[0][j]));
          }
          if (RES) st_h4(hp16, o[0]);
        }
      };
      auto st4 = [](float* dst, const float (&v)[NPF]) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      };
      auto store_dg = [&](int b, const float (&o)[5][NPF]) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < NPF; ++j) bsum[i][j] += o[i][j];
        float* gp = p.gates + (((size_t)t * B + b) * D + d) * GH + u0 + uq;
#pragma unroll
        for (int g = 0; g < G; ++g) st4(gp + g * H, o[g]);
        if (RNN == DS2_RNN_GRU) st4(p.aux + (((size_t)d * T + t) * B + b) * H + u0 + uq, o[4]);
        if (RES && p.dgn16) {
          // precision-16 GEMM operands, produced where the values are: the row-major copy (dX) as 8-byte stores, the
          // transposed copies (dW_ih, dW_hh: the reduction index t*B+b must be contiguous) as 2-byte stores — 8
          // consecutive batch columns of a (gate, unit) come from 8 lanes of a warp and are merged in L2
          const size_t TBs = (size_t)T * B, col = (size_t)t * B + b;
          __half* r16 = p.dgn16 + (col * D + d) * GH + u0 + uq;
          __half* t16 = p.dgn16T + ((size_t)d * GH + u0 + uq) * TBs + col;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            __half h[NPF];
#pragma unroll
            for (int j = 0; j < NPF; ++j) {
              h[j] = to_half_sat(o[g][j] * nscale);
              t16[((size_t)g * H + j) * TBs] = h[j];
            }
            const __half2 lo = __halves2half2(h[0], h[1]), hi = __halves2half2(h[2], h[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned int*>(&lo);
            pk.y = *reinterpret_cast<const unsigned int*>(&hi);
            *reinterpret_cast<uint2*>(r16 + g * H) = pk;
          }
          if (RNN == DS2_RNN_GRU) {
            __half* a16 = p.auxn16T + ((size_t)d * H + u0 + uq) * TBs + col;
#pragma unroll
            for (int j = 0; j < NPF; ++j) a16[(size_t)j * TBs] = to_half_sat(o[4][j] * nscale);
          }
        }
      };
      // LL: fetch the scaled fp16 gate gradients dGh[t] of ALL units of this CTA's K range (written by every CTA of
      // the direction) into the swizzled K-major tile of the next step's MMAs; same packet / group scheme as the
      // forward sweep (fetch_h).  Before the last group is handed over, the per-(CTA, warp) maxima of the step are
      // polled too and their maximum is left in wmax[next parity][warp] for the scale of the next step.
      auto fetch_dg = [&](int t_src) {
        const size_t ld = (size_t)D * GH;
        const __half* src = p.dg16 + (size_t)t_src * B * ld + (size_t)d * GH + kbase;
        const uint32_t vb = smem_u32(smem + NKR * A_BYTES);
        for (int g = 0; g < NG; ++g) {
          const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
          const int npk = (c1 - c0) * 8, total = B * npk;
          for (int base = 0; base < total; base += 128 * 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int idx = base + k * 128 + e;
              if (idx < total) v[k] = ld_relaxed_v4(src + (size_t)(idx / npk) * ld + (size_t)(c0 * 8 + idx % npk) * 8);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int idx = base + k * 128 + e;
              if (idx < total) {
                const int row = idx / npk, pc = c0 * 8 + idx % npk;
                ll_wait(v[k], src + (size_t)row * ld + (size_t)pc * 8, p.err);
                st_shared_v4(vb + (uint32_t)((pc >> 3) * B_BYTES + row * 128 + (((pc & 7) ^ (row & 7)) << 4)), v[k]);
              }
            }
          }
          if (g == NG - 1) {
            const unsigned int* mp = p.gmeta + ((size_t)d * T + step) * nmeta;
            unsigned int m = 0u;
            for (int i = e; i < nmeta; i += 128) m = max(m, ll_wait_u32(mp + i, p.err));
            m = __reduce_max_sync(0xffffffffu, m);
            if (lane == 0) wmax[((step + 1) & 1) * 4 + q] = m;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(full + g);
          if (e == 0 && g == 0) trace_stamp(p.trace, p.T, step + 1, 1);
        }
      };
      // the non-resident variants stream the fp32 gate gradients themselves: nothing can be deferred there
      const bool defer = RES && (LL || p.defer) && single;
      float sv[5][NPF];
      if (single) {
        finish4(b_own, kc, pdy, pvalid, sv);
        if (!defer && b_own < B) store_dg(b_own, sv);
      } else {
        for (int b = b_own; b < B; b += 32) {
          load_coefs(b, kc, pdy, pvalid);
          finish4(b, kc, pdy, pvalid, sv);
          store_dg(b, sv);
        }
      }
      if (RES) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if (LL) {           // one word per (CTA, epilogue warp): no intra-CTA reduction on the critical path
          unsigned int lb = __float_as_uint(lmax);
          if (lb >= 0x7f800000u) lb = 0x7f7fffffu;                    // inf / nan: finite, and never the sentinel
          if (lane == 0) st_relaxed_u32(p.gmeta + ((size_t)d * T + step) * nmeta + cta_in_dir * 4 + q, lb);
        } else if (lane == 0) {
          atomicMax(cta_max, __float_as_uint(lmax));                  // non-negative floats order like uints
        }
      }
      if (e == 0) trace_stamp(p.trace, p.T, step, 8);
      if (!LL) {
        named_bar_sync(1, 128);
        if (e == 0) {
          trace_stamp(p.trace, p.T, step, 9);
          if (RES) {
            atomicMax(p.gmax + (size_t)d * (T + 1) + step + 1, cta_max[0]);
            cta_max[0] = 0u;
          }
          fence_proxy_async_global();
          trace_stamp(p.trace, p.T, step, 10);
          red_release(ctr, 1u);
          trace_stamp(p.trace, p.T, step, 11);
          trace_stamp_ns(p.trace, p.T, step, 12);
        }
      } else if (e == 0) {
        trace_stamp(p.trace, p.T, step, 11);
        trace_stamp_ns(p.trace, p.T, step, 12);
      }
      if (defer) {
        if (b_own < B) store_dg(b_own, sv);
        if (e == 0) trace_stamp(p.trace, p.T, step, 13);
      }
      if (LL && step + 1 < T) {        // the other CTAs' gate gradients land while the fp32 stores above drain
        if (e == 0) trace_stamp(p.trace, p.T, step + 1, 0);
        fetch_dg(t);
      }
    }
    if (p.dbias[d]) {
      // lanes with equal (lane & 3) hold the same 4 units for different batch columns: reduce over lane bits 2..4,
      // then the four warps through shared memory in a fixed order (the exchange tiles are free now).  Every
      // (gate, unit) of these 16 units belongs to this CTA alone: plain stores, bit-repeatable, no atomics.
      named_bar_sync(1, 128);                              // everybody is done reading `part`
      float* red = part;                                   // [4 warps][5][16]
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = bsum[i][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (lane < 4) red[(q * 5 + i) * 16 + 4 * lane + j] = v;
        }
      }
      named_bar_sync(1, 128);
      if (e < 80) {
        const int i = e / 16, u = e % 16;
        const float v = (red[(0 * 5 + i) * 16 + u] + red[(1 * 5 + i) * 16 + u]) +
                        (red[(2 * 5 + i) * 16 + u] + red[(3 * 5 + i) * 16 + u]);
        if (i < G) p.dbias[d][(size_t)i * H + u0 + u] = v;
        else if (RNN == DS2_RNN_GRU && i == 4) p.dbias_hn[d][u0 + u] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // nobody exits while a peer may still read its tile
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Forward sweep, split-K variant (LSTM / GRU, resident fp16 weights).  A 2-CTA cluster owns 32 hidden units of
// one direction; CTA `rank` multiplies the G*32 gate rows (MMA M = 128, rows ordered [owner CTA][gate][unit])
// with its half of K = H, so a step issues H/32 MMAs instead of H/16 — the single-thread MMA issue chain is the
// longest part of a step.  TMEM lanes 0..63 hold the rows CTA 0 finishes, lanes 64..127 those of CTA 1: every
// epilogue warp pushes its 32 accumulator rows into the owner's shared memory with st.async (complete_tx on
// the owner's mbarrier); the owner adds the two partial tiles, the input projection and the biases, and runs
// gates + cell update for its 16 units in one pass (thread = 4 consecutive units x one batch column).
// NKR_T: compile-time number of K chunks per CTA (0 = runtime): with constant chunk offsets the MMA descriptors are
// "uniform base + immediate" and the issue loop needs no vector arithmetic / R2UR per instruction.
template <int RNN, bool LL, int NKR_T = 0>
__global__ void __launch_bounds__(rp::THREADS, 1) rnn_fwd_splitk_kernel(const __grid_constant__ PersistParams p) {
  using namespace rp;
  using namespace tc;
  constexpr int G = RNN == DS2_RNN_LSTM ? 4 : 3;
  constexpr int AW = 128 * 128;                          // bytes of one K chunk of the weight tile (128 rows)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still __shared__
  const int NB = p.NB, B = p.B, T = p.T, H = p.H, D = p.D;
  const int B_BYTES = NB * 128;
  const int NBp = NB + 1;
  const int NKR = NKR_T ? NKR_T : H / 128;               // 64-wide fp16 chunks of this CTA's K half
  const int NG = grp_count(NKR);
  float* part = reinterpret_cast<float*>(smem + NKR * (AW + B_BYTES));   // [2 sources] 64-row partial tiles (xt_* layout)
  float* cst = part + 2 * xt_slice(64, NB);                              // [16][NBp] cell (LSTM) / hidden (GRU) state
  int* lens_s = reinterpret_cast<int*>(cst + UT * NBp);
  uint64_t* full = reinterpret_cast<uint64_t*>(lens_s + ((NB + 1) & ~1));   // one per group of 4 chunks (<= 8)
  uint64_t* wbar = full + 8;
  uint64_t* accum_bar = wbar + 1;
  uint64_t* part_bar = accum_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = blockIdx.x & 1, cl = blockIdx.x >> 1;
  const int NTc = H / 32;                                // clusters per direction
  const int d = p.d0 + cl / NTc, U0 = (cl % NTc) * 32;   // first unit of the cluster
  const int u0 = U0 + rank * UT;                         // the 16 units this CTA finishes
  const int GH = G * H;
  unsigned int* ctr = p.bar + 32 * d;
  const unsigned int n_arrive = (unsigned int)(NTc * 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[d]);
    if (!LL) tma_prefetch_desc(&p.tmV[d]);
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], LL ? 4 : 1);   // LL: one arrival per loader (= epilogue) warp
    mbar_init(wbar, 1);
    mbar_init(accum_bar, 1);
    mbar_init(part_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < UT * NBp; i += THREADS) cst[i] = 0.f;
  for (int i = threadIdx.x; i < NB; i += THREADS) lens_s[i] = i < B ? p.len[i] : 0;
  if (LL) {   // batch-padding rows of the streamed tile are never loaded: keep them finite (their columns are unused)
    uint8_t* hb = smem + NKR * AW;
    for (int i = threadIdx.x; i < NKR * (NB - B) * 8; i += THREADS) {
      const int c = i / ((NB - B) * 8), r = B + (i / 8) % (NB - B), j = i % 8;
      *reinterpret_cast<uint4*>(hb + c * B_BYTES + r * 128 + j * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                                     // the peer's mbarriers are initialised
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0 && threadIdx.x == 0) {               // literal TMEM address 0 in the MMAs, see the other kernels
    *(volatile int*)p.err = 2;
    printf("ds2: unexpected TMEM base %u (block %d)\n", tmem_base, blockIdx.x);
  }
  const int kc0 = rank * NKR;                             // first K chunk (of H/64) of this CTA

  if (warp == 0) {
    if (lane == 0) {
      // weights once: per chunk the rows of the two owners, [gate][16 units] each (rows >= 16 G of a half unused)
      mbar_arrive_expect_tx(wbar, (uint32_t)(NKR * 2 * G * UT * 128));
      for (int c = 0; c < NKR; ++c) {
        tma_load_3d(smem + c * AW, &p.tmW[d], wbar, (kc0 + c) * 64, U0, 0);
        tma_load_3d(smem + c * AW + 64 * 128, &p.tmW[d], wbar, (kc0 + c) * 64, U0 + UT, 0);
      }
      uint8_t* hbuf = smem + NKR * AW;
      for (int step = 1; !LL && step < T; ++step) {       // LL: the epilogue warps fetch h_{t-1} themselves (below)
        const int t = d == 0 ? step : T - 1 - step;
        const int tp = d == 0 ? t - 1 : t + 1;
        grid_wait_counter(ctr, n_arrive * (unsigned int)step, p.err);
        fence_proxy_async_global();
        trace_stamp(p.trace, p.T, step, 0);
        for (int g = 0; g < NG; ++g) {
          uint64_t* fb = full + g;
          const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
          if (p.box3) {
            mbar_arrive_expect_tx(fb, (uint32_t)(4 * B_BYTES));
            tma_load_3d(hbuf + c0 * B_BYTES, &p.tmV3[d], fb, 0, tp * B, kc0 + c0);
          } else {
            mbar_arrive_expect_tx(fb, (uint32_t)((c1 - c0) * B * 128));
            for (int c = c0; c < c1; ++c) tma_load_2d(hbuf + c * B_BYTES, &p.tmV[d], fb, (kc0 + c) * 64, tp * B);
          }
        }
        trace_stamp(p.trace, p.T, step, 1);
      }
    }
  } else if (warp == 1) {
    {   // the whole warp runs the issue loop converged (mma_f16_w: elect.sync inside): 32 MMAs per step in ~1.2k
        // cycles instead of ~2.4k from a single divergent lane
      const uint32_t idesc = instr_desc(FMT_F16, 128, NB);
      const uint64_t a_base = warp_uniform(smem_desc_sw128(smem_u32(smem)));
      const uint64_t b_base = warp_uniform(smem_desc_sw128(smem_u32(smem + NKR * AW)));
      const uint64_t a_step = (uint64_t)(AW >> 4), b_step = (uint64_t)(B_BYTES >> 4);
      mbar_wait(wbar, 0);
      uint32_t ph = 0;
      for (int step = 1; step < T; ++step) {
        if constexpr (NKR_T > 0) {
          constexpr int NGT = (NKR_T + 3) / 4;
#pragma unroll
          for (int g = 0; g < NGT; ++g) {
            mbar_wait(full + g, ph);
            tc_fence_after();
            if (g == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
            if (g == NGT - 1 && lane == 0) trace_stamp(p.trace, p.T, step, 3);
#pragma unroll
            for (int c = 4 * g; c < (4 * g + 4 < NKR_T ? 4 * g + 4 : NKR_T); ++c) {
              const uint64_t ad = a_base + (uint64_t)c * a_step, bd = b_base + (uint64_t)c * b_step;
              mma_f16_w(0u, ad, bd, idesc, c > 0);
              mma_f16_w(0u, ad + 2, bd + 2, idesc, 1);
              mma_f16_w(0u, ad + 4, bd + 4, idesc, 1);
              mma_f16_w(0u, ad + 6, bd + 6, idesc, 1);
            }
          }
        } else {
          for (int g = 0; g < NG; ++g) {
            mbar_wait(full + g, ph);
            tc_fence_after();
            if (g == 0 && lane == 0) trace_stamp(p.trace, p.T, step, 2);
            const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
            if (c1 == NKR && lane == 0) trace_stamp(p.trace, p.T, step, 3);
            for (int c = c0; c < c1; ++c) {
              const uint64_t ad = a_base + (uint64_t)c * a_step, bd = b_base + (uint64_t)c * b_step;
              mma_f16_w(0u, ad, bd, idesc, c > 0);
              mma_f16_w(0u, ad + 2, bd + 2, idesc, 1);
              mma_f16_w(0u, ad + 4, bd + 4, idesc, 1);
              mma_f16_w(0u, ad + 6, bd + 6, idesc, 1);
            }
          }
        }
        mma_commit_w(accum_bar);
        if (lane == 0) trace_stamp(p.trace, p.T, step, 4);
        ph ^= 1;
      }
    }
  } else {
    const int q = warp % 4;
    const int e = threadIdx.x - 64;          // 0..127
    // TMEM lane (= row) 32q + lane belongs to CTA q/2; inside that CTA's 64-row slice it is row 32 (q&1) + lane
    const int dst_cta = q >> 1;
    const int my_row = 32 * (q & 1) + lane;
    const uint32_t dst_row = mapa_u32(smem_u32(part), (uint32_t)dst_cta) + (uint32_t)(rank * xt_slice(64, NB) * 4);
    const uint32_t dst_bar = mapa_u32(smem_u32(part_bar), (uint32_t)dst_cta);
    const uint32_t part_tx = (uint32_t)(2 * 64 * NB * 4);
    const int uq = 4 * (e & 3), b_own = e >> 2;
    // biases of this thread's 4 units: x-side + h-side summed, except the GRU n gate (r multiplies the h side)
    float bsum[G][4], bhn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float bx = p.b_ih[d][g * H + u0 + uq + j], bh = p.b_hh[d][g * H + u0 + uq + j];
        bsum[g][j] = (RNN == DS2_RNN_GRU && g == 2) ? bx : bx + bh;
        if (RNN == DS2_RNN_GRU && g == 2) bhn[j] = bh;
      }
      if (RNN != DS2_RNN_GRU) bhn[j] = 0.f;
    }
    auto ld4 = [](const float* src, float (&v)[4]) {
      const float4 x = *reinterpret_cast<const float4*>(src);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    };
    auto st4 = [](float* dst, const float (&v)[4]) { *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]); };
    uint32_t acc_phase = 0, part_phase = 0;
    const bool single = B <= 32;
    int step_for_trace = 0;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? step : T - 1 - step;
      // input projections of this thread's cells: independent of the recurrence, fetched before the MMA wait
      float gx[G][4];
      if (single && b_own < B) {
#pragma unroll
        for (int g = 0; g < G; ++g) ld4(p.gates + (((size_t)t * B + b_own) * D + d) * GH + (size_t)g * H + u0 + uq, gx[g]);
      }
      if (step > 0) {
        if (e == 0) mbar_arrive_expect_tx(part_bar, part_tx);   // arm this step's phase (the peer may already have sent)
        mbar_wait(accum_bar, acc_phase);
        tc_fence_after();
        if (e == 0) trace_stamp(p.trace, p.T, step, 5);
        for (int cb = 0; cb < NB; cb += 32) {
          float acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, acc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c0 = cb + 4 * i;                    // NB is a multiple of 8: a group of 4 columns is in or out
            if (c0 < NB)
              st_async_v4(dst_row + (uint32_t)(xt_off(64, my_row, c0) * 4), acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3], dst_bar);
          }
        }
        acc_phase ^= 1;
        tc_fence_before();
        if (e == 0) trace_stamp(p.trace, p.T, step, 6);
        mbar_wait_cluster(part_bar, part_phase, p.err);   // both partial tiles of this CTA's rows have landed
        part_phase ^= 1;
      }
      if (e == 0) trace_stamp(p.trace, p.T, step, 7);
      // ---- gates + cell update: o[0..G-1] saved gate values, o[4] aux (LSTM c / GRU h_n + b_hn), o[5] h
      auto cell4 = [&](int b, const float (&x)[G][4], float (&o)[6][4]) {
        const bool valid = t < lens_s[b];
        float pre[G][4], hn[4], hval[4];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float r = 0.f;
            if (step > 0) {
              const float* pr = part + xt_off(64, g * UT + uq + j, b);
              r = pr[0] + pr[xt_slice(64, NB)];
            }
            if (RNN == DS2_RNN_GRU && g == 2) { hn[j] = r + bhn[j]; pre[g][j] = x[g][j] + bsum[g][j]; }
            else pre[g][j] = (x[g][j] + r) + bsum[g][j];
          }
        if (RNN == DS2_RNN_LSTM) {
          float a[4][4], cval[4], th[4];
          // i, f, o: sigmoid; g: tanh = 2 sigmoid(2x) - 1  (one EX2 + one RCP per value, stage by stage)
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[g][j] = ex2_ftz((g == 2 ? -2.f * LOG2E : -LOG2E) * pre[g][j]);
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[g][j] = rcp_ftz(1.f + a[g][j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a[2][j] = fmaf(2.f, a[2][j], -1.f);
            cval[j] = valid ? fmaf(a[1][j], cst[(uq + j) * NBp + b], a[0][j] * a[2][j]) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) th[j] = ex2_ftz(-2.f * LOG2E * cval[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) th[j] = rcp_ftz(1.f + th[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            hval[j] = valid ? a[3][j] * fmaf(2.f, th[j], -1.f) : 0.f;
            if (valid) cst[(uq + j) * NBp + b] = cval[j];
#pragma unroll
            for (int g = 0; g < 4; ++g) o[g][j] = valid ? a[g][j] : 0.f;
            o[4][j] = cval[j];
          }
        } else {
          float a[2][4], nv[4];
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[g][j] = ex2_ftz(-LOG2E * pre[g][j]);
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[g][j] = rcp_ftz(1.f + a[g][j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) nv[j] = ex2_ftz(-2.f * LOG2E * fmaf(a[0][j], hn[j], pre[2][j]));
#pragma unroll
          for (int j = 0; j < 4; ++j) nv[j] = rcp_ftz(1.f + nv[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float nval = valid ? fmaf(2.f, nv[j], -1.f) : 0.f;
            const float hprev = cst[(uq + j) * NBp + b];
            hval[j] = valid ? fmaf(a[1][j], hprev - nval, nval) : 0.f;
            if (valid) cst[(uq + j) * NBp + b] = hval[j];
            o[0][j] = valid ? a[0][j] : 0.f; o[1][j] = valid ? a[1][j] : 0.f; o[2][j] = nval; o[3][j] = 0.f;
            o[4][j] = valid ? hn[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[5][j] = hval[j];
        const __half2 lo = __floats2half2_rn(hval[0], hval[1]), hi = __floats2half2_rn(hval[2], hval[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const unsigned int*>(&lo);
        pk.y = *reinterpret_cast<const unsigned int*>(&hi);
        __half* hdst = p.h16 + (((size_t)d * T + t) * B + b) * H + u0 + uq;
        if (LL) st_relaxed_v2(hdst, pk.x, pk.y);          // the values are their own ready flags (|h| < 1: never 0xFFFF)
        else *reinterpret_cast<uint2*>(hdst) = pk;
      };
      auto store4 = [&](int b, const float (&o)[6][4]) {
        const size_t so = (((size_t)d * T + t) * B + b) * H + u0 + uq;
        float* gp = p.gates + (((size_t)t * B + b) * D + d) * GH + u0 + uq;
        st4(p.aux + so, o[4]);
        if (p.training) {
#pragma unroll
          for (int g = 0; g < G; ++g) st4(gp + g * H, o[g]);
        }
        st4(p.hseq + so, o[5]);
      };
      // LL: fetch h_t of ALL units of this CTA's K half (written by the 32 owner CTAs of those units) into the
      // swizzled K-major tile of the next step's MMA.  Called by the 128 epilogue threads after accum_bar of this
      // step completed, i.e. when the MMAs that read the tile are done.  One 16-byte packet = 8 consecutive k of
      // one batch row; consecutive threads take consecutive packets of a row (512 contiguous bytes per warp
      // instruction); up to 8 polls in flight per thread; a group of 4 K chunks is handed to the MMA thread as soon
      // as it is complete (fence.proxy.async: generic-proxy stores -> async-proxy reads of tcgen05.mma).
      auto fetch_h = [&](int t_src) {
        const __half* src = p.h16 + ((size_t)d * T + t_src) * B * H + (size_t)kc0 * 64;
        const uint32_t hb = smem_u32(smem + NKR * AW);
        for (int g = 0; g < NG; ++g) {
          const int c0 = grp_begin(g), c1 = min(NKR, grp_begin(g + 1));
          const int npk = (c1 - c0) * 8, total = B * npk;
          for (int base = 0; base < total; base += 128 * 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int idx = base + k * 128 + e;
              if (idx < total) v[k] = ld_relaxed_v4(src + (size_t)(idx / npk) * H + (size_t)(c0 * 8 + idx % npk) * 8);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int idx = base + k * 128 + e;
              if (idx < total) {
                const int row = idx / npk, pc = c0 * 8 + idx % npk;
                ll_wait(v[k], src + (size_t)row * H + (size_t)pc * 8, p.err);
                st_shared_v4(hb + (uint32_t)((pc >> 3) * B_BYTES + row * 128 + (((pc & 7) ^ (row & 7)) << 4)), v[k]);
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(full + g);
          if (e == 0 && g == 0) trace_stamp(p.trace, p.T, step_for_trace, 1);
        }
      };
      const bool defer = LL || (p.defer && single);
      float sv[6][4];
      if (single) {
        if (b_own < B) {
          cell4(b_own, gx, sv);
          if (!defer) store4(b_own, sv);
        }
      } else {
        for (int b = b_own; b < B; b += 32) {
#pragma unroll
          for (int g = 0; g < G; ++g) ld4(p.gates + (((size_t)t * B + b) * D + d) * GH + (size_t)g * H + u0 + uq, gx[g]);
          cell4(b, gx, sv);
          store4(b, sv);
        }
      }
      if (e == 0) trace_stamp(p.trace, p.T, step, 8);
      if (!LL) {
        named_bar_sync(1, 128);        // CTA-scope: every epilogue thread's stores happen-before thread 0's release
        if (e == 0) {
          trace_stamp(p.trace, p.T, step, 9);
          fence_proxy_async_global();
          trace_stamp(p.trace, p.T, step, 10);
          red_release(ctr, 1u);
          trace_stamp(p.trace, p.T, step, 11);
          trace_stamp_ns(p.trace, p.T, step, 12);
        }
      } else if (e == 0) {
        trace_stamp(p.trace, p.T, step, 11);
        trace_stamp_ns(p.trace, p.T, step, 12);
      }
      if (defer && single) {
        if (b_own < B) store4(b_own, sv);
        if (e == 0) trace_stamp(p.trace, p.T, step, 13);
      }
      if (LL && step + 1 < T) {        // the other CTAs' h_t lands while the fp32 stores above drain
        step_for_trace = step + 1;
        if (e == 0) trace_stamp(p.trace, p.T, step + 1, 0);
        fetch_h(t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // nobody exits while the peer may still write into its tile
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

static size_t fwd_splitk_smem_bytes(int NB, int H) {
  using namespace rp;
  size_t NBp = NB + 1;
  return 1024 + (size_t)(H / 128) * (128 * 128 + (size_t)NB * 128) + (2 * xt_slice(64, NB) + UT * NBp + NB + 4) * sizeof(float) +
         12 * sizeof(uint64_t) + 64;
}

// returns 1 when the shape / device does not take this variant (the caller then uses the 16-unit resident kernel)
template <int RNN, bool LL>
static int launch_fwd_splitk(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  constexpr int G = RNN == DS2_RNN_LSTM ? 4 : 3;
  if (a.H % 128 != 0 || a.H / 128 > 32) return 1;
  if (!vec_ok(a.gates, a.hseq, a.aux) || !a.aux) return 1;
  if (ws_bytes < res_ws_bytes(G, a.T, a.B, a.H, a.D)) return 1;
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 15) / 16 * 16;   // MMA M = 128 needs N % 16 == 0
  p.H = a.H; p.D = a.D; p.NT = a.H / 32; p.G = G;
  p.training = a.training;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux;
  p.trace = trace_ptr_from_env("DS2_TRACE_FWD");
  p.defer = sweep_defer_default();
  set_acc_layout(p);
  p.err = static_cast<int*>(ws);
  p.bar = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 128);
  __half* w16 = reinterpret_cast<__half*>(static_cast<char*>(ws) + 4096);
  p.h16 = reinterpret_cast<__half*>(static_cast<char*>(ws) + 4096 + align_up((size_t)a.D * G * a.H * a.H * 2, 256));
  const size_t smem = one_cta_per_sm(fwd_splitk_smem_bytes(p.NB, a.H));
  if (smem > 227 * 1024) return 1;
  auto kern = (a.H == 1024) ? rnn_fwd_splitk_kernel<RNN, LL, 8> : rnn_fwd_splitk_kernel<RNN, LL, 0>;   // H = 1024: unrolled issue loop
  static DeviceOnce attr_once;
  if (attr_once.first()) {   // BOTH instantiations: whichever shape comes first must not leave the other without its opt-in
    DS2_CHECK_CUDA(cudaFuncSetAttribute(rnn_fwd_splitk_kernel<RNN, LL, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DS2_CHECK_CUDA(cudaFuncSetAttribute(rnn_fwd_splitk_kernel<RNN, LL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  int grid = a.D * p.NT * 2, launches = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = noncoop_cluster_launch() ? 1 : 2;     // see the backward launcher (Nsight Compute)
  int max_clusters = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (oe != cudaSuccess) { (void)cudaGetLastError(); return 1; }
  if (max_clusters * 2 < grid) {
    if (max_clusters * 2 < p.NT * 2) return 1;
    grid = p.NT * 2;
    launches = a.D;
    cfg.gridDim = dim3(grid);
  }
  const size_t wn = (size_t)G * a.H * a.H;
  for (int d = 0; d < a.D; ++d) {
    p.b_ih[d] = a.b_ih[d];
    p.b_hh[d] = a.b_hh[d];
    DS2_LAUNCH(f32_to_f16_kernel, 148 * 4, 256, 0, st, wn, a.w_hh[d], w16 + (size_t)d * wn);
    int rc = make_tmap_f16(&p.tmW[d], w16 + (size_t)d * wn, 3, a.H, a.H, G, (size_t)a.H, (size_t)a.H * a.H, 64, UT, G);
    if (rc) return rc;
    rc = make_tmap_f16(&p.tmV[d], p.h16 + (size_t)d * a.T * a.B * a.H, 2, a.H, a.T * a.B, 1, (size_t)a.H, 0, 64, a.B, 1);
    if (rc) return rc;
    p.box3 = (a.H / 128) % 4 == 0;
    if (p.box3) {
      rc = make_tmap_f16(&p.tmV3[d], p.h16 + (size_t)d * a.T * a.B * a.H, 3, 64, a.T * a.B, a.H / 64, (size_t)a.H, 64, 64,
                         p.NB, 4);
      if (rc) return rc;
    }
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096, st));
  if (LL)   // every 2-byte element of the h stream is its own "not yet written" flag
    DS2_CHECK_CUDA(cudaMemsetAsync(p.h16, 0xFF, (size_t)a.D * a.T * a.B * a.H * sizeof(__half), st));
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) {
      (void)cudaGetLastError();
      if (li == 0) {
        fprintf(stderr, "ds2_b200: WARNING split-K forward sweep launch failed (%s); using the 16-unit kernel\n",
                cudaGetErrorString(le));
        return 1;
      }
      set_error("split-K forward sweep: second launch failed: %s", cudaGetErrorString(le));
      return DS2_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  return DS2_OK;
}

static size_t splitk_smem_bytes(int NB, int CL) {
  using namespace rp;
  size_t NBp = NB + 1;
  return 1024 + (size_t)STAGES * ((size_t)UT * CL * 128 + (size_t)NB * 128) +
         ((size_t)CL * xt_slice(UT, NB) + UT * NBp + NB + 16) * sizeof(float) + (2 * STAGES + 3) * sizeof(uint64_t) + 64;
}

// max |x| over n floats -> atomicMax on float bits (x >= 0 after fabs)
__global__ void absmax_kernel(size_t n, const float* __restrict__ x, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x % 32 == 0) atomicMax(out, __float_as_uint(m));
}

// out[t] = float bits of max |x[t, :]| over the n floats of row t (one CTA per row, x >= 0 after fabs -> uint order)
__global__ void absmax_rows_kernel(size_t n, const float* __restrict__ x, unsigned int* __restrict__ out) {
  __shared__ float red[8];
  const float* row = x + (size_t)blockIdx.x * n;
  float m = 0.f;
  for (size_t i = threadIdx.x * 4; i + 3 < n; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (size_t i = (n & ~(size_t)3) + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x / 32); ++w) m = fmaxf(m, red[w]);
    out[blockIdx.x] = __float_as_uint(m);
  }
}

static size_t splitk_res_smem_bytes(int NB, int Kc, int CL) {
  using namespace rp;
  size_t NBp = NB + 1;
  return 1024 + (size_t)(Kc / 64) * ((size_t)UT * CL * 128 + (size_t)NB * 128) +
         ((size_t)CL * xt_slice(UT, NB) + UT * NBp + NB + 16) * sizeof(float) + (32 + STAGES + 3) * sizeof(uint64_t) + 64;
}
// workspace of the resident backward: [4 KB control][gmax D*(T+1) uints][W^T fp16: D*H*GH][dg16: T*B*D*GH]
// (+ [dymax: T uints][gmeta: D*T*(H/16)*4 uints] after gmax: H/16 CTAs per direction, 4 epilogue warps each)
static size_t splitk_res_ws_bytes(int G, int T, int B, int H, int D) {
  const size_t GH = (size_t)G * H;
  return 4096 + align_up((size_t)D * (T + 1) * 4, 256) + align_up((size_t)T * 4, 256) +
         align_up((size_t)D * T * (H / 16) * 4 * 4, 256) + align_up((size_t)D * H * GH * 2, 256) +
         align_up((size_t)T * B * D * GH * 2, 256);
}

template <int RNN, int CL, bool LL>
static int launch_bwd_splitk_resident(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  const int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  const int GH = G * a.H;
  constexpr int UM = UT * CL;
  if (a.H % UM != 0 || GH % CL != 0 || (GH / CL) % 64 != 0 || (GH / CL) / 64 > 120) return 1;   // <= 32 group barriers
  if (CL == 8 && ((a.B + 7) / 8 * 8) % 16 != 0) return 1;                                       // M = 128: N % 16 == 0
  if (!vec_ok(a.gates, a.hseq, a.aux, a.dy)) return 1;
  if (ws_bytes < splitk_res_ws_bytes(G, a.T, a.B, a.H, a.D)) return 1;
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 7) / 8 * 8; p.H = a.H; p.D = a.D; p.NT = a.H / UM; p.G = G;
  p.training = 1;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux; p.dy = a.dy;
  p.trace = trace_ptr_from_env("DS2_TRACE_BWD");
  p.defer = sweep_defer_default();
  for (int d = 0; d < a.D; ++d) { p.dbias[d] = a.dbias[d]; p.dbias_hn[d] = a.dbias_hn[d]; }
  if (a.f16_dg && a.f16_dgT && a.f16_scale && (RNN != DS2_RNN_GRU || a.f16_auxT)) {
    p.dgn16 = static_cast<__half*>(a.f16_dg);
    p.dgn16T = static_cast<__half*>(a.f16_dgT);
    p.auxn16T = static_cast<__half*>(a.f16_auxT);
    p.nscale = a.f16_scale;
  }
  set_acc_layout(p);
  char* base = static_cast<char*>(ws);
  p.err = reinterpret_cast<int*>(base);
  p.bar = reinterpret_cast<unsigned int*>(base + 128);
  size_t off = 4096;
  p.gmax = reinterpret_cast<unsigned int*>(base + off); off += align_up((size_t)a.D * (a.T + 1) * 4, 256);
  p.dymax = reinterpret_cast<unsigned int*>(base + off); off += align_up((size_t)a.T * 4, 256);
  p.gmeta = reinterpret_cast<unsigned int*>(base + off);
  const size_t gmeta_bytes = (size_t)a.D * a.T * (a.H / 16) * 4 * 4;
  off += align_up(gmeta_bytes, 256);
  __half* wT16 = reinterpret_cast<__half*>(base + off); off += align_up((size_t)a.D * a.H * GH * 2, 256);
  p.dg16 = reinterpret_cast<__half*>(base + off);
  const size_t smem = one_cta_per_sm(splitk_res_smem_bytes(p.NB, GH / CL, CL));
  if (smem > 227 * 1024) return 1;
  // H = 1024 (the BASELINE shapes): compile-time chunk count -> unrolled issue loop
  constexpr int NKU = (G * 1024 / CL) / 64;   // chunks per CTA at H = 1024: 16 / 12 / 4 (CL 4), 8 / 6 / 2 (CL 8)
  auto kern = (!LL && a.H == 1024) ? rnn_bwd_splitk_kernel<RNN, true, CL, LL, NKU> : rnn_bwd_splitk_kernel<RNN, true, CL, LL, 0>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {   // BOTH instantiations (see the forward launcher)
    DS2_CHECK_CUDA(cudaFuncSetAttribute(rnn_bwd_splitk_kernel<RNN, true, CL, LL, NKU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DS2_CHECK_CUDA(cudaFuncSetAttribute(rnn_bwd_splitk_kernel<RNN, true, CL, LL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  int grid = a.D * p.NT * CL, launches = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CL; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  // Nsight Compute cannot replay a cooperative cluster launch: DS2_SPLITK_NONCOOP=1 (profiling only) drops the
  // cooperative attribute; co-residency is still checked with cudaOccupancyMaxActiveClusters below.
  cfg.numAttrs = noncoop_cluster_launch() ? 1 : 2;
  int max_clusters = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (oe != cudaSuccess) { (void)cudaGetLastError(); return 1; }
  if (max_clusters * CL < grid) {
    // 8-CTA clusters are only worth it when both directions run concurrently (B200: 16 clusters of 8 do not fit the
    // GPCs, measured: 10.4k cycles per step instead of 11.7k, but the two directions then run back to back)
    if (CL == 8 || max_clusters * CL < p.NT * CL) return 1;
    grid = p.NT * CL;
    launches = a.D;
    cfg.gridDim = dim3(grid);
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096 + align_up((size_t)a.D * (a.T + 1) * 4, 256), st));
  const size_t wn = (size_t)a.H * GH;
  const bool cached = a.w_hhT16[0] && (a.D == 1 || a.w_hhT16[1]);   // fp16 W_hh^T left by the forward pass
  if (!cached) {
    int rc = materialize_w_hh(a, st);
    if (rc) return rc;
  }
  for (int d = 0; d < a.D; ++d) {
    const __half* wsrc = static_cast<const __half*>(a.w_hhT16[d]);
    if (!cached) {
      DS2_LAUNCH(f32_to_f16_kernel, 148 * 4, 256, 0, st, wn, a.w_hh[d], wT16 + (size_t)d * wn);
      wsrc = wT16 + (size_t)d * wn;
    }
    int rc = make_tmap_f16(&p.tmW[d], wsrc, 2, GH, a.H, 1, (size_t)GH, 0, 64, UM, 1);
    if (rc) return rc;
    rc = make_tmap_f16(&p.tmV[d], p.dg16, 2, a.D * GH, a.T * a.B, 1, (size_t)a.D * GH, 0, 64, a.B, 1);
    if (rc) return rc;
    p.box3 = ((GH / CL) / 64) % 4 == 0;
    if (p.box3) {
      rc = make_tmap_f16(&p.tmV3[d], p.dg16, 3, 64, a.T * a.B, a.D * GH / 64, (size_t)a.D * GH, 64, 64, p.NB, 4);
      if (rc) return rc;
    }
  }
  // per-time-step max |dY[t]|: the step-0 scale, and part of every later step's scale (spiky upstream gradients)
  DS2_LAUNCH(absmax_rows_kernel, a.T, 256, 0, st, (size_t)a.B * a.H, a.dy, p.dymax);
  if (LL) {   // every 2-byte element of the gate-gradient stream / every word of the maxima is its own ready flag
    DS2_CHECK_CUDA(cudaMemsetAsync(p.gmeta, 0xFF, gmeta_bytes, st));
    DS2_CHECK_CUDA(cudaMemsetAsync(p.dg16, 0xFF, (size_t)a.T * a.B * a.D * GH * sizeof(__half), st));
  }
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) {
      (void)cudaGetLastError();
      if (li == 0) {
        if (CL != 8)
          fprintf(stderr, "ds2_b200: WARNING resident split-K backward sweep launch failed (%s); using a slower variant\n",
                  cudaGetErrorString(le));
        return 1;
      }
      set_error("resident split-K backward sweep: second launch failed: %s", cudaGetErrorString(le));
      return DS2_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  if (a.dbias_done && a.dbias[0]) *a.dbias_done = 1;
  if (a.f16_done && p.dgn16) *a.f16_done = 1;
  return DS2_OK;
}

template <int RNN, int CL>
static int launch_bwd_splitk(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  const int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  const int GH = G * a.H;
  if (!getenv("DS2_NO_RESIDENT")) {
    // DS2_BWD_LL=1: flag-in-data exchange instead of grid barrier + TMA of dGh[t_next] (measured slower, see the
    // forward launcher)
    int rc = env_flag("DS2_BWD_LL", 0) ? launch_bwd_splitk_resident<RNN, CL, true>(a, ws, ws_bytes, st)
                                       : launch_bwd_splitk_resident<RNN, CL, false>(a, ws, ws_bytes, st);
    if (rc != 1) return rc;
  }
  constexpr int UM = UT * CL;
  if (a.H % UM != 0 || (GH / CL) % BK != 0 || GH % CL != 0) return 1;
  { int mrc = materialize_w_hh(a, st); if (mrc) return mrc; }   // this kernel streams the fp32 W_hh^T
  if (CL == 8 && ((a.B + 7) / 8 * 8) % 16 != 0) return 1;
  if (!vec_ok(a.gates, a.hseq, a.aux, a.dy)) return 1;
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 7) / 8 * 8; p.H = a.H; p.D = a.D; p.NT = a.H / UM; p.G = G;
  p.training = 1;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux; p.dy = a.dy;
  p.trace = trace_ptr_from_env("DS2_TRACE_BWD");
  if (ws_bytes < 4096) return 1;
  set_acc_layout(p);
  p.err = static_cast<int*>(ws);
  p.bar = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 128);
  const size_t smem = one_cta_per_sm(splitk_smem_bytes(p.NB, CL));
  if (smem > 227 * 1024) return 1;
  auto kern = rnn_bwd_splitk_kernel<RNN, false, CL>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    DS2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  int grid = a.D * p.NT * CL, launches = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CL; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  // Nsight Compute cannot replay a cooperative cluster launch: DS2_SPLITK_NONCOOP=1 (profiling only) drops the
  // cooperative attribute; co-residency is still checked with cudaOccupancyMaxActiveClusters below.
  cfg.numAttrs = noncoop_cluster_launch() ? 1 : 2;
  int max_clusters = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (oe != cudaSuccess) { (void)cudaGetLastError(); return 1; }
  if (max_clusters * CL < grid) {                        // all clusters must be co-resident (grid barrier)
    if (CL == 8 || max_clusters * CL < p.NT * CL) return 1;
    grid = p.NT * CL;                                    // one launch per direction
    launches = a.D;
    cfg.gridDim = dim3(grid);
  }
  for (int d = 0; d < a.D; ++d) {
    int rc = make_tmap_2d(&p.tmW[d], a.w_hh[d], a.H, GH, GH, UM, BK);   // W_hh^T (H, G*H): UM unit rows per box
    if (rc) return rc;
    rc = make_tmap_2d(&p.tmV[d], a.gates, a.T * a.B, a.D * GH, a.D * GH, a.B, BK);
    if (rc) return rc;
    if (RNN == DS2_RNN_GRU) {
      rc = make_tmap_2d(&p.tmV2[d], a.aux + (size_t)d * a.T * a.B * a.H, a.T * a.B, a.H, a.H, a.B, BK);
      if (rc) return rc;
    }
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096, st));
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) {
      (void)cudaGetLastError();
      if (li == 0) return 1;                             // e.g. cooperative+cluster launch refused: 16-unit kernel
      set_error("split-K backward sweep: second launch failed: %s", cudaGetErrorString(le));
      return DS2_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  return DS2_OK;
}

template <int RNN>
static int launch_bwd(const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace rp;
  const int G = RNN == DS2_RNN_LSTM ? 4 : (RNN == DS2_RNN_GRU ? 3 : 1);
  const int GH = G * a.H;
  { int mrc = materialize_w_hh(a, st); if (mrc) return mrc; }   // fp32 W_hh^T through TMA
  PersistParams p{};
  p.T = a.T; p.B = a.B; p.NB = (a.B + 7) / 8 * 8; p.H = a.H; p.D = a.D; p.NT = a.H / UT; p.G = G;
  p.training = 1;
  p.len = a.len; p.gates = a.gates; p.hseq = a.hseq; p.aux = a.aux; p.dy = a.dy;
  p.trace = trace_ptr_from_env("DS2_TRACE_BWD");
  if (ws_bytes < 4096) return 1;
  set_acc_layout(p);
  p.err = static_cast<int*>(ws);
  p.bar = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 128);
  const size_t smem = one_cta_per_sm(fwd_smem_bytes(p.NB));
  auto kern = rnn_bwd_persist_kernel<RNN>;
  static DeviceOnce attr_once;
  const int num_sms = device_sm_count();
  if (attr_once.first()) {
    DS2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.done();
  }
  if (smem > 227 * 1024) return 1;
  int max_blocks_per_sm = 0;
  DS2_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, kern, THREADS, smem));
  int grid = a.D * p.NT, launches = 1;
  if (max_blocks_per_sm < 1) return 1;
  if (grid > max_blocks_per_sm * num_sms) {           // both directions do not fit: one launch per direction
    if (p.NT > max_blocks_per_sm * num_sms) return 1;
    grid = p.NT;
    launches = a.D;
  }
  for (int d = 0; d < a.D; ++d) {
    // a.w_hh[d] is the transposed recurrent matrix (H, G*H) here
    int rc = make_tmap_2d(&p.tmW[d], a.w_hh[d], a.H, GH, GH, UT, BK);
    if (rc) return rc;
    // gate gradients: rows (t,b), full row width D*GH (the direction offset is a coordinate)
    rc = make_tmap_2d(&p.tmV[d], a.gates, a.T * a.B, a.D * GH, a.D * GH, a.B, BK);
    if (rc) return rc;
    if (RNN == DS2_RNN_GRU) {
      rc = make_tmap_2d(&p.tmV2[d], a.aux + (size_t)d * a.T * a.B * a.H, a.T * a.B, a.H, a.H, a.B, BK);
      if (rc) return rc;
    }
  }
  DS2_CHECK_CUDA(cudaMemsetAsync(ws, 0, 4096, st));
  for (int li = 0; li < launches; ++li) {
    p.d0 = li;
    void* args[] = {&p};
    DS2_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(THREADS), args, smem, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  DS2_LAUNCH(sweep_check_kernel, 1, 1, 0, st, p.err);
  return DS2_OK;
}

int rnn_sweep_bwd_tc(int rnn, const SeqArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (a.H % 32 != 0 || a.B > 256 || a.T < 2) return 1;
  {
    int rc = 1;
    if (!getenv("DS2_NO_SPLITK")) {
      // 8-CTA clusters (128 units, K/8 per CTA) halve the MMA chain of a step; 4-CTA clusters take the shapes
      // they do not (H % 128, K/8 not a multiple of 64, or 8-CTA clusters that do not fit the GPCs)
      if (env_flag("DS2_SPLITK_CL", 8) == 8) {
        if (rnn == DS2_RNN_LSTM) rc = launch_bwd_splitk<DS2_RNN_LSTM, 8>(a, ws, ws_bytes, st);
        else if (rnn == DS2_RNN_GRU) rc = launch_bwd_splitk<DS2_RNN_GRU, 8>(a, ws, ws_bytes, st);
        else rc = launch_bwd_splitk<DS2_RNN_TANH, 8>(a, ws, ws_bytes, st);
      }
      if (rc == 1) {
        if (rnn == DS2_RNN_LSTM) rc = launch_bwd_splitk<DS2_RNN_LSTM, 4>(a, ws, ws_bytes, st);
        else if (rnn == DS2_RNN_GRU) rc = launch_bwd_splitk<DS2_RNN_GRU, 4>(a, ws, ws_bytes, st);
        else rc = launch_bwd_splitk<DS2_RNN_TANH, 4>(a, ws, ws_bytes, st);
      }
    }
    if (rc != 1) return rc;   // done or a hard error; 1 = not eligible -> 16-unit kernel below
  }
  if (rnn == DS2_RNN_LSTM) return launch_bwd<DS2_RNN_LSTM>(a, ws, ws_bytes, st);
  if (rnn == DS2_RNN_GRU) return launch_bwd<DS2_RNN_GRU>(a, ws, ws_bytes, st);
  return launch_bwd<DS2_RNN_TANH>(a, ws, ws_bytes, st);
}


}  // namespace ds2
