// The GEMM / LayerNorm chain of one decode step as PERSISTENT multi-phase kernels (bf16 tensor-core mode).
//
// A decode step of the 6-layer post-LN decoder is 38 skinny GEMMs (M = 2B <= 64 rollout rows) and 19 reduce +
// LayerNorm passes between 12 attention launches.  As separate dependent launches each of them cost 3-6 us of
// launch boundary + cold ramp for 0.2-0.5 us of work (round 1: 0.24 ms of a 0.56 ms step for 10 % of its bytes).
// Here every run of GEMM / LayerNorm work between two attention kernels is ONE launch of decode_chain_kernel:
//
//   * 144 CTAs (one per SM, all co-resident), 192 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
//     warps 2-5 = epilogue (M = 64 MMA: 16 rows per TMEM sub-partition), all six warps in the LayerNorm / embedding phases;
//   * a launch interprets a list of PHASES (ChainPhase, built once per rollout shape by the engine); consecutive
//     phases are separated by a grid barrier (one atomic + an acquire spin per CTA) instead of a kernel boundary;
//   * a GEMM phase gives each CTA one [64 x BN] output tile over a 768-deep K slice: the CTA's whole weight slice
//     (BN x 768 bf16 = 24 / 48 KiB) is requested by TMA TWO PHASES AHEAD into one of two shared-memory buffers -
//     weights are constants of the rollout, so they stream from HBM while the preceding phases (or the preceding
//     attention kernel, ahead of the programmatic-dependency wait) are still running; after the barrier only the
//     [64 x 768] activation tile (L2-resident, 96 KiB) has to arrive before the 48 MMAs run;
//   * K = 3072 (FFN down) is four 768-deep splits into fp32 partials; every 768-wide projection that is followed by a
//     LayerNorm leaves its accumulator as an fp32 partial and the next phase reduces partials + bias (+GELU) +
//     residual and normalises, one row per CTA.  The residual stream can be carried in fp32 beside the bf16 copy that
//     feeds the next GEMM (ChainPhase::out_f32 / residual_f32).
//
// Reference semantics: BertLayer / BertSelfOutput / BertOutput / BertLMPredictionHead.transform
// (SP/models/bert/modeling_bert.py:287-298,330-356,379-421,471-501).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int CT = 192;                          // threads
constexpr int C_ROWS = 64, C_BK = 64, C_K = 768, C_KB = C_K / C_BK;
constexpr int C_AKB = C_ROWS * C_BK * 2;         // 8 KiB: one k-block of the activation tile
constexpr int C_ABYTES = C_KB * C_AKB;           // 96 KiB
constexpr int C_WBYTES = 32 * C_K * 2;           // 48 KiB per weight buffer (BN <= 32)
constexpr int C_SMEM = C_ABYTES + 2 * C_WBYTES + 1024 /*barriers*/ + 1024 /*alignment*/;
constexpr int C_H = 768;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major operand tile, 128-byte swizzle, 8-row groups 1024 B apart (same encoding as gemm_tcgen05.cu make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16: D = f32, A = B = bf16, K-major, M = 64, N = bn.  M = 64 (not 128 with an unused upper half, as the
// skinny GEMM does): the MMA phase of a tile is bound by reading the A operand from shared memory, and the 64-row
// shape reads half as much.  Accumulator row m then lives in TMEM lane (m % 16) + 32 * (m / 16).
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (static_cast<uint32_t>(64 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// All CTAs of the launch are co-resident (one per SM): arrive with a release, spin relaxed, then one acquire fence.
// bar.sync orders the CTA's stores before thread 0's release (cumulative at gpu scope) and thread 0's acquire before
// the CTA's loads; the proxy fences order the generic-proxy stores of the producers before the TMA (async-proxy) reads.
// (A two-level version - per-group counters forwarding to a root - measured SLOWER, 2.3 vs 1.5 us from the last
// arrival to the release: the barrier is bound by L2 round trips and fences, not by atomics serialising on one
// address.)  Counters only grow within a launch (generation g completes at g * gridDim.x); the last CTA to leave the
// kernel zeroes them.  bar[0] arrivals, bar[32] exits.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned gen) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned target = gen * gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  tc_fence_after();
}

constexpr int kMaxPhases = 7;
// The phase list travels in the kernel's parameter block: TMA fetches a tensor map from param/constant space at full
// speed, while the first version (descriptors in a global-memory array) paid a ~0.35 us descriptor fetch per TMA
// instruction: the 12 loads of an activation tile took 4.4 us (CXRM_CHAIN_TRACE).
struct ChainArgs {
  ChainPhase phases[kMaxPhases];
  int count;
  int R;
  const int* done;          // rollout finished: the whole launch is a no-op
  unsigned* bar;            // grid-barrier counters (see grid_sync), all 0 between launches
  const int* cur_token;
  const int* cur_type;
  const int* cur_pos;
  unsigned long long* trace;   // debug (CXRM_CHAIN_TRACE): kChainTraceSlots %globaltimer stamps per CTA of this launch
};
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// sub-stamps of GEMM phase i (any one thread): 1 weights in, 2 first / 6 last activation k-block in, 3 MMAs issued,
// 4 accumulator complete, 5 activation loads issued
#define GTRACE(sub)                                                                                              \
  do {                                                                                                           \
    if (a.trace) a.trace[static_cast<long long>(blockIdx.x) * kChainTraceSlots + 2 + 8 * i + (sub)] = gtimer(); \
  } while (0)
#define CTRACE(slot)                                                                                        \
  do {                                                                                                      \
    if (a.trace && threadIdx.x == 0) a.trace[static_cast<long long>(blockIdx.x) * kChainTraceSlots + (slot)] = gtimer(); \
  } while (0)

__device__ __forceinline__ float block_sum_192(float x, float* red) {
  x = warp_sum(x);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  return ((red[0] + red[1]) + (red[2] + red[3])) + (red[4] + red[5]);
}

// row statistics + normalisation of v[4] (columns 4*tid .. +3 of a 768-wide row), two-pass in fp32
__device__ __forceinline__ void ln_store(const ChainPhase& p, int m, float (&v)[4], float* red) {
  const int c = threadIdx.x * 4;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
  const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + c));
  const float mean = block_sum_192((v[0] + v[1]) + (v[2] + v[3]), red) * (1.0f / C_H);
  float d2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) d2 += (v[j] - mean) * (v[j] - mean);
  const float rstd = rsqrtf(block_sum_192(d2, red) * (1.0f / C_H) + p.eps);
  const float o0 = (v[0] - mean) * rstd * gm.x + bt.x, o1 = (v[1] - mean) * rstd * gm.y + bt.y;
  const float o2 = (v[2] - mean) * rstd * gm.z + bt.z, o3 = (v[3] - mean) * rstd * gm.w + bt.w;
  __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), b = __floats2bfloat162_rn(o2, o3);
  uint2 st;
  st.x = *reinterpret_cast<uint32_t*>(&a);
  st.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(static_cast<bf16*>(p.out) + static_cast<long long>(m) * p.ldo + c) = st;
  if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + static_cast<long long>(m) * C_H + c) = make_float4(o0, o1, o2, o3);
}

__global__ void __launch_bounds__(CT, 1) decode_chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float red[8];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = base;
  uint8_t* sW = base + C_ABYTES;                    // two buffers of C_WBYTES
  uint64_t* afull = reinterpret_cast<uint64_t*>(base + C_ABYTES + 2 * C_WBYTES);   // [C_KB]
  uint64_t* wfull = afull + C_KB;                   // [2]
  uint64_t* tfull = wfull + 2;                      // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  CTRACE(0);
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int i = 0; i < C_KB + 3; ++i) mbar_init(&afull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(kFull, *tmem_slot, 0);

  // ---- weight prefetch (producer thread only): this CTA's slice of GEMM phase `pi` -> buffer (ordinal & 1) ----------
  int pf_phase = 0;        // next phase index to look at for a GEMM this CTA takes part in
  int pf_ord = 0;          // ordinal (among this CTA's GEMMs) of the next prefetch
  auto prefetch_next_w = [&]() {
    for (; pf_phase < a.count; ++pf_phase) {
      const ChainPhase& q = a.phases[pf_phase];
      if (q.type != CH_GEMM || cta >= q.n_tiles * q.nsplit) continue;
      const int n0 = (cta % q.n_tiles) * q.bn, k0 = (cta / q.n_tiles) * q.kslice, nkb = q.kslice / C_BK;
      uint8_t* dst = sW + (pf_ord & 1) * C_WBYTES;
      uint64_t* bar = &wfull[pf_ord & 1];
      mbar_expect_tx(bar, static_cast<uint32_t>(q.bn) * q.kslice * 2);
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(dst + kb * q.bn * 128, &q.tmB, k0 + kb * C_BK, n0, bar);
      ++pf_ord;
      ++pf_phase;
      return;
    }
  };
  const bool producer = (warp == 0 && lane == 0);
  if (producer) {          // weights are rollout constants: the first two slices stream ahead of the dependency wait
    prefetch_next_w();
    prefetch_next_w();
  }
  pdl_wait();
  CTRACE(1);
  const bool skip = a.done != nullptr && *a.done != 0;
  if (skip) {
    if (producer)          // shared memory must not be released with bulk copies in flight
      for (int o = 0; o < pf_ord; ++o) mbar_wait(&wfull[o & 1], 0);
  } else {
    int ord = 0;           // ordinal of this CTA's next GEMM (every thread keeps the same count)
    unsigned gen = 0;      // grid barriers passed
#pragma unroll 1
    for (int i = 0; i < a.count; ++i) {
      const ChainPhase& p = a.phases[i];
      if (i > 0) {
        CTRACE(2 + 8 * (i - 1) + 7);   // end of phase i - 1 (this CTA)
        ++gen;
        grid_sync(a.bar, gen);
      }
      CTRACE(2 + 8 * i);               // start of phase i
      if (p.type == CH_GEMM) {
        if (cta >= p.n_tiles * p.nsplit) continue;
        const int n0 = (cta % p.n_tiles) * p.bn, split = cta / p.n_tiles, k0 = split * p.kslice, nkb = p.kslice / C_BK;
        const int buf = ord & 1;
        const uint32_t par = ord & 1, wpar = (ord >> 1) & 1;
        if (warp == 0) {
          if (lane == 0) {
            // (the same tile through the LSU - 128 threads x 48 16-byte loads into the swizzled layout - measured
            // SLOWER than these 12 TMA boxes: 5.7 vs 3.5 us until the operands are in)
            // ONE 3-D box: 12 k-block tiles of 64 x 64 (twelve 2-D boxes arrived one L2 round trip after the other:
            // 3.4 us for the 96 KiB; CXRM_CHAIN_TRACE)
            mbar_expect_tx(&afull[0], static_cast<uint32_t>(nkb) * C_AKB);
            tma_load_3d(sA, &p.tmA, 0, 0, k0 / C_BK, &afull[0]);
            GTRACE(5);
            mbar_wait(tfull, par);       // the MMAs have read this phase's weight buffer: refill it two GEMMs ahead
            prefetch_next_w();
          }
        } else if (warp == 1) {
          // The whole warp walks the loop (warp-uniform control flow, operands in uniform registers) and one elected
          // lane issues: with the loop inside `if (lane == 0)` the compiler could not prove the descriptors uniform
          // and wrapped EVERY tcgen05.mma in an ELECT / 6 x R2UR.BROADCAST waterfall - ~110 cycles per MMA, 2 us of
          // the 3.5 us a tile took (CXRM_CHAIN_TRACE, SASS).
          const uint32_t idesc = make_idesc(p.bn);
          const bool leader = elect_one();
          mbar_wait(&wfull[buf], wpar);
          mbar_wait(&afull[0], par);
          if (leader) GTRACE(2);
          tc_fence_after();
          const uint32_t wb = smem_u32(sW) + static_cast<uint32_t>(buf * C_WBYTES);
          const uint32_t ab = smem_u32(sA);
          const uint32_t wstep = static_cast<uint32_t>(p.bn) * 128u;
#pragma unroll 1
          for (int kb = 0; kb < nkb; ++kb) {
            const uint64_t da = make_desc(ab + static_cast<uint32_t>(kb * C_AKB)), db = make_desc(wb + static_cast<uint32_t>(kb) * wstep);
            // FOUR accumulators (TMEM columns 32j ..): the k-steps of a k-block go to different ones, so consecutive
            // MMAs are independent; the epilogue adds the four
#pragma unroll
            for (int k = 0; k < C_BK / 16; ++k)
              if (leader)
                umma(tmem + static_cast<uint32_t>(64 * k), da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                     kb != 0 ? 1u : 0u);
          }
          if (leader) {
            umma_commit(tfull);
            GTRACE(3);
          }
          __syncwarp();
        } else {
          // warps 2..5: sub-partition q = warp % 4 holds rows 16q .. 16q + 15 in its lanes 0..15
          const int q = warp & 3;
          const int m = (lane < 16) ? q * 16 + lane : C_ROWS;      // lanes 16..31 hold nothing
          const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16);
          const int nchunk = p.bn / 16;
          float bs[32];
          if (p.epi != CE_PARTIAL) {
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (q4 * 4 < p.bn) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * q4));
              bs[4 * q4] = b4.x; bs[4 * q4 + 1] = b4.y; bs[4 * q4 + 2] = b4.z; bs[4 * q4 + 3] = b4.w;
            }
          }
          mbar_wait(tfull, par);
          tc_fence_after();
          if (threadIdx.x == 64) GTRACE(4);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            if (ch >= nchunk) break;     // warp-uniform
            uint32_t r[16];
            tmem_ld16(taddr + static_cast<uint32_t>(ch * 16), r);
#pragma unroll
            for (int acc = 1; acc < 4; ++acc) {
              uint32_t r2[16];
              tmem_ld16(taddr + static_cast<uint32_t>(64 * acc + ch * 16), r2);
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
            if (m >= a.R) continue;
            const int nb = n0 + ch * 16;
            if (p.epi == CE_PARTIAL) {
              float4* dst = reinterpret_cast<float4*>(p.partial + (static_cast<long long>(split) * C_ROWS + m) * p.N + nb);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                dst[q4] = make_float4(__uint_as_float(r[4 * q4]), __uint_as_float(r[4 * q4 + 1]), __uint_as_float(r[4 * q4 + 2]),
                                      __uint_as_float(r[4 * q4 + 3]));
            } else {
              bf16* cp = static_cast<bf16*>(p.out) + static_cast<long long>(m) * p.ldo + nb;
#pragma unroll
              for (int q8 = 0; q8 < 2; ++q8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[j] = __uint_as_float(r[8 * q8 + j]) + bs[(ch & 1) * 16 + 8 * q8 + j];   // (bias epilogues: bn <= 32)
                  if (p.epi == CE_BF16_GELU) v[j] = gelu_fast(v[j]);
                }
                Vec16<bf16> ov;
                ov.pack(v);
                ov.store(cp + 8 * q8);
              }
            }
          }
        }
        ++ord;
      } else if (cta < a.R) {
        // ---- row phases: one row per CTA, four columns per thread -------------------------------------------------
        const int m = cta, c = threadIdx.x * 4;
        float v[4];
        if (p.type == CH_LN) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          // split-K partials, four loads in flight, summed in split order (reproducible)
#pragma unroll 1
          for (int s0 = 0; s0 < p.nsplit; s0 += 4) {
            float4 part[4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if (s0 + s < p.nsplit)
                part[s] = __ldcg(reinterpret_cast<const float4*>(p.partial + (static_cast<long long>(s0 + s) * C_ROWS + m) * C_H + c));
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if (s0 + s < p.nsplit) {
                acc.x += part[s].x; acc.y += part[s].y; acc.z += part[s].z; acc.w += part[s].w;
              }
          }
          float rs[4] = {0.f, 0.f, 0.f, 0.f};
          if (p.residual_f32) {
            const float4 r4 = __ldcg(reinterpret_cast<const float4*>(p.residual_f32 + static_cast<long long>(m) * C_H + c));
            rs[0] = r4.x; rs[1] = r4.y; rs[2] = r4.z; rs[3] = r4.w;
          } else if (p.residual) {
            const uint2 rr = __ldcg(reinterpret_cast<const uint2*>(static_cast<const bf16*>(p.residual) + static_cast<long long>(m) * C_H + c));
            rs[0] = __uint_as_float(rr.x << 16); rs[1] = __uint_as_float(rr.x & 0xffff0000u);
            rs[2] = __uint_as_float(rr.y << 16); rs[3] = __uint_as_float(rr.y & 0xffff0000u);
          }
          const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[0] = acc.x + b4.x; v[1] = acc.y + b4.y; v[2] = acc.z + b4.z; v[3] = acc.w + b4.w;
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] += rs[j];
          if (p.round_pre) {     // the reference under autocast rounds the pre-LayerNorm sum to bf16 (a GEMM output)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
          }
        } else {                 // CH_EMBED: (word + type) + position, as BertEmbeddings.forward associates it
          const bf16* w = p.word + static_cast<long long>(a.cur_token[m]) * C_H + c;
          const bf16* t = p.type_emb + static_cast<long long>(a.cur_type[m]) * C_H + c;
          const bf16* ps = p.pos_emb + static_cast<long long>(a.cur_pos[m]) * C_H + c;
          const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w)), tv = __ldg(reinterpret_cast<const uint2*>(t)),
                      pv = __ldg(reinterpret_cast<const uint2*>(ps));
          auto lo = [](uint32_t u) { return __uint_as_float(u << 16); };
          auto hi = [](uint32_t u) { return __uint_as_float(u & 0xffff0000u); };
          v[0] = (lo(wv.x) + lo(tv.x)) + lo(pv.x);
          v[1] = (hi(wv.x) + hi(tv.x)) + hi(pv.x);
          v[2] = (lo(wv.y) + lo(tv.y)) + lo(pv.y);
          v[3] = (hi(wv.y) + hi(tv.y)) + hi(pv.y);
        }
        ln_store(p, m, v, red);
      }
    }
    __syncthreads();
    CTRACE(2 + 8 * (a.count - 1) + 7);
    // leave the barrier counters at zero for the next launch: the last CTA to get here knows everyone passed every barrier
    if (threadIdx.x == 0 && a.count > 1) {
      __threadfence();
      const unsigned prev = atomicAdd(a.bar + 32, 1u);
      if (prev == gridDim.x - 1) {
        a.bar[0] = 0;
        a.bar[32] = 0;
        __threadfence();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

}  // namespace

int decode_chain_ctas() { return 144; }

bool decode_chain_available() {
  static int ok = -1;
  if (ok < 0) {
    ok = 0;
    // OPT-IN (CXRM_CHAIN=1).  Measured on B200 at the benchmark shape: both the full chain (all
    // GEMM / LayerNorm work between two attention kernels in one launch: rollout 152.8 ms) and the pair form used now
    // (one launch per split-K GEMM + LayerNorm pair: 160.0 ms) are SLOWER than the PDL-chained stand-alone kernels
    // (147.4 ms): a resident 198 KB CTA per SM keeps the NEXT kernel from becoming resident early, so its weight / K/V
    // prefetch ahead of the dependency wait - what the PDL chain lives on - is lost.
    if (std::getenv("CXRM_CHAIN") != nullptr) {
      int dev = 0, sms = 0, per_sm = 0;
      if (cudaGetDevice(&dev) == cudaSuccess &&
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
          cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_chain_kernel, CT, C_SMEM) == cudaSuccess)
        ok = (per_sm >= 1 && sms >= decode_chain_ctas()) ? 1 : 0;   // the grid barrier needs every CTA resident
      cudaGetLastError();
    }
  }
  return ok == 1;
}

void decode_chain(const ChainPhase* phases, int count, int R, const RolloutState& st, unsigned* bar, int n_ctas,
                  cudaStream_t stream, unsigned long long* trace) {
  CXRM_CHECK(n_ctas >= 1 && n_ctas <= decode_chain_ctas(), "decode_chain grid");
  CXRM_CHECK(decode_chain_available(), "decode chain kernel unavailable on this device");
  CXRM_CHECK(count >= 1 && count <= kMaxPhases && R >= 1 && R <= C_ROWS && 2 + 8 * count <= kChainTraceSlots, "decode_chain shape");
  static_assert(sizeof(ChainArgs) <= 4000, "kernel parameter block");
  ChainArgs a;
  std::memset(&a, 0, sizeof(a));
  a.trace = trace;
  for (int i = 0; i < count; ++i) a.phases[i] = phases[i];
  a.count = count; a.R = R; a.done = st.done; a.bar = bar;
  a.cur_token = st.cur_token; a.cur_type = st.cur_type; a.cur_pos = st.cur_pos;
  launch_chain(decode_chain_kernel, dim3(n_ctas), dim3(CT), C_SMEM, stream, a);
  check_launch("decode_chain");
}

}  // namespace cxrm
