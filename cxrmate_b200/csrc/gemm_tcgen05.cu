// bf16 GEMM on the 5th-generation tensor cores of sm_100a:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) -> shared memory ring
//   -> tcgen05.mma (one elected thread, operands from shared-memory descriptors)
//   -> fp32 accumulator in TMEM -> tcgen05.ld -> fused epilogue
//   (bias, exact GELU, residual add, bf16 or fp32 store).
//
// C[M,N] = epi(A[M,K] . W[N,K]^T).  Both operands are K-major, which is the
// natural layout of nn.Linear (x[M,K], weight[N,K]); out-of-range rows/columns
// and the K tail are zero-filled by TMA.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer, warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int BM = 128, BK = 64, MAX_STAGES = 4, NTHREADS = 192;
constexpr int A_BYTES = BM * BK * 2;   // 16 KiB

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
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(i)                                                                                         \
  do {                                                                                                   \
    if (g.trace) g.trace[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = gtimer(); \
  } while (0)
// one lane of a converged warp (the loops around it stay warp-uniform, so descriptors / TMEM addresses live in uniform
// registers: a loop run by `if (lane == 0)` makes every operand of UTCHMMA / UTMALDG divergent and ptxas wraps each
// instruction in an ELECT + R2UR.BROADCAST waterfall loop)
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

// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64) = 2)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) |
         (static_cast<uint32_t>(BM >> 4) << 24);
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
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Fused epilogue of one accumulator row segment: v[0..31] = D[m, nb .. nb+31] (fp32) ->
// + bias -> act -> (+ residual) -> store as bf16 / fp32 / head-major bf16.
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, float (&v)[32], long long m, int nb, int vec_ok) {
  const bf16* __restrict__ R = static_cast<const bf16*>(g.residual);
  const bool full_chunk = (nb + 32 <= g.N);
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += (full_chunk || nb + j < g.N) ? g.bias[nb + j] : 0.f;
  }
  if (g.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
  }
  if (g.c_head_stride > 0) {
    // head-major K/V cache store: 32 consecutive columns never straddle a 64-wide head
    bf16* cp = static_cast<bf16*>(g.C) + static_cast<long long>(nb / 64) * g.c_head_stride + m * 64 + (nb % 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Vec16<bf16> ov;
      ov.pack(v + 8 * q);
      ov.store(cp + 8 * q);
    }
  } else if (full_chunk && vec_ok) {
    if (R) {
      const bf16* rp = R + m * g.ldr + nb;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Vec16<bf16> rv;
        rv.load(rp + 8 * q);
        float rf[8];
        rv.unpack(rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
      }
    }
    if (g.out_f32) {
      float* cp = static_cast<float*>(g.C) + m * g.ldc + nb;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
      bf16* cp = static_cast<bf16*>(g.C) + m * g.ldc + nb;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Vec16<bf16> ov;
        ov.pack(v + 8 * q);
        ov.store(cp + 8 * q);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = nb + j;
      if (n >= g.N) continue;
      float o = v[j];
      if (R) o += to_f(R[m * g.ldr + n]);
      if (g.out_f32)
        static_cast<float*>(g.C)[m * g.ldc + n] = o;
      else
        static_cast<bf16*>(g.C)[m * g.ldc + n] = from_f<bf16>(o);
    }
  }
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

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;   // power of two >= BN
  static constexpr int smem(int stages) { return stages * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + BN * 4 /*bias*/; }
};

// EPI: 0 = general epilogue (every GemmArgs flag), 1 = TMA-store epilogue only (bf16 row-major output, bias, optional
// residual), 2 = the same with GELU.  The specialisations exist for code size: the general epilogue of the 256-wide
// tile is 210 KB of straight-line SASS (8 unrolled 32-column chunks x every store variant), far beyond the
// instruction cache, and its speed then depends on code placement (the same GEMM measured 145 us and 305 us in two
// builds that differed elsewhere).
template <int BN, int EPI = 0>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmC, GemmArgs g,
                                                           int vec_ok, int STAGES, int tma_store) {
  // STAGES (1..4) is a launch parameter: short-K GEMMs (CvT stage 1: K = 64) take one stage of shared memory, so
  // several CTAs share an SM and one CTA's epilogue overlaps another's loads and MMAs.
  using C_ = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): set-up below overlaps the previous kernel's tail
  if (g.skip_flag) {
    pdl_wait();
    if (*g.skip_flag) return;
  }

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * C_::STAGE_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);   // [BN]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = (g.K + BK - 1) / BK;
  if (threadIdx.x == 0) TRACE(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_slot, 0);   // provably warp-uniform: no R2UR waterfall around each tcgen05.mma
  if (threadIdx.x == 0) TRACE(1);
  pdl_wait();   // barriers, TMEM and descriptors are ready; operands / residual / output belong to the previous kernels

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], C_::STAGE_BYTES);
        uint8_t* a_dst = tiles + s * C_::STAGE_BYTES;
        tma_load_2d(a_dst, &tmA, kb * BK, m0, &full[s]);
        tma_load_2d(a_dst + A_BYTES, &tmB, kb * BK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (kb == 0) TRACE(2);
        const uint32_t a_addr = smem_u32(tiles + s * C_::STAGE_BYTES);
        const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // +32 bytes per 16-element K step inside the 128-byte swizzle atom: start-address field += 2
          umma(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
               (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);     // accumulator complete
      TRACE(3);
    }
  } else {
    // ---- epilogue: 4 warps x 32 TMEM lanes ---------------------------------------
    // Everything the epilogue needs from global memory is fetched WHILE the mainloop runs: the tile's bias slice into
    // shared memory, this thread's residual row into registers.  (Loading them per 32-column chunk after the
    // accumulator was ready serialised two L2 round trips per chunk: ncu showed the tensor pipe 1-4 % busy.)
    const int quarter = warp % 4;
    const int et = threadIdx.x - 64;   // 0..127
    const long long m = static_cast<long long>(m0) + quarter * 32 + lane;
    const bool row_ok = m < g.M;
    for (int j = et; j < BN; j += 128) sbias[j] = (g.bias && n0 + j < g.N) ? g.bias[n0 + j] : 0.f;
    constexpr int RPRE = BN < 128 ? BN : 128;   // residual columns held in registers (the rest are loaded in the loop)
    uint4 rres[RPRE / 8];
    const bf16* __restrict__ R = static_cast<const bf16*>(g.residual);
    // (EPI 2, the GELU epilogue, keeps its chunk loop rolled for code size: no register-resident residual there)
    const bool res_pre = EPI != 2 && R != nullptr && vec_ok && row_ok && g.c_head_stride == 0 && (n0 + BN <= g.N);
    if (res_pre) {
#pragma unroll
      for (int q = 0; q < RPRE / 8; ++q) rres[q] = __ldg(reinterpret_cast<const uint4*>(R + m * g.ldr + n0 + 8 * q));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // bias slice visible to the four epilogue warps
    if (et == 0) TRACE(4);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (et == 0) TRACE(5);
#pragma unroll(EPI == 2 ? 1 : BN / 32)
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= g.N) break;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
      if (EPI == 0 && !row_ok && !tma_store) continue;   // (TMA store: every thread fills its panel row; out-of-range rows are clipped)
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sbias[c0 + j];
      const void* res_left = g.residual;     // residual still to be added by the store path (nullptr: done / none)
      if (EPI == 2 || (EPI == 0 && g.act == ACT_GELU)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
      }
      if (res_pre && c0 < RPRE) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          Vec16<bf16> rv;
          rv.raw = rres[c0 / 8 + q];
          float rf[8];
          rv.unpack(rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
        }
        res_left = nullptr;
      }
      if (EPI != 0 || tma_store) {
        // Row-per-thread 16-byte global stores hit 32 different rows per warp instruction (the globaltimer trace showed
        // the epilogue at 3-11 us per tile, several times the mainloop).  Instead the thread's 64 bytes go into a
        // 64B-swizzled [128 rows x 32 cols] panel in shared memory (the pipeline stages are idle by now) ...
        if (res_left && row_ok) {
          const bf16* rp = static_cast<const bf16*>(res_left) + m * g.ldr + n0 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec16<bf16> rv;
            rv.load(rp + 8 * q);
            float rf[8];
            rv.unpack(rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
          }
        }
        const int row = quarter * 32 + lane;
        uint8_t* panel = tiles + (c0 / 32) * (BM * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          Vec16<bf16> ov;
          ov.pack(v + 8 * q);
          *reinterpret_cast<uint4*>(panel + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = ov.raw;
        }
      } else if constexpr (EPI == 0) {
        GemmArgs gl = g;
        gl.bias = nullptr;
        gl.act = ACT_NONE;
        gl.residual = res_left;
        epilogue_chunk(gl, v, m, n0 + c0, vec_ok);
      }
      if (EPI != 0 || tma_store) {
        // ... and one thread hands the panel to the TMA unit: a coalesced, asynchronous, bounds-clipped store.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmC)),
                       "r"(smem_u32(tiles + (c0 / 32) * (BM * 64))), "r"(n0 + c0), "r"(m0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if ((EPI != 0 || tma_store) && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem read out before exit
    if (et == 0) TRACE(6);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(7);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// Persistent variant for the big-M GEMMs of the encoder / prefill / reward model (TMA-store epilogues only).
// One CTA per SM walks the output tiles (M fastest, so the CTAs running together share a weight slice in L2):
//   warp 0      TMA producer: runs AHEAD through the shared-memory ring across tile boundaries
//   warp 1      MMA issuer: accumulates tile i into TMEM buffer i % 2
//   warps 2..5  epilogue of tile i (TMEM -> registers -> bias / GELU / residual -> swizzled panels -> TMA store)
//               while the MMAs of tile i + 1 fill the other TMEM buffer.
// The per-tile kernel above gives every CTA one tile: its loads start cold and its epilogue has nothing to hide
// behind, which is why the K = 192 / 384 contractions of CvT ran at 180-330 TF/s.
// =====================================================================================================
constexpr int P_MAX_STAGES = 8;
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN>
struct PCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING = (BN / 32) * (BM * 64);                       // one 64B-swizzled panel per 32 columns
  static constexpr int TMEM_COLS = BN <= 16 ? 32 : BN <= 32 ? 64 : BN <= 64 ? 128 : BN <= 128 ? 256 : 512;   // 2 accumulators
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
  static constexpr int stages() {
    int s = (220 * 1024 - STAGING - 2048 - BN * 4) / STAGE_BYTES;
    return s > P_MAX_STAGES ? P_MAX_STAGES : s;
  }
  static constexpr int smem() { return stages() * STAGE_BYTES + STAGING + 1024 /*alignment*/ + 512 /*barriers*/ + BN * 4; }
};

template <int BN, int EPI>   // EPI 1: bias (+ residual) -> bf16;  2: bias -> GELU (+ residual) -> bf16
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      const __grid_constant__ CUtensorMap tmC, GemmArgs g,
                                                                      int vec_ok, int tiles_m, int n_tiles) {
  using C_ = PCfg<BN>;
  constexpr int STAGES = C_::stages();
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* staging = tiles + STAGES * C_::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + C_::STAGING);
  uint64_t* empty = full + P_MAX_STAGES;
  uint64_t* tfull = empty + P_MAX_STAGES;      // [2] accumulator complete
  uint64_t* tempty = tfull + 2;                // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);   // [BN]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_kb = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_slot, 0);   // provably warp-uniform: no R2UR waterfall around each tcgen05.mma
  pdl_wait();   // (encoder chain) set-up done under the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      int it = 0;   // k-blocks issued so far, across tiles
#pragma unroll 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], C_::STAGE_BYTES);
          uint8_t* a_dst = tiles + s * C_::STAGE_BYTES;
          tma_load_2d(a_dst, &tmA, kb * BK, m0, &full[s]);
          tma_load_2d(a_dst + A_BYTES, &tmB, kb * BK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int it = 0, t = 0;
#pragma unroll 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator (first use: passes)
        tc_fence_after();
        const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * C_::ACC_STRIDE);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + s * C_::STAGE_BYTES);
          const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    const int quarter = warp % 4;
    const int et = threadIdx.x - 64;   // 0..127
    const bf16* __restrict__ R = static_cast<const bf16*>(g.residual);
    int t = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
      const long long m = static_cast<long long>(m0) + quarter * 32 + lane;
      const bool row_ok = m < g.M;
      // the panels still belong to the TMA stores of the previous tile until they have been read out
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");   // (also: every thread is done with the previous tile's sbias)
      for (int j = et; j < BN; j += 128) sbias[j] = (g.bias && n0 + j < g.N) ? g.bias[n0 + j] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull[buf], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * C_::ACC_STRIDE);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= g.N) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld32(acc + static_cast<uint32_t>(c0), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sbias[c0 + j];
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        }
        if (R && row_ok) {
          const bf16* rp = R + m * g.ldr + n0 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec16<bf16> rv;
            rv.load(rp + 8 * q);
            float rf[8];
            rv.unpack(rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
          }
        }
        const int row = quarter * 32 + lane;
        uint8_t* panel = staging + (c0 / 32) * (BM * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          Vec16<bf16> ov;
          ov.pack(v + 8 * q);
          *reinterpret_cast<uint4*>(panel + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = ov.raw;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmC)),
                       "r"(smem_u32(panel)), "r"(n0 + c0), "r"(m0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      // every tcgen05.ld of this tile has completed (tmem_ld32 waits): hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty[buf]);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem read out before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// Weight-resident persistent GEMM for the short-K, huge-M contractions of CvT (K <= 384: q/k/v/o projections, MLP-up,
// conv embeddings of stage 1; M = 4.6 k .. 295 k rows).  With a ring that streams A AND B, a 128 x BN tile of such a GEMM
// loads as many bytes from L2 as it has time to compute on (K = 384, BN = 128: 196 KB per 12.6 MFLOP = 64 FLOP/B, an
// L2 ceiling of ~390 TF/s; measured 215-330 TF/s).  Here a CTA owns ONE column slice of the output for its whole life:
//   * the slice's weights [BN, K] are loaded ONCE into shared memory (<= 147 KB) and stay there;
//   * warp 0 streams only A tiles (128 x 64, 16 KB stages) through a 4-8 deep ring, running ahead across tiles;
//   * warp 1 issues the MMAs into two alternating TMEM accumulators;
//   * 8 epilogue warps in two groups (group g takes the 32-column chunks g, g + 2, ...): TMEM -> registers -> bias /
//     GELU / residual (the next chunk's residual row is requested before the current chunk is processed) -> one of two
//     64B-swizzled staging panels per group -> TMA store.
// CTA c works on slice c % n_slices and on the M tiles c / n_slices, + grid / n_slices, ...: the CTAs of one M tile run
// at the same time, so A comes from HBM once and from L2 n_slices times.
// =====================================================================================================
constexpr int W_GROUPS = 2, W_THREADS = 64 + 128 * W_GROUPS, W_MAX_STAGES = 8, W_PANEL = BM * 64;

template <int BN>
struct WCfg {
  static constexpr int TMEM_COLS = BN <= 16 ? 32 : BN <= 32 ? 64 : BN <= 64 ? 128 : BN <= 128 ? 256 : 512;   // 2 accumulators
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
};
inline size_t wres_smem(int bn, int nkb, int stages) {
  return static_cast<size_t>(nkb) * bn * 128 + static_cast<size_t>(stages) * A_BYTES + W_GROUPS * 2 * W_PANEL + 1024 + 512 + bn * 4;
}

template <int BN, int EPI>   // EPI 1: bias (+ residual) -> bf16;  2: bias -> GELU (+ residual) -> bf16
__global__ void __launch_bounds__(W_THREADS, 1) gemm_tc_wres_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmC, GemmArgs g,
                                                                    int tiles_m, int n_slices, int STAGES) {
  using C_ = WCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // debug stamps (cxrm_test_set_gemm_trace): 16 slots per CTA - 0 entry, 1 set-up done, 2 weights resident, 3 first A
  // tile arrived, 4 last MMA issued, 5 first accumulator seen by the epilogue, 6 first tile stored, 7 last tile
  // stored, 8 exit, 9 / 10 ns the MMA warp waited for A data / for a drained accumulator, 11 ns epilogue group 0
  // waited for accumulators, 12 tiles of this CTA
#define WTRACE(i) do { if (g.trace) g.trace[static_cast<long long>(blockIdx.x) * 16 + (i)] = gtimer(); } while (0)
#define WTRACE_ADD(i, v) do { if (g.trace) g.trace[static_cast<long long>(blockIdx.x) * 16 + (i)] += (v); } while (0)
  const int num_kb = (g.K + BK - 1) / BK;
  pdl_launch_dependents();   // (encoder chain) the next kernel may start its set-up
  if (threadIdx.x == 0) WTRACE(0);
  uint8_t* wres = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ring = wres + num_kb * (BN * 128);
  uint8_t* staging = ring + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + W_GROUPS * 2 * W_PANEL);
  uint64_t* empty = full + W_MAX_STAGES;
  uint64_t* tfull = empty + W_MAX_STAGES;      // [2] accumulator complete
  uint64_t* tempty = tfull + 2;                // [2] accumulator drained by the epilogue
  uint64_t* wfull = tempty + 2;                // weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);   // [BN]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int slice = blockIdx.x % n_slices, mt0 = blockIdx.x / n_slices, mstride = gridDim.x / n_slices;
  const int n0 = slice * BN;

  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4 * W_GROUPS);   // one arrival per epilogue warp
    }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < BN; j += W_THREADS) sbias[j] = g.bias ? g.bias[n0 + j] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_slot, 0);
  if (threadIdx.x == 0) WTRACE(1);

  if (warp == 0) {
    // TMA producer: the whole warp walks the loops (uniform control flow), one elected lane issues
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      mbar_expect_tx(wfull, static_cast<uint32_t>(num_kb * BN * 128));
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(wres + kb * (BN * 128), &tmB, kb * BK, n0, wfull);
    }
    __syncwarp();
    pdl_wait();   // the weights above are constants; the activations belong to the previous kernel
    int it = 0;
#pragma unroll 1
    for (int mt = mt0; mt < tiles_m; mt += mstride) {
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[s], A_BYTES);
          tma_load_2d(ring + s * A_BYTES, &tmA, kb * BK, mt * BM, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform loops, one elected lane issues the MMAs and their commits
    constexpr uint32_t idesc = make_idesc(BN);
    pdl_wait();
    mbar_wait(wfull, 0);
    tc_fence_after();
    if (lane == 0) WTRACE(2);
    const uint32_t w_addr = smem_u32(wres);
    const uint32_t ring_addr = smem_u32(ring);
    int it = 0, t = 0;
    unsigned long long wait_data = 0, wait_acc = 0, t_first = 0;   // trace only: kept in registers, stored once
#pragma unroll 1
    for (int mt = mt0; mt < tiles_m; mt += mstride, ++t) {
      const int buf = t & 1;
      unsigned long long w0 = g.trace ? gtimer() : 0;
      mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      if (g.trace) wait_acc += gtimer() - w0;
      const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * C_::ACC_STRIDE);
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        w0 = g.trace ? gtimer() : 0;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        if (g.trace) {
          const unsigned long long now = gtimer();
          wait_data += now - w0;
          if (it == 0) t_first = now;
        }
        const uint64_t da = make_desc(ring_addr + static_cast<uint32_t>(s * A_BYTES));
        const uint64_t db = make_desc(w_addr + static_cast<uint32_t>(kb * (BN * 128)));
#ifdef CXRM_WRES_KBTRACE
        const long long c0 = clock64();
#endif
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
#ifdef CXRM_WRES_KBTRACE
          const long long c1 = clock64();
#endif
          umma_commit(&empty[s]);
          if (kb == num_kb - 1) umma_commit(&tfull[buf]);
#ifdef CXRM_WRES_KBTRACE
          if (g.trace && t == 1 && kb < 6) {   // debug build only: clocks of tile 1 (k-block start, MMAs issued, commits issued)
            unsigned long long* q = g.trace + 148 * 16 + static_cast<long long>(blockIdx.x) * 32 + kb * 4;
            q[0] = static_cast<unsigned long long>(c0);
            q[1] = static_cast<unsigned long long>(c1);
            q[2] = static_cast<unsigned long long>(clock64());
          }
#endif
        }
        __syncwarp();
      }
    }
    if (lane == 0 && g.trace) {
      WTRACE(4);
      g.trace[static_cast<long long>(blockIdx.x) * 16 + 3] = t_first;
      WTRACE_ADD(9, wait_data);
      WTRACE_ADD(10, wait_acc);
      WTRACE_ADD(12, static_cast<unsigned long long>(t));
    }
  } else {
    const int quarter = warp % 4;                 // TMEM lane quarter this warp may read
    const int grp = (warp - 2) / 4;               // epilogue group
    const int et = threadIdx.x - 64 - grp * 128;  // 0..127 within the group
    const int bar_id = 1 + grp;
    uint8_t* panels = staging + grp * 2 * W_PANEL;
    const bf16* __restrict__ R = static_cast<const bf16*>(g.residual);
    constexpr int NCH = BN / 32;
    int t = 0, pi = 0;
    unsigned long long wait_epi = 0;
    pdl_wait();   // residual rows and the output buffer belong to the previous kernels
#pragma unroll 1
    for (int mt = mt0; mt < tiles_m; mt += mstride, ++t) {
      const int buf = t & 1;
      const int m0 = mt * BM;
      const long long m = static_cast<long long>(m0) + quarter * 32 + lane;
      const bool row_ok = m < g.M;
      const bf16* rrow = (R && row_ok) ? R + m * g.ldr + n0 : nullptr;
      uint4 rnext[4];
      if (rrow && grp < NCH) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rnext[q] = __ldg(reinterpret_cast<const uint4*>(rrow + grp * 32 + 8 * q));
      }
      const bool tr = g.trace && grp == 0 && et == 0;
      const unsigned long long w1 = tr ? gtimer() : 0;
      mbar_wait(&tfull[buf], (t >> 1) & 1);
      tc_fence_after();
      if (tr) {
        wait_epi += gtimer() - w1;
        if (t == 0) WTRACE(5);
      }
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * C_::ACC_STRIDE);
#pragma unroll 1
      for (int ci = grp; ci < NCH; ci += W_GROUPS) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(acc + static_cast<uint32_t>(c0), r);
        uint4 rcur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
        if (rrow && ci + W_GROUPS < NCH) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rnext[q] = __ldg(reinterpret_cast<const uint4*>(rrow + (ci + W_GROUPS) * 32 + 8 * q));
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sbias[c0 + j];
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        }
        if (rrow) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec16<bf16> rv;
            rv.raw = rcur[q];
            float rf[8];
            rv.unpack(rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
          }
        }
        // the panel about to be written was handed to the TMA unit two chunks ago: its read must have finished
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const int row = quarter * 32 + lane;
        uint8_t* panel = panels + pi * W_PANEL;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          Vec16<bf16> ov;
          ov.pack(v + 8 * q);
          *reinterpret_cast<uint4*>(panel + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = ov.raw;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (et == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmC)),
                       "r"(smem_u32(panel)), "r"(n0 + c0), "r"(m0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        pi ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty[buf]);
      if (tr && t == 0) WTRACE(6);
    }
    if (g.trace && grp == 0 && et == 0) {
      WTRACE(7);
      WTRACE_ADD(11, wait_epi);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem read out before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::TMEM_COLS)
                 : "memory");
  }
  if (threadIdx.x == 0) WTRACE(8);
#undef WTRACE
#undef WTRACE_ADD
}

// =====================================================================================================
// Skinny GEMM for the decode steps: M <= 64 rows (2B rollout rows), weights streamed from HBM once.
// Such a GEMM is bound by the latency of its TMA round trips, not by the MMAs: the generic kernel's
// 4-stage ring needs K/64/4 serial round trips (9.7 us for K = 768, 19.4 us for K = 3072 measured by ncu).
// Here a CTA (a) loads only 64 A rows per stage (8 KiB; the MMA still spans 128 rows - rows 64..127 of the
// accumulator are never read), (b) keeps up to 16 stages in flight so that a whole K slice is requested at
// once, and (c) splits K across blockIdx.y so that enough CTAs pull weights concurrently.  With split-K the
// fp32 partial tiles go to `partial[split][64][N]`; splitk_ln_kernel below reduces them and applies
// bias / GELU / residual / LayerNorm in the same pass (every split GEMM of the decoder is followed by one).
// =====================================================================================================
constexpr int kSkMaxSplit = 8;   // split-K partials of the skinny GEMM (workspace: gemm_skinny_partial_floats)
constexpr int SK_ROWS = 64, SK_A_BYTES = SK_ROWS * BK * 2, SK_MAX_STAGES = 16, SK_SLACK = A_BYTES - SK_A_BYTES;

// EPI selects a LEAN epilogue.  The decode-step kernels are latency-bound launches whose code is fetched cold on
// every launch (~330 KB of distinct kernels per step against a 128 KB instruction cache; ncu: `no_instruction` is the
// third-largest stall of the skinny GEMM), so the chain runs specialisations that carry only the path they execute:
//   0 general (every flag of GemmArgs; test hooks and odd shapes)   1 split-K fp32 partial tile
//   2 bias -> bf16, 16-byte stores   3 bias -> GELU -> bf16   4 bias -> fp32 (LM head; last tile may be partial)
// 1-3 need whole tiles (N % BN == 0) and 16-byte aligned rows, no residual, no head-major store.
template <int BN, int EPI = 0>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_skinny_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB, GemmArgs g,
                                                                  int stages, int kb_per_split,
                                                                  float* __restrict__ partial, int vec_ok) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = SK_A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();   // the next kernel of the chain may start its own prologue now

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + stages * STAGE_BYTES + SK_SLACK);
  uint64_t* empty = full + SK_MAX_STAGES;
  uint64_t* tmem_full = empty + SK_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.x * BN, split = blockIdx.y;
  const int num_kb = (g.K + BK - 1) / BK;
  const int kb0 = split * kb_per_split;
  const int nkb = min(kb_per_split, num_kb - kb0);   // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
#pragma unroll 1   // (code size: every launch of this latency-bound kernel fetches its instructions cold)
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_slot, 0);   // provably warp-uniform: no R2UR waterfall around each tcgen05.mma
  // After a skip (every rollout row finished) only the loads already requested are drained; no MMA, no stores.
  const int npre = min(nkb, stages);

  if (warp == 0) {
    // (uniform control flow + one elected lane per issue: operands stay in uniform registers, see elect_one())
    // weight tiles do not depend on the previous kernel: request the first ring-full BEFORE the dependency wait
    if (elect_one()) {
#pragma unroll 1
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&full[i], STAGE_BYTES);
        tma_load_2d(tiles + i * STAGE_BYTES + SK_A_BYTES, &tmB, (kb0 + i) * BK, n0, &full[i]);
      }
    }
    __syncwarp();
    const bool skip = g.skip_flag && *g.skip_flag;   // stable for the whole step (see splitk_ln_kernel): read ahead of the wait
    pdl_wait();
    if (elect_one()) {
#pragma unroll 1
      for (int i = 0; i < npre; ++i) tma_load_2d(tiles + i * STAGE_BYTES, &tmA, (kb0 + i) * BK, 0, &full[i]);
    }
    __syncwarp();
    if (!skip) {
#pragma unroll 1
      for (int i = npre; i < nkb; ++i) {
        const int s = i % stages;
        const uint32_t ph = (i / stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* a_dst = tiles + s * STAGE_BYTES;
          tma_load_2d(a_dst, &tmA, (kb0 + i) * BK, 0, &full[s]);
          tma_load_2d(a_dst + SK_A_BYTES, &tmB, (kb0 + i) * BK, n0, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const bool skip = g.skip_flag && *g.skip_flag;   // stable for the whole step (see splitk_ln_kernel): read ahead of the wait
    pdl_wait();
    const int n_do = skip ? npre : nkb;
    constexpr uint32_t idesc = make_idesc(BN);
    const uint32_t tiles_addr = smem_u32(tiles);
#pragma unroll 1
    for (int i = 0; i < n_do; ++i) {
      const int s = i % stages;
      const uint32_t ph = (i / stages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t a_addr = tiles_addr + static_cast<uint32_t>(s * STAGE_BYTES);
      const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + SK_A_BYTES);
      if (elect_one()) {
        if (!skip) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                 (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (i == n_do - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else if (EPI != 0 && warp % 4 < 2) {
   if constexpr (EPI != 0) {
    // ---- lean epilogues (see the template comment) ------------------------------------------------
    const int quarter = warp % 4;
    constexpr int W = BN >= 32 ? 32 : BN;
    // (the 128-wide LM-head tile fetches its bias per 32-column chunk: 128 prefetched values would be 128 registers)
    constexpr bool kBiasAhead = EPI != 1 && BN <= 64;
    float bpre[kBiasAhead ? BN : 32];
    if (kBiasAhead) {
#pragma unroll
      for (int j = 0; j < BN; ++j) bpre[j] = (g.bias && (EPI != 4 || n0 + j < g.N)) ? g.bias[n0 + j] : 0.f;
    }
    const bool skip = g.skip_flag && *g.skip_flag;   // stable for the whole step (see splitk_ln_kernel): read ahead of the wait
    pdl_wait();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const long long m = quarter * 32 + lane;
    const bool row_ok = m < g.M && !skip;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (skip || (EPI == 4 && n0 + c0 >= g.N)) break;
      uint32_t r[32];
      if (EPI != 1 && !kBiasAhead) {
#pragma unroll
        for (int j = 0; j < 32; ++j) bpre[j] = (g.bias && n0 + c0 + j < g.N) ? g.bias[n0 + c0 + j] : 0.f;
      }
      if (BN >= 32) tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
      else tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
      if (!row_ok) continue;
      const int nb = n0 + c0;
      const int cb = kBiasAhead ? c0 : 0;   // where this chunk's bias sits in bpre
      if (EPI == 1) {
        float4* dst = reinterpret_cast<float4*>(partial + (static_cast<long long>(split) * SK_ROWS + m) * g.N + nb);
#pragma unroll
        for (int q = 0; q < W / 4; ++q)
          dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                               __uint_as_float(r[4 * q + 3]));
      } else if (EPI == 2 || EPI == 3) {
        bf16* cp = static_cast<bf16*>(g.C) + m * g.ldc + nb;
#pragma unroll
        for (int q = 0; q < W / 8; ++q) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = __uint_as_float(r[8 * q + j]) + bpre[cb + 8 * q + j];
            if (EPI == 3) v[j] = gelu_fast(v[j]);
          }
          Vec16<bf16> ov;
          ov.pack(v);
          ov.store(cp + 8 * q);
        }
      } else {
        float* cp = static_cast<float*>(g.C) + m * g.ldc + nb;
        if (nb + 32 <= g.N) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(cp + 4 * q) =
                make_float4(__uint_as_float(r[4 * q]) + bpre[cb + 4 * q], __uint_as_float(r[4 * q + 1]) + bpre[cb + 4 * q + 1],
                            __uint_as_float(r[4 * q + 2]) + bpre[cb + 4 * q + 2], __uint_as_float(r[4 * q + 3]) + bpre[cb + 4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < g.N) cp[j] = __uint_as_float(r[j]) + bpre[cb + j];
        }
      }
    }
   }
  } else if (warp % 4 < 2) {
    // ---- epilogue: warps 4,5 own TMEM lanes 0..63 = the 64 real rows ---------------------------
    const int quarter = warp % 4;
    // the bias is a constant of the chain: fetch this CTA's slice before the dependency wait / the accumulator
    constexpr int BW = BN >= 32 ? BN : 32;
    if constexpr (EPI == 0) {
    float bpre[BW];
    const bool pre_bias = g.bias != nullptr && partial == nullptr;
    if (pre_bias) {
#pragma unroll
      for (int j = 0; j < BW; ++j) bpre[j] = (j < BN && n0 + j < g.N) ? g.bias[n0 + j] : 0.f;
    }
    const bool skip = g.skip_flag && *g.skip_flag;   // stable for the whole step (see splitk_ln_kernel): read ahead of the wait
    pdl_wait();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const long long m = quarter * 32 + lane;
    const bool row_ok = m < g.M && !skip;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= g.N || skip) break;
      uint32_t r[32];
      if (BN >= 32) {
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
      } else {
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
#pragma unroll
        for (int j = 16; j < 32; ++j) r[j] = 0u;
      }
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int nb = n0 + c0;
      if (partial) {
        float* dst = partial + (static_cast<long long>(split) * SK_ROWS + m) * g.N + nb;
        constexpr int W = BN >= 32 ? 32 : BN;
        if (nb + W <= g.N && (g.N % 4) == 0) {
#pragma unroll
          for (int q = 0; q < W / 4; ++q)
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (nb + j < g.N) dst[j] = v[j];
        }
      } else {
        GemmArgs gl = g;
        if (BN < 32) gl.N = min(g.N, nb + BN);   // columns beyond this CTA's tile belong to its neighbour
        if (pre_bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bpre[c0 + j];
          gl.bias = nullptr;
        }
        epilogue_chunk(gl, v, m, nb, (BN >= 32) ? vec_ok : 0);
      }
    }
    }
  } else {
    pdl_wait();   // every thread of a chain kernel passes the dependency wait
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}



// out[m, :] = LayerNorm( act( sum_s partial[s][m][:] + bias ) + residual[m, :] ) for m < M; one CTA per row.
__global__ void __launch_bounds__(256) splitk_ln_kernel(const float* __restrict__ partial, int nsplit, int N,
                                                        const float* __restrict__ bias, int act,
                                                        const bf16* __restrict__ residual, int ldr,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        bf16* __restrict__ out, int ldo,
                                                        const int* __restrict__ skip_flag,
                                                        const float* __restrict__ res_gamma,
                                                        const float* __restrict__ res_beta) {
  __shared__ float sh[8];
  pdl_launch_dependents();
  const int m = blockIdx.x, tid = threadIdx.x;
  const int c = tid * 4;
  const bool on = c < N;   // N % 4 == 0, N <= 1024
  // parameters are constants of the chain: fetch them while the producing GEMM is still running
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f), gm = bs, bt = bs, rg = bs, rb = bs;
  if (on) {
    if (bias) bs = *reinterpret_cast<const float4*>(bias + c);
    gm = *reinterpret_cast<const float4*>(gamma + c);
    bt = *reinterpret_cast<const float4*>(beta + c);
    if (res_gamma) {
      rg = *reinterpret_cast<const float4*>(res_gamma + c);
      rb = *reinterpret_cast<const float4*>(res_beta + c);
    }
  }
  // the flag is written by the PREVIOUS step's last kernel and a step opens with a fully serialised launch: it is
  // stable for the whole step, so its load rides ahead of the dependency wait instead of gating the partial loads
  const bool skip = skip_flag && *skip_flag;
  pdl_wait();
  if (skip) return;
  auto block_sum = [&](float x) {
    x = warp_sum(x);
    __syncthreads();
    if (tid % kWarp == 0) sh[tid / kWarp] = x;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    return t;
  };
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  float rln[4] = {0.f, 0.f, 0.f, 0.f};   // residual as a recomputed LayerNorm output
  if (res_gamma) {
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (on) {
      const uint2 rr = *reinterpret_cast<const uint2*>(residual + static_cast<long long>(m) * ldr + c);
      x[0] = __uint_as_float(rr.x << 16); x[1] = __uint_as_float(rr.x & 0xffff0000u);
      x[2] = __uint_as_float(rr.y << 16); x[3] = __uint_as_float(rr.y & 0xffff0000u);
    }
    const float rmean = block_sum(x[0] + x[1] + x[2] + x[3]) / N;
    float q2 = 0.f;
    if (on) {
#pragma unroll
      for (int j = 0; j < 4; ++j) q2 += (x[j] - rmean) * (x[j] - rmean);
    }
    const float rrstd = rsqrtf(block_sum(q2) / N + eps);
    const float g4[4] = {rg.x, rg.y, rg.z, rg.w}, b4[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) rln[j] = __bfloat162float(__float2bfloat16_rn((x[j] - rmean) * rrstd * g4[j] + b4[j]));
  }
  if (on) {
    // all loads first (one L2 round trip instead of a dependent chain), then the arithmetic
    float4 p[kSkMaxSplit];
    uint2 rr = make_uint2(0u, 0u);
#pragma unroll
    for (int s = 0; s < kSkMaxSplit; ++s)
      p[s] = (s < nsplit) ? __ldcg(reinterpret_cast<const float4*>(partial + (static_cast<long long>(s) * SK_ROWS + m) * N + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    if (residual && !res_gamma) rr = *reinterpret_cast<const uint2*>(residual + static_cast<long long>(m) * ldr + c);
#pragma unroll
    for (int s = 0; s < kSkMaxSplit; ++s) {
      v[0] += p[s].x; v[1] += p[s].y; v[2] += p[s].z; v[3] += p[s].w;
    }
    v[0] += bs.x; v[1] += bs.y; v[2] += bs.z; v[3] += bs.w;
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
    }
    if (res_gamma) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += rln[j];
    } else if (residual) {
      v[0] += __uint_as_float(rr.x << 16); v[1] += __uint_as_float(rr.x & 0xffff0000u);
      v[2] += __uint_as_float(rr.y << 16); v[3] += __uint_as_float(rr.y & 0xffff0000u);
    }
    // the reference rounds the pre-LN sum to the storage type (bf16 GEMM output) before normalising
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
  }
  const float mean = block_sum(on ? v[0] + v[1] + v[2] + v[3] : 0.f) / N;
  float d2 = 0.f;
  if (on) {
#pragma unroll
    for (int j = 0; j < 4; ++j) d2 += (v[j] - mean) * (v[j] - mean);
  }
  const float var = block_sum(d2) / N;
  const float rstd = rsqrtf(var + eps);
  if (on) {
    const float o0 = (v[0] - mean) * rstd * gm.x + bt.x, o1 = (v[1] - mean) * rstd * gm.y + bt.y;
    const float o2 = (v[2] - mean) * rstd * gm.z + bt.z, o3 = (v[3] - mean) * rstd * gm.w + bt.w;
    __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), b = __floats2bfloat162_rn(o2, o3);
    uint2 st;
    st.x = *reinterpret_cast<uint32_t*>(&a);
    st.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + static_cast<long long>(m) * ldo + c) = st;
  }
}


// ---- host side -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [rows, cols] with row pitch ld elements; box = [64 cols, box_rows rows]; 128-byte swizzle
CUtensorMap make_map(const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}

// C tile for the TMA-store epilogue: bf16 [rows, cols], box = 32 columns (64 B) x 128 rows, 64-byte swizzle
CUtensorMap make_map_out(const void* ptr, long long rows, long long cols, long long ld) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(BM)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled (output) failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}

// Tile width and ring depth.  Shared memory per CTA = stages * (16 KiB + BN * 128 B); the aim is two or more CTAs
// per SM (one CTA's epilogue then overlaps another's mainloop) unless K is long enough for the mainloop to dominate.
void pick_tile(int M, int N, int K, int* bn_out, int* stages_out) {
  const int num_kb = ceil_div(K, BK);
  int bn;
  if (N <= 32) bn = 32;
  else if (N <= 64) bn = 64;
  else if (N % 128 == 0) bn = (num_kb >= 16 && N % 256 == 0) ? 256 : 128;
  else if (N % 96 == 0) bn = 96;
  else if (N <= 128) bn = 128;
  else bn = (N % 192 == 0 && N <= 768) ? 192 : 256;
  if (M <= 256 && N >= 512) bn = 32;   // few rows: many narrow tiles so that enough CTAs pull weights concurrently
  int stages = std::min(num_kb, bn >= 192 ? 4 : 3);
  *bn_out = bn;
  *stages_out = std::max(stages, 1);
}

int persist_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    CXRM_CUDA_CHECK(cudaGetDevice(&dev));
    CXRM_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

template <int BN>
void launch(const GemmArgs& g, int stages, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::smem(MAX_STAGES)));
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::smem(MAX_STAGES)));
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::smem(MAX_STAGES)));
    configured = true;
  }
  const CUtensorMap ta = make_map(g.A, g.M, g.K, g.lda, BM);
  const CUtensorMap tb = make_map(g.W, g.N, g.K, g.ldw, BN);
  const int esz = g.out_f32 ? 4 : 2;
  int vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((static_cast<long long>(g.ldc) * esz) % 16 == 0);
  if (g.residual)
    vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(g.residual) % 16 == 0) && (g.ldr % 8 == 0);
  // TMA-store epilogue: bf16 row-major output, 16-byte aligned rows, whole 32-column panels inside N's padding rules
  // (a partial last panel is clipped by the tensor map); the staging panels reuse the pipeline stages.
  const int tma_store = (!g.out_f32 && g.c_head_stride == 0 && vec_ok && g.N % 8 == 0 &&
                         stages * Cfg<BN>::STAGE_BYTES >= BM * BN * 2 && (!g.residual || g.N % 32 == 0)) ? 1 : 0;
  const CUtensorMap tc = tma_store ? make_map_out(g.C, g.M, g.N, g.ldc) : ta;
  dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN));
  // enough tiles for every SM to overlap epilogues with mainloops: the persistent kernel (residual rows must be
  // 16-byte loadable there: guaranteed by vec_ok, which tma_store implies)
  static const bool persist_ok = std::getenv("CXRM_NO_PERSIST_GEMM") == nullptr;
  const int n_tiles = static_cast<int>(grid.x * grid.y);
  // ... and a mainloop long enough to hide an epilogue behind: with K < 384 and narrow N the tile is all epilogue, and
  // one persistent CTA per SM has only 4 epilogue warps where 2-3 co-resident one-tile CTAs have 8-12 (measured:
  // [256x64] 82 -> 121 us, [192x192] 17 -> 19.5 us persistent; [1536x384] 62 -> 57, [3072x768] 194 -> 160 us)
  if (persist_ok && tma_store && n_tiles >= 2 * persist_sms() && (g.N % 32 == 0) && (g.K >= 384 || g.N >= 512)) {
    static bool pconf = false;
    if (!pconf) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::smem()));
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::smem()));
      pconf = true;
    }
    const int ctas = std::min(n_tiles, persist_sms());
    if (g.act == ACT_GELU)
      launch_chain(gemm_tc_persist_kernel<BN, 2>, dim3(ctas), dim3(NTHREADS), PCfg<BN>::smem(), stream, ta, tb, tc, g, vec_ok, static_cast<int>(grid.x), n_tiles);
    else
      launch_chain(gemm_tc_persist_kernel<BN, 1>, dim3(ctas), dim3(NTHREADS), PCfg<BN>::smem(), stream, ta, tb, tc, g, vec_ok, static_cast<int>(grid.x), n_tiles);
    check_launch("gemm_tcgen05_persist");
    return;
  }
  const size_t smem = Cfg<BN>::smem(stages);
  if (tma_store && g.act == ACT_GELU)
    launch_chain(gemm_tc_kernel<BN, 2>, grid, dim3(NTHREADS), smem, stream, ta, tb, tc, g, vec_ok, stages, tma_store);
  else if (tma_store)
    launch_chain(gemm_tc_kernel<BN, 1>, grid, dim3(NTHREADS), smem, stream, ta, tb, tc, g, vec_ok, stages, tma_store);
  else
    launch_chain(gemm_tc_kernel<BN, 0>, grid, dim3(NTHREADS), smem, stream, ta, tb, tc, g, vec_ok, stages, tma_store);
  check_launch("gemm_tcgen05");
}


// ---- weight-resident path -----------------------------------------------------------------------------
// Picks the widest column slice whose weights fit beside a >= 4-stage A ring; 0: not eligible.
int wres_pick(const GemmArgs& g, int* stages_out, int* n_slices_out, int* grid_out) {
  static const bool off = std::getenv("CXRM_NO_WRES_GEMM") != nullptr;
  if (off || g.out_f32 || g.c_head_stride != 0 || g.K > 384 || g.N % 32 != 0 || g.skip_flag) return 0;
  const int esz = 2;
  bool vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((static_cast<long long>(g.ldc) * esz) % 16 == 0);
  if (g.residual) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(g.residual) % 16 == 0) && (g.ldr % 8 == 0);
  if (!vec_ok) return 0;
  const int nkb = ceil_div(g.K, BK), tiles_m = ceil_div(g.M, BM);
  const int cand[6] = {256, 192, 128, 96, 64, 32};
  for (int bn : cand) {
    if (g.N % bn != 0) continue;
    if (bn < 128 && bn != g.N) break;      // never slice an output narrower than 128 columns (N <= 128 itself is one slice)
    const int n_slices = g.N / bn;
    if (n_slices > persist_sms()) continue;
    // >= 3 A stages: a 192-wide slice of a K = 384 weight (147 KB) leaves room for three, and the wider tile is worth it
    // (95 % instead of 66 % of the tensor pipe per MMA: [18464 x 1536 x 384] 46.8 -> 39.4 us, [18432 x 768 x 384] 22.8 -> 19.1)
    static const int min_stages = std::getenv("CXRM_WRES_MINSTAGES") ? std::atoi(std::getenv("CXRM_WRES_MINSTAGES")) : 3;
    int stages = W_MAX_STAGES;
    while (stages >= min_stages && wres_smem(bn, nkb, stages) > 227 * 1024) --stages;
    if (stages < min_stages) continue;
    const int grid = (persist_sms() / n_slices) * n_slices;
    if (tiles_m < 2 * (grid / n_slices)) continue;     // too few row tiles per CTA to amortise this slice's weights: try narrower
    *stages_out = stages; *n_slices_out = n_slices; *grid_out = grid;
    return bn;
  }
  return 0;
}

template <int BN>
void launch_wres(const GemmArgs& g, int stages, int n_slices, int grid, cudaStream_t stream) {
  const int nkb = ceil_div(g.K, BK);
  const size_t smem = wres_smem(BN, nkb, stages);
  static size_t configured = 0;
  if (smem > configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_wres_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_wres_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  const CUtensorMap ta = make_map(g.A, g.M, g.K, g.lda, BM);
  const CUtensorMap tb = make_map(g.W, g.N, g.K, g.ldw, BN);
  const CUtensorMap tc = make_map_out(g.C, g.M, g.N, g.ldc);
  const int tiles_m = ceil_div(g.M, BM);
  if (g.act == ACT_GELU)
    launch_chain(gemm_tc_wres_kernel<BN, 2>, dim3(grid), dim3(W_THREADS), smem, stream, ta, tb, tc, g, tiles_m, n_slices, stages);
  else
    launch_chain(gemm_tc_wres_kernel<BN, 1>, dim3(grid), dim3(W_THREADS), smem, stream, ta, tb, tc, g, tiles_m, n_slices, stages);
  check_launch("gemm_tcgen05_wres");
}

// ---- skinny path (M <= 64) -------------------------------------------------------------------------
template <int BN>
void launch_skinny(const GemmArgs& g, int stages, int nsplit, int kb_per_split, float* partial, cudaStream_t stream) {
  // tiles | slack | barriers (512 B) | alignment
  const size_t smem = static_cast<size_t>(stages) * (SK_A_BYTES + BN * BK * 2) + SK_SLACK + 1024 + 512;
  const CUtensorMap ta = make_map(g.A, g.M, g.K, g.lda, SK_ROWS);
  const CUtensorMap tb = make_map(g.W, g.N, g.K, g.ldw, BN);
  const int esz = g.out_f32 ? 4 : 2;
  int vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((static_cast<long long>(g.ldc) * esz) % 16 == 0);
  if (g.residual) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(g.residual) % 16 == 0) && (g.ldr % 8 == 0);
  dim3 grid(ceil_div(g.N, BN), nsplit);
  // lean epilogue when the call fits one (always the case for the decode step), else the general kernel
  int epi = 0;
  static const bool lean = std::getenv("CXRM_NO_LEAN_EPILOGUE") == nullptr;
  if (lean && g.c_head_stride == 0 && !g.residual && vec_ok) {
    if (partial && g.N % BN == 0) epi = 1;
    else if (!partial && !g.out_f32 && g.N % BN == 0 && g.M <= SK_ROWS) epi = g.act == ACT_GELU ? 3 : 2;
    else if (!partial && g.out_f32 && g.act == ACT_NONE && BN >= 64) epi = 4;
  }
  auto go = [&](auto kern) {
    // keyed by the function: the instantiations share one function-pointer type, so a static in this lambda would too
    static std::unordered_map<const void*, size_t> configured;
    size_t& have = configured[reinterpret_cast<const void*>(kern)];
    if (smem > have) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      have = smem;
    }
    launch_chain(kern, grid, dim3(NTHREADS), smem, stream, ta, tb, g, stages, kb_per_split, partial, vec_ok);
  };
  switch (epi) {
    case 1: go(gemm_tc_skinny_kernel<BN, 1>); break;
    case 2: go(gemm_tc_skinny_kernel<BN, 2>); break;
    case 3: go(gemm_tc_skinny_kernel<BN, 3>); break;
    case 4: go(gemm_tc_skinny_kernel<BN, 4>); break;
    default: go(gemm_tc_skinny_kernel<BN, 0>); break;
  }
  check_launch("gemm_tcgen05_skinny");
}

// the 128-wide tile exists for the fp32 LM head only (EPI 4); anything else falls back to 64-wide tiles
template <int BN>
void launch_skinny_lm(const GemmArgs& g, int stages, int nsplit, int kb_per_split, float* partial, cudaStream_t stream) {
  const int esz = g.out_f32 ? 4 : 2;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((static_cast<long long>(g.ldc) * esz) % 16 == 0);
  static const bool lean = std::getenv("CXRM_NO_LEAN_EPILOGUE") == nullptr;
  if (!(lean && g.c_head_stride == 0 && !g.residual && vec_ok && !partial && g.out_f32 && g.act == ACT_NONE && nsplit == 1)) {
    launch_skinny<64>(g, std::min(stages + 2, kb_per_split), nsplit, kb_per_split, partial, stream);
    return;
  }
  const size_t smem = static_cast<size_t>(stages) * (SK_A_BYTES + BN * BK * 2) + SK_SLACK + 1024 + 512;
  const CUtensorMap ta = make_map(g.A, g.M, g.K, g.lda, SK_ROWS);
  const CUtensorMap tb = make_map(g.W, g.N, g.K, g.ldw, BN);
  static size_t have = 0;
  if (smem > have) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_skinny_kernel<BN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    have = smem;
  }
  launch_chain(gemm_tc_skinny_kernel<BN, 4>, dim3(ceil_div(g.N, BN), 1), dim3(NTHREADS), smem, stream, ta, tb, g, stages, kb_per_split, partial, 1);
  check_launch("gemm_tcgen05_skinny_lm");
}

}  // namespace

CUtensorMap make_tensor_map_bf16(const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                                 int box_cols) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}


// A row-major bf16 matrix [rows, cols] (row pitch ld) seen as [cols / 64 k-blocks][rows][64]: one TMA box
// {64, box_rows, box_kblocks} lands in shared memory as box_kblocks consecutive K-major 128-byte-swizzled tiles of
// box_rows x 64 - the operand layout of tcgen05.mma - with ONE instruction instead of one per k-block.
CUtensorMap make_tensor_map_bf16_kblocks(const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                                         int box_kblocks) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 64)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_kblocks)};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled (k-block view) failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}

int gemm_skinny_supported(const GemmArgs& g) {
  if (gemm_tcgen05_supported(g) != 0) return 1;
  if (g.M > SK_ROWS) return 2;
  return 0;
}

// nsplit_out: number of K splits written to `partial` (0 when the result was stored directly through the epilogue)
// LM head tile width: every tcgen05.mma costs >= 93 clocks whatever its width (section 4e), so 128-wide tiles halve the
// MMA time of the 64-wide ones: 14.2 -> 11.9 us of the step's critical path (tools/trace_step.py); 235 tiles, two CTAs
// per SM.  (208-wide tiles - 145 tiles, one CTA on each of 145 SMs - were slower, 13.8 us: nothing overlaps the epilogue.)
static int lm_head_bn() {
  static const int forced = std::getenv("CXRM_LM_BN") ? std::atoi(std::getenv("CXRM_LM_BN")) : 0;
  return forced == 64 ? 64 : 128;
}
int gemm_skinny_tile_n(const GemmArgs& g, bool split_allowed) {
  if (g.N >= 8192) return lm_head_bn();   // LM head: 128-wide tiles halve the MMA instructions (each >= 93 clocks, section 4e)
  if (!split_allowed && ceil_div(g.N, 32) < 48 && g.c_head_stride == 0) return 16;   // no split possible: narrower tiles
  return 32;
}

void gemm_tcgen05_skinny(const GemmArgs& g, float* partial, int* nsplit_out, cudaStream_t stream) {
  CXRM_CHECK(gemm_skinny_supported(g) == 0, "shape not supported by the skinny tcgen05 GEMM");
  const int num_kb = ceil_div(g.K, BK);
  const int bn = gemm_skinny_tile_n(g, partial != nullptr);
  int stages = bn == 128 ? 4 : bn == 64 ? 6 : SK_MAX_STAGES;   // LM head: two CTAs per SM
  // split-K: every tcgen05.mma of this kernel costs ~93 clocks whatever its width (the 128 x 16 A slice is read from shared
  // memory per instruction: tools/mb_mma.cu), so a CTA's floor is 4 * k-blocks * 93 clocks - 2.3 us for K = 768 unsplit.
  // The GEMMs that feed the reduce + LayerNorm kernel split as far as kSkMaxSplit partials of >= kb_min k-blocks.
  // (The direct-epilogue GEMMs - QKV, cross-Q, FFN-up - were tried with split-K across a 2- / 4-CTA cluster reduced
  // through distributed shared memory: correct, but 152.3 / 159.6 ms of rollout against 146.7 without - the cluster
  // launch and its barrier cost more than the 24-36 MMAs they save; removed)
  static const int max_split = std::getenv("CXRM_SK_MAXSPLIT") ? std::max(1, std::min(kSkMaxSplit, std::atoi(std::getenv("CXRM_SK_MAXSPLIT")))) : kSkMaxSplit;
  static const int kb_min = std::getenv("CXRM_SK_KBMIN") ? std::max(1, std::atoi(std::getenv("CXRM_SK_KBMIN"))) : 3;   // measured: 8 x >= 3 k-blocks, -3.2 ms per SCST step vs 4 x >= 6
  int nsplit = 1;
  if (partial && ceil_div(g.N, bn) < 64) nsplit = std::max(1, std::min(max_split, num_kb / kb_min));
  const int kb_per_split = ceil_div(num_kb, nsplit);
  nsplit = ceil_div(num_kb, kb_per_split);
  stages = std::min(stages, kb_per_split);
  {   // experiment knob (tools/exp_two_stream.py): cap the ring depth so that the CTA fits beside an attention CTA
    static const int cap = std::getenv("CXRM_SK_STAGES") ? std::max(2, std::atoi(std::getenv("CXRM_SK_STAGES"))) : SK_MAX_STAGES;
    stages = std::min(stages, cap);
  }
  if (nsplit_out) *nsplit_out = partial ? nsplit : 0;
  switch (bn) {
    case 16: launch_skinny<16>(g, stages, nsplit, kb_per_split, partial, stream); break;
    case 64: launch_skinny<64>(g, stages, nsplit, kb_per_split, partial, stream); break;
    case 128: launch_skinny_lm<128>(g, stages, nsplit, kb_per_split, partial, stream); break;
    default: launch_skinny<32>(g, stages, nsplit, kb_per_split, partial, stream); break;
  }
}

void splitk_ln(const float* partial, int nsplit, int M, int N, const float* bias, int act, const void* residual,
               int ldr, const float* gamma, const float* beta, float eps, void* out, int ldo, const int* skip_flag,
               cudaStream_t stream, const float* res_gamma, const float* res_beta) {
  CXRM_CHECK(!res_gamma || (residual && res_beta), "splitk_ln: residual LayerNorm needs the pre-LN residual rows");
  CXRM_CHECK(N % 4 == 0 && N <= 1024 && M <= SK_ROWS && ldo % 4 == 0 && (!residual || ldr % 4 == 0), "splitk_ln shape");
  // (a one-warp-per-row version - every load in flight at once, shuffle reductions - measured 7.6 ms per SCST step
  // SLOWER than this 256-thread block per row: 191.9 vs 184.3 ms)
  launch_chain(splitk_ln_kernel, dim3(M), dim3(256), 0, stream, partial, nsplit, N, bias, act,
               static_cast<const bf16*>(residual), ldr, gamma, beta, eps, static_cast<bf16*>(out), ldo, skip_flag, res_gamma,
               res_beta);
  check_launch("splitk_ln");
}

size_t gemm_skinny_partial_floats(int N) { return static_cast<size_t>(kSkMaxSplit) * SK_ROWS * N; }

namespace {
}  // namespace

int gemm_tcgen05_supported(const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 1;
  if (g.K % 8 != 0 || g.lda % 8 != 0 || g.ldw % 8 != 0) return 2;
  if (reinterpret_cast<uintptr_t>(g.A) % 16 != 0 || reinterpret_cast<uintptr_t>(g.W) % 16 != 0) return 3;
  if (ceil_div(g.N, 32) > 65535) return 4;
  if (g.c_head_stride > 0 && (g.N % 64 != 0 || g.out_f32 || g.residual || reinterpret_cast<uintptr_t>(g.C) % 16 != 0 ||
                              g.c_head_stride % 8 != 0))
    return 5;
  return 0;
}

void gemm_tcgen05(const GemmArgs& g, cudaStream_t stream) {
  CXRM_CHECK(gemm_tcgen05_supported(g) == 0, "shape/alignment not supported by the tcgen05 GEMM");
  int bn = 128, stages = 3;
  {
    int wst = 0, wsl = 0, wgrid = 0;
    switch (wres_pick(g, &wst, &wsl, &wgrid)) {
      case 32: launch_wres<32>(g, wst, wsl, wgrid, stream); return;
      case 64: launch_wres<64>(g, wst, wsl, wgrid, stream); return;
      case 96: launch_wres<96>(g, wst, wsl, wgrid, stream); return;
      case 128: launch_wres<128>(g, wst, wsl, wgrid, stream); return;
      case 192: launch_wres<192>(g, wst, wsl, wgrid, stream); return;
      case 256: launch_wres<256>(g, wst, wsl, wgrid, stream); return;
      default: break;
    }
  }
  pick_tile(g.M, g.N, g.K, &bn, &stages);
  switch (bn) {
    case 32: launch<32>(g, stages, stream); break;
    case 64: launch<64>(g, stages, stream); break;
    case 96: launch<96>(g, stages, stream); break;
    case 128: launch<128>(g, stages, stream); break;
    case 192: launch<192>(g, stages, stream); break;
    default: launch<256>(g, stages, stream); break;
  }
}

}  // namespace cxrm
