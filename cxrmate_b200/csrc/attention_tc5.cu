// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the dense, unmasked attention of the
// CvT encoder (reference: HF modeling_cvt.py:166-244; scale C^-1/2, heads of 64).  Replaces the mma.sync kernel of
// attention.cu for those calls (ncu there: HMMA + LDSM only, tensor pipe 33 %, 197 TF/s).
//
// One CTA = 128 queries of one (image, head); key tiles of 192 (CvT's key counts 2304 / 576 / 145 are 12 / 3 / 1 tiles).
// Per tile:
//   warp 0   TMA: K tile and V tile, both [192 keys x 64 dims] exactly as they lie in memory (one buffer each: K_(i+1)
//            lands during softmax_i / P.V_i, V_(i+1) during S_(i+1) and its softmax)
//   warp 1   S = Q.K^T   : 4 x tcgen05.mma 128 x 192 x 16 into TMEM columns 0..191 (N = 192 keeps the pipe 95 % busy; the
//                          measured floor of ~93 clocks per MMA makes N <= 128 no cheaper)
//            O_t = P.V   : 12 x tcgen05.mma 128 x 64 x 16 into TMEM columns 192..255; P from shared memory (K-major), V as
//                          an MN-MAJOR B operand (instruction-descriptor bit 16): rows = keys = K, 128 contiguous bytes =
//                          the 64 output dimensions, so no transposed copy of V exists (a first version transposed V per
//                          head in a separate kernel: 42 launches, 0.6 ms per 100 images)
//            a ragged last tile issues N (and P.V k-steps) for its own keys only, rounded up to 16
//   warps 2-5 (one query row per thread): tcgen05.ld of S (two passes over 32-column chunks: row maximum, then
//            p = 2^(s.scale.log2e - m) -> bf16 -> the K-major, 128-byte-swizzled P tile in shared memory), then
//            o = o.alpha + O_t from TMEM.  The running output lives in registers, so nothing in TMEM is rescaled.
// A CTA is serial over its tiles (S -> softmax -> P.V -> accumulate; S_(i+1) overlaps the accumulate); two CTAs per SM
// (112 KB of shared memory, 256 TMEM columns each) overlap one's softmax with the other's MMAs.
#include <cuda.h>

#include <mutex>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int QT = 128, KT = 192, HD5 = 64, NTH = 192;
constexpr int KB5 = KT / 64;                                  // 64-key k-blocks of the P.V product
constexpr int Q_BYTES = QT * 128, K_BYTES = KT * 128, VT_BYTES = KB5 * 64 * 128, P_BYTES = KB5 * QT * 128;
constexpr int SMEM5 = Q_BYTES + K_BYTES + VT_BYTES + P_BYTES + 256;   // 112.25 KB: 2 x (this + 1 KB reserved) fits an SM's 228 KB
constexpr int S_COLS = KT, TM_COLS = 256;                     // TMEM: S in columns 0..191, the tile's P.V in 192..255
static_assert(S_COLS + HD5 <= TM_COLS, "TMEM budget");

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect1() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// K-major operand tile, 128-byte swizzle: rows of 64 bf16, 8-row groups 1024 B apart (same form as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t desc5(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc5(int n) {   // kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void mma5(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit5(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// pass 1 of a tile: row maximum of the raw scores (MASK: the tile is ragged, keys >= nk do not exist)
template <bool MASK>
__device__ __forceinline__ float tile_row_max(uint32_t t_s, int n_chunks, int nk) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < n_chunks; ++c) {
    uint32_t r[32];
    ld32(t_s + c * 32, r);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float a = (!MASK || c * 32 + j < nk) ? __uint_as_float(r[j]) : -INFINITY;
      const float b = (!MASK || c * 32 + j + 1 < nk) ? __uint_as_float(r[j + 1]) : -INFINITY;
      mx = fmaxf(mx, fmaxf(a, b));
    }
  }
  return mx;
}
// pass 2: p = 2^(s * sl2 - m) -> bf16 -> the swizzled K-major P tile; returns the row sum
template <bool MASK>
__device__ __forceinline__ float tile_probs(uint32_t t_s, uint8_t* sP, int row, int n_chunks, int nk, float sl2, float m_new) {
  float ps0 = 0.f, ps1 = 0.f;
#pragma unroll 1
  for (int c = 0; c < n_chunks; ++c) {
    uint32_t r[32];
    ld32(t_s + c * 32, r);
    uint8_t* prow = sP + (c >> 1) * (QT * 128) + row * 128;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float e = ex2f(fmaf(__uint_as_float(r[g8 * 8 + j]), sl2, -m_new));
        p[j] = (!MASK || c * 32 + g8 * 8 + j < nk) ? e : 0.f;
      }
      ps0 += (p[0] + p[1]) + (p[2] + p[3]);
      ps1 += (p[4] + p[5]) + (p[6] + p[7]);
      const int chunk = (c & 1) * 4 + g8;              // 16-byte chunk of the row inside its 64-key k-block
      *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
          make_uint4(pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]));
    }
  }
  return ps0 + ps1;
}

// registers are allocated per 4 warps: 2 CTAs / SM need <= 128 registers per thread, which is what a 256-thread bound asks for
__global__ void __launch_bounds__(256, 2) attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                               const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ o,
                                                               long long o_ts, int Lq, int Lk, int heads, float sl2,
                                                               const int* __restrict__ q_offset, const int* __restrict__ lq_per_batch,
                                                               const int* __restrict__ kv_offset, const int* __restrict__ lk_per_batch) {
  extern __shared__ uint8_t smem_raw5[];   // no static shared memory in this kernel: the dynamic window starts 1 KB aligned
  pdl_launch_dependents();
  uint8_t* sm = smem_raw5;
  if ((s32(sm) & 1023u) != 0) __trap();    // the 128-byte swizzle atoms need 1 KB alignment
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + Q_BYTES;              // [KT keys][128 B]
  uint8_t* sV = sK + K_BYTES;              // [KT keys][128 B] (MN-major B operand of P.V)
  uint8_t* sP = sV + VT_BYTES;             // [KB5 k-blocks][128 rows][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mb_init(&bars[i], i == 6 ? 4 : 1);   // p_full: one arrival per softmax warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(kFull, *tmem_slot, 0);
  pdl_wait();   // q / k / v (and the offsets of a packed batch) are the previous kernels' outputs
  // packed (ragged) batches: sequence b owns rows q_row0 .. + Lq of q / o and k_row0 .. + Lk of k / v
  const int q_row0 = q_offset ? q_offset[b] : b * Lq, k_row0 = kv_offset ? kv_offset[b] : b * Lk;
  if (lq_per_batch) Lq = lq_per_batch[b];
  if (lk_per_batch) Lk = lk_per_batch[b];
  const bool live = q0 < Lq && Lk > 0;            // a tile past the end of a short sequence only tears down
  const int n_tiles = live ? (Lk + KT - 1) / KT : 0;

  if (warp == 0) {
    // K and V^T have one buffer each: K of tile i+1 lands while tile i is in softmax / P.V (its slot is free once S_i is
    // done), V^T of tile i+1 while S_(i+1) and its softmax run
    if (live && elect1()) {
      mb_expect(q_full, Q_BYTES);
      tma2d(sQ, &tmQ, h * HD5, q_row0 + q0, q_full);
    }
    __syncwarp();
    for (int i = 0; i < n_tiles; ++i) {
      mb_wait(k_empty, (i & 1) ^ 1);
      if (elect1()) {
        mb_expect(k_full, K_BYTES);
        tma2d(sK, &tmK, h * HD5, k_row0 + i * KT, k_full);
      }
      __syncwarp();
      mb_wait(v_empty, (i & 1) ^ 1);
      if (elect1()) {
        mb_expect(v_full, VT_BYTES);
        tma2d(sV, &tmV, h * HD5, k_row0 + i * KT, v_full);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id_o = idesc5(HD5) | (1u << 16);   // B (= V) is MN-major
    const uint64_t dq = desc5(s32(sQ)), dk = desc5(s32(sK));
    const uint32_t aP = s32(sP), aV = s32(sV);
    if (live) mb_wait(q_full, 0);
    for (int i = 0; i < n_tiles; ++i) {
      // a ragged last tile computes only the keys it has, rounded up to the MMA's N / K granularity of 16
      const int n16 = (min(KT, Lk - i * KT) + 15) & ~15;
      const uint32_t id_s = idesc5(n16);
      mb_wait(k_full, i & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect1()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma5(tmem, dq + 2 * k, dk + 2 * k, id_s, k != 0);
        commit5(s_full);
        commit5(k_empty);
      }
      __syncwarp();
      mb_wait(v_full, i & 1);
      mb_wait(p_full, i & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect1()) {
        for (int ks = 0; ks < n16 / 16; ++ks) {
          const int kb = ks >> 2, k = ks & 3;   // P: 64-key k-blocks of [128 rows x 128 B]; V: 16 keys = 16 rows of 128 B per step
          mma5(tmem + S_COLS, desc5(aP + kb * (QT * 128)) + 2 * k, desc5(aV + ks * (16 * 128)), id_o, ks != 0);
        }
        commit5(o_full);
        commit5(v_empty);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp % 4;
    const int row = quarter * 32 + lane;                 // query row of the tile = TMEM lane
    const uint32_t t_s = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_s + S_COLS;
    float o_acc[HD5];
#pragma unroll
    for (int d = 0; d < HD5; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int i = 0; i < n_tiles; ++i) {
      const int nk = min(KT, Lk - i * KT);               // valid keys of this tile (>= 1)
      const int n_chunks = (nk + 31) / 32;               // 32-key chunks that hold a valid key (they cover n16)
      const bool ragged = (nk & 31) != 0;
      mb_wait(s_full, i & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float mx = ragged ? tile_row_max<true>(t_s, n_chunks, nk) : tile_row_max<false>(t_s, n_chunks, nk);
      const float m_new = fmaxf(m_run, mx * sl2);        // sl2 > 0, so the maximum commutes with the scale
      const float alpha = ex2f(m_run - m_new);           // first tile: 2^(-inf) = 0
      m_run = m_new;
      const float ps = ragged ? tile_probs<true>(t_s, sP, row, n_chunks, nk, sl2, m_new)
                              : tile_probs<false>(t_s, sP, row, n_chunks, nk, sl2, m_new);
      l_run = fmaf(l_run, alpha, ps);
      // P is read by the tensor core through the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(p_full);
      mb_wait(o_full, i & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        ld32(t_o + c * 32, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) o_acc[c * 32 + j] = fmaf(o_acc[c * 32 + j], alpha, __uint_as_float(r[j]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (live && q0 + row < Lq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* dst = o + (static_cast<long long>(q_row0) + q0 + row) * o_ts + h * HD5;
#pragma unroll
      for (int d = 0; d < HD5; d += 8) {
        *reinterpret_cast<uint4*>(dst + d) =
            make_uint4(pack2(o_acc[d] * inv, o_acc[d + 1] * inv), pack2(o_acc[d + 2] * inv, o_acc[d + 3] * inv),
                       pack2(o_acc[d + 4] * inv, o_acc[d + 5] * inv), pack2(o_acc[d + 6] * inv, o_acc[d + 7] * inv));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TM_COLS) : "memory");
}

typedef CUresult (*EncodeFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn5 encode5() {
  static EncodeFn5 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    fn = reinterpret_cast<EncodeFn5>(p);
  });
  return fn;
}
CUtensorMap map5(const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode5()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (attention) failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}

}  // namespace

int attention_tc5_supported(const AttnArgs& a) {
  static const bool off = std::getenv("CXRM_NO_TC5_ATTN") != nullptr;
  if (off) return 9;
  if (a.key_mask || a.causal || a.kv_batch_mod) return 1;
  const bool packed = a.q_offset || a.Lq_per_batch || a.kv_offset || a.Lk_per_batch;
  if (packed && !(a.q_offset && a.Lq_per_batch && a.kv_offset && a.Lk_per_batch && a.total_q > 0 && a.total_kv > 0)) return 1;
  const long long C = static_cast<long long>(a.heads) * 64;
  if (a.q_hs != 64 || a.k_hs != 64 || a.v_hs != 64 || a.o_hs != 64) return 2;
  if (a.q_ts < C || a.k_ts < C || a.v_ts < C || a.o_ts < C) return 3;
  if (!packed && (a.q_bs != a.Lq * a.q_ts || a.k_bs != a.Lk * a.k_ts || a.v_bs != a.Lk * a.v_ts || a.o_bs != a.Lq * a.o_ts)) return 4;   // batches contiguous
  if (a.q_ts % 8 || a.k_ts % 8 || a.v_ts % 8 || a.o_ts % 8) return 5;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.o)) return 6;
  if (a.Lq < 64 || a.Lk < 16) return 7;              // tiny problems stay on the mma.sync kernel
  if (a.batch > 65535 || a.heads > 65535) return 8;
  return 0;
}

void attention_tc5(const AttnArgs& a, cudaStream_t stream) {
  CXRM_CHECK(attention_tc5_supported(a) == 0, "attention_tc5: unsupported arguments");
  if (a.batch <= 0 || a.Lq <= 0) return;
  static bool configured = false;
  if (!configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM5));
    // 2 CTAs / SM need the 228 KB carveout (ncu Occupancy: block limit 2 by registers and by shared memory)
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const long long C = static_cast<long long>(a.heads) * 64;
  const bool packed = a.q_offset != nullptr;
  const long long rows_q = packed ? a.total_q : static_cast<long long>(a.batch) * a.Lq;
  const long long rows_kv = packed ? a.total_kv : static_cast<long long>(a.batch) * a.Lk;
  const CUtensorMap tq = map5(a.q, rows_q, C, a.q_ts, QT);
  const CUtensorMap tk = map5(a.k, rows_kv, C, a.k_ts, KT);
  const CUtensorMap tv = map5(a.v, rows_kv, C, a.v_ts, KT);
  launch_chain(attention_tc5_kernel, dim3(ceil_div(a.Lq, QT), a.heads, a.batch), dim3(NTH), SMEM5, stream, tq, tk, tv,
               static_cast<bf16*>(a.o), a.o_ts, a.Lq, a.Lk, a.heads, a.scale * 1.4426950408889634f, a.q_offset, a.Lq_per_batch,
               a.kv_offset, a.Lk_per_batch);
  check_launch("attention_tc5");
}

}  // namespace cxrm
