// Multi-query attention, fp32 math, online softmax (flash style: scores never
// reach HBM).  One kernel covers the four dense-attention sites of the path:
//   * CvT attention, no mask, scale = embed_dim^-0.5  (HF modeling_cvt.py:236-243)
//   * CXR-BERT self-attention with key padding mask    (HF modeling_bert.py:115-139)
//   * decoder prefill self-attention: causal AND key padding (modeling_bert.py:628-691)
//   * decoder prefill cross-attention over the (ragged) encoder K/V cache
// Masked keys are skipped, which equals the reference's additive finfo.min mask
// whenever a row has at least one visible key (always true on this path).
#include "kernels.h"

namespace cxrm {

namespace {

constexpr int BQ = 64, BKV = 64, D = 64, NT = 256, LDT = 68;  // LDT: padded stride, keeps float4 alignment

struct Smem {
  float Qt[D][LDT];     // [d][query]
  float Kt[D][LDT];     // [d][key]
  float Vs[BKV][LDT];   // [key][d]
  float Pt[BKV][LDT];   // [key][query]
};

template <typename T>
__device__ __forceinline__ void load16(const T* p, float* f);
template <>
__device__ __forceinline__ void load16<float>(const float* p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
    f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
  }
}
template <>
__device__ __forceinline__ void load16<bf16>(const bf16* p, float* f) {
  Vec16<bf16> a, b;
  a.load(p);
  b.load(p + 8);
  a.unpack(f);
  b.unpack(f + 8);
}

template <typename T>
__global__ void __launch_bounds__(NT) attention_simt_kernel(AttnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int kvb = a.kv_batch_mod > 0 ? b % a.kv_batch_mod : b;
  const int Lk = a.Lk_per_batch ? a.Lk_per_batch[kvb] : a.Lk;
  const long long koff = a.kv_offset ? a.kv_offset[kvb] : 0;
  const int Lq = a.Lq_per_batch ? a.Lq_per_batch[b] : a.Lq;
  if (q0 >= Lq) return;   // packed queries: the grid is sized for the longest row
  const long long qtok0 = a.q_offset ? a.q_offset[b] : 0;
  const T* __restrict__ Q = static_cast<const T*>(a.q) + (a.q_offset ? qtok0 * a.q_ts : b * a.q_bs) + h * a.q_hs;
  const T* __restrict__ K = static_cast<const T*>(a.k) + kvb * a.k_bs + h * a.k_hs + koff * a.k_ts;
  const T* __restrict__ V = static_cast<const T*>(a.v) + kvb * a.v_bs + h * a.v_hs + koff * a.v_ts;
  const uint8_t* __restrict__ km =
      a.key_mask ? a.key_mask + static_cast<long long>(a.key_mask_per_q_batch ? b : kvb) * a.key_mask_ld : nullptr;

  // loader mapping: row = tid/4 (0..63), 16 consecutive d per thread
  const int lrow = tid >> 2, ld0 = (tid & 3) * 16;
  {
    float f[16];
    if (q0 + lrow < Lq) {
      load16<T>(Q + static_cast<long long>(q0 + lrow) * a.q_ts + ld0, f);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) sm.Qt[ld0 + i][lrow] = f[i];
  }

  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  // causal: keys beyond the last query of this tile are never visible
  int k_end = Lk;
  if (a.causal) {
    const int last_q = min(q0 + BQ, Lq) - 1 + a.q_pos_offset;
    k_end = min(k_end, last_q + 1);
  }

  for (int k0 = 0; k0 < k_end; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Qt stores on the first pass)
    {
      float f[16];
      const bool ok = (k0 + lrow) < Lk;
      if (ok) {
        load16<T>(K + static_cast<long long>(k0 + lrow) * a.k_ts + ld0, f);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) sm.Kt[ld0 + i][lrow] = f[i];
      if (ok) {
        load16<T>(V + static_cast<long long>(k0 + lrow) * a.v_ts + ld0, f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&sm.Vs[lrow][ld0 + 4 * i]) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 16
    for (int d = 0; d < D; ++d) {
      const float4 qa = *reinterpret_cast<const float4*>(&sm.Qt[d][ty * 4]);
      const float4 kb = *reinterpret_cast<const float4*>(&sm.Kt[d][tx * 4]);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }

    // mask + online softmax
    bool kvalid[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kj = k0 + tx * 4 + j;
      kvalid[j] = kj < Lk && (!km || km[kj] != 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx * 4 + j;
        const bool vis = kvalid[j] && (!a.causal || kj <= qi + a.q_pos_offset);
        s[i][j] = vis ? s[i][j] * a.scale : -INFINITY;
        tmax = fmaxf(tmax, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(kFull, tmax, off));
      const float m_new = fmaxf(m[i], tmax);
      const float alpha = (m_new == -INFINITY) ? 1.f : expf(m[i] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (s[i][j] == -INFINITY) ? 0.f : expf(s[i][j] - m_new);
        s[i][j] = p;
        psum += p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(kFull, psum, off);
      l[i] = l[i] * alpha + psum;
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&sm.Pt[tx * 4 + j][ty * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();

#pragma unroll 16
    for (int c = 0; c < BKV; ++c) {
      const float4 pa = *reinterpret_cast<const float4*>(&sm.Pt[c][ty * 4]);
      const float4 vb = *reinterpret_cast<const float4*>(&sm.Vs[c][tx * 4]);
      const float pv[4] = {pa.x, pa.y, pa.z, pa.w}, vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
    }
  }

  T* __restrict__ O = static_cast<T*>(a.o) + (a.q_offset ? qtok0 * a.o_ts : b * a.o_bs) + h * a.o_hs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= Lq) continue;
    const float inv = l[i] > 0.f ? 1.f / l[i] : 0.f;
    T* dst = O + static_cast<long long>(qi) * a.o_ts + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = from_f<T>(o[i][j] * inv);
  }
}


// =====================================================================================================
// bf16 tensor-core version (flash-attention-2 data flow): 64 queries per CTA (4 warps x 16 rows), K/V tiles of
// 64 keys double-buffered in shared memory by cp.async (16-byte chunks, XOR-swizzled so that the ldmatrix
// reads are bank-conflict free), S = Q.K^T and O += P.V by mma.sync.m16n8k16 with fp32 accumulation, online
// softmax in registers, probabilities rounded to bf16 for the second product (as the reference's bf16
// attention does).  Same AttnArgs contract as the SIMT kernel.
// =====================================================================================================
namespace tc {

constexpr int TBQ = 64, TBK = 64, TNT = 128;
constexpr int TILE_B = 64 * 128;   // one 64 x 64 bf16 tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 2^x in one MUFU instruction (relative error 2^-22; the probabilities are rounded to bf16 right after); 2^-inf = +0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of (row r, 16-byte chunk c) in a 64-row x 128-byte tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

// 64 rows x 64 bf16 from global (row stride ts elements) into a swizzled tile; rows >= n_rows are zero-filled
__device__ __forceinline__ void load_tile(uint32_t dst, const bf16* src, long long ts, int n_rows) {
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int idx = threadIdx.x + x * TNT;   // 512 chunks
    const int r = idx >> 3, c = idx & 7;
    const bool ok = r < n_rows;
    cp_async16(dst + swz(r, c), ok ? static_cast<const void*>(src + r * ts + c * 8) : static_cast<const void*>(src), ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(TNT) attention_mma_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain)
  pdl_wait();
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  const uint32_t sQ = smem_u32(sm), sK = sQ + TILE_B, sV = sK + 2 * TILE_B;   // K, V double-buffered
  const int tid = threadIdx.x, warp = tid / kWarp, lane = tid % kWarp;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, mr = lane & 7;
  const int q0 = blockIdx.x * TBQ, h = blockIdx.y, b = blockIdx.z;
  const int kvb = a.kv_batch_mod > 0 ? b % a.kv_batch_mod : b;
  const int Lk = a.Lk_per_batch ? a.Lk_per_batch[kvb] : a.Lk;
  const long long koff = a.kv_offset ? a.kv_offset[kvb] : 0;
  const int Lq = a.Lq_per_batch ? a.Lq_per_batch[b] : a.Lq;
  if (q0 >= Lq) return;   // packed queries: the grid is sized for the longest row (whole CTA leaves, after the PDL wait)
  const long long qtok0 = a.q_offset ? a.q_offset[b] : 0;
  const bf16* __restrict__ Q = static_cast<const bf16*>(a.q) + (a.q_offset ? qtok0 * a.q_ts : b * a.q_bs) + h * a.q_hs;
  const bf16* __restrict__ K = static_cast<const bf16*>(a.k) + kvb * a.k_bs + h * a.k_hs + koff * a.k_ts;
  const bf16* __restrict__ V = static_cast<const bf16*>(a.v) + kvb * a.v_bs + h * a.v_hs + koff * a.v_ts;
  const uint8_t* __restrict__ km =
      a.key_mask ? a.key_mask + static_cast<long long>(a.key_mask_per_q_batch ? b : kvb) * a.key_mask_ld : nullptr;

  int k_end = Lk;
  if (a.causal) k_end = min(k_end, min(q0 + TBQ, Lq) + a.q_pos_offset);   // keys beyond the tile's last query
  const int n_tiles = (k_end + TBK - 1) / TBK;

  load_tile(sQ, Q + static_cast<long long>(q0) * a.q_ts, a.q_ts, Lq - q0);
  if (n_tiles > 0) {
    load_tile(sK, K, a.k_ts, Lk);
    load_tile(sV, V, a.v_ts, Lk);
  }
  cp_commit();

  float o[8][4], m_[2] = {-INFINITY, -INFINITY}, l_[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[nt][j] = 0.f;
  uint32_t qf[4][4];
  const float sl2 = a.scale * 1.4426950408889634f;   // scores are kept in the log2 domain
  const int qi0 = q0 + warp * 16 + g;                // this lane's rows: qi0 and qi0 + 8

  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_tiles) {
      const int kb = (it + 1) * TBK;
      load_tile(sK + (buf ^ 1) * TILE_B, K + static_cast<long long>(kb) * a.k_ts, a.k_ts, Lk - kb);
      load_tile(sV + (buf ^ 1) * TILE_B, V + static_cast<long long>(kb) * a.v_ts, a.v_ts, Lk - kb);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(sQ + swz(warp * 16 + (mi & 1) * 8 + mr, ks * 2 + (mi >> 1)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    const uint32_t kt = sK + buf * TILE_B, vt = sV + buf * TILE_B;
    const int k0 = it * TBK;
    // ---- S = Q.K^T ----
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
      uint32_t bb[4];
      ldsm_x4(kt + swz(nt * 8 + mr, mi), bb[0], bb[1], bb[2], bb[3]);
      mma_bf16(s[nt], qf[0][0], qf[0][1], qf[0][2], qf[0][3], bb[0], bb[1]);
      mma_bf16(s[nt], qf[1][0], qf[1][1], qf[1][2], qf[1][3], bb[2], bb[3]);
      ldsm_x4(kt + swz(nt * 8 + mr, 4 + mi), bb[0], bb[1], bb[2], bb[3]);
      mma_bf16(s[nt], qf[2][0], qf[2][1], qf[2][2], qf[2][3], bb[0], bb[1]);
      mma_bf16(s[nt], qf[3][0], qf[3][1], qf[3][2], qf[3][3], bb[2], bb[3]);
    }
    // ---- mask + online softmax (rows g and g+8; s[nt][0..1] row g, s[nt][2..3] row g+8; cols nt*8 + 2t, +1) ----
    // Scores stay RAW; the scale and the running maximum enter through one FFMA per element:
    // p = 2^(s * sl2 - m), m = running maximum in the log2 domain.  (The softmax is 3/4 of this kernel's
    // instructions - ncu: tensor pipe 33 %, issue slots 61 % - so every per-element instruction counts.)
    const bool need_mask = km != nullptr || (k0 + TBK > Lk) || (a.causal && (k0 + TBK - 1 > q0 + warp * 16 + a.q_pos_offset));
    float mx[2] = {-INFINITY, -INFINITY};
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kj = k0 + nt * 8 + 2 * t + (j & 1);
          const int qi = qi0 + (j >> 1) * 8;
          const bool vis = kj < Lk && (!km || km[kj] != 0) && (!a.causal || kj <= qi + a.q_pos_offset);
          s[nt][j] = vis ? s[nt][j] : -INFINITY;
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) mx[j >> 1] = fmaxf(mx[j >> 1], s[nt][j]);
    float alpha[2], msub[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(kFull, mx[rr], 1));
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(kFull, mx[rr], 2));
      const float m_new = fmaxf(m_[rr], mx[rr] * sl2);             // sl2 > 0: the maximum commutes with the scale
      msub[rr] = (m_new == -INFINITY) ? 0.f : m_new;               // every key masked so far: 2^(-inf - 0) = 0
      alpha[rr] = (m_[rr] == -INFINITY) ? 0.f : ex2_approx(m_[rr] - msub[rr]);
      m_[rr] = m_new;
    }
    float ps[2] = {0.f, 0.f};
    uint32_t pa[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pv[j] = ex2_approx(fmaf(s[nt][j], sl2, -msub[j >> 1]));     // masked: fma(-inf) = -inf -> 0
        ps[j >> 1] += pv[j];
      }
      pa[nt][0] = pack_bf16(pv[0], pv[1]);
      pa[nt][1] = pack_bf16(pv[2], pv[3]);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      ps[rr] += __shfl_xor_sync(kFull, ps[rr], 1);
      ps[rr] += __shfl_xor_sync(kFull, ps[rr], 2);
      l_[rr] = l_[rr] * alpha[rr] + ps[rr];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
    }
    // ---- O += P.V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bb[4];
        ldsm_x4_t(vt + swz(kk * 16 + (mi & 1) * 8 + mr, 2 * dp + (mi >> 1)), bb[0], bb[1], bb[2], bb[3]);
        mma_bf16(o[2 * dp], pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1], bb[0], bb[1]);
        mma_bf16(o[2 * dp + 1], pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1], bb[2], bb[3]);
      }
    }
    __syncthreads();   // this buffer is overwritten by the load issued in the next iteration
  }
  if (n_tiles == 0) cp_wait<0>();

  bf16* __restrict__ O = static_cast<bf16*>(a.o) + (a.q_offset ? qtok0 * a.o_ts : b * a.o_bs) + h * a.o_hs;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int qi = qi0 + rr * 8;
    if (qi >= Lq) continue;
    const float inv = l_[rr] > 0.f ? 1.f / l_[rr] : 0.f;
    bf16* dst = O + static_cast<long long>(qi) * a.o_ts + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][2 * rr] * inv, o[nt][2 * rr + 1] * inv);
  }
}

}  // namespace tc

}  // namespace

template <typename T>
void attention_simt(const AttnArgs& a, cudaStream_t stream) {
  if (a.batch <= 0 || a.Lq <= 0) return;
  static bool configured = false;
  if (!configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(attention_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(Smem))));
    configured = true;
  }
  dim3 grid(ceil_div(a.Lq, BQ), a.heads, a.batch);
  CXRM_CHECK(grid.z <= 65535 && grid.y <= 65535, "attention batch too large for grid.z");
  attention_simt_kernel<T><<<grid, NT, sizeof(Smem), stream>>>(a);
  check_launch("attention_simt");
}

int attention_mma_supported(const AttnArgs& a) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const long long st[] = {a.q_bs, a.q_hs, a.q_ts, a.k_bs, a.k_hs, a.k_ts, a.v_bs, a.v_hs, a.v_ts};
  for (long long x : st)
    if (x % 8 != 0) return 1;                       // 16-byte cp.async chunks
  if (a.o_bs % 2 != 0 || a.o_hs % 2 != 0 || a.o_ts % 2 != 0) return 2;
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || reinterpret_cast<uintptr_t>(a.o) % 4 != 0) return 3;
  if (a.kv_offset && a.k_ts % 8 != 0) return 4;
  return 0;
}

void attention_mma(const AttnArgs& a, cudaStream_t stream) {
  if (a.batch <= 0 || a.Lq <= 0) return;
  CXRM_CHECK(attention_mma_supported(a) == 0, "strides/alignment not supported by the tensor-core attention");
  constexpr int smem = 5 * tc::TILE_B + 128;
  static bool configured = false;
  if (!configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(tc::attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(a.Lq, tc::TBQ), a.heads, a.batch);
  CXRM_CHECK(grid.z <= 65535 && grid.y <= 65535, "attention batch too large for grid.z");
  launch_chain(tc::attention_mma_kernel, grid, dim3(tc::TNT), smem, stream, a);
  check_launch("attention_mma");
}

template void attention_simt<float>(const AttnArgs&, cudaStream_t);
template void attention_simt<bf16>(const AttnArgs&, cudaStream_t);

}  // namespace cxrm
