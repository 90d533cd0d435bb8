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
  const T* __restrict__ Q = static_cast<const T*>(a.q) + b * a.q_bs + h * a.q_hs;
  const T* __restrict__ K = static_cast<const T*>(a.k) + kvb * a.k_bs + h * a.k_hs + koff * a.k_ts;
  const T* __restrict__ V = static_cast<const T*>(a.v) + kvb * a.v_bs + h * a.v_hs + koff * a.v_ts;
  const uint8_t* __restrict__ km =
      a.key_mask ? a.key_mask + static_cast<long long>(a.key_mask_per_q_batch ? b : kvb) * a.key_mask_ld : nullptr;

  // loader mapping: row = tid/4 (0..63), 16 consecutive d per thread
  const int lrow = tid >> 2, ld0 = (tid & 3) * 16;
  {
    float f[16];
    if (q0 + lrow < a.Lq) {
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
    const int last_q = min(q0 + BQ, a.Lq) - 1 + a.q_pos_offset;
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

  T* __restrict__ O = static_cast<T*>(a.o) + b * a.o_bs + h * a.o_hs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= a.Lq) continue;
    const float inv = l[i] > 0.f ? 1.f / l[i] : 0.f;
    T* dst = O + static_cast<long long>(qi) * a.o_ts + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = from_f<T>(o[i][j] * inv);
  }
}

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

template void attention_simt<float>(const AttnArgs&, cudaStream_t);
template void attention_simt<bf16>(const AttnArgs&, cudaStream_t);

}  // namespace cxrm
