// Backward-pass kernels of the teacher-forced decoder step:
// the SCST training step recomputes the sampled rollout's log-probs in ONE teacher-forced pass and backpropagates
// the REINFORCE loss (reference scst/gen_prompt.py:331-366), instead of keeping the autograd graph of 255 cached decode
// steps alive as the reference does; the same pass with a cross-entropy head is the teacher-forced training step
// (reference longitudinal/gt_prompt.py:186-249).  Eval-mode arithmetic (no dropout), as parity is defined.
//
// Everything dense goes through the engine's GEMM dispatcher (tcgen05 in bf16 mode, strict fp32 FMA in validation
// mode): dX = dY.W and dW = dY^T.X are brought into the dispatcher's K-major form by explicit transposes.  This file
// holds what is not a GEMM: transposes, column sums (bias gradients), LayerNorm / GELU backward, the fused loss +
// dlogits head (cross-entropy, or REINFORCE over the top-k-masked scores), attention backward, embedding scatter.
#include "kernels.h"

namespace cxrm {

namespace {

constexpr int HD = 64;

// ---- transpose ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, long long ld_in, T* __restrict__ out, long long ld_out,
                                 long long rows, int cols) {
  __shared__ T tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.y) * 32;
  const int c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * ld_in + c] : from_f<T>(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<long long>(c) * ld_out + r] = tile[threadIdx.x][i];
  }
}

// ---- column sums: out[n] (+)= sum_m x[m, n] -------------------------------------------------------------------
// one block per 32 columns, fixed summation order (reproducible); optional second operand: sum_m x[m,n] * y[m,n]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, long long rows, int cols,
                                                     float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x % 32, lane_r = threadIdx.x / 32;
  float s = 0.f;
  if (c < cols)
    for (long long r = lane_r; r < rows; r += 8) s += to_f(x[r * ldx + c]);
  sh[lane_r][threadIdx.x % 32] = s;
  __syncthreads();
  if (lane_r == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ---- LayerNorm backward ------------------------------------------------------------------------------------------
// y = (x - mean) * rstd * gamma + beta.  One warp per row: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// Also writes (mean, rstd) per row for the parameter-gradient pass.
template <typename T, int C>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                     const float* __restrict__ gamma, float eps, T* __restrict__ dx,
                                                     float2* __restrict__ stats, long long rows) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (row >= rows) return;
  constexpr int PER = C / kWarp;
  float xv[PER], gv[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + i * kWarp;
    xv[i] = to_f(x[row * C + c]);
    s += xv[i];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) q += (xv[i] - mean) * (xv[i] - mean);
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + i * kWarp;
    xv[i] = (xv[i] - mean) * rstd;
    gv[i] = to_f(dy[row * C + c]) * gamma[c];
    sg += gv[i];
    sgx += gv[i] * xv[i];
  }
  sg = warp_sum(sg) / C;
  sgx = warp_sum(sgx) / C;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + i * kWarp;
    dx[row * C + c] = from_f<T>(rstd * (gv[i] - sg - xv[i] * sgx));
  }
  if (lane == 0 && stats) stats[row] = make_float2(mean, rstd);
}
// dgamma[c] (+)= sum_m dy * xhat, dbeta[c] (+)= sum_m dy
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_params_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                            const float2* __restrict__ stats, long long rows, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int accumulate) {
  __shared__ float sg[8][33], sb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x % 32, lane_r = threadIdx.x / 32;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (long long r = lane_r; r < rows; r += 8) {
      const float2 st = stats[r];
      const float d = to_f(dy[r * C + c]);
      a += d * (to_f(x[r * C + c]) - st.x) * st.y;
      b += d;
    }
  sg[lane_r][threadIdx.x % 32] = a;
  sb[lane_r][threadIdx.x % 32] = b;
  __syncthreads();
  if (lane_r == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ta += sg[i][threadIdx.x];
      tb += sb[i][threadIdx.x];
    }
    dgamma[c] = accumulate ? dgamma[c] + ta : ta;
    dbeta[c] = accumulate ? dbeta[c] + tb : tb;
  }
}

// ---- GELU (exact erf) forward / backward ---------------------------------------------------------------------------
template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = from_f<T>(gelu_erf(to_f(x[i])));
}
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = to_f(x[i]);
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  dx[i] = from_f<T>(to_f(dy[i]) * (cdf + v * pdf));
}
template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ a, const T* __restrict__ b, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) a[i] = from_f<T>(to_f(a[i]) + to_f(b[i]));
}
__global__ void scale_kernel(float* __restrict__ a, float s, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) a[i] *= s;
}

// ---- loss head ------------------------------------------------------------------------------------------------------
// One block per token row of fp32 logits [rows, V].
//   kind 0 (cross-entropy, ignore_index):  loss_row = -log_softmax(z)[target];  dz = (softmax(z) - onehot) * w,  w = 1 / n_counted
//   kind 1 (REINFORCE over the top-k-masked scores, scst/gen_prompt.py:350-364): s = z / temperature, entries below the
//          k-th largest are -inf (ties kept, HF TopKLogitsWarper); lp = log_softmax(s)[target];
//          loss_row = -lp * adv[row / L] / R;  dz = -(adv / R / temperature) * (onehot - softmax(s)) on the survivors, 0 elsewhere.
// dz is written in T; loss_row goes to row_loss[row] (summed afterwards in a fixed order).
constexpr int LNT = 512;
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

template <typename T>
__global__ void __launch_bounds__(LNT) loss_head_kernel(const float* __restrict__ logits, int V, const int* __restrict__ targets,
                                                        int ignore_index, int kind, const float* __restrict__ adv, int L,
                                                        float w_ce, float w_rl, int top_k, float temperature,
                                                        T* __restrict__ dz, float* __restrict__ row_loss) {
  extern __shared__ unsigned keys[];   // [V] sortable keys of the (temperature-scaled) row
  __shared__ float redf[LNT / 32];
  __shared__ int hist[256];
  __shared__ unsigned sh_prefix;
  __shared__ int sh_kk;
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* z = logits + row * V;
  T* d = dz + row * V;
  const int tgt = targets[row];
  if (tgt == ignore_index) {
    for (int i = tid; i < V; i += LNT) d[i] = from_f<T>(0.f);
    if (tid == 0) row_loss[row] = 0.f;
    return;
  }
  const float inv_t = (kind == 1 && temperature != 1.0f) ? 1.0f / temperature : 1.0f;
  float mx = -INFINITY;
  for (int i = tid; i < V; i += LNT) {
    float s = z[i];
    if (inv_t != 1.0f) s = s / temperature;
    keys[i] = f2key(s);
    mx = fmaxf(mx, s);
  }
  auto block_max = [&](float x) {
    x = warp_max(x);
    __syncthreads();
    if (tid % 32 == 0) redf[tid / 32] = x;
    __syncthreads();
    float t = -INFINITY;
#pragma unroll
    for (int i = 0; i < LNT / 32; ++i) t = fmaxf(t, redf[i]);
    return t;
  };
  auto block_sum = [&](float x) {
    x = warp_sum(x);
    __syncthreads();
    if (tid % 32 == 0) redf[tid / 32] = x;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < LNT / 32; ++i) t += redf[i];
    return t;
  };
  mx = block_max(mx);
  unsigned thr = 0;   // everything survives
  if (kind == 1 && top_k > 0 && top_k < V) {
    // k-th largest key: 4 x 8-bit radix select, MSB first
    if (tid == 0) {
      sh_prefix = 0;
      sh_kk = top_k;
    }
    unsigned mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned prefix = sh_prefix;
      for (int i = tid; i < V; i += LNT)
        if ((keys[i] & mask) == prefix) atomicAdd(&hist[(keys[i] >> shift) & 255], 1);
      __syncthreads();
      if (tid == 0) {
        int kk = sh_kk, cum = 0;
        for (int bin = 255; bin >= 0; --bin) {
          if (cum + hist[bin] >= kk) {
            sh_prefix = prefix | (static_cast<unsigned>(bin) << shift);
            sh_kk = kk - cum;
            break;
          }
          cum += hist[bin];
        }
      }
      mask |= 255u << shift;
      __syncthreads();
    }
    thr = sh_prefix;
  }
  float es = 0.f;
  for (int i = tid; i < V; i += LNT)
    if (keys[i] >= thr) es += expf(key2f(keys[i]) - mx);
  es = block_sum(es);
  const float lse = mx + logf(es);
  const float st = key2f(keys[tgt]);
  const bool tgt_alive = keys[tgt] >= thr;   // a REINFORCE target is a sampled id: always among the survivors
  float coef;
  if (kind == 0) {
    coef = w_ce;
    if (tid == 0) row_loss[row] = -(st - lse) * w_ce;
  } else {
    const float a = adv[row / L] * w_rl;
    coef = a * inv_t;                    // d(-a * lp)/dz = a * inv_t * (softmax - onehot)
    if (tid == 0) row_loss[row] = tgt_alive ? -(st - lse) * a : 0.f;
  }
  for (int i = tid; i < V; i += LNT) {
    float g = 0.f;
    if (keys[i] >= thr) g = coef * expf(key2f(keys[i]) - lse);
    if (i == tgt && tgt_alive) g -= coef;
    d[i] = from_f<T>(g);
  }
}
__global__ void sum_rows_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[256];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}
__global__ void count_targets_kernel(const int* __restrict__ t, long long n, int ignore_index, int* __restrict__ out) {
  __shared__ int sh[256];
  int s = 0;
  for (long long i = threadIdx.x; i < n; i += 256) s += t[i] != ignore_index;
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ---- attention backward ----------------------------------------------------------------------------------------------
// Two kernels, both with a fixed summation order (no atomics):
//   A  one thread pair per QUERY (32 head dims each): pass 1 recomputes the row's log-sum-exp and D = <dO, O>, pass 2
//      forms p = exp(s - lse), ds = p (<dO, v> - D) and accumulates dQ; lse / D are stored for B.
//   B  one thread pair per KEY: loops over the queries that see the key: dV += p dO, dK += ds q * scale.
// Visibility as in the forward kernels: key j of kv-batch b' visible to query i iff j < Lk[b'] (ragged caches),
// key_mask[.., j] != 0, and j <= i + q_pos_offset when causal.
struct AttnBwdArgs {
  AttnArgs f;              // forward description (q, k, v, o = forward OUTPUT, strides, masks)
  const void* dO;          // like o
  void* dQ;                // like q (strides of q)
  void* dK; void* dV;      // nullptr: skip kernel B.  Element (kv batch b', head h, key j): dK[b' * g_bs + h * g_hs + (koff + j) * g_ts]
  long long g_bs, g_hs, g_ts;
  float* lse; float* D;    // [batch, heads, Lq]
};

template <typename T>
__device__ __forceinline__ void load32(const T* p, float* f) {
#pragma unroll
  for (int i = 0; i < 32; i += Vec16<T>::N) {
    Vec16<T> v;
    v.load(p + i);
    v.unpack(f + i);
  }
}
template <typename T>
__device__ __forceinline__ void store32(T* p, const float* f) {
#pragma unroll
  for (int i = 0; i < 32; i += Vec16<T>::N) {
    Vec16<T> v;
    v.pack(f + i);
    v.store(p + i);
  }
}

constexpr int ABQ = 64, ABK = 64, ABT = 128;   // queries per block (pairs of threads), keys per staged tile

template <typename T>
__global__ void __launch_bounds__(ABT) attn_bwd_q_kernel(AttnBwdArgs a) {
  __shared__ float Ks[ABK][HD + 1], Vs[ABK][HD + 1];
  __shared__ uint8_t vis[ABK];
  const AttnArgs& f = a.f;
  const int tid = threadIdx.x, qi = tid >> 1, half = tid & 1;
  const int q0 = blockIdx.x * ABQ, h = blockIdx.y, b = blockIdx.z;
  const int kvb = f.kv_batch_mod > 0 ? b % f.kv_batch_mod : b;
  const int Lk = f.Lk_per_batch ? f.Lk_per_batch[kvb] : f.Lk;
  const long long koff = f.kv_offset ? f.kv_offset[kvb] : 0;
  const T* K = static_cast<const T*>(f.k) + kvb * f.k_bs + h * f.k_hs + koff * f.k_ts;
  const T* V = static_cast<const T*>(f.v) + kvb * f.v_bs + h * f.v_hs + koff * f.v_ts;
  const uint8_t* km = f.key_mask ? f.key_mask + static_cast<long long>(f.key_mask_per_q_batch ? b : kvb) * f.key_mask_ld : nullptr;
  const int i = q0 + qi;
  const bool q_ok = i < f.Lq;
  float q[32], dO[32], dQ[32];
  float Dv = 0.f;
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = dO[d] = dQ[d] = 0.f;
  if (q_ok) {
    load32<T>(static_cast<const T*>(f.q) + b * f.q_bs + h * f.q_hs + static_cast<long long>(i) * f.q_ts + half * 32, q);
    load32<T>(static_cast<const T*>(a.dO) + b * f.o_bs + h * f.o_hs + static_cast<long long>(i) * f.o_ts + half * 32, dO);
    float o[32];
    load32<T>(static_cast<const T*>(f.o) + b * f.o_bs + h * f.o_hs + static_cast<long long>(i) * f.o_ts + half * 32, o);
#pragma unroll
    for (int d = 0; d < 32; ++d) Dv += dO[d] * o[d];
  }
  Dv += __shfl_xor_sync(kFull, Dv, 1);
  int k_end = Lk;
  if (f.causal) k_end = min(k_end, min(q0 + ABQ, f.Lq) + f.q_pos_offset);
  const int my_end = f.causal ? min(Lk, i + f.q_pos_offset + 1) : Lk;
  float m = -INFINITY, l = 0.f, lse = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    for (int k0 = 0; k0 < k_end; k0 += ABK) {
      __syncthreads();
      for (int x = tid; x < ABK * (HD / 8); x += ABT) {      // 8 dims per thread-iteration
        const int kr = x / (HD / 8), d0 = (x % (HD / 8)) * 8;
        float kf[8], vf[8];
        if (k0 + kr < Lk) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            kf[j] = to_f(K[static_cast<long long>(k0 + kr) * f.k_ts + d0 + j]);
            vf[j] = to_f(V[static_cast<long long>(k0 + kr) * f.v_ts + d0 + j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) kf[j] = vf[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          Ks[kr][d0 + j] = kf[j];
          Vs[kr][d0 + j] = vf[j];
        }
      }
      for (int x = tid; x < ABK; x += ABT) vis[x] = (k0 + x < Lk) && (!km || km[k0 + x]);
      __syncthreads();
      const int kn = min(ABK, k_end - k0);
      for (int kr = 0; kr < kn; ++kr) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) s += q[d] * Ks[kr][half * 32 + d];
        s += __shfl_xor_sync(kFull, s, 1);
        s *= f.scale;
        const bool v_ok = q_ok && vis[kr] && (k0 + kr) < my_end;
        if (pass == 0) {
          if (v_ok) {
            const float mn = fmaxf(m, s);
            l = l * expf(m - mn) + expf(s - mn);
            m = mn;
          }
        } else {
          float dp = 0.f;
#pragma unroll
          for (int d = 0; d < 32; ++d) dp += dO[d] * Vs[kr][half * 32 + d];
          dp += __shfl_xor_sync(kFull, dp, 1);
          if (v_ok) {
            const float p = expf(s - lse);
            const float ds = p * (dp - Dv) * f.scale;
#pragma unroll
            for (int d = 0; d < 32; ++d) dQ[d] += ds * Ks[kr][half * 32 + d];
          }
        }
      }
    }
    if (pass == 0) lse = (l > 0.f) ? m + logf(l) : 0.f;
  }
  if (q_ok) {
    store32<T>(static_cast<T*>(a.dQ) + b * f.q_bs + h * f.q_hs + static_cast<long long>(i) * f.q_ts + half * 32, dQ);
    if (half == 0) {
      const long long sidx = (static_cast<long long>(b) * f.heads + h) * f.Lq + i;
      a.lse[sidx] = lse;
      a.D[sidx] = Dv;
    }
  }
}

// grid: (key blocks, heads, kv batches).  q batches that read kv batch b': b' , b' + mod, ... (kv_batch_mod) or b' itself.
template <typename T>
__global__ void __launch_bounds__(ABT) attn_bwd_kv_kernel(AttnBwdArgs a) {
  __shared__ float Qs[ABQ][HD + 1], dOs[ABQ][HD + 1];
  __shared__ float sl[ABQ], sD[ABQ];
  const AttnArgs& f = a.f;
  const int tid = threadIdx.x, ki = tid >> 1, half = tid & 1;
  const int k0 = blockIdx.x * ABK, h = blockIdx.y, kvb = blockIdx.z;
  const int Lk = f.Lk_per_batch ? f.Lk_per_batch[kvb] : f.Lk;
  if (k0 >= Lk) return;
  const long long koff = f.kv_offset ? f.kv_offset[kvb] : 0;
  const int j = k0 + ki;
  const bool k_ok = j < Lk;
  float kx[32], vx[32], dK[32], dV[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) kx[d] = vx[d] = dK[d] = dV[d] = 0.f;
  if (k_ok) {
    load32<T>(static_cast<const T*>(f.k) + kvb * f.k_bs + h * f.k_hs + (koff + j) * f.k_ts + half * 32, kx);
    load32<T>(static_cast<const T*>(f.v) + kvb * f.v_bs + h * f.v_hs + (koff + j) * f.v_ts + half * 32, vx);
  }
  const int n_qb = f.kv_batch_mod > 0 ? f.batch / f.kv_batch_mod : 1;
  for (int rep = 0; rep < n_qb; ++rep) {
    const int b = f.kv_batch_mod > 0 ? kvb + rep * f.kv_batch_mod : kvb;
    const uint8_t* km = f.key_mask ? f.key_mask + static_cast<long long>(f.key_mask_per_q_batch ? b : kvb) * f.key_mask_ld : nullptr;
    const bool key_vis = k_ok && (!km || km[j]);
    // causal: queries i >= j - q_pos_offset see key j; the first query block that can see any key of this tile
    const int q_begin = f.causal ? max(0, k0 - f.q_pos_offset) / ABQ * ABQ : 0;
    for (int q0 = q_begin; q0 < f.Lq; q0 += ABQ) {
      __syncthreads();
      for (int x = tid; x < ABQ * (HD / 8); x += ABT) {
        const int qr = x / (HD / 8), d0 = (x % (HD / 8)) * 8;
        const bool ok = q0 + qr < f.Lq;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          Qs[qr][d0 + jj] = ok ? to_f(static_cast<const T*>(f.q)[b * f.q_bs + h * f.q_hs + static_cast<long long>(q0 + qr) * f.q_ts + d0 + jj]) : 0.f;
          dOs[qr][d0 + jj] = ok ? to_f(static_cast<const T*>(a.dO)[b * f.o_bs + h * f.o_hs + static_cast<long long>(q0 + qr) * f.o_ts + d0 + jj]) : 0.f;
        }
      }
      for (int x = tid; x < ABQ; x += ABT) {
        const bool ok = q0 + x < f.Lq;
        const long long sidx = (static_cast<long long>(b) * f.heads + h) * f.Lq + q0 + x;
        sl[x] = ok ? a.lse[sidx] : 0.f;
        sD[x] = ok ? a.D[sidx] : 0.f;
      }
      __syncthreads();
      const int qn = min(ABQ, f.Lq - q0);
      for (int qr = 0; qr < qn; ++qr) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          s += Qs[qr][half * 32 + d] * kx[d];
          dp += dOs[qr][half * 32 + d] * vx[d];
        }
        s += __shfl_xor_sync(kFull, s, 1);
        dp += __shfl_xor_sync(kFull, dp, 1);
        const bool vis = key_vis && (!f.causal || j <= q0 + qr + f.q_pos_offset);
        if (vis) {
          const float p = expf(s * f.scale - sl[qr]);
          const float ds = p * (dp - sD[qr]) * f.scale;
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            dV[d] += p * dOs[qr][half * 32 + d];
            dK[d] += ds * Qs[qr][half * 32 + d];
          }
        }
      }
    }
  }
  if (k_ok) {
    store32<T>(static_cast<T*>(a.dK) + kvb * a.g_bs + h * a.g_hs + (koff + j) * a.g_ts + half * 32, dK);
    store32<T>(static_cast<T*>(a.dV) + kvb * a.g_bs + h * a.g_hs + (koff + j) * a.g_ts + half * 32, dV);
  }
}

// ---- embedding backward: dtable[idx[row], :] += dx[row, :] (fp32 atomics; rows of one id collide) --------------------
template <typename T>
__global__ void scatter_add_rows_kernel(const T* __restrict__ dx, const int* __restrict__ idx, float* __restrict__ table,
                                        long long rows, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  atomicAdd(&table[static_cast<long long>(idx[r]) * C + c], to_f(dx[i]));
}

// (word[id] + type[tt]) + pos[p], the pre-LayerNorm embedding sum (kept for the LayerNorm backward)
template <typename T>
__global__ void embed_sum_kernel(const int* __restrict__ ids, const int* __restrict__ types, const int* __restrict__ pos,
                                 const T* __restrict__ word, const T* __restrict__ type_emb, const T* __restrict__ pos_emb,
                                 T* __restrict__ out, long long rows, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  const float e = (to_f(word[static_cast<long long>(ids[r]) * C + c]) + to_f(type_emb[static_cast<long long>(types[r]) * C + c])) +
                  to_f(pos_emb[static_cast<long long>(pos[r]) * C + c]);
  out[i] = from_f<T>(e);
}

inline unsigned grid1d(long long n, int block) { return static_cast<unsigned>(ceil_div_ll(n, block)); }

}  // namespace

template <typename T>
void transpose(const T* in, long long ld_in, T* out, long long ld_out, long long rows, int cols, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return;
  dim3 grid(ceil_div(cols, 32), static_cast<unsigned>(ceil_div_ll(rows, 32)));
  CXRM_CHECK(grid.y <= 65535u * 1024u, "transpose rows");
  transpose_kernel<T><<<grid, dim3(32, 8), 0, stream>>>(in, ld_in, out, ld_out, rows, cols);
  check_launch("transpose");
}
template <typename T>
void colsum(const T* x, long long ldx, long long rows, int cols, float* out, bool accumulate, cudaStream_t stream) {
  colsum_kernel<T><<<ceil_div(cols, 32), 256, 0, stream>>>(x, ldx, rows, cols, out, accumulate ? 1 : 0);
  check_launch("colsum");
}
template <typename T>
void layernorm_bwd(const T* x, const T* dy, const float* gamma, float eps, T* dx, float2* stats, float* dgamma, float* dbeta,
                   bool accumulate, long long rows, int C, cudaStream_t stream) {
  CXRM_CHECK(C == 768 || C == 128, "layernorm_bwd: C must be 768 or 128");
  CXRM_CHECK(stats || !dgamma, "layernorm_bwd: parameter gradients need the statistics buffer");
  if (C == 768)
    ln_bwd_kernel<T, 768><<<grid1d(rows * kWarp, 256), 256, 0, stream>>>(x, dy, gamma, eps, dx, stats, rows);
  else
    ln_bwd_kernel<T, 128><<<grid1d(rows * kWarp, 256), 256, 0, stream>>>(x, dy, gamma, eps, dx, stats, rows);
  check_launch("ln_bwd");
  if (dgamma) {
    ln_bwd_params_kernel<T><<<ceil_div(C, 32), 256, 0, stream>>>(x, dy, stats, rows, C, dgamma, dbeta, accumulate ? 1 : 0);
    check_launch("ln_bwd_params");
  }
}
template <typename T>
void gelu_fwd(const T* x, T* y, long long n, cudaStream_t stream) {
  gelu_fwd_kernel<T><<<grid1d(n, 256), 256, 0, stream>>>(x, y, n);
  check_launch("gelu_fwd");
}
template <typename T>
void gelu_bwd(const T* x, const T* dy, T* dx, long long n, cudaStream_t stream) {
  gelu_bwd_kernel<T><<<grid1d(n, 256), 256, 0, stream>>>(x, dy, dx, n);
  check_launch("gelu_bwd");
}
template <typename T>
void add_inplace(T* a, const T* b, long long n, cudaStream_t stream) {
  add_inplace_kernel<T><<<grid1d(n, 256), 256, 0, stream>>>(a, b, n);
  check_launch("add_inplace");
}
void scale_f32(float* a, float s, long long n, cudaStream_t stream) {
  scale_kernel<<<grid1d(n, 256), 256, 0, stream>>>(a, s, n);
  check_launch("scale");
}
template <typename T>
void loss_head(const float* logits, long long rows, int V, const int* targets, int ignore_index, int kind, const float* adv,
               int L, int R, int top_k, float temperature, T* dz, float* row_loss, int* n_counted, float* loss_out,
               cudaStream_t stream) {
  CXRM_CHECK(static_cast<size_t>(V) * 4 <= 200 * 1024, "loss_head: vocabulary too large for shared memory");
  static size_t configured = 0;
  const size_t smem = static_cast<size_t>(V) * sizeof(unsigned);
  if (smem > configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(loss_head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  float w_ce = 1.0f;
  if (kind == 0) {   // CrossEntropyLoss(ignore_index): mean over the counted targets (one small read-back)
    count_targets_kernel<<<1, 256, 0, stream>>>(targets, rows, ignore_index, n_counted);
    check_launch("count_targets");
    int n = 0;
    CXRM_CUDA_CHECK(cudaMemcpyAsync(&n, n_counted, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CXRM_CUDA_CHECK(cudaStreamSynchronize(stream));
    w_ce = n > 0 ? 1.0f / n : 0.f;
  }
  loss_head_kernel<T><<<static_cast<unsigned>(rows), LNT, smem, stream>>>(logits, V, targets, ignore_index, kind, adv, L, w_ce,
                                                                           1.0f / R, top_k, temperature, dz, row_loss);
  check_launch("loss_head");
  sum_rows_kernel<<<1, 256, 0, stream>>>(row_loss, rows, loss_out);
  check_launch("sum_rows");
}
template <typename T>
void attention_bwd(const AttnArgs& f, const void* dO, void* dQ, void* dK, void* dV, long long g_bs, long long g_hs,
                   long long g_ts, float* lse, float* D, cudaStream_t stream) {
  AttnBwdArgs a;
  a.f = f; a.dO = dO; a.dQ = dQ; a.dK = dK; a.dV = dV; a.lse = lse; a.D = D;
  a.g_bs = g_bs; a.g_hs = g_hs; a.g_ts = g_ts;
  attn_bwd_q_kernel<T><<<dim3(ceil_div(f.Lq, ABQ), f.heads, f.batch), ABT, 0, stream>>>(a);
  check_launch("attn_bwd_q");
  if (dK) {
    const int kvb = f.kv_batch_mod > 0 ? f.kv_batch_mod : f.batch;
    attn_bwd_kv_kernel<T><<<dim3(ceil_div(f.Lk, ABK), f.heads, kvb), ABT, 0, stream>>>(a);
    check_launch("attn_bwd_kv");
  }
}
template <typename T>
void scatter_add_rows(const T* dx, const int* idx, float* table, long long rows, int C, cudaStream_t stream) {
  scatter_add_rows_kernel<T><<<grid1d(rows * C, 256), 256, 0, stream>>>(dx, idx, table, rows, C);
  check_launch("scatter_add_rows");
}

template <typename T>
void embed_sum(const int* ids, const int* types, const int* pos, const T* word, const T* type_emb, const T* pos_emb, T* out,
               long long rows, int C, cudaStream_t stream) {
  embed_sum_kernel<T><<<grid1d(rows * C, 256), 256, 0, stream>>>(ids, types, pos, word, type_emb, pos_emb, out, rows, C);
  check_launch("embed_sum");
}

#define INST(T)                                                                                                          \
  template void transpose<T>(const T*, long long, T*, long long, long long, int, cudaStream_t);                         \
  template void colsum<T>(const T*, long long, long long, int, float*, bool, cudaStream_t);                             \
  template void layernorm_bwd<T>(const T*, const T*, const float*, float, T*, float2*, float*, float*, bool, long long, \
                                 int, cudaStream_t);                                                                     \
  template void gelu_fwd<T>(const T*, T*, long long, cudaStream_t);                                                      \
  template void gelu_bwd<T>(const T*, const T*, T*, long long, cudaStream_t);                                           \
  template void add_inplace<T>(T*, const T*, long long, cudaStream_t);                                                  \
  template void loss_head<T>(const float*, long long, int, const int*, int, int, const float*, int, int, int, float, T*, \
                             float*, int*, float*, cudaStream_t);                                                        \
  template void attention_bwd<T>(const AttnArgs&, const void*, void*, void*, void*, long long, long long, long long,    \
                                 float*, float*, cudaStream_t);                                                         \
  template void scatter_add_rows<T>(const T*, const int*, float*, long long, int, cudaStream_t);                           \
  template void embed_sum<T>(const int*, const int*, const int*, const T*, const T*, const T*, T*, long long, int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace cxrm
