// Decode-step kernels of the KV-cached rollout: rollout state set-up from the
// prompt, one-token self- and cross-attention over the caches (HBM-bound,
// 128-bit loads, warp shuffles), the fused greedy / top-k multinomial head with
// log-prob gather and per-row state update, and small helpers.
//
// Reference semantics restated here:
//   modelling_longitudinal.py:274-283  mask = ids != mask_token_id, pos = relu(cumsum(mask)-1)
//   modelling_longitudinal.py:297-364  token types (full rule at prefill, `_past` rule per step)
//   HF generation/logits_process.py:577-587  top-k: scores < kth_largest -> -inf (ties kept)
//   HF generation/utils.py:2788-2805   softmax -> multinomial(1) == argmax(p/q), q~Exp(1); finished rows emit PAD
//   scst/gen_prompt.py:350-355         log-prob of the sampled id under log_softmax(top-k-masked scores), PAD ignored
#include <curand_kernel.h>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int H = 768, HD = 64, NH = 12;

// =============================================================================
// rollout_init
// =============================================================================
__global__ void rollout_init_kernel(RolloutState st, RolloutParams p, const int* __restrict__ prompt, int* pre_ids,
                                    int* pre_types, int* pre_pos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) {
    *st.step = 0;
    *st.done = 0;
    *st.arrive = 0;
  }
  if (r >= p.R) return;
  const int blk = r / p.B, study = r % p.B;
  const int* row = prompt + static_cast<long long>(study) * p.P;
  const int ns = p.n_special[blk];
  // first occurrence (argmax of equality; 0 when absent) of every special token
  int cols[kMaxSpecial];
  bool ok[kMaxSpecial];
  unsigned seen = 0;
  for (int i = 0; i < ns; ++i) {
    int first = 0;
    bool found = false;
    for (int c = 0; c < p.P; ++c) {
      if (row[c] == p.special_ids[blk][i]) {
        if (!found) first = c;
        found = true;
      }
    }
    if (found) seen |= 1u << i;
    cols[i] = first + 1;
    ok[i] = (cols[i] != 1) && (cols[i] < p.P);
  }
  int cum = 0;
  for (int c = 0; c < p.P; ++c) {
    const int id = row[c];
    const bool valid = p.mask_token_id < 0 || id != p.mask_token_id;
    cum += valid ? 1 : 0;
    int tt = p.sections[blk][0];
    for (int i = 0; i < ns; ++i)
      if (ok[i] && c >= cols[i]) tt = p.sections[blk][i + 1];
    st.seq[static_cast<long long>(r) * p.Lmax + c] = id;
    st.key_valid[static_cast<long long>(r) * p.Lmax + c] = valid ? 1 : 0;
    pre_ids[static_cast<long long>(r) * p.P + c] = id;
    pre_types[static_cast<long long>(r) * p.P + c] = tt;
    pre_pos[static_cast<long long>(r) * p.P + c] = max(cum - 1, 0);
  }
  for (int c = p.P; c < p.Lmax; ++c) {
    st.seq[static_cast<long long>(r) * p.Lmax + c] = p.pad;
    st.key_valid[static_cast<long long>(r) * p.Lmax + c] = 0;
  }
  st.n_valid[r] = cum;
  st.seen[r] = seen;
  st.cur_len[r] = p.P;
  st.cur_token[r] = p.pad;
  st.cur_type[r] = 0;
  st.cur_pos[r] = 0;
  st.finished[r] = 0;
}

// =============================================================================
// sampling head
// =============================================================================
constexpr int SNT = 512;

__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct ValIdx {
  float v;
  int i;
};
// larger value wins; ties -> lower index (torch.argmax returns the first maximal index)
__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ValIdx block_argmax(ValIdx x, ValIdx* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(kFull, x.v, o);
    y.i = __shfl_xor_sync(kFull, x.i, o);
    x = better(x, y);
  }
  const int w = threadIdx.x / kWarp, l = threadIdx.x % kWarp;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    x = (l < SNT / kWarp) ? sh[l] : ValIdx{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ValIdx y;
      y.v = __shfl_xor_sync(kFull, x.v, o);
      y.i = __shfl_xor_sync(kFull, x.i, o);
      x = better(x, y);
    }
    if (l == 0) sh[0] = x;
  }
  __syncthreads();
  x = sh[0];
  return x;
}
__device__ float block_sum(float x, float* sh) {
  x = warp_sum(x);
  const int w = threadIdx.x / kWarp, l = threadIdx.x % kWarp;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    x = (l < SNT / kWarp) ? sh[l] : 0.f;
    x = warp_sum(x);
    if (l == 0) sh[0] = x;
  }
  __syncthreads();
  return sh[0];
}

// Exp(1) draw for (row-step stream, vocab index): counter-based, independent of the thread mapping
__device__ __forceinline__ float philox_exp(unsigned long long seed, unsigned long long stream_id, unsigned idx) {
  curandStatePhilox4_32_10_t s;
  curand_init(seed, stream_id, static_cast<unsigned long long>(idx), &s);
  return -logf(curand_uniform(&s));   // curand_uniform is in (0,1]
}

__global__ void __launch_bounds__(SNT) sample_step_kernel(RolloutState st, RolloutParams p,
                                                          const float* __restrict__ logits, int ldl,
                                                          const float* __restrict__ exp_noise) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw);   // [V]
  __shared__ int hist[256];
  __shared__ ValIdx sh_vi[SNT / kWarp];
  __shared__ float sh_f[SNT / kWarp];
  __shared__ unsigned sh_prefix;
  __shared__ int sh_kk;
  __shared__ int sh_cnt;

  if (*st.done) return;
  const int r = blockIdx.x, tid = threadIdx.x;
  const int t = *st.step;
  const int blk = r / p.B;
  const int mode = p.mode_of_block[blk];   // 0 sample, 1 greedy
  const int V = p.V;
  const float* lrow = logits + static_cast<long long>(r) * ldl;
  const bool was_finished = st.finished[r] != 0;
  const float inv_temp = (mode == 0 && p.temperature != 1.0f) ? 1.0f / p.temperature : 1.0f;

  int next = p.pad;
  float lp = 0.f, margin = 0.f;
  int n_surv = 0;

  if (!was_finished) {
    // ---- stage the row as sortable keys, find the maximum ---------------------
    ValIdx best{-INFINITY, 0x7fffffff};
    for (int i = tid; i < V; i += SNT) {
      float s = lrow[i];
      if (inv_temp != 1.0f) s = s / p.temperature;
      keys[i] = f2key(s);
      best = better(best, ValIdx{s, i});
    }
    best = block_argmax(best, sh_vi);
    const float smax = best.v;

    unsigned thr_key = 0;   // everything survives
    if (mode == 0 && p.top_k > 0 && p.top_k < V) {
      // ---- radix select of the k-th largest key (4 x 8 bits, MSB first) ---------
      if (tid == 0) {
        sh_prefix = 0;
        sh_kk = p.top_k;
      }
      unsigned mask = 0;
      for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = sh_prefix;
        for (int i = tid; i < V; i += SNT) {
          const unsigned k = keys[i];
          if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
        }
        __syncthreads();
        if (tid < kWarp) {
          // lane l owns bins 255-8l .. 248-8l (descending)
          int loc = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) loc += hist[255 - 8 * tid - j];
          int inc = loc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, inc, o);
            if (tid >= o) inc += y;
          }
          const int before = inc - loc;
          const int kk = sh_kk;
          const bool mine = before < kk && kk <= inc;
          if (mine) {
            int cum = before;
            for (int j = 0; j < 8; ++j) {
              const int b = 255 - 8 * tid - j;
              if (cum + hist[b] >= kk) {
                sh_prefix = prefix | (static_cast<unsigned>(b) << shift);
                sh_kk = kk - cum;
                break;
              }
              cum += hist[b];
            }
          }
        }
        mask |= 255u << shift;
        __syncthreads();
      }
      thr_key = sh_prefix;
    }

    // ---- softmax over the survivors -------------------------------------------
    if (tid == 0) sh_cnt = 0;
    float esum = 0.f;
    for (int i = tid; i < V; i += SNT) {
      const unsigned k = keys[i];
      if (k >= thr_key) esum += expf(key2f(k) - smax);
    }
    esum = block_sum(esum, sh_f);

    if (mode == 1) {
      next = best.i;
      lp = -logf(esum);   // s_max - smax - log(sum)
      // margin: top1 - top2 logit
      ValIdx second{-INFINITY, 0x7fffffff};
      for (int i = tid; i < V; i += SNT)
        if (i != best.i) second = better(second, ValIdx{key2f(keys[i]), i});
      second = block_argmax(second, sh_vi);
      margin = smax - second.v;
      n_surv = V;
    } else {
      // ---- exponential race among the survivors: argmax (e/sum) / q --------------
      const int nrow = r % p.B;
      const float* qrow = exp_noise ? exp_noise + (static_cast<long long>(t) * p.B + nrow) * V : nullptr;
      const unsigned long long stream_id = static_cast<unsigned long long>(t) * p.R + r;
      ValIdx win{-INFINITY, 0x7fffffff};
      const long long slot0 = (static_cast<long long>(r) * p.Tmax + t) * kTopKCap;
      for (int i = tid; i < V; i += SNT) {
        const unsigned k = keys[i];
        if (k < thr_key) continue;
        const float s = key2f(k);
        const float prob = expf(s - smax) / esum;
        const float q = qrow ? qrow[i] : philox_exp(p.seed, stream_id, static_cast<unsigned>(i));
        win = better(win, ValIdx{prob / q, i});
        const int slot = atomicAdd(&sh_cnt, 1);
        if (slot < kTopKCap) {
          st.topk_idx[slot0 + slot] = i;
          st.topk_val[slot0 + slot] = s;
        }
      }
      win = block_argmax(win, sh_vi);
      next = win.i;
      lp = key2f(keys[next]) - smax - logf(esum);
      ValIdx second{-INFINITY, 0x7fffffff};
      for (int i = tid; i < V; i += SNT) {
        const unsigned k = keys[i];
        if (k < thr_key || i == win.i) continue;
        const float prob = expf(key2f(k) - smax) / esum;
        const float q = qrow ? qrow[i] : philox_exp(p.seed, stream_id, static_cast<unsigned>(i));
        second = better(second, ValIdx{prob / q, i});
      }
      second = block_argmax(second, sh_vi);
      margin = (win.v - second.v) / win.v;
      n_surv = sh_cnt;
    }
  }

  // ---- per-row state update (thread 0) ---------------------------------------
  if (tid == 0) {
    const long long ro = static_cast<long long>(r);
    const int slot = p.P + t;
    st.seq[ro * p.Lmax + slot] = next;
    st.logprob[ro * p.Tmax + t] = (next != p.pad) ? lp : 0.f;
    st.margin[ro * p.Tmax + t] = margin;
    st.topk_cnt[ro * p.Tmax + t] = n_surv;
    if (!was_finished) {
      const bool valid = p.mask_token_id < 0 || next != p.mask_token_id;
      st.key_valid[ro * p.Lmax + slot] = valid ? 1 : 0;
      const int nv = st.n_valid[r] + (valid ? 1 : 0);
      st.n_valid[r] = nv;
      st.cur_pos[r] = max(nv - 1, 0);
      // `_past` rule: the last listed special token seen strictly before this token decides its type
      const unsigned seen = st.seen[r];
      int tt = p.sections[blk][0];
      for (int i = 0; i < p.n_special[blk]; ++i)
        if (seen & (1u << i)) tt = p.sections[blk][i + 1];
      st.cur_type[r] = tt;
      unsigned add = 0;
      for (int i = 0; i < p.n_special[blk]; ++i)
        if (next == p.special_ids[blk][i]) add |= 1u << i;
      st.seen[r] = seen | add;
      st.cur_token[r] = next;
      st.cur_len[r] = slot;
      if (next == p.eos) st.finished[r] = 1;
    }
    __threadfence();
    const unsigned prev = atomicAdd(st.arrive, 1u);
    if (prev == static_cast<unsigned>(p.R) - 1) {
      // last block of this step: all rows' flags are visible
      __threadfence();
      bool all = true;
      for (int i = 0; i < p.R; ++i) all = all && (reinterpret_cast<volatile uint8_t*>(st.finished)[i] != 0);
      *st.step = t + 1;
      if (all || t + 1 >= p.Tmax) *st.done = 1;
      *st.arrive = 0;
    }
  }
}

// =============================================================================
// one-token self-attention
// =============================================================================
template <typename T>
__global__ void __launch_bounds__(128) decode_self_attn_kernel(const T* __restrict__ qkv, T* __restrict__ kcache,
                                                               T* __restrict__ vcache, T* __restrict__ ctx,
                                                               RolloutState st, int Lmax) {
  constexpr int NT = 128;
  constexpr int VN = Vec16<T>::N;        // elements per 16-byte vector
  constexpr int LPK = HD / VN;           // lanes per key
  constexpr int NG = NT / LPK;           // key groups per block
  __shared__ float qs[HD];
  __shared__ float sc[512];
  __shared__ float red[NG][HD];
  __shared__ float sh_red[NT / kWarp];
  if (*st.done) return;
  const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  if (st.finished[r]) return;
  const int L = st.cur_len[r];           // slot of the new token; keys 0..L
  const T* qrow = qkv + static_cast<long long>(r) * 3 * H + h * HD;
  T* kbase = kcache + static_cast<long long>(r) * Lmax * H + h * HD;
  T* vbase = vcache + static_cast<long long>(r) * Lmax * H + h * HD;
  if (tid < HD) {
    qs[tid] = to_f(qrow[tid]);
    kbase[static_cast<long long>(L) * H + tid] = qrow[H + tid];
  } else {
    vbase[static_cast<long long>(L) * H + (tid - HD)] = qrow[2 * H + (tid - HD)];
  }
  __syncthreads();
  const uint8_t* kv = st.key_valid + static_cast<long long>(r) * Lmax;
  const int g = tid / LPK, sub = tid % LPK;
  float qf[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) qf[i] = qs[sub * VN + i];
  const int n = L + 1;
  // warp-uniform trip count: every lane takes part in the group shuffles
  for (int jb = 0; jb < n; jb += NG) {
    const int j = jb + g;
    const bool in = j < n;
    float kf[VN];
    if (in) {
      Vec16<T> kvv;
      kvv.load(kbase + static_cast<long long>(j) * H + sub * VN);
      kvv.unpack(kf);
    } else {
#pragma unroll
      for (int i = 0; i < VN; ++i) kf[i] = 0.f;
    }
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < VN; ++i) d = fmaf(qf[i], kf[i], d);
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
    if (in && sub == 0) sc[j] = kv[j] ? d * 0.125f : -INFINITY;
  }
  __syncthreads();
  // max and sum
  float mx = -INFINITY;
  for (int j = tid; j < n; j += NT) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (tid % kWarp == 0) sh_red[tid / kWarp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(sh_red[0], sh_red[1]), fmaxf(sh_red[2], sh_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < n; j += NT) {
    const float e = (sc[j] == -INFINITY) ? 0.f : expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (tid % kWarp == 0) sh_red[tid / kWarp] = sum;
  __syncthreads();
  sum = sh_red[0] + sh_red[1] + sh_red[2] + sh_red[3];
  float acc[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) acc[i] = 0.f;
  for (int j = g; j < n; j += NG) {
    const float pj = sc[j];
    Vec16<T> vv;
    vv.load(vbase + static_cast<long long>(j) * H + sub * VN);
    float vf[VN];
    vv.unpack(vf);
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = fmaf(pj, vf[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < VN; ++i) red[g][sub * VN + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) o += red[gg][tid];
    ctx[static_cast<long long>(r) * H + h * HD + tid] = from_f<T>(sum > 0.f ? o / sum : 0.f);
  }
}

// =============================================================================
// one-token cross-attention (NQ rows of one study share every K/V load)
// =============================================================================
template <typename T, int NQ>
__global__ void __launch_bounds__(256) decode_cross_attn_kernel(const T* __restrict__ q, const T* __restrict__ kc,
                                                                const T* __restrict__ vc, int ld,
                                                                T* __restrict__ ctx,
                                                                const int* __restrict__ kv_off,
                                                                const int* __restrict__ kv_len, RolloutState st,
                                                                int B, int nsplit, float* __restrict__ ws) {
  constexpr int NT = 256;
  constexpr int VN = Vec16<T>::N;
  constexpr int LPK = HD / VN;
  constexpr int NG = NT / LPK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);      // [NQ][chunk]
  __shared__ float qs[NQ][HD];
  __shared__ float red[NG][NQ][HD];
  __shared__ float sh_red[NQ][NT / kWarp];
  if (*st.done) return;
  const int b = blockIdx.x, h = blockIdx.y, s = blockIdx.z, tid = threadIdx.x;
  bool all_fin = true;
#pragma unroll
  for (int i = 0; i < NQ; ++i) all_fin = all_fin && st.finished[b + i * B];
  if (all_fin) return;
  const int len = kv_len[b];
  const int chunk = ceil_div(len, nsplit);
  const int j0 = s * chunk, j1 = min(len, j0 + chunk);
  const int n = max(j1 - j0, 0);
  if (tid < NQ * HD) qs[tid / HD][tid % HD] = to_f(q[static_cast<long long>(b + (tid / HD) * B) * H + h * HD + tid % HD]);
  __syncthreads();
  const T* kbase = kc + (static_cast<long long>(kv_off[b]) + j0) * ld + h * HD;
  const T* vbase = vc + (static_cast<long long>(kv_off[b]) + j0) * ld + h * HD;
  const int g = tid / LPK, sub = tid % LPK;
  float qf[NQ][VN];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) qf[i][e] = qs[i][sub * VN + e];

  // pass 1: scores (4 keys in flight per thread)
  constexpr int U = 4;
  // warp-uniform trip count: every lane takes part in the group shuffles
  for (int jb = 0; jb < n; jb += NG * U) {
    Vec16<T> kv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * NG + g;
      if (j < n) {
        kv[u].load_nc(kbase + static_cast<long long>(j) * ld + sub * VN);
      } else {
        float z[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) z[e] = 0.f;
        kv[u].pack(z);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * NG + g;
      float kf[VN];
      kv[u].unpack(kf);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < VN; ++e) d = fmaf(qf[i][e], kf[e], d);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
        if (j < n && sub == 0) sc[i * chunk + j] = d * 0.125f;
      }
    }
  }
  __syncthreads();
  float mx[NQ], sum[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float m = -INFINITY;
    for (int j = tid; j < n; j += NT) m = fmaxf(m, sc[i * chunk + j]);
    m = warp_max(m);
    if (tid % kWarp == 0) sh_red[i][tid / kWarp] = m;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NT / kWarp; ++w) m = fmaxf(m, sh_red[i][w]);
    mx[i] = m;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float sm = 0.f;
    for (int j = tid; j < n; j += NT) {
      const float e = expf(sc[i * chunk + j] - mx[i]);
      sc[i * chunk + j] = e;
      sm += e;
    }
    sm = warp_sum(sm);
    if (tid % kWarp == 0) sh_red[i][tid / kWarp] = sm;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float sm = 0.f;
#pragma unroll
    for (int w = 0; w < NT / kWarp; ++w) sm += sh_red[i][w];
    sum[i] = sm;
  }
  // pass 2: weighted sum of V
  float acc[NQ][VN];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[i][e] = 0.f;
  for (int jb = 0; jb < n; jb += NG * U) {
    Vec16<T> vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * NG + g;
      if (j < n) vv[u].load_nc(vbase + static_cast<long long>(j) * ld + sub * VN);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * NG + g;
      if (j >= n) continue;
      float vf[VN];
      vv[u].unpack(vf);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const float pj = sc[i * chunk + j];
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[i][e] = fmaf(pj, vf[e], acc[i][e]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) red[g][i][sub * VN + e] = acc[i][e];
  __syncthreads();
  if (tid < NQ * HD) {
    const int i = tid / HD, d = tid % HD;
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) o += red[gg][i][d];
    const int r = b + i * B;
    const float m_i = (i == 0) ? mx[0] : mx[NQ - 1];
    const float s_i = (i == 0) ? sum[0] : sum[NQ - 1];
    if (nsplit == 1) {
      ctx[static_cast<long long>(r) * H + h * HD + d] = from_f<T>(s_i > 0.f ? o / s_i : 0.f);
    } else {
      float* w = ws + ((static_cast<long long>(r) * NH + h) * nsplit + s) * (HD + 2);
      w[2 + d] = o;
      if (d == 0) {
        w[0] = (n > 0) ? m_i : -INFINITY;
        w[1] = s_i;
      }
    }
  }
}

template <typename T>
__global__ void cross_combine_kernel(const float* __restrict__ ws, T* __restrict__ ctx, RolloutState st, int nsplit) {
  if (*st.done) return;
  const int r = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
  if (st.finished[r]) return;
  const float* w = ws + (static_cast<long long>(r) * NH + h) * nsplit * (HD + 2);
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, w[s * (HD + 2)]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = w[s * (HD + 2)];
    if (ms == -INFINITY) continue;
    const float a = expf(ms - m);
    l += a * w[s * (HD + 2) + 1];
    o += a * w[s * (HD + 2) + 2 + d];
  }
  ctx[static_cast<long long>(r) * H + h * HD + d] = from_f<T>(l > 0.f ? o / l : 0.f);
}

template <typename T>
__global__ void prefill_store_kv_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                                        int R, int P, int Lmax) {
  constexpr int VN = Vec16<T>::N;
  const int cv = H / VN;
  const long long total = static_cast<long long>(R) * P * cv * 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * VN;
    long long t = i / cv;
    const int which = static_cast<int>(t % 2);
    t /= 2;
    const int pcol = static_cast<int>(t % P);
    const long long r = t / P;
    Vec16<T> v;
    v.load(qkv + (r * P + pcol) * 3 * H + (1 + which) * H + c);
    v.store((which ? vcache : kcache) + (r * Lmax + pcol) * H + c);
  }
}

template <typename T>
__global__ void take_last_kernel(const T* __restrict__ x, T* __restrict__ out, int R, int P, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(R) * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  out[i] = x[(r * P + (P - 1)) * C + c];
}

__global__ void cosine_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                   int n, int C) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, lane = threadIdx.x % kWarp;
  if (row >= n) return;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int c = lane; c < C; c += kWarp) {
    const float x = a[static_cast<long long>(row) * C + c], y = b[static_cast<long long>(row) * C + c];
    ab = fmaf(x, y, ab);
    aa = fmaf(x, x, aa);
    bb = fmaf(y, y, bb);
  }
  ab = warp_sum(ab);
  aa = warp_sum(aa);
  bb = warp_sum(bb);
  // torch.nn.functional.cosine_similarity: x.y / (max(|x|, eps) * max(|y|, eps)), eps = 1e-8
  if (lane == 0) out[row] = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
}

}  // namespace

void rollout_init(const RolloutState& st, const RolloutParams& p, const int* prompt_ids, int* pre_ids, int* pre_types,
                  int* pre_pos, cudaStream_t stream) {
  rollout_init_kernel<<<ceil_div(p.R, 64), 64, 0, stream>>>(st, p, prompt_ids, pre_ids, pre_types, pre_pos);
  check_launch("rollout_init");
}

void sample_step(const RolloutState& st, const RolloutParams& p, const float* logits, int ldl, const float* exp_noise,
                 cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(p.V) * sizeof(unsigned);
  static size_t configured = 0;
  if (smem > configured) {
    CXRM_CHECK(smem <= 200 * 1024, "vocabulary too large for the sampling kernel's shared-memory staging");
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(sample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    configured = smem;
  }
  sample_step_kernel<<<p.R, SNT, smem, stream>>>(st, p, logits, ldl, exp_noise);
  check_launch("sample_step");
}

template <typename T>
void decode_self_attention(const T* qkv, T* kcache, T* vcache, T* ctx, const RolloutState& st, int R, int Lmax,
                           cudaStream_t stream) {
  CXRM_CHECK(Lmax <= 512, "decode_self_attention supports at most 512 cached tokens");
  decode_self_attn_kernel<T><<<dim3(R, NH), 128, 0, stream>>>(qkv, kcache, vcache, ctx, st, Lmax);
  check_launch("decode_self_attention");
}

size_t decode_cross_ws_bytes(int R, int nsplit) {
  return static_cast<size_t>(R) * NH * nsplit * (HD + 2) * sizeof(float);
}

template <typename T>
void decode_cross_attention(const T* q, const T* kc, const T* vc, int ld, T* ctx, const int* kv_off,
                            const int* kv_len, const RolloutState& st, int R, int B, int max_len, int nsplit, float* ws,
                            cudaStream_t stream) {
  const int nq = R / B;
  CXRM_CHECK(nq == 1 || nq == 2, "decode_cross_attention: 1 or 2 rows per study");
  const int chunk = ceil_div(max_len, nsplit);
  const size_t smem = static_cast<size_t>(nq) * chunk * sizeof(float);
  CXRM_CHECK(smem <= 160 * 1024, "cross-attention chunk too large: raise nsplit");
  auto launch = [&](auto kern) {
    static size_t configured = 0;
    if (smem > configured) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured = smem;
    }
    kern<<<dim3(B, NH, nsplit), 256, smem, stream>>>(q, kc, vc, ld, ctx, kv_off, kv_len, st, B, nsplit, ws);
  };
  if (nq == 1)
    launch(decode_cross_attn_kernel<T, 1>);
  else
    launch(decode_cross_attn_kernel<T, 2>);
  check_launch("decode_cross_attention");
  if (nsplit > 1) {
    cross_combine_kernel<T><<<dim3(R, NH), HD, 0, stream>>>(ws, ctx, st, nsplit);
    check_launch("cross_combine");
  }
}

template <typename T>
void prefill_store_kv(const T* qkv, T* kcache, T* vcache, int R, int P, int Lmax, cudaStream_t stream) {
  const long long total = static_cast<long long>(R) * P * (H / Vec16<T>::N) * 2;
  if (total <= 0) return;
  long long grid = ceil_div_ll(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  prefill_store_kv_kernel<T><<<static_cast<unsigned>(grid), 256, 0, stream>>>(qkv, kcache, vcache, R, P, Lmax);
  check_launch("prefill_store_kv");
}

template <typename T>
void take_last_token(const T* x, T* out, int R, int P, int C, cudaStream_t stream) {
  const long long total = static_cast<long long>(R) * C;
  take_last_kernel<T><<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(x, out, R, P, C);
  check_launch("take_last_token");
}

void cosine_rows(const float* a, const float* b, float* out, int n, int C, cudaStream_t stream) {
  if (n <= 0) return;
  cosine_rows_kernel<<<ceil_div(n * kWarp, 128), 128, 0, stream>>>(a, b, out, n, C);
  check_launch("cosine_rows");
}

#define INST(T)                                                                                                  \
  template void decode_self_attention<T>(const T*, T*, T*, T*, const RolloutState&, int, int, cudaStream_t);      \
  template void decode_cross_attention<T>(const T*, const T*, const T*, int, T*, const int*, const int*,          \
                                          const RolloutState&, int, int, int, int, float*, cudaStream_t);         \
  template void prefill_store_kv<T>(const T*, T*, T*, int, int, int, cudaStream_t);                               \
  template void take_last_token<T>(const T*, T*, int, int, int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace cxrm
