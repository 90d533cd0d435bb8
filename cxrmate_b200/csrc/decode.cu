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

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int H = 768, HD = 64, NH = 12;

// =============================================================================
// rollout_init
// =============================================================================
__global__ void rollout_init_kernel(RolloutState st, RolloutParams p, const int* __restrict__ prompt, int* pre_ids,
                                    int* pre_types, int* pre_pos, uint8_t* pre_valid) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) {
    *st.step = 0;
    *st.done = 0;
    *st.arrive = 0;
    *st.seed = p.seed;
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
    pre_valid[static_cast<long long>(r) * p.P + c] = valid ? 1 : 0;   // key mask of the prompt pass (column layout)
    pre_ids[static_cast<long long>(r) * p.P + c] = id;
    pre_types[static_cast<long long>(r) * p.P + c] = tt;
    pre_pos[static_cast<long long>(r) * p.P + c] = max(cum - 1, 0);
  }
  // The self-attention cache holds only the VISIBLE prompt tokens, compacted: token (r, c) sits at slot pos = cum - 1
  // (a masked key has softmax weight exactly 0 in the reference - additive finfo.min - so dropping it changes
  // nothing, and right-padded prompts stop costing K/V reads on every decode step); generated tokens follow.
  for (int c = 0; c < p.Lmax; ++c) st.key_valid[static_cast<long long>(r) * p.Lmax + c] = c < cum ? 1 : 0;
  for (int c = p.P; c < p.Lmax; ++c) st.seq[static_cast<long long>(r) * p.Lmax + c] = p.pad;
  st.n_valid[r] = cum;
  st.seen[r] = seen;
  st.cur_len[r] = cum - 1;   // last occupied cache slot; the sampling head advances it per emitted token
  st.cur_token[r] = p.pad;
  st.cur_type[r] = 0;
  st.cur_pos[r] = 0;
  st.finished[r] = 0;
}

// =============================================================================
// packed prompt pass (kernels.h PromptPack)
// =============================================================================
__global__ void pack_prompt_count_kernel(PromptPack pk, const uint8_t* __restrict__ pre_valid, int R, int P) {
  __shared__ int s_lq[1024];
  const int r = threadIdx.x;
  int nv = 0, extra = 0;
  if (r < R) {
    const uint8_t* v = pre_valid + static_cast<long long>(r) * P;
    for (int c = 0; c < P; ++c) nv += v[c] ? 1 : 0;
    extra = v[P - 1] ? 0 : 1;          // the last column is masked: it still is the query of the first new token
    pk.row_lk[r] = nv;
    pk.row_lq[r] = nv + extra;
    s_lq[r] = nv + extra;
  }
  __syncthreads();
  if (r == 0) {
    int off = 0;
    for (int i = 0; i < R; ++i) {
      pk.row_off[i] = off;
      off += s_lq[i];
      pk.last_idx[i] = off - 1;
    }
    *pk.total = off;
  }
}
// one warp per row: ballot compaction of the visible columns, 32 columns per round
__global__ void __launch_bounds__(32) pack_prompt_fill_kernel(PromptPack pk, const int* __restrict__ pre_ids,
                                                              const int* __restrict__ pre_types, const int* __restrict__ pre_pos,
                                                              const uint8_t* __restrict__ pre_valid, int P) {
  const int r = blockIdx.x, lane = threadIdx.x;
  const long long base = static_cast<long long>(r) * P;
  int out = pk.row_off[r];
  for (int c0 = 0; c0 < P; c0 += 32) {
    const int c = c0 + lane;
    const bool vis = c < P && pre_valid[base + c] != 0;
    const unsigned m = __ballot_sync(kFull, vis);
    if (vis) {
      const int i = out + __popc(m & ((1u << lane) - 1u));
      pk.ids[i] = pre_ids[base + c];
      pk.types[i] = pre_types[base + c];
      pk.pos[i] = pre_pos[base + c];
      pk.tok_row[i] = r;
      pk.tok_slot[i] = pre_pos[base + c];      // compact cache: a visible token sits at slot cumsum(mask) - 1 = its position
      pk.tok_cache[i] = 1;
    }
    out += __popc(m);
  }
  if (lane == 0 && pre_valid[base + P - 1] == 0) {
    pk.ids[out] = pre_ids[base + P - 1];
    pk.types[out] = pre_types[base + P - 1];
    pk.pos[out] = pre_pos[base + P - 1];
    pk.tok_row[out] = r;
    pk.tok_slot[out] = 0;
    pk.tok_cache[out] = 0;
  }
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

// k-th largest of arr[0..n) (sortable keys in shared memory): 4 x 8-bit radix select, MSB first.
// Block-wide; hist/sh_prefix/sh_kk are shared scratch.  Returns the key of the k-th largest element.
__device__ unsigned radix_kth(const unsigned* arr, int n, int k, int* hist, unsigned* sh_prefix, int* sh_kk) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    *sh_prefix = 0;
    *sh_kk = k;
  }
  unsigned mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = *sh_prefix;
    for (int i = tid; i < n; i += SNT) {
      const unsigned key = arr[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
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
      const int kk = *sh_kk;
      if (before < kk && kk <= inc) {
        int cum = before;
        for (int j = 0; j < 8; ++j) {
          const int bin = 255 - 8 * tid - j;
          if (cum + hist[bin] >= kk) {
            *sh_prefix = prefix | (static_cast<unsigned>(bin) << shift);
            *sh_kk = kk - cum;
            break;
          }
          cum += hist[bin];
        }
      }
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  return *sh_prefix;
}

__global__ void __launch_bounds__(SNT) sample_step_kernel(RolloutState st, RolloutParams p,
                                                          const float* __restrict__ logits, int ldl,
                                                          const float* __restrict__ exp_noise) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw);   // [V]
  constexpr int kCandCap = 1024;                             // elements >= the lower bound of the k-th value
  constexpr int kListCap = 256;                              // survivors staged for the race (top-k + ties)
  __shared__ int hist[256];
  __shared__ ValIdx sh_vi[SNT / kWarp];
  __shared__ float sh_f[SNT / kWarp];
  __shared__ unsigned tmax[SNT];
  __shared__ int cand[kCandCap];
  __shared__ int sv_idx[kListCap];
  __shared__ unsigned sh_prefix;
  __shared__ int sh_kk;
  __shared__ int sh_cnt, sh_ncand;

  pdl_launch_dependents();
  // the rollout state was written by the PREVIOUS step's sampling kernel (or the prompt pass) and a step opens with a
  // fully serialised launch: everything but the logits is readable ahead of the dependency wait
  const bool all_done = *st.done != 0;
  const int r = blockIdx.x, tid = threadIdx.x;
  const int t = *st.step;
  const int blk = r / p.B;
  const int mode = p.mode_of_block[blk];   // 0 sample, 1 greedy
  const int V = p.V;
  const float* lrow = logits + static_cast<long long>(r) * ldl;
  const bool was_finished = st.finished[r] != 0;
  const float inv_temp = (mode == 0 && p.temperature != 1.0f) ? 1.0f / p.temperature : 1.0f;
  const bool diag = p.want_margin != 0;

  int next = p.pad;
  float lp = 0.f, margin = 0.f;
  int n_surv = 0;
  // the row state thread 0 updates at the end: requested now, so the loads ride under the sampling work
  int pre_len = 0, pre_nv = 0;
  unsigned pre_seen = 0;
  if (tid == 0 && !was_finished) {
    pre_len = st.cur_len[r];
    pre_nv = st.n_valid[r];
    pre_seen = st.seen[r];
  }
  pdl_wait();
  if (all_done) return;

  if (!was_finished) {
    // ---- pass 1: stage the row as sortable keys; per-thread and block maximum ---------------------
    ValIdx best{-INFINITY, 0x7fffffff};
    if ((V % 4) == 0 && (ldl % 4) == 0 && (reinterpret_cast<uintptr_t>(logits) % 16) == 0) {
      // 128-bit streaming loads, eight in flight per thread (a scalar loop serialises ~60 L2 round trips)
      const float4* l4 = reinterpret_cast<const float4*>(lrow);
      const int nv4 = V / 4;
      constexpr int U = 8;   // 16-byte loads in flight per thread: the 120 KB row arrives in two L2 round trips
      for (int base = tid; base < nv4; base += U * SNT) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = base + u * SNT;
          v[u] = (idx < nv4) ? __ldcs(l4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = base + u * SNT;
          if (idx >= nv4) continue;
          float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint4 kq;
          unsigned* kp = reinterpret_cast<unsigned*>(&kq);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (inv_temp != 1.0f) e[j] = e[j] / p.temperature;
            kp[j] = f2key(e[j]);
            best = better(best, ValIdx{e[j], 4 * idx + j});
          }
          *reinterpret_cast<uint4*>(keys + 4 * idx) = kq;
        }
      }
    } else {
      for (int i = tid; i < V; i += SNT) {
        float s = lrow[i];
        if (inv_temp != 1.0f) s = s / p.temperature;
        keys[i] = f2key(s);
        best = better(best, ValIdx{s, i});
      }
    }
    tmax[tid] = (best.i != 0x7fffffff) ? f2key(best.v) : 0u;
    best = block_argmax(best, sh_vi);   // (its barriers also publish keys / tmax)
    const float smax = best.v;

    const bool filter = (mode == 0 && p.top_k > 0 && p.top_k < V);
    unsigned thr_key = 0;   // everything survives
    bool listed = false;    // sv_idx / sh_cnt hold exactly the survivors
    if (tid == 0) {
      sh_cnt = 0;
      sh_ncand = 0;
    }
    __syncthreads();
    if (filter) {
      bool exact = false;
      if (p.top_k <= SNT / 2) {
        // The k largest per-thread maxima are k distinct elements, so the k-th largest of the 512 maxima is a LOWER
        // bound of the k-th largest logit: one more pass collects the few elements above it, and the exact
        // threshold (ties kept, as TopKLogitsWarper does) is found among those candidates only.
        const unsigned lb = radix_kth(tmax, SNT, p.top_k, hist, &sh_prefix, &sh_kk);
        for (int i = tid; i < V; i += SNT)
          if (keys[i] >= lb) {
            const int slot = atomicAdd(&sh_ncand, 1);
            if (slot < kCandCap) cand[slot] = i;
          }
        __syncthreads();
        const int nc = sh_ncand;
        if (nc <= kCandCap) {
          for (int x = tid; x < nc; x += SNT) {
            const unsigned v = keys[cand[x]];
            int gt = 0, ge = 0;
            for (int y = 0; y < nc; ++y) {
              const unsigned u = keys[cand[y]];
              gt += u > v;
              ge += u >= v;
            }
            if (gt < p.top_k && p.top_k <= ge) sh_prefix = v;   // every such thread writes the same value
          }
          __syncthreads();
          thr_key = sh_prefix;
          for (int x = tid; x < nc; x += SNT)
            if (keys[cand[x]] >= thr_key) {
              const int slot = atomicAdd(&sh_cnt, 1);
              if (slot < kListCap) sv_idx[slot] = cand[x];
            }
          __syncthreads();
          exact = true;
          listed = sh_cnt <= kListCap;
        }
      }
      if (!exact) {   // very flat rows (more than kCandCap elements above the bound) or a large k: full radix select
        thr_key = radix_kth(keys, V, p.top_k, hist, &sh_prefix, &sh_kk);
        if (tid == 0) sh_cnt = 0;
        __syncthreads();
        for (int i = tid; i < V; i += SNT)
          if (keys[i] >= thr_key) atomicAdd(&sh_cnt, 1);
        __syncthreads();
      }
    }

    // ---- softmax denominator over the survivors ------------------------------------------------
    float esum = 0.f;
    if (listed) {
      for (int x = tid; x < sh_cnt; x += SNT) esum += expf(key2f(keys[sv_idx[x]]) - smax);
    } else {
      for (int i = tid; i < V; i += SNT) {
        const unsigned k = keys[i];
        if (k >= thr_key) esum += expf(key2f(k) - smax);
      }
    }
    esum = block_sum(esum, sh_f);

    if (mode == 1) {
      next = best.i;
      lp = -logf(esum);   // s_max - smax - log(sum)
      if (diag) {         // margin: top1 - top2 logit
        ValIdx second{-INFINITY, 0x7fffffff};
        for (int i = tid; i < V; i += SNT)
          if (i != best.i) second = better(second, ValIdx{key2f(keys[i]), i});
        second = block_argmax(second, sh_vi);
        margin = smax - second.v;
      }
      n_surv = V;
    } else {
      // ---- exponential race among the survivors: argmax (e/sum) / q --------------
      const int nrow = r % p.B;
      const float* qrow = exp_noise ? exp_noise + (static_cast<long long>(t) * p.B + nrow) * V : nullptr;
      const unsigned long long stream_id = static_cast<unsigned long long>(t) * p.R + r;
      const unsigned long long seed = *st.seed;
      const long long slot0 = (static_cast<long long>(r) * p.Tmax + t) * kTopKCap;
      n_surv = filter ? sh_cnt : V;
      ValIdx win{-INFINITY, 0x7fffffff};
      auto ratio = [&](int i) {
        const float prob = expf(key2f(keys[i]) - smax) / esum;
        const float q = qrow ? qrow[i] : philox_exp(seed, stream_id, static_cast<unsigned>(i));
        return prob / q;
      };
      if (listed) {
        for (int x = tid; x < n_surv; x += SNT) {
          const int i = sv_idx[x];
          win = better(win, ValIdx{ratio(i), i});
          if (x < kTopKCap) {
            st.topk_idx[slot0 + x] = i;
            st.topk_val[slot0 + x] = key2f(keys[i]);
          }
        }
      } else {   // no top-k filter, or more ties at the threshold than the list holds: scan the row
        for (int i = tid; i < V; i += SNT)
          if (keys[i] >= thr_key) win = better(win, ValIdx{ratio(i), i});
      }
      win = block_argmax(win, sh_vi);
      next = win.i;
      lp = key2f(keys[next]) - smax - logf(esum);
      if (diag) {
        ValIdx second{-INFINITY, 0x7fffffff};
        for (int i = tid; i < V; i += SNT) {
          if (keys[i] < thr_key || i == win.i) continue;
          second = better(second, ValIdx{ratio(i), i});
        }
        second = block_argmax(second, sh_vi);
        margin = (win.v - second.v) / win.v;
      }
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
      const int cslot = pre_len + 1;           // cache slot of the emitted token (visible prompt tokens + emitted so far)
      st.key_valid[ro * p.Lmax + cslot] = valid ? 1 : 0;
      const int nv = pre_nv + (valid ? 1 : 0);
      st.n_valid[r] = nv;
      st.cur_pos[r] = max(nv - 1, 0);
      // `_past` rule: the last listed special token seen strictly before this token decides its type
      const unsigned seen = pre_seen;
      int tt = p.sections[blk][0];
      for (int i = 0; i < p.n_special[blk]; ++i)
        if (seen & (1u << i)) tt = p.sections[blk][i + 1];
      st.cur_type[r] = tt;
      unsigned add = 0;
      for (int i = 0; i < p.n_special[blk]; ++i)
        if (next == p.special_ids[blk][i]) add |= 1u << i;
      st.seen[r] = seen | add;
      st.cur_token[r] = next;
      st.cur_len[r] = cslot;
      if (next == p.eos) st.finished[r] = 1;
    }
    // the arrival ticket also counts the finished rows (high half): the last block knows whether every row is finished
    // without reading the flags back (a short-circuit loop of R volatile loads was ~10 us of serial L2 round trips at
    // the very end of every decode step)
    const unsigned fin_now = (was_finished || next == p.eos) ? 1u : 0u;
    __threadfence();
    const unsigned prev = atomicAdd(st.arrive, 1u | (fin_now << 16));
    if ((prev & 0xffffu) == static_cast<unsigned>(p.R) - 1) {
      const bool all = (prev >> 16) + fin_now == static_cast<unsigned>(p.R);
      *st.step = t + 1;
      if (all || t + 1 >= p.Tmax) *st.done = 1;
      *st.arrive = 0;
    }
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
                  int* pre_pos, uint8_t* pre_valid, cudaStream_t stream) {
  rollout_init_kernel<<<ceil_div(p.R, 64), 64, 0, stream>>>(st, p, prompt_ids, pre_ids, pre_types, pre_pos, pre_valid);
  check_launch("rollout_init");
}

void pack_prompt(const PromptPack& pk, const int* pre_ids, const int* pre_types, const int* pre_pos, const uint8_t* pre_valid,
                 int R, int P, cudaStream_t stream) {
  CXRM_CHECK(R >= 1 && R <= 1024 && P >= 1, "pack_prompt shape");
  pack_prompt_count_kernel<<<1, 1024, 0, stream>>>(pk, pre_valid, R, P);
  check_launch("pack_prompt_count");
  pack_prompt_fill_kernel<<<R, 32, 0, stream>>>(pk, pre_ids, pre_types, pre_pos, pre_valid, P);
  check_launch("pack_prompt_fill");
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
  launch_chain(sample_step_kernel, dim3(p.R), dim3(SNT), smem, stream, st, p, logits, ldl, exp_noise);
  check_launch("sample_step");
}

// Test hook (cxrm_test_sample): one call of the sampling head on caller-provided logits, every row a SAMPLE row of a
// fresh rollout whose step counter is `step` - the in-kernel Philox stream is (seed, step * R + row), the draw for
// vocabulary entry i is counter i of that stream (cxrm.h).
void sample_rows_test(const float* logits, int R, int V, int top_k, float temperature, unsigned long long seed, int step,
                      int Tmax, int* out_tokens, float* out_logprob, cudaStream_t stream) {
  CXRM_CHECK(R >= 1 && V >= 1 && step >= 0 && step < Tmax, "sample_rows_test shape");
  const int Lmax = Tmax + 16;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~static_cast<size_t>(255); return o; };
  const size_t o_tok = take(sizeof(int) * R), o_len = take(sizeof(int) * R), o_type = take(sizeof(int) * R),
               o_pos = take(sizeof(int) * R), o_nv = take(sizeof(int) * R), o_seen = take(sizeof(unsigned) * R),
               o_fin = take(R), o_kv = take(static_cast<size_t>(R) * Lmax), o_seq = take(sizeof(int) * R * Lmax),
               o_lp = take(sizeof(float) * R * Tmax), o_mg = take(sizeof(float) * R * Tmax),
               o_ti = take(sizeof(int) * R * Tmax * kTopKCap), o_tv = take(sizeof(float) * R * Tmax * kTopKCap),
               o_tc = take(sizeof(int) * R * Tmax), o_sc = take(256);
  char* base = nullptr;
  CXRM_CUDA_CHECK(cudaMallocAsync(&base, off, stream));
  CXRM_CUDA_CHECK(cudaMemsetAsync(base, 0, off, stream));
  RolloutState st{};
  st.cur_token = reinterpret_cast<int*>(base + o_tok); st.cur_len = reinterpret_cast<int*>(base + o_len);
  st.cur_type = reinterpret_cast<int*>(base + o_type); st.cur_pos = reinterpret_cast<int*>(base + o_pos);
  st.n_valid = reinterpret_cast<int*>(base + o_nv); st.seen = reinterpret_cast<unsigned*>(base + o_seen);
  st.finished = reinterpret_cast<uint8_t*>(base + o_fin); st.key_valid = reinterpret_cast<uint8_t*>(base + o_kv);
  st.seq = reinterpret_cast<int*>(base + o_seq); st.logprob = reinterpret_cast<float*>(base + o_lp);
  st.margin = reinterpret_cast<float*>(base + o_mg); st.topk_idx = reinterpret_cast<int*>(base + o_ti);
  st.topk_val = reinterpret_cast<float*>(base + o_tv); st.topk_cnt = reinterpret_cast<int*>(base + o_tc);
  st.step = reinterpret_cast<int*>(base + o_sc); st.done = st.step + 1;
  st.arrive = reinterpret_cast<unsigned*>(st.step + 2);
  st.seed = reinterpret_cast<unsigned long long*>(base + o_sc + 64);
  CXRM_CUDA_CHECK(cudaMemcpyAsync(st.step, &step, sizeof(int), cudaMemcpyHostToDevice, stream));
  CXRM_CUDA_CHECK(cudaMemcpyAsync(st.seed, &seed, sizeof(seed), cudaMemcpyHostToDevice, stream));
  RolloutParams p{};
  p.R = R; p.B = R; p.P = 1; p.Lmax = Lmax; p.Tmax = Tmax; p.V = V;
  p.mode_of_block[0] = 0; p.mode_of_block[1] = 0;
  p.mask_token_id = -1; p.eos = -1; p.pad = -2; p.top_k = top_k; p.temperature = temperature; p.seed = seed;
  sample_step(st, p, logits, V, nullptr, stream);
  CXRM_CUDA_CHECK(cudaMemcpy2DAsync(out_tokens, sizeof(int), st.seq + 1 + step, sizeof(int) * Lmax, sizeof(int), R,
                                    cudaMemcpyDeviceToDevice, stream));
  if (out_logprob)
    CXRM_CUDA_CHECK(cudaMemcpy2DAsync(out_logprob, sizeof(float), st.logprob + step, sizeof(float) * Tmax, sizeof(float), R,
                                      cudaMemcpyDeviceToDevice, stream));
  CXRM_CUDA_CHECK(cudaFreeAsync(base, stream));
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

// REINFORCE loss of one rollout batch (reference scst/gen_prompt.py:350-364):
//   loss = mean_b( -sum_t logprob[b, t] * advantage[b] ),
// logprob = log-softmax of the top-k-masked scores at the sampled id, 0 at PAD positions (sample_step stores it so),
// which is nll_loss(log_softmax(scores), ids, ignore_index=pad, 'none').sum(-1) * reward, .mean().
// One block; fixed summation order (thread b sums its row, then a shared-memory tree): reproducible.
__global__ void __launch_bounds__(256) reinforce_loss_kernel(const float* __restrict__ logprob, int ld,
                                                             const float* __restrict__ advantage, int B, int T,
                                                             float* __restrict__ loss) {
  __shared__ float sh[256];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    float srow = 0.f;
    for (int t = 0; t < T; ++t) srow += logprob[static_cast<long long>(b) * ld + t];
    acc += -srow * advantage[b];
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = sh[0] / B;
}

void reinforce_loss(const float* logprob, int ld, const float* advantage, int B, int T, float* loss, cudaStream_t stream) {
  CXRM_CHECK(B >= 1 && T >= 1 && ld >= T, "reinforce_loss shape");
  reinforce_loss_kernel<<<1, 256, 0, stream>>>(logprob, ld, advantage, B, T, loss);
  check_launch("reinforce_loss");
}

#define INST(T) template void take_last_token<T>(const T*, T*, int, int, int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace cxrm
