// Beam search on top of the KV-cached rollout (reference test_step: generate(num_beams = num_test_beams),
// modules/lightning_modules/longitudinal/gt_prompt.py:344-362, single.py:552-562).
//
// The rows of the rollout are the running beams, beam-major: row = beam * B + study, every row a "virtual study" of the
// decode-step kernels (the cross-attention unit table of beam j points at the encoder K/V of its real study, so the
// K/V cache is shared, not replicated).  Per step, after the LM head:
//
//   beam_step_kernel   one CTA per study: log_softmax of each beam's fp32 logits + accumulated score, the 2*nb best
//                      continuations over the flattened (beam, token) axis, HF's bookkeeping of running / finished
//                      hypotheses and its early-stop heuristic (transformers generation/utils.py `_beam_search`:
//                      `_get_top_k_continuations`, `_get_running_beams_for_next_iteration`, `_update_finished_beams`,
//                      `_check_early_stop_heuristic`; default flags: early_stopping False), the gather of the running
//                      sequences and per-row decode state from the source beams;
//   beam_kv_gather / beam_kv_scatter   `Cache.reorder_cache(beam_idx)`: the GENERATED slots of the self-attention
//                      K/V of every row whose source beam differs move through a scratch copy (the prompt slots are
//                      identical across the beams of a study and never move).
#include <cmath>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int BNT = 512;
constexpr float NEG = -1.0e9f;
constexpr int NHB = 12, HDB = 64;

struct Cand {
  float v;
  int i;
};
// total order of the candidates: larger value first, ties by lower flat index
__device__ __forceinline__ bool before(Cand a, Cand b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

__device__ float blk_max(float x, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(kFull, x, o));
  const int w = threadIdx.x / kWarp, l = threadIdx.x % kWarp;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  x = sh[0];
  for (int i = 1; i < BNT / kWarp; ++i) x = fmaxf(x, sh[i]);
  return x;
}
__device__ float blk_sum(float x, float* sh) {
  x = warp_sum(x);
  const int w = threadIdx.x / kWarp, l = threadIdx.x % kWarp;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < BNT / kWarp; ++i) s += sh[i];   // fixed order: every thread gets the same bits
  return s;
}
__device__ Cand blk_best(Cand x, Cand* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Cand y{__shfl_xor_sync(kFull, x.v, o), __shfl_xor_sync(kFull, x.i, o)};
    if (before(y, x)) x = y;
  }
  const int w = threadIdx.x / kWarp, l = threadIdx.x % kWarp;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  x = sh[0];
  for (int i = 1; i < BNT / kWarp; ++i)
    if (before(sh[i], x)) x = sh[i];
  return x;
}

__global__ void beam_init_kernel(BeamState bs, int B, int nb, int Tmax, int fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    *bs.n_slots = 0;
    *bs.arrive = 0;
    *bs.cnt_can = 0;
    *bs.cnt_hit = 0;
  }
  if (i < B) bs.can_improve[i] = 1;
  if (i < B * nb) {
    bs.run_score[i] = (i % nb == 0) ? 0.f : NEG;
    bs.fin_score[i] = NEG;
    bs.fin_len[i] = 0;
    bs.is_fin[i] = 0;
  }
  for (long long x = i; x < static_cast<long long>(B) * nb * Tmax; x += static_cast<long long>(gridDim.x) * blockDim.x)
    bs.fin_seq[x] = fill;
}

__global__ void __launch_bounds__(BNT) beam_step_kernel(RolloutState st, RolloutParams p, BeamState bs,
                                                        const float* __restrict__ logits, int ldl) {
  __shared__ float sh_f[BNT / kWarp];
  __shared__ Cand sh_c[BNT / kWarp];
  __shared__ float s_lse[kMaxBeams], s_rs[kMaxBeams];
  __shared__ Cand s_cand[2 * kMaxBeams];
  __shared__ int s_keep[kMaxBeams], s_finsrc[kMaxBeams], s_src[2 * kMaxBeams], s_tok[2 * kMaxBeams];
  __shared__ int s_oldlen[kMaxBeams], s_oldnv[kMaxBeams];
  __shared__ unsigned s_oldseen[kMaxBeams];
  __shared__ int s_oldfin[kMaxBeams][kBeamMaxT], s_oldrun[kMaxBeams][kBeamMaxT];
  __shared__ uint8_t s_oldkv[kMaxBeams][kBeamMaxT];

  const int tid = threadIdx.x, b = blockIdx.x;
  const int nb = p.beams, K = 2 * nb, B = p.R / nb, V = p.V;
  if (*st.done) {
    if (b == 0 && tid == 0) *bs.n_slots = 0;   // nothing moved in this (skipped) step: no cache reorder
    return;
  }
  const int t = *st.step;
  const int fill = p.pad ? p.pad : p.eos;      // HF: output_fill_value = pad_token_id or eos_token_id[0]

  // ---- A: log-sum-exp of every running beam's logits -------------------------------------------------------------
  for (int j = 0; j < nb; ++j) {
    const float* row = logits + static_cast<long long>(j * B + b) * ldl;
    float m = -INFINITY;
    for (int v = tid; v < V; v += BNT) m = fmaxf(m, row[v]);
    m = blk_max(m, sh_f);
    float s = 0.f;
    for (int v = tid; v < V; v += BNT) s += expf(row[v] - m);
    s = blk_sum(s, sh_f);
    if (tid == 0) {
      s_lse[j] = m + logf(s);
      s_rs[j] = bs.run_score[b * nb + j];
    }
  }
  __syncthreads();

  // ---- B: the K best (beam, token) continuations, best first -----------------------------------------------------
  Cand prev{INFINITY, -1};
  for (int k = 0; k < K; ++k) {
    Cand best{-INFINITY, 0x7fffffff};
    for (int j = 0; j < nb; ++j) {
      const float* row = logits + static_cast<long long>(j * B + b) * ldl;
      const float lse = s_lse[j], rs = s_rs[j];
      for (int v = tid; v < V; v += BNT) {
        const Cand c{(row[v] - lse) + rs, j * V + v};
        if (before(prev, c) && before(c, best)) best = c;
      }
    }
    best = blk_best(best, sh_c);
    if (tid == 0) s_cand[k] = best;
    prev = best;
  }

  // ---- stage the state of the old beams ----------------------------------------------------------------------------
  const int len_old = st.cur_len[b];            // identical for every beam of the study
  const int gen0 = len_old - t + 1;             // cache slot of the first generated token
  for (int x = tid; x < nb * t; x += BNT) {
    const int j = x / t, c = x % t;
    const long long r = j * B + b;
    s_oldfin[j][c] = bs.fin_seq[(static_cast<long long>(b) * nb + j) * p.Tmax + c];
    s_oldrun[j][c] = st.seq[r * p.Lmax + p.P + c];
    s_oldkv[j][c] = st.key_valid[r * p.Lmax + gen0 + c];
  }
  if (tid < nb) {
    const int r = tid * B + b;
    s_oldlen[tid] = st.cur_len[r];
    s_oldnv[tid] = st.n_valid[r];
    s_oldseen[tid] = st.seen[r];
  }
  __syncthreads();

  // ---- C: bookkeeping (one thread; a handful of candidates) --------------------------------------------------------
  __shared__ float s_newrs[kMaxBeams], s_newfs[kMaxBeams];
  __shared__ int s_newfl[kMaxBeams];
  __shared__ uint8_t s_newif[kMaxBeams];
  if (tid == 0) {
    const bool hitmax = t + 1 >= p.Tmax;
    const float div = static_cast<float>(pow(static_cast<double>(t + 1), static_cast<double>(p.length_penalty)));
    bool hit[2 * kMaxBeams];
    float adj[2 * kMaxBeams];
    bool all_hit = true;
    for (int k = 0; k < K; ++k) {
      s_src[k] = s_cand[k].i / V;
      s_tok[k] = s_cand[k].i % V;
      hit[k] = s_tok[k] == p.eos || hitmax;
      all_hit = all_hit && hit[k];
      adj[k] = s_cand[k].v + (hit[k] ? NEG : 0.f);
    }
    // next running beams: the nb best by adjusted score (stable in k)
    bool used[2 * kMaxBeams] = {};
    for (int i = 0; i < nb; ++i) {
      int pick = -1;
      for (int k = 0; k < K; ++k)
        if (!used[k] && (pick < 0 || adj[k] > adj[pick])) pick = k;
      used[pick] = true;
      s_keep[i] = pick;
      s_newrs[i] = adj[pick];
    }
    // finished set: best nb of {old finished} U {top-nb continuations that hit}
    const bool can = bs.can_improve[b] != 0;
    float ms[3 * kMaxBeams];
    bool mf[3 * kMaxBeams];
    int ml[3 * kMaxBeams];
    for (int j = 0; j < nb; ++j) {
      ms[j] = bs.fin_score[b * nb + j];
      mf[j] = bs.is_fin[b * nb + j] != 0;
      ml[j] = bs.fin_len[b * nb + j];
    }
    for (int k = 0; k < K; ++k) {
      const bool just = hit[k] && k < nb;
      float s = s_cand[k].v / div;
      s = s + (can ? 0.f : NEG);
      s = s + (just ? 0.f : NEG);
      ms[nb + k] = s;
      mf[nb + k] = just;
      ml[nb + k] = t + 1;
    }
    bool usedm[3 * kMaxBeams] = {};
    float minfin = INFINITY;
    for (int i = 0; i < nb; ++i) {
      int pick = -1;
      for (int m = 0; m < nb + K; ++m)
        if (!usedm[m] && (pick < 0 || ms[m] > ms[pick])) pick = m;
      usedm[pick] = true;
      s_finsrc[i] = pick;
      s_newfs[i] = ms[pick];
      s_newif[i] = mf[pick] ? 1 : 0;
      s_newfl[i] = ml[pick];
      minfin = fminf(minfin, ms[pick]);
    }
    // early-stop heuristic (early_stopping False): can the best running beam still beat the worst finished one?
    const float best_running = s_newrs[0] / div;
    bool any = false;
    for (int i = 0; i < nb; ++i) any = any || (best_running > (s_newif[i] ? minfin : NEG));
    const bool can_new = can && any;
    bs.can_improve[b] = can_new ? 1 : 0;
    if (can_new) atomicAdd(bs.cnt_can, 1);
    if (all_hit) atomicAdd(bs.cnt_hit, 1);
  }
  __syncthreads();

  // ---- D: gather sequences, finished set and per-row state ----------------------------------------------------------
  for (int x = tid; x < nb * (t + 1); x += BNT) {
    const int i = x / (t + 1), c = x % (t + 1);
    const int k = s_keep[i];
    const long long r = i * B + b;
    st.seq[r * p.Lmax + p.P + c] = c < t ? s_oldrun[s_src[k]][c] : s_tok[k];
    if (c < t) st.key_valid[r * p.Lmax + gen0 + c] = s_oldkv[s_src[k]][c];
    const int m = s_finsrc[i];
    int fv;
    if (m < nb) fv = c < t ? s_oldfin[m][c] : fill;
    else fv = c < t ? s_oldrun[s_src[m - nb]][c] : s_tok[m - nb];
    bs.fin_seq[(static_cast<long long>(b) * nb + i) * p.Tmax + c] = fv;
  }
  if (tid < nb) {
    const int i = tid, k = s_keep[i], src = s_src[k], next = s_tok[k];
    const int r = i * B + b;
    bs.run_score[b * nb + i] = s_newrs[i];
    bs.fin_score[b * nb + i] = s_newfs[i];
    bs.fin_len[b * nb + i] = s_newfl[i];
    bs.is_fin[b * nb + i] = s_newif[i];
    bs.src_row[r] = src * B + b;
    // the decode state of the emitted token, exactly as sample_step_kernel advances it (decode.cu)
    const bool valid = p.mask_token_id < 0 || next != p.mask_token_id;
    const int cslot = s_oldlen[src] + 1;
    st.key_valid[static_cast<long long>(r) * p.Lmax + cslot] = valid ? 1 : 0;
    const int nv = s_oldnv[src] + (valid ? 1 : 0);
    st.n_valid[r] = nv;
    st.cur_pos[r] = max(nv - 1, 0);
    const unsigned seen = s_oldseen[src];
    int tt = p.sections[0][0];
    for (int q = 0; q < p.n_special[0]; ++q)
      if (seen & (1u << q)) tt = p.sections[0][q + 1];
    st.cur_type[r] = tt;
    unsigned add = 0;
    for (int q = 0; q < p.n_special[0]; ++q)
      if (next == p.special_ids[0][q]) add |= 1u << q;
    st.seen[r] = seen | add;
    st.cur_token[r] = next;
    st.cur_len[r] = cslot;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned prev_arr = atomicAdd(bs.arrive, 1u);
    if (prev_arr == static_cast<unsigned>(B) - 1) {
      __threadfence();
      const int can_cnt = *reinterpret_cast<volatile int*>(bs.cnt_can);
      const int hit_cnt = *reinterpret_cast<volatile int*>(bs.cnt_hit);
      *st.step = t + 1;
      *bs.n_slots = t;                             // generated cache slots that exist now (the new token is not cached yet)
      if (can_cnt == 0 || hit_cnt == B || t + 1 >= p.Tmax) *st.done = 1;
      *bs.cnt_can = 0;
      *bs.cnt_hit = 0;
      *bs.arrive = 0;
    }
  }
}

// rows whose source beam differs: scratch[l][kv][r][h][x] <- cache[l][kv][src_row[r]][h][gen0 + x], x < n_slots
template <typename T, bool SCATTER>
__global__ void beam_kv_move_kernel(T* kc, T* vc, T* scratch, const int* __restrict__ src_row, const int* __restrict__ cur_len,
                                    const int* __restrict__ n_slots_p, int R, int Lmax, int Tmax, int layers,
                                    long long layer_stride) {
  constexpr int VPR = HDB * sizeof(T) / 16;       // 16-byte vectors per (slot, head) row
  const int n = *n_slots_p;
  if (n <= 0) return;
  const long long total = static_cast<long long>(layers) * 2 * R * NHB * n * VPR;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % VPR);
    long long u = i / VPR;
    const int x = static_cast<int>(u % n); u /= n;
    const int h = static_cast<int>(u % NHB); u /= NHB;
    const int r = static_cast<int>(u % R); u /= R;
    const int kv = static_cast<int>(u % 2);
    const int l = static_cast<int>(u / 2);
    const int src = src_row[r];
    if (src == r) continue;
    const int gen0 = cur_len[r] - n;              // cur_len is the slot of the token emitted by this step
    T* cache = (kv ? vc : kc) + l * layer_stride;
    const long long so = ((((static_cast<long long>(l) * 2 + kv) * R + r) * NHB + h) * Tmax + x) * HDB;
    if (!SCATTER) {
      const long long co = ((static_cast<long long>(src) * NHB + h) * Lmax + gen0 + x) * HDB;
      reinterpret_cast<uint4*>(scratch + so)[q] = reinterpret_cast<const uint4*>(cache + co)[q];
    } else {
      const long long co = ((static_cast<long long>(r) * NHB + h) * Lmax + gen0 + x) * HDB;
      reinterpret_cast<uint4*>(cache + co)[q] = reinterpret_cast<const uint4*>(scratch + so)[q];
    }
  }
}

__global__ void beam_finalize_kernel(RolloutState st, BeamState bs, int B, int nb, int P, int Lmax, int Tmax, int* out_seq,
                                     float* out_score, int* out_len) {
  const int b = blockIdx.x;
  const int Lout = P + Tmax;
  for (int c = threadIdx.x; c < Lout; c += blockDim.x)
    out_seq[static_cast<long long>(b) * Lout + c] =
        c < P ? st.seq[static_cast<long long>(b) * Lmax + c] : bs.fin_seq[static_cast<long long>(b) * nb * Tmax + (c - P)];
  if (threadIdx.x == 0) {
    if (out_score) out_score[b] = bs.fin_score[b * nb];
    if (out_len) out_len[b] = bs.is_fin[b * nb] ? bs.fin_len[b * nb] : 0;
  }
}

}  // namespace

size_t beam_scratch_elems(int R, int Tmax, int layers) {
  return static_cast<size_t>(layers) * 2 * R * NHB * Tmax * HDB;
}

void beam_init(const BeamState& bs, int B, int nb, int Tmax, int fill, cudaStream_t stream) {
  CXRM_CHECK(nb >= 2 && nb <= kMaxBeams && Tmax <= kBeamMaxT, "beam search: 2..8 beams, at most 256 new tokens");
  beam_init_kernel<<<ceil_div(B * nb * 32, 256), 256, 0, stream>>>(bs, B, nb, Tmax, fill);
  check_launch("beam_init");
}

void beam_step(const RolloutState& st, const RolloutParams& p, const BeamState& bs, const float* logits, int ldl,
               cudaStream_t stream) {
  beam_step_kernel<<<p.R / p.beams, BNT, 0, stream>>>(st, p, bs, logits, ldl);
  check_launch("beam_step");
}

template <typename T>
void beam_reorder_kv(T* kcache, T* vcache, T* scratch, const RolloutState& st, const BeamState& bs, int R, int Lmax, int Tmax,
                     int layers, long long layer_stride, cudaStream_t stream) {
  const int grid = 148 * 8;
  beam_kv_move_kernel<T, false><<<grid, 256, 0, stream>>>(kcache, vcache, scratch, bs.src_row, st.cur_len, bs.n_slots, R, Lmax,
                                                         Tmax, layers, layer_stride);
  check_launch("beam_kv_gather");
  beam_kv_move_kernel<T, true><<<grid, 256, 0, stream>>>(kcache, vcache, scratch, bs.src_row, st.cur_len, bs.n_slots, R, Lmax,
                                                        Tmax, layers, layer_stride);
  check_launch("beam_kv_scatter");
}

void beam_finalize(const RolloutState& st, const BeamState& bs, int B, int nb, int P, int Lmax, int Tmax, int* out_seq,
                   float* out_score, int* out_len, cudaStream_t stream) {
  beam_finalize_kernel<<<B, 256, 0, stream>>>(st, bs, B, nb, P, Lmax, Tmax, out_seq, out_score, out_len);
  check_launch("beam_finalize");
}

template void beam_reorder_kv<float>(float*, float*, float*, const RolloutState&, const BeamState&, int, int, int, int, long long,
                                     cudaStream_t);
template void beam_reorder_kv<bf16>(bf16*, bf16*, bf16*, const RolloutState&, const BeamState&, int, int, int, int, long long,
                                    cudaStream_t);

}  // namespace cxrm
