// Host-side launchers of every CUDA kernel on the SCST rollout path.
// All launchers are asynchronous on `stream`; pointers are device pointers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace cxrm {

enum Act : int { ACT_NONE = 0, ACT_GELU = 1 };

// C[M,N] = epi(A[M,K] . W[N,K]^T): + bias[N] (fp32, nullable) -> act -> + residual[M,N] (T, nullable);
// stored as T, or as fp32 when out_f32.  K % 8 == 0, lda/ldw % 8 == 0.
struct GemmArgs {
  const void* A; int lda;
  const void* W; int ldw;
  void* C; int ldc;
  int M, N, K;
  const float* bias;
  int act;
  const void* residual; int ldr;
  int out_f32;
  const int* skip_flag;   // nullable device flag: kernel returns immediately when *skip_flag != 0
  // > 0: head-major store (K/V caches): element (m, n) goes to C[(n / 64) * c_head_stride + m * 64 + n % 64]
  long long c_head_stride;
  unsigned long long* trace;   // nullable debug buffer: 8 globaltimer stamps per CTA (gemm_tc_kernel only)
};
// column tile width the skinny kernel will use for g
int gemm_skinny_tile_n(const GemmArgs& g, bool split_allowed);

// strict-fp32 FMA path (validation mode; also the bf16-storage SIMT debug path)
template <typename T> void gemm_simt(const GemmArgs& g, cudaStream_t stream);
// tcgen05 / TMEM / TMA path (bf16 operands, fp32 accumulate)
void gemm_tcgen05(const GemmArgs& g, cudaStream_t stream);
// returns 0 when the tcgen05 path can take this shape
int gemm_tcgen05_supported(const GemmArgs& g);
// decode-step GEMMs (M <= 64): deep TMA ring, optional split-K into fp32 partials [nsplit][64][N]
// (partial == nullptr: direct fused epilogue).  *nsplit_out = splits written (0 = direct).
int gemm_skinny_supported(const GemmArgs& g);
void gemm_tcgen05_skinny(const GemmArgs& g, float* partial, int* nsplit_out, cudaStream_t stream);
size_t gemm_skinny_partial_floats(int N);
// out[M, N] (bf16) = LayerNorm(act(sum_s partial[s] + bias) + residual): consumer of the split-K partials
// res_gamma/res_beta (nullable): the residual is itself a LayerNorm output that was never stored: `residual` then holds
// the PRE-LayerNorm rows and LN(residual) * res_gamma + res_beta (same eps, rounded to bf16) is added instead.
void splitk_ln(const float* partial, int nsplit, int M, int N, const float* bias, int act, const void* residual,
               int ldr, const float* gamma, const float* beta, float eps, void* out, int ldo, const int* skip_flag,
               cudaStream_t stream, const float* res_gamma = nullptr, const float* res_beta = nullptr);

// ---- elementwise / normalisation (elementwise.cu) -------------------------
template <typename T>
void layernorm(const T* x, int ldx, T* y, int ldy, const float* gamma, const float* beta, long long rows, int C,
               float eps, cudaStream_t stream);

// pixels [*,3,H,W] fp32 NCHW; img_idx[n] selects the source image of output image n (nullable = identity).
// out [n_img*Ho*Wo, Kpad], K index = (cin*kh + ky)*kw + kx, zero padded to Kpad.
template <typename T>
void im2col_pixels(const float* pixels, const int* img_idx, T* out, int n_img, int H, int W, int ksz, int stride,
                   int pad, int Kpad, cudaStream_t stream);
// tokens [n_img, H*W, C] (token-major == NHWC); out [n_img*Ho*Wo, ksz*ksz*C], K index = (ky*ksz + kx)*C + c.
template <typename T>
void im2col_tokens(const T* in, T* out, int n_img, int H, int W, int C, int ksz, int stride, int pad,
                   cudaStream_t stream);
// CvT attention front end: LayerNorm (gamma/beta/eps) -> depth-wise 3x3 (pad 1) + folded BatchNorm for the q/k/v
// convolutional projections, fused.  x [n_img, cls+H*W, C] -> q [n_img, cls+H*W, C] (stride 1); k,v [n_img, cls+Hk*Wk, C]
// (stride 2).  w [3][9][C] fp32 (q,k,v; tap-major), scale/shift [3][C] fp32.  cls rows bypass the convolution.
// stats: scratch of 2 floats per token (mean, rstd).
template <typename T>
void ln_dwconv_qkv(const T* x, T* q, T* k, T* v, float* stats, const float* gamma, const float* beta, float eps,
                   const float* w, const float* scale, const float* shift, int n_img, int H, int W, int C, int cls,
                   cudaStream_t stream);
// x[n_img, 1+HW, C] <- cat(cls_token[C], tokens[n_img, HW, C])
template <typename T>
void cat_cls(const T* tokens, const float* cls_token, T* out, int n_img, int HW, int C, cudaStream_t stream);
// out[n_img, HW, C] <- in[n_img, 1+HW, C][:, 1:]
template <typename T>
void drop_cls(const T* in, T* out, int n_img, int HW, int C, cudaStream_t stream);
// dst rows <- src rows gathered: dst[i, :] = src[idx[i], :] (idx < 0 -> zeros)
template <typename T>
void gather_rows(const T* src, const int* idx, T* dst, long long n_rows, int C, cudaStream_t stream);
// dst[idx[i], :] = src[i, :]
template <typename T>
void scatter_rows(const T* src, const int* idx, T* dst, long long n_rows, int C, cudaStream_t stream);
template <typename TS, typename TD>
void cast_copy(const TS* src, TD* dst, long long n, cudaStream_t stream);
void fill_zero(void* p, size_t bytes, cudaStream_t stream);
// BERT embeddings: (word[id] + type[tt]) + pos[p] -> LayerNorm.  One row per token.
template <typename T>
void embed_ln(const int* ids, const int* types, const int* pos, const T* word, const T* type_emb, const T* pos_emb,
              const float* gamma, const float* beta, T* out, long long rows, int C, float eps, cudaStream_t stream);
// weight preparation
void bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
             float* shift, int C, cudaStream_t stream);
// W[out,in] += s * B[out,r] . A[r,in]
void lora_merge(float* W, const float* A, const float* B, int n_out, int n_in, int r, float s, cudaStream_t stream);
// dst[rows, cols_dst] (T) <- src[rows, cols_src] fp32 with zero padding of the extra columns
template <typename T>
void pack_matrix(const float* src, T* dst, int rows, int cols_src, int cols_dst, cudaStream_t stream);
// conv weight [Cout, Cin, k, k] fp32 -> [Cout, k, k, Cin] T
template <typename T>
void pack_conv_khwc(const float* src, T* dst, int Cout, int Cin, int ksz, cudaStream_t stream);
// dw weight [C,1,3,3] fp32 -> [9][C] fp32
void pack_dw(const float* src, float* dst, int C, cudaStream_t stream);

// ---- attention (attention.cu) ----------------------------------------------
struct AttnArgs {
  const void* q; const void* k; const void* v; void* o;
  long long q_bs, q_hs, q_ts;    // batch / head / token strides (elements); head dim is 64, unit stride
  long long k_bs, k_hs, k_ts;
  long long v_bs, v_hs, v_ts;
  long long o_bs, o_hs, o_ts;
  int batch, heads, Lq, Lk;
  const int* Lk_per_batch;       // nullable: per kv-batch key count (<= Lk)
  const uint8_t* key_mask;       // nullable: [kv_batch or batch][key_mask_ld], 1 = visible
  int key_mask_ld;
  int key_mask_per_q_batch;      // 1: key_mask indexed by q batch; 0: by kv batch
  int causal;                    // key j visible to query i iff j <= i + q_pos_offset
  int q_pos_offset;
  int kv_batch_mod;              // > 0: kv batch = q batch % kv_batch_mod
  const int* kv_offset;          // nullable: per kv-batch token offset added to the key index (ragged caches)
  float scale;
  // packed (ragged) queries, nullable: batch b's queries are tokens q_offset[b] .. + Lq_per_batch[b] of q / o (q_bs / o_bs
  // unused); Lq is then only the upper bound that sizes the grid
  const int* q_offset;
  const int* Lq_per_batch;
  // optional: rows that exist in the packed q / o and k / v buffers (lets the TMA-based kernel bound its tensor maps;
  // 0 = unknown, that kernel then declines packed calls)
  long long total_q, total_kv;
};
template <typename T> void attention_simt(const AttnArgs& a, cudaStream_t stream);
// bf16 tensor-core (mma.sync m16n8k16) flash attention with the same contract; returns 0 from _supported when usable
int attention_mma_supported(const AttnArgs& a);
void attention_mma(const AttnArgs& a, cudaStream_t stream);
// bf16 tcgen05 / TMEM / TMA flash attention for dense (unmasked, non-causal, unpacked) calls with >= 64 queries
// (attention_tc5.cu)
int attention_tc5_supported(const AttnArgs& a);
void attention_tc5(const AttnArgs& a, cudaStream_t stream);

}  // namespace cxrm

// ---- decode-step kernels and rollout state (decode.cu) -----------------------
namespace cxrm {

constexpr int kMaxSpecial = 8;
constexpr int kTopKCap = 64;     // slots per (row, step) for the surviving (index, score) pairs

// Per-row rollout state, all device memory owned by the engine.
struct RolloutState {
  int* cur_token;        // [R] token fed at the next decoder step
  int* cur_len;          // [R] cache slot of that token (= tokens already cached)
  int* cur_type;         // [R] its token-type id
  int* cur_pos;          // [R] its position id
  int* n_valid;          // [R] number of non-masked tokens so far (incl. the fed one)
  unsigned* seen;        // [R] bit i set: special_ids[i] occurred before the fed token
  uint8_t* finished;     // [R]
  uint8_t* key_valid;    // [R, Lmax] self-attention key validity (ids != mask_token_id)
  int* seq;              // [R, Lmax] prompt + generated ids
  float* logprob;        // [R, Tmax] log-prob of the emitted token under the (top-k-masked) distribution
  float* margin;         // [R, Tmax] decision margin (tie diagnosis)
  int* topk_idx;         // [R, Tmax, kTopKCap]
  float* topk_val;       // [R, Tmax, kTopKCap]
  int* topk_cnt;         // [R, Tmax] number of survivors (may exceed kTopKCap on ties)
  int* step;             // scalar: decode steps executed so far
  int* done;             // scalar: 1 once every row is finished or the budget is spent
  unsigned* arrive;      // scalar: block-arrival counter of the sampling kernel
  unsigned long long* seed;  // scalar: Philox seed of this rollout (device-resident so the step graph is reusable)
};

struct RolloutParams {
  int R, B;              // rows (= B * n_modes), studies
  int P;                 // prompt length (columns)
  int Lmax;              // row stride of seq / key_valid
  int Tmax;              // max new tokens
  int V;
  int n_special[2];      // per mode (0 = sample rows [0,B), 1 = greedy rows [B,2B) when both run)
  int special_ids[2][kMaxSpecial];
  int sections[2][kMaxSpecial + 1];
  int mode_of_block[2];  // mode (0 sample / 1 greedy) of row block 0 and 1
  int mask_token_id;     // -1: none (all keys valid)
  int eos, pad;
  int top_k;
  float temperature;
  unsigned long long seed;
  int want_margin;       // 1: also compute the decision margins (two more passes over the row; diagnostics)
  int beams;             // 0: greedy / sampling heads; nb >= 2: beam search, rows = beam * (R / nb) + study (beam.cu)
  float length_penalty;  // beam search: finished score = sum_logprob / generated_len ** length_penalty
};

// state from the prompt (reference modelling_longitudinal.py:274-282): types (full rule), positions,
// key validity, `seen` bitmask; copies the prompt into seq; also emits the [R,P] prefill inputs.
// pre_valid [R,P]: key mask of the prompt pass.  The self-attention cache is COMPACT: only visible prompt tokens are
// stored (token (r, c) at slot pre_pos[r, c]), st.key_valid / st.cur_len are in cache-slot coordinates.
void rollout_init(const RolloutState& st, const RolloutParams& p, const int* prompt_ids, int* pre_ids, int* pre_types,
                  int* pre_pos, uint8_t* pre_valid, cudaStream_t stream);

// logits [R,V] fp32 (row stride ldl) -> next token per row (greedy argmax or top-k multinomial),
// log-prob, survivors, and the state for the next decoder step.  exp_noise: nullable [Tmax, B, V].
void sample_step(const RolloutState& st, const RolloutParams& p, const float* logits, int ldl, const float* exp_noise,
                 cudaStream_t stream);

// ---- beam search (beam.cu): HF `_beam_search` bookkeeping on the device ------------------------------------------
constexpr int kMaxBeams = 8, kBeamMaxT = 256;
struct BeamState {
  float* run_score;      // [B, nb] accumulated log-prob of the running beams
  float* fin_score;      // [B, nb] length-penalised scores of the finished set, best first (-1e9: empty slot)
  int* fin_len;          // [B, nb] generated length of each finished hypothesis
  uint8_t* is_fin;       // [B, nb]
  uint8_t* can_improve;  // [B] early-stop heuristic still unsatisfied
  int* fin_seq;          // [B, nb, Tmax] generated tokens of the finished set
  int* src_row;          // [R] row whose self K/V (generated slots) this row continues from
  int* n_slots;          // scalar: generated cache slots to move in this step's reorder (0: none)
  unsigned* arrive;      // scalar
  int* cnt_can;          // scalar: studies that can still improve (this step)
  int* cnt_hit;          // scalar: studies whose every continuation hit a stopping criterion (this step)
};
size_t beam_scratch_elems(int R, int Tmax, int layers);
void beam_init(const BeamState& bs, int B, int nb, int Tmax, int fill, cudaStream_t stream);
// logits [R, V] fp32 of the running beams -> next running beams (st: tokens, sequences, decode state), finished set
void beam_step(const RolloutState& st, const RolloutParams& p, const BeamState& bs, const float* logits, int ldl,
               cudaStream_t stream);
// Cache.reorder_cache(beam_idx) for the generated slots of the self-attention caches [layers][R][12][Lmax][64]
template <typename T>
void beam_reorder_kv(T* kcache, T* vcache, T* scratch, const RolloutState& st, const BeamState& bs, int R, int Lmax, int Tmax,
                     int layers, long long layer_stride, cudaStream_t stream);
// best finished hypothesis per study -> out_seq [B, P + Tmax] (prompt + generated, fill = pad or eos), score, length
void beam_finalize(const RolloutState& st, const BeamState& bs, int B, int nb, int P, int Lmax, int Tmax, int* out_seq,
                   float* out_score, int* out_len, cudaStream_t stream);

// test hook: the sampling head on caller logits [R,V], all rows sample rows, Philox draws of (seed, step) -> tokens [R]
void sample_rows_test(const float* logits, int R, int V, int top_k, float temperature, unsigned long long seed, int step,
                      int Tmax, int* out_tokens, float* out_logprob, cudaStream_t stream);

// ---- one-token attention over the head-major K/V caches (decode_attn.cu) ----------------------------------
// Work units of the cross-attention: (study, chunk of <= CH encoder tokens); built on the host at
// cxrm_prefill_cross_kv, read by every decode step.  The grid is sized for max_units so that the captured
// CUDA graph does not depend on the batch's image counts.
struct CrossUnits {
  const int* study;      // [max_units]
  const int* j0;         // [max_units] first token of the chunk in the compact cache (kv_off[study] + chunk*CH)
  const int* n;          // [max_units] tokens in the chunk
  const int* chunk;      // [max_units] chunk index within its study
  const int* n_chunks;   // [B] chunks per study
  const int* first_unit; // [B] index of the study's first unit (units are ordered by study, then chunk)
  const int* n_units;    // scalar: live units
  int max_units, max_chunks;
};
// TMA descriptors of the bf16 caches (built once by the engine; nullptr in fp32 mode):
//   cross: rows = layers * 2 * 12 * tok_cap of 64 bf16, box 192 rows;  self_k / self_v: rows = layers * R * 12 * Lmax, box 64 rows
struct AttnMaps {
  CUtensorMap cross, self_k, self_v;
  int self_rows_per_layer;
};
// 2-D bf16 tensor map [rows, cols] (row pitch ld elements), box [box_cols, box_rows], 128-byte swizzle (gemm_tcgen05.cu)
CUtensorMap make_tensor_map_bf16(const void* ptr, long long rows, long long cols, long long ld, int box_rows, int box_cols);
// the same matrix as [cols / 64][rows][64]: box {64, box_rows, box_kblocks} = consecutive K-major swizzled k-block tiles
CUtensorMap make_tensor_map_bf16_kblocks(const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                                         int box_kblocks);
int decode_attn_chunk(size_t elem_size);                   // CH: keys per unit (192 bf16 / 96 fp32)
size_t decode_attn_ws_floats(int rows, int max_chunks);    // fp32 partials (max, sum, out[64]) per (row, head, chunk)

// self-attention of each row's new token over its cache [R][12][Lmax][64] (this layer); appends the new K/V.
// qkv [R, 3*768] (q | k | v); ctx [R,768]; ws / tickets: partials and per-(row, head) arrival counters (zeroed once).
template <typename T>
void decode_self_attention(const T* qkv, T* kcache, T* vcache, T* ctx, const RolloutState& st, int R, int P, int Lmax,
                           float* ws, unsigned* tickets, const AttnMaps* maps, int layer, cudaStream_t stream);

// cross-attention of the R/B rows of each study over that study's encoder K/V, kc/vc [12][tokens][64] of this
// layer (head_stride = tokens * 64).  q [R, ldq]; rows of study b: b, b+B.
template <typename T>
void decode_cross_attention(const T* q, int ldq, const T* kc, const T* vc, long long head_stride, T* ctx,
                            const CrossUnits& cu, const RolloutState& st, int R, int B, float* ws, unsigned* tickets,
                            const AttnMaps* maps, int layer, cudaStream_t stream);

// qkv [R*P, 3*768] -> head-major kcache/vcache [R][12][Lmax][64]: token (r, c) goes to slot[r*P + c] when valid[r*P + c]
// (masked prompt tokens are not cached)
// row_of (nullable): token t belongs to row row_of[t] instead of t / P (packed prompts: R = 1, P = tokens)
template <typename T>
void prefill_store_kv(const T* qkv, T* kcache, T* vcache, const int* slot, const uint8_t* valid, int R, int P, int Lmax,
                      cudaStream_t stream, const int* row_of = nullptr);

// Packed prompt pass: the visible tokens of every row, in order, plus - where the last prompt column is masked - that
// column as a query-only token (the reference takes the first new token's logits from the LAST column of the padded
// prompt, modelling_longitudinal.py:251-295: a PAD-token query at the position of the last visible token that attends
// the visible keys only).  pack_prompt_count: per-row counts, offsets and the total (one block); pack_prompt_fill: the
// packed ids / types / positions and, per packed token, its row, cache slot and whether it is cached.
struct PromptPack {
  int* ids; int* types; int* pos;      // [<= R * P] packed decoder inputs
  int* tok_row; int* tok_slot;         // [<= R * P] row and self-attention cache slot of each packed token
  uint8_t* tok_cache;                  // [<= R * P] 1: a visible token (cached), 0: the query-only last column
  int* row_off; int* row_lq; int* row_lk; int* last_idx;   // [R] offset, queries, keys, packed index of the row's last query
  int* total;                          // scalar: packed tokens
};
void pack_prompt(const PromptPack& pk, const int* pre_ids, const int* pre_types, const int* pre_pos, const uint8_t* pre_valid,
                 int R, int P, cudaStream_t stream);


// ---- persistent GEMM / LayerNorm chain of the decode step (decode_chain.cu) ---------------------------------------
enum ChainPhaseType : int { CH_GEMM = 0, CH_LN = 1, CH_EMBED = 2 };
enum ChainEpi : int { CE_PARTIAL = 0, CE_BF16 = 1, CE_BF16_GELU = 2 };
// One phase of a chain launch; a launch's list (<= 7 phases) travels in the kernel's parameter block.
struct alignas(128) ChainPhase {
  CUtensorMap tmA;           // GEMM: activations [R, K_total] bf16 as k-blocks: box {64, 64 rows, kslice / 64 k-blocks}
  CUtensorMap tmB;           // GEMM: weights [N, K_total] bf16, box bn rows x 64 columns
  int type;                  // ChainPhaseType
  int n_tiles, nsplit, bn;   // GEMM: column tiles, K splits, tile width (16 | 32 | 64); n_tiles * nsplit <= CTAs
  int kslice;                // GEMM: K elements per split (multiple of 64, <= 768; bn * kslice * 2 <= 48 KiB)
  int epi;                   // ChainEpi
  int N;                     // GEMM: output columns (row pitch of the partials)
  int ldo;                   // row pitch of `out` (bf16 elements)
  int act;                   // LN: activation applied to partial sum + bias (ACT_GELU: LM-head transform)
  int round_pre;             // LN: round the pre-LayerNorm sum to bf16 first (what bf16 autocast does)
  float eps;
  const float* bias;         // GEMM epilogue (CE_BF16*) or LN
  void* out;                 // bf16 output (GEMM CE_BF16*, LN, EMBED)
  float* partial;            // GEMM CE_PARTIAL: destination [nsplit][64][N]; LN: source (N == 768)
  const void* residual;      // LN: bf16 [R, 768] or nullptr
  const float* residual_f32; // LN: fp32 residual stream (takes precedence over `residual`)
  float* out_f32;            // LN / EMBED: optional fp32 copy of the output (residual stream)
  const float* gamma;
  const float* beta;
  const bf16* word;          // EMBED tables
  const bf16* type_emb;
  const bf16* pos_emb;
};
bool decode_chain_available();   // false: CXRM_NO_CHAIN set, or the device cannot hold one CTA per phase item
int decode_chain_ctas();
// one launch interpreting phases[0 .. count): bar = kChainBarWords zeroed unsigned counters owned by the caller
// trace (nullable, debug): kChainTraceSlots %globaltimer stamps per CTA: [0] entry, [1] after the dependency wait,
// [2 + 8i] start of phase i, [+1..+6] GEMM sub-stamps (decode_chain.cu GTRACE), [+7] end of phase i
constexpr int kChainTraceSlots = 64;
constexpr int kChainBarWords = 64 + 32 * 16;   // grid-barrier counters of decode_chain (zeroed once by the caller)
constexpr int kChainMaxSplit = 12;
void decode_chain(const ChainPhase* host_phases, int count, int R, const RolloutState& st, unsigned* bar, int n_ctas,
                  cudaStream_t stream, unsigned long long* trace = nullptr);

// copy rows [r, P-1] of x [R*P, C] into out [R, C]
template <typename T>
void take_last_token(const T* x, T* out, int R, int P, int C, cudaStream_t stream);

// loss[0] = mean_b( -sum_t logprob[b * ld + t] * advantage[b] )  (reinforce_loss, scst/gen_prompt.py:350-364)
void reinforce_loss(const float* logprob, int ld, const float* advantage, int B, int T, float* loss, cudaStream_t stream);

// ---- backward-pass kernels of the teacher-forced step (train.cu) ----------------------------------------------------
template <typename T>
void transpose(const T* in, long long ld_in, T* out, long long ld_out, long long rows, int cols, cudaStream_t stream);
// out[n] (+)= sum_m x[m, n]
template <typename T>
void colsum(const T* x, long long ldx, long long rows, int cols, float* out, bool accumulate, cudaStream_t stream);
// dx of y = LayerNorm(x) (stats: scratch [rows] (mean, rstd)); dgamma / dbeta (nullable) fp32, optionally accumulated
template <typename T>
void layernorm_bwd(const T* x, const T* dy, const float* gamma, float eps, T* dx, float2* stats, float* dgamma, float* dbeta,
                   bool accumulate, long long rows, int C, cudaStream_t stream);
template <typename T> void gelu_fwd(const T* x, T* y, long long n, cudaStream_t stream);
template <typename T> void gelu_bwd(const T* x, const T* dy, T* dx, long long n, cudaStream_t stream);
template <typename T> void add_inplace(T* a, const T* b, long long n, cudaStream_t stream);
void scale_f32(float* a, float s, long long n, cudaStream_t stream);
// logits [rows, V] fp32 -> loss (sum of row losses) and dz [rows, V] (T).  kind 0: cross-entropy, mean over targets !=
// ignore_index; kind 1: REINFORCE, -adv[row / L] / R * log_softmax(top-k-masked z / temperature)[target].
template <typename T>
void loss_head(const float* logits, long long rows, int V, const int* targets, int ignore_index, int kind, const float* adv,
               int L, int R, int top_k, float temperature, T* dz, float* row_loss, int* n_counted, float* loss_out,
               cudaStream_t stream);
// backward of the attention described by f (f.o = the forward output); dQ is laid out like q; dK == nullptr: dQ only,
// else element (kv batch b', head h, key j) of dK / dV sits at b' * g_bs + h * g_hs + (kv_offset[b'] + j) * g_ts.
// lse / D: scratch [batch, heads, Lq]
template <typename T>
void attention_bwd(const AttnArgs& f, const void* dO, void* dQ, void* dK, void* dV, long long g_bs, long long g_hs,
                   long long g_ts, float* lse, float* D, cudaStream_t stream);
// table[idx[r], :] += dx[r, :]
template <typename T>
void scatter_add_rows(const T* dx, const int* idx, float* table, long long rows, int C, cudaStream_t stream);

template <typename T>
void embed_sum(const int* ids, const int* types, const int* pos, const T* word, const T* type_emb, const T* pos_emb, T* out,
               long long rows, int C, cudaStream_t stream);

// Resize(size) + CenterCrop(size) + ToTensor + Normalize of one uint8 image [H, W, channels] (device, row pitch in
// bytes) -> out [3, size, size] fp32, bit-exact with the reference's torchvision / Pillow pipeline (preprocess.cu)
void preprocess_image(const uint8_t* img_dev, int H, int W, int channels, long long pitch, int size, const float* mean,
                      const float* stdv, float* out, cudaStream_t stream);

// cosine similarity of rows: out[i] = <a_i, b_i> / (max(|a_i|, eps) * max(|b_i|, eps))  (torch eps 1e-8)
void cosine_rows(const float* a, const float* b, float* out, int n, int C, cudaStream_t stream);

}  // namespace cxrm
