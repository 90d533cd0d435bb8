/*
 * cxrm.h - C ABI of the B200-native CXRMate SCST rollout engine (libcxrm.so).
 *
 * This is the drop-in boundary for the one hot path named in BASELINE.json:
 * encode -> KV-cached rollout (greedy baseline + top-k multinomial sample) ->
 * CXR-BERT cosine reward.  The reference implements that path in Python on
 * top of Hugging Face transformers; it has no FFI of its own, so each entry
 * point below names the reference call it replaces (paths relative to the
 * reference repository; SP = site-packages/transformers 5.5.0).
 *
 * Conventions
 *  - every function returns 0 on success and a negative cxrm_status on
 *    failure; the message is available from cxrm_last_error(engine);
 *  - pointers marked "dev" are CUDA device pointers on the engine's device,
 *    row-major, densely packed unless a stride is named; "host" pointers are
 *    ordinary host memory;
 *  - `stream` is a cudaStream_t cast to void* (torch.cuda.current_stream().cuda_stream);
 *    calls are asynchronous on that stream unless stated otherwise;
 *  - one engine per process/GPU, not re-entrant; the engine owns all of its
 *    scratch memory (sized from cxrm_config at creation) and never takes
 *    ownership of caller buffers;
 *  - there is no CPU fallback: a missing/unsupported GPU is an error.
 */
#ifndef CXRM_H_
#define CXRM_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CXRM_API __attribute__((visibility("default")))
#else
#define CXRM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cxrm_engine cxrm_engine;

typedef enum cxrm_status {
  CXRM_OK = 0,
  CXRM_ERR_INVALID = -1,   /* bad argument / shape / state */
  CXRM_ERR_CUDA = -2,      /* CUDA runtime or driver failure */
  CXRM_ERR_WEIGHT = -3,    /* unknown, missing or mis-shaped weight tensor */
  CXRM_ERR_INTERNAL = -4
} cxrm_status;

typedef enum cxrm_dtype {
  CXRM_F32 = 0,            /* fp32 validation mode: strict fp32 FMA arithmetic, no tensor cores */
  CXRM_BF16 = 1            /* bf16 storage/operands, fp32 accumulation, tcgen05 tensor cores */
} cxrm_dtype;

typedef enum cxrm_mode {
  CXRM_GREEDY = 1,         /* argmax head (generate(num_beams=1, do_sample=False)) */
  CXRM_SAMPLE = 2,         /* top-k multinomial head (generate(do_sample=True, top_k=k)) */
  CXRM_BOTH = 3            /* rows [0,B) sample, rows [B,2B) greedy; one decode loop, shared cross-attention K/V */
} cxrm_mode;

typedef struct cxrm_config {
  int dtype;               /* cxrm_dtype */
  int image_h, image_w;    /* 384 x 384; any multiple of 16 */
  int max_studies;         /* B  */
  int max_images;          /* N per study (reference: max_images_per_study = 5) */
  int max_prompt;          /* P <= 256 */
  int max_new_tokens;      /* T <= 255 (decoder_max_len - 1) */
  int vocab;               /* 30000 */
  int cvt_depth[3];        /* {1, 4, 16} = microsoft/cvt-21-384-22k */
  int dec_layers;          /* 6 */
  int rwd_layers;          /* 12; 0 = no reward model */
  int rwd_vocab;           /* 30522 */
  int rwd_max_len;         /* 512 */
  int rwd_max_seqs;        /* sequences per cxrm_reward_embed call (>= 3 * B to batch sample+greedy+labels) */
  int enc_chunk;           /* most images encoded per pass (0 = default 64); the valid images split into equal passes */
  int use_tensor_cores;    /* bf16 only: 1 = tcgen05 GEMMs (default), 0 = SIMT debug path */
  int use_cuda_graph;      /* 1 = replay the decode step as a CUDA graph */
  int max_train_tokens;    /* rows x tokens of the largest cxrm_train_step batch; 0 = no training workspace (default) */
} cxrm_config;

/* Fills `cfg` with the named architecture (cxrmate, config 4 of BASELINE.json). */
CXRM_API void cxrm_default_config(cxrm_config* cfg);

CXRM_API int cxrm_create(const cxrm_config* cfg, int device, cxrm_engine** out);
CXRM_API void cxrm_destroy(cxrm_engine* e);
/* Message of the last failure on this engine (valid until the next call). e == NULL: creation error. */
CXRM_API const char* cxrm_last_error(const cxrm_engine* e);
/* Bytes of device memory the engine allocated for scratch + caches (informational). */
CXRM_API size_t cxrm_workspace_bytes(const cxrm_engine* e);
/* Number of kernels launched by this library since it was loaded (for bench.py's gpu_launches).  The counter is
 * process-wide: the contract is one engine per process / GPU. */
CXRM_API uint64_t cxrm_launch_count(const cxrm_engine* e);

/*
 * Weights.  `name` is a key of the reference model's state_dict (SURVEY.md
 * Appendix D; e.g. "encoder.cvt.encoder.stages.0.embedding.convolution_embeddings.projection.weight",
 * "decoder.bert.encoder.layer.3.attention.self.query.weight",
 * "decoder.bert.encoder.layer.3.attention.self.query.lora_A.weight"), or for
 * the reward model a key of CXR-BERT prefixed with "reward." (e.g.
 * "reward.bert.encoder.layer.0.attention.self.query.weight",
 * "reward.cls_projection_head.dense_to_hidden.weight").  Data is fp32;
 * `on_device` says where `data` lives.  The engine copies; the caller keeps
 * ownership.  cxrm_finalize_weights packs everything (BatchNorm folded to
 * scale/shift, LoRA merged: W + (alpha/r) B A, q|k|v concatenated, conv weights
 * re-laid out, cast to the engine dtype) and frees the fp32 staging copies.
 * Replaces: PreTrainedModel.load_state_dict / from_pretrained
 * (reference tools/stages.py:78-82, modelling_longitudinal.py:152-154).
 */
CXRM_API int cxrm_load_weight(cxrm_engine* e, const char* name, const float* data, const int64_t* shape, int ndim,
                     int on_device);
CXRM_API int cxrm_finalize_weights(cxrm_engine* e);

/*
 * Encoder.  Replaces MultiCvtWithProjectionHead.forward
 * (modelling_longitudinal.py:56-90; modelling_multi.py:53-87; single:
 * modelling_single.py:53-78 with N = 1).
 *   pixels     dev fp32 [B, N, 3, H, W]
 *   memory_out dev (engine dtype: fp32 or bf16) [B, N*T, 768], T = (H/16)*(W/16); may be NULL
 *   mask_out   dev uint8 [B, N*T]: 1 where pixels[b,n,0,0,0] != 0; may be NULL
 * Zero-padded images are not encoded: their memory rows are 0 and masked
 * (the reference encodes them and masks them; decoder outputs are identical).
 * The result is also kept inside the engine for cxrm_prefill_cross_kv(NULL, NULL).
 * Synchronises the stream once (the valid-image list is read back).
 */
CXRM_API int cxrm_encode(cxrm_engine* e, const float* pixels, int B, int N, void* memory_out, uint8_t* mask_out, void* stream);

/*
 * Cross-attention K/V of all decoder layers for the current studies, computed
 * once and shared by every following rollout / forward call.  Replaces the
 * first-step `key(encoder_hidden_states)`, `value(...)` of BertCrossAttention
 * (SP/models/bert/modeling_bert.py:247-262) that the reference recomputes in
 * each generate() call.
 *   memory dev (engine dtype) [B, S, 768] and mask dev uint8 [B, S] (NULL mask = all visible),
 *   or both NULL to use the engine's own cxrm_encode result.
 * Masked tokens are dropped from the cache (attention is permutation
 * invariant over keys).  Synchronises the stream once.
 */
CXRM_API int cxrm_prefill_cross_kv(cxrm_engine* e, const void* memory, const uint8_t* mask, int B, int S, void* stream);

typedef struct cxrm_rollout_args {
  int mode;                        /* cxrm_mode */
  int B, P;                        /* studies, prompt columns */
  const int32_t* prompt_ids;       /* dev [B, P] right-padded prompt (no auto-prepended BOS) */
  int mask_token_id;               /* keys equal to this id are masked and do not advance positions; -1 = none */
  /* token-type sections per head (modelling_longitudinal.py:280-282, scst/gen_prompt.py:209-215,282) */
  int n_special_sample;  int special_sample[8];  int sections_sample[9];
  int n_special_greedy;  int special_greedy[8];  int sections_greedy[9];
  int max_new_tokens;              /* T */
  int eos_token_id, pad_token_id;
  int top_k;                       /* 0 = no top-k filtering */
  float temperature;
  const float* exp_noise;          /* dev [T, B, V] Exp(1) draws (validation mode: torch exponential_) or NULL */
  uint64_t seed;                   /* Philox seed when exp_noise == NULL */
  /* outputs, all dev, rows R = B (single mode) or 2B (CXRM_BOTH: sample rows first) */
  int32_t* sequences;              /* [R, P + T] prompt + generated ids, PAD-filled */
  float* logprobs;                 /* [R, T] log-prob of each emitted token (0 where PAD); may be NULL */
  float* margins;                  /* [R, T] decision margin per step (diagnostic); may be NULL */
  int32_t* topk_idx;               /* [R, T, 64] surviving vocabulary ids of the sample rows; may be NULL */
  float* topk_val;                 /* [R, T, 64] their (temperature-scaled) scores; may be NULL */
  int32_t* topk_cnt;               /* [R, T] number of survivors; may be NULL */
  float* last_logits;              /* [R, V] fp32 logits of the last executed step (diagnostic); may be NULL */
  int32_t* steps_out;              /* host int: number of executed steps (= len(scores) in HF); synchronises if non-NULL */
} cxrm_rollout_args;

/*
 * The rollout.  Replaces GenerationMixin.generate/_sample
 * (SP/generation/utils.py:2658-2841) driven by the reference's
 * prepare_inputs_for_generation (modelling_longitudinal.py:251-295), as called
 * from scst/gen_prompt.py:206-224 (greedy) and :279-300 (sample).
 * Requires cxrm_prefill_cross_kv for the same B.
 */
CXRM_API int cxrm_rollout(cxrm_engine* e, const cxrm_rollout_args* a, void* stream);

/*
 * Beam search.  Replaces generate(num_beams = num_test_beams) of the reference's test_step
 * (modules/lightning_modules/longitudinal/gt_prompt.py:344-362, gen_prompt.py:184-200, single.py:552-562,
 * multi.py:265-275), i.e. HF GenerationMixin._beam_search (SP/generation/utils.py:3076-3385) with the default flags
 * (early_stopping False, num_return_sequences 1, no logits processors) driven by the reference's
 * prepare_inputs_for_generation.  The running beams are rows of the KV-cached rollout (the encoder K/V of a study is
 * shared by its beams, the generated part of the self-attention cache is reordered every step), so
 * B * num_beams <= max_studies of the engine.  Requires cxrm_prefill_cross_kv for the same B.
 */
typedef struct cxrm_beam_args {
  int B, P;                        /* studies, prompt columns */
  const int32_t* prompt_ids;       /* dev [B, P] right-padded prompt (no auto-prepended BOS) */
  int mask_token_id;               /* as cxrm_rollout_args; -1 = none */
  int n_special;  int special[8];  int sections[9];
  int num_beams;                   /* 2..8 */
  int max_new_tokens;              /* T <= 255 */
  int eos_token_id, pad_token_id;
  float length_penalty;            /* HF default 1.0: score = sum_logprob / generated_len ** length_penalty */
  /* outputs, all dev */
  int32_t* sequences;              /* [B, P + T] prompt + best finished hypothesis, filled with pad (eos when pad == 0, as HF) */
  float* scores;                   /* [B] its score (HF sequences_scores); may be NULL */
  int32_t* lengths;                /* [B] its generated length (tokens up to and including EOS); may be NULL */
  int32_t* steps_out;              /* host int: decode steps executed; synchronises if non-NULL */
} cxrm_beam_args;
CXRM_API int cxrm_rollout_beam(cxrm_engine* e, const cxrm_beam_args* a, void* stream);

/*
 * Image preprocessing on the GPU.  Replaces the reference's test_transforms
 * (modules/lightning_modules/single.py:248-262, multi.py:89-103: Resize(size) -> CenterCrop([size, size]) -> ToTensor ->
 * Normalize(mean, std)) applied to the decoded image of data/dicom_id.py:91-92 (`convert('RGB')`; a grey image,
 * channels == 1, fills the three output channels).  Pillow's antialiased bilinear resampler is reproduced bit for bit
 * (22-bit fixed point, uint8 intermediate).  img: uint8 [H, W, channels] with `row_pitch` bytes per row, on the device
 * (img_on_device != 0) or on the host (copied inside the call); out: dev fp32 [3, size, size], e.g. slot (b, n) of the
 * pixel tensor cxrm_encode takes.  JPEG decoding stays with the caller.
 */
CXRM_API int cxrm_preprocess_image(cxrm_engine* e, const uint8_t* img, int H, int W, int channels, long long row_pitch,
                                   int img_on_device, int size, const float* mean3, const float* std3, float* out,
                                   void* stream);

/*
 * Teacher-forced decoder forward.  Replaces
 * LongitudinalPromptMultiCXREncoderDecoderModel.forward with encoder_outputs
 * given (modelling_longitudinal.py:173-249).
 *   ids, token_type_ids, position_ids dev int32 [R, L]; key_mask dev uint8 [R, L] (decoder_attention_mask)
 *   row r attends the encoder K/V of study r % B
 *   logits_out dev fp32 [R, L, V], or [R, V] when last_only
 */
CXRM_API int cxrm_decoder_forward(cxrm_engine* e, const int32_t* ids, const int32_t* token_type_ids,
                         const int32_t* position_ids, const uint8_t* key_mask, int R, int L, int B, int last_only,
                         float* logits_out, void* stream);

/*
 * Reward model.  Replaces CXRBERTReward.reward
 * (tools/rewards/cxrbert.py:23-73) after tokenisation:
 *   ids dev int32 [n, L] (padding='longest'), lens dev int32 [n] -> emb_out dev fp32 [n, 128]
 *   (the projected [CLS] embedding, element [2] of the hub model's tuple);
 *   cxrm_cosine: out[i] = cosine_similarity(a[i], b[i]) (torch eps 1e-8).
 * cxrm_reward = embed(pred) , embed(label), cosine.
 */
CXRM_API int cxrm_reward_embed(cxrm_engine* e, const int32_t* ids, const int32_t* lens, int n, int L, float* emb_out,
                      void* stream);
CXRM_API int cxrm_cosine(cxrm_engine* e, const float* a, const float* b, int n, int dim, float* out, void* stream);
/* REINFORCE loss of the SCST step (reference scst/gen_prompt.py:331-366 `reinforce_loss`):
 *   loss_out[0] = mean_b( -sum_t logprobs[b * ld + t] * advantage[b] )
 * logprobs: dev fp32, the SAMPLE rows of cxrm_rollout's `logprobs` output (log-softmax of the top-k-masked scores at
 * the sampled id, 0 at PAD positions = nll_loss(..., ignore_index=pad)); advantage dev fp32 [B]; loss_out dev fp32 [1]. */
CXRM_API int cxrm_reinforce_loss(cxrm_engine* e, const float* logprobs, int ld, const float* advantage, int B, int T,
                        float* loss_out, void* stream);
CXRM_API int cxrm_reward(cxrm_engine* e, const int32_t* pred_ids, const int32_t* pred_lens, int L_pred,
                const int32_t* label_ids, const int32_t* label_lens, int L_label, int n, float* reward_out,
                void* stream);

/*
 * Whole SCST rollout step with HOST buffers (the end-to-end call bench.py
 * times): H2D pixels + prompts -> encode -> cross K/V -> CXRM_BOTH rollout ->
 * device-side id bridge (decoder id -> reward-model id via `id_map`, sections
 * split as split_and_decode_sections does: scst/gen_prompt.py:233-240,312-317)
 * -> CXR-BERT embeddings of sample, greedy and label reports -> cosine rewards,
 * advantage = sample - baseline (scst/gen_prompt.py:241) -> D2H.
 * The id bridge stands in for the CPU text round trip (BPE decode + WordPiece
 * encode), for which no vocabularies exist offline.
 *   pixels host fp32 [B,N,3,H,W]; prompt_ids host int32 [B,P]; label_ids host int32 [B,L_label];
 *   label_lens host int32 [B]; id_map dev int32 [vocab] (set once with cxrm_set_id_map)
 *   outputs host: sequences int32 [2B, P+T], logprobs fp32 [2B, T], reward/baseline/advantage fp32 [B]
 * Padding images (pixels[b, n, 0, 0, 0] == 0, modelling_longitudinal.py:83) are detected on the host and never
 * copied: only the valid images cross PCIe, compacted, one encoder chunk at a time on the engine's copy stream, so
 * chunk c + 1 is in flight while chunk c is encoded.  Pinned host memory makes those copies asynchronous; pageable
 * memory works but serialises them.  Synchronous: returns when the outputs are in the host buffers.
 */
/* n = decoder vocabulary; decoder ids < n_special are special tokens and are dropped (skip_special_tokens=True);
 * every id_map[i >= n_special], cls_id and sep_id must lie in [0, rwd_vocab): CXRM_ERR_INVALID otherwise. */
CXRM_API int cxrm_set_id_map(cxrm_engine* e, const int32_t* id_map_host, int n, int cls_id, int sep_id, int bos_id,
                    int sep_dec_id, int n_special);
/* The device-side text bridge on its own: generated rows [R, L] (prompt included or not) -> reward-model ids
 * out_ids dev int32 [R, L_out] = [CLS] map(findings) map(impression) [SEP], zero padded; out_lens dev int32 [R].
 * Sections as split_and_decode_sections (modelling_longitudinal.py:413-457) with specials [BOS, SEP, EOS]: section j is
 * ids[first col of special j-1 : first col of special j] (absent or at column 0 -> to the end); ids < n_special dropped. */
CXRM_API int cxrm_bridge_ids(cxrm_engine* e, const int32_t* sequences, int R, int L, int eos_token_id, int32_t* out_ids,
                    int32_t* out_lens, int L_out, void* stream);
CXRM_API int cxrm_scst_step_host(cxrm_engine* e, const float* pixels, int B, int N, const int32_t* prompt_ids, int P,
                        const cxrm_rollout_args* rollout_template, const int32_t* label_ids,
                        const int32_t* label_lens, int L_label, int32_t* sequences, float* logprobs, float* reward,
                        float* baseline, float* advantage, int32_t* steps_out, void* stream);

/* Same step with every input and output buffer already resident in device memory (bench.py's `value`). */
CXRM_API int cxrm_scst_step_device(cxrm_engine* e, const float* pixels, int B, int N, const int32_t* prompt_ids, int P,
                          const cxrm_rollout_args* rollout_template, const int32_t* label_ids,
                          const int32_t* label_lens, int L_label, int32_t* sequences, float* logprobs, float* reward,
                          float* baseline, float* advantage, int32_t* steps_out, void* stream);

/*
 * Teacher-forced decoder forward + backward: the training half of the SCST step and the plain teacher-forced step.
 * Replaces, for the decoder, `loss.backward()` of
 *   - longitudinal/gt_prompt.py:186-249 (cross-entropy on the radiologist report, loss_kind 0), and
 *   - scst/gen_prompt.py:331-366 (`reinforce_loss`, loss_kind 1): the reference backpropagates through the autograd
 *     graph of the 255 cached decode steps of the sampled rollout; here the sampled sequence is run ONCE teacher-forced
 *     (identical arithmetic in eval mode: SURVEY.md finding 9) and the gradient of
 *     mean_b( -sum_t log_softmax(top-k-masked scores)[b, t, id] * advantage[b] ) flows back through that pass.
 * Requires cxrm_prefill_cross_kv for the same studies (row r attends study r % B); the encoder is frozen.
 * Gradients are written, fp32, into ONE flat buffer (`grads`) laid out by cxrm_grad_info; with lora_only = 1 only the
 * LoRA A / B matrices of the self-attention query / key projections (modelling_longitudinal.py:163-170) get gradients,
 * otherwise every decoder parameter does (the LoRA update is then part of the merged weight).  Eval-mode arithmetic:
 * no dropout.
 *
 * Stages (for overlapping the gradient all-reduce with the backward pass): stage -1 runs everything;
 * otherwise call stages 0 .. cxrm_train_stages() - 1 in order with the same arguments: stage 0 = forward + loss + LM-head
 * backward, stage 1 + i = decoder layer (last - i), the last stage = embeddings.  When stage s returns, the slice of
 * `grads` belonging to the slots with stage == s is final.
 */
typedef struct cxrm_train_args {
  int R, L;                        /* rows, tokens per row (R * L a multiple of 8, <= max_train_tokens) */
  const int32_t* ids;              /* dev [R, L] decoder input ids */
  const int32_t* token_type_ids;   /* dev [R, L] */
  const int32_t* position_ids;     /* dev [R, L] */
  const uint8_t* key_mask;         /* dev [R, L] decoder_attention_mask */
  const int32_t* targets;          /* dev [R, L] id to predict at each position; ignore_index where nothing is counted */
  int ignore_index;
  int loss_kind;                   /* 0 cross-entropy (mean over counted targets), 1 REINFORCE */
  const float* advantage;          /* dev [R] (loss_kind 1) */
  int top_k;                       /* loss_kind 1: the sampling head's top-k (0 = none) */
  float temperature;               /* loss_kind 1 */
  int lora_only;
  float* loss_out;                 /* dev [1] */
  float* grads;                    /* dev fp32 [cxrm_grad_total(e, lora_only)] */
} cxrm_train_args;
CXRM_API int cxrm_train_step(cxrm_engine* e, const cxrm_train_args* a, int stage, void* stream);
CXRM_API int cxrm_train_stages(const cxrm_engine* e);
/* layout of the flat gradient buffer: slot i has the reference's state_dict name (e.g.
 * "decoder.bert.encoder.layer.3.attention.self.query.lora_A.weight"), an element offset, a size and the stage that
 * finishes it; slots are ordered by stage, so every stage owns one contiguous range. */
CXRM_API int cxrm_grad_count(const cxrm_engine* e, int lora_only);
CXRM_API int64_t cxrm_grad_total(const cxrm_engine* e, int lora_only);
CXRM_API int cxrm_grad_info(const cxrm_engine* e, int lora_only, int i, char* name, size_t name_len, int64_t* offset,
                   int64_t* numel, int* stage);

/* Device time (ms, CUDA events on the step's stream) of the phases of the LAST cxrm_scst_step_* call:
 * out5 = {encode, cross K/V, rollout, reward + result copies, prompt pass (prefill + first token; part of rollout)}. */
CXRM_API int cxrm_last_phase_ms(cxrm_engine* e, float* out5);

/*
 * Per-kernel-class timing: while enabled every kernel launch is bracketed by a CUDA event pair on its stream
 * (CUDA-graph replay is bypassed so each launch is visible).  cxrm_profile_report synchronises, writes a JSON
 * object {"<phase>.<kernel class>": {"ms": total, "n": launches}, ...} into buf and clears the records.
 */
CXRM_API int cxrm_set_profile(cxrm_engine* e, int on);
CXRM_API int cxrm_profile_report(cxrm_engine* e, char* buf, size_t len);

/* Standalone GEMM entry used by the kernel tests: C = A[M,K] . W[N,K]^T (+bias, act, +residual).
 * impl: 0 = SIMT fp32-FMA, 1 = tcgen05, 2 = tcgen05 skinny (M <= 64, deep TMA ring).  dtype: cxrm_dtype of A/W/C/residual. */
CXRM_API int cxrm_test_gemm(int impl, int dtype, const void* A, const void* W, void* C, int M, int N, int K,
                   const float* bias, int act, const void* residual, int out_f32, void* stream);
/* Decode-step pair used by the kernel tests: skinny split-K tcgen05 GEMM (M <= 64, bf16) into fp32 partials, then the
 * fused reduce + bias + act + residual + LayerNorm kernel.  partial_ws: dev fp32 [8 * 64 * N] (up to 8 K-splits). */
CXRM_API int cxrm_test_gemm_ln(const void* A, const void* W, void* out, int M, int N, int K, const float* bias, int act,
                      const void* residual, const float* gamma, const float* beta, float eps, float* partial_ws,
                      void* stream);
/* Micro-benchmark switch: launch the decode-chain kernels reached through the test hooks as programmatic dependent
 * launches (the engine sets this itself inside a decode step). */
CXRM_API void cxrm_test_set_pdl(int on);
/* Debug: when non-NULL, cxrm_test_gemm(impl 1) writes 8 %globaltimer stamps per CTA into dev_buf (phase timeline). */
CXRM_API void cxrm_test_set_gemm_trace(unsigned long long* dev_buf);
/* Standalone attention entry used by the kernel tests (q,k,v,o: [batch, L, heads*64] token-major). */
CXRM_API int cxrm_test_attention(int dtype, const void* q, const void* k, const void* v, void* o, int batch, int heads,
                        int Lq, int Lk, const uint8_t* key_mask, int causal, float scale, void* stream);
/* The same for a PACKED (ragged) self-attention batch, as the CXR-BERT reward uses it (engine.cu reward_embed): qkv dev
 * [total, 3 * heads*64] (q | k | v per token), sequence i = rows offsets[i] .. offsets[i] + lens[i] (dev int32 [n_seq],
 * lens <= Lmax), every key of a sequence visible to every query of it; o dev [total, heads*64]. */
CXRM_API int cxrm_test_attention_packed(int dtype, const void* qkv, void* o, int n_seq, int heads, int Lmax,
                                        const int32_t* offsets, const int32_t* lens, long long total, float scale,
                                        void* stream);

/* Sampling head on caller-provided logits dev fp32 [R, V]: every row is a sample row of a fresh rollout at decode step
 * `step` (< Tmax).  Philox4x32-10 contract of the in-kernel draw (exp_noise == NULL): stream = (seed, subsequence
 * step * R + row), the Exp(1) variate of vocabulary entry i is -log(curand_uniform) at offset i of that stream, and the
 * token is argmax_i softmax(top-k-masked scores)_i / q_i.  out_tokens dev int32 [R]; out_logprob dev fp32 [R] or NULL. */
CXRM_API int cxrm_test_sample(const float* logits, int R, int V, int top_k, float temperature, uint64_t seed, int step, int Tmax,
                     int32_t* out_tokens, float* out_logprob, void* stream);
/* LayerNorm of rows [rows, C] (eps as given), dtype = cxrm_dtype of x / y. */
CXRM_API int cxrm_test_layernorm(int dtype, const void* x, void* y, const float* gamma, const float* beta, long long rows,
                        int C, float eps, void* stream);
/* CvT attention front end (HF modeling_cvt.py:124-141,215-228,371-377): x [n_img, cls + H*W, C] ->
 * q [n_img, cls + H*W, C], k, v [n_img, cls + Hk*Wk, C] = BatchNorm_eval(dwconv3x3(LayerNorm(x))) with stride 1 / 2
 * (BatchNorm folded to scale/shift [3][C]; w [3][9][C] tap-major; stats: scratch of 2 floats per token). */
CXRM_API int cxrm_test_ln_dwconv(int dtype, const void* x, void* q, void* k, void* v, float* stats, const float* gamma,
                        const float* beta, float eps, const float* w, const float* scale, const float* shift,
                        int n_img, int H, int W, int C, int cls, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CXRM_H_ */
