// extern "C" boundary (include/cxrm.h): exception -> status translation only.
#include <cstring>
#include <memory>
#include <string>

#include "engine.cuh"

using namespace cxrm;

struct cxrm_engine {
  EngineBase* impl = nullptr;
};

static thread_local std::string g_create_error;

#define CXRM_GUARD(e, body)                                   \
  if (!(e) || !(e)->impl) return CXRM_ERR_INVALID;            \
  try {                                                       \
    body;                                                     \
    return CXRM_OK;                                           \
  } catch (const cxrm::WeightError& ex) {                     \
    (e)->impl->last_error = ex.what();                        \
    return CXRM_ERR_WEIGHT;                                   \
  } catch (const cxrm::CudaError& ex) {                       \
    (e)->impl->last_error = ex.what();                        \
    return CXRM_ERR_CUDA;                                     \
  } catch (const std::exception& ex) {                        \
    (e)->impl->last_error = ex.what();                        \
    return CXRM_ERR_INVALID;                                  \
  } catch (...) {                                             \
    (e)->impl->last_error = "unknown exception";              \
    return CXRM_ERR_INTERNAL;                                 \
  }

extern "C" {

void cxrm_default_config(cxrm_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->dtype = CXRM_BF16;
  c->image_h = c->image_w = 384;
  c->max_studies = 32;
  c->max_images = 5;
  c->max_prompt = 256;
  c->max_new_tokens = 255;
  c->vocab = 30000;
  c->cvt_depth[0] = 1; c->cvt_depth[1] = 4; c->cvt_depth[2] = 16;
  c->dec_layers = 6;
  c->rwd_layers = 12;
  c->rwd_vocab = 30522;
  c->rwd_max_len = 512;
  c->rwd_max_seqs = 96;
  c->enc_chunk = 64;
  c->use_tensor_cores = 1;
  c->use_cuda_graph = 1;
  c->max_train_tokens = 0;
}

int cxrm_create(const cxrm_config* cfg, int device, cxrm_engine** out) {
  if (!cfg || !out) return CXRM_ERR_INVALID;
  *out = nullptr;
  try {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
      throw std::runtime_error("no CUDA device: the engine has no CPU fallback");
    std::unique_ptr<cxrm_engine> e(new cxrm_engine());
    e->impl = make_engine(*cfg, device);     // throws: the wrapper is released by the unique_ptr
    *out = e.release();
    return CXRM_OK;
  } catch (const cxrm::CudaError& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_INVALID;
  } catch (...) {
    g_create_error = "unknown exception";
    return CXRM_ERR_INTERNAL;
  }
}

void cxrm_destroy(cxrm_engine* e) {
  if (!e) return;
  delete e->impl;
  delete e;
}

const char* cxrm_last_error(const cxrm_engine* e) {
  if (!e || !e->impl) return g_create_error.c_str();
  return e->impl->last_error.c_str();
}

size_t cxrm_workspace_bytes(const cxrm_engine* e) { return (e && e->impl) ? e->impl->workspace_bytes() : 0; }
uint64_t cxrm_launch_count(const cxrm_engine* e) { (void)e; return g_launch_count; }

int cxrm_load_weight(cxrm_engine* e, const char* name, const float* data, const int64_t* shape, int ndim,
                     int on_device) {
  CXRM_GUARD(e, e->impl->load_weight(name, data, shape, ndim, on_device != 0));
}
int cxrm_finalize_weights(cxrm_engine* e) { CXRM_GUARD(e, e->impl->finalize_weights()); }

int cxrm_encode(cxrm_engine* e, const float* pixels, int B, int N, void* memory_out, uint8_t* mask_out, void* stream) {
  CXRM_GUARD(e, e->impl->encode(pixels, B, N, memory_out, mask_out, static_cast<cudaStream_t>(stream)));
}
int cxrm_prefill_cross_kv(cxrm_engine* e, const void* memory, const uint8_t* mask, int B, int S, void* stream) {
  CXRM_GUARD(e, e->impl->prefill_cross_kv(memory, mask, B, S, static_cast<cudaStream_t>(stream)));
}
int cxrm_rollout(cxrm_engine* e, const cxrm_rollout_args* a, void* stream) {
  if (!a) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->rollout(*a, static_cast<cudaStream_t>(stream)));
}
int cxrm_preprocess_image(cxrm_engine* e, const uint8_t* img, int H, int W, int channels, long long row_pitch, int img_on_device,
                          int size, const float* mean3, const float* std3, float* out, void* stream) {
  CXRM_GUARD(e, {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!img || !mean3 || !std3 || H < 1 || W < 1) throw std::runtime_error("cxrm_preprocess_image: bad arguments");
    const uint8_t* dev = img;
    uint8_t* staged = nullptr;
    if (!img_on_device) {
      const size_t bytes = static_cast<size_t>(H) * row_pitch;
      CXRM_CUDA_CHECK(cudaMallocAsync(&staged, bytes, s));
      CXRM_CUDA_CHECK(cudaMemcpyAsync(staged, img, bytes, cudaMemcpyHostToDevice, s));
      dev = staged;
    }
    try {
      preprocess_image(dev, H, W, channels, row_pitch, size, mean3, std3, out, s);
    } catch (...) {
      if (staged) cudaFreeAsync(staged, s);
      throw;
    }
    if (staged) CXRM_CUDA_CHECK(cudaFreeAsync(staged, s));
  });
}
int cxrm_rollout_beam(cxrm_engine* e, const cxrm_beam_args* a, void* stream) {
  if (!a) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->rollout_beam(*a, static_cast<cudaStream_t>(stream)));
}
int cxrm_decoder_forward(cxrm_engine* e, const int32_t* ids, const int32_t* tt, const int32_t* pos,
                         const uint8_t* key_mask, int R, int L, int B, int last_only, float* logits_out, void* stream) {
  CXRM_GUARD(e, e->impl->decoder_forward(ids, tt, pos, key_mask, R, L, B, last_only != 0, logits_out,
                                         static_cast<cudaStream_t>(stream)));
}
int cxrm_reward_embed(cxrm_engine* e, const int32_t* ids, const int32_t* lens, int n, int L, float* emb_out,
                      void* stream) {
  CXRM_GUARD(e, e->impl->reward_embed(ids, lens, n, L, emb_out, static_cast<cudaStream_t>(stream)));
}
int cxrm_reinforce_loss(cxrm_engine* e, const float* logprobs, int ld, const float* advantage, int B, int T, float* loss_out,
                        void* stream) {
  CXRM_GUARD(e, reinforce_loss(logprobs, ld, advantage, B, T, loss_out, static_cast<cudaStream_t>(stream)));
}

int cxrm_cosine(cxrm_engine* e, const float* a, const float* b, int n, int dim, float* out, void* stream) {
  CXRM_GUARD(e, cosine_rows(a, b, out, n, dim, static_cast<cudaStream_t>(stream)));
}
int cxrm_reward(cxrm_engine* e, const int32_t* pred_ids, const int32_t* pred_lens, int L_pred,
                const int32_t* label_ids, const int32_t* label_lens, int L_label, int n, float* reward_out,
                void* stream) {
  CXRM_GUARD(e, {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* emb = nullptr;
    CXRM_CUDA_CHECK(cudaMallocAsync(&emb, static_cast<size_t>(2) * n * 128 * sizeof(float), s));
    try {
      e->impl->reward_embed(pred_ids, pred_lens, n, L_pred, emb, s);
      e->impl->reward_embed(label_ids, label_lens, n, L_label, emb + static_cast<size_t>(n) * 128, s);
      cosine_rows(emb, emb + static_cast<size_t>(n) * 128, reward_out, n, 128, s);
    } catch (...) {
      cudaFreeAsync(emb, s);
      throw;
    }
    CXRM_CUDA_CHECK(cudaFreeAsync(emb, s));
  });
}
int cxrm_set_id_map(cxrm_engine* e, const int32_t* id_map_host, int n, int cls_id, int sep_id, int bos_id,
                    int sep_dec_id, int n_special) {
  CXRM_GUARD(e, e->impl->set_id_map(id_map_host, n, cls_id, sep_id, bos_id, sep_dec_id, n_special));
}
int cxrm_bridge_ids(cxrm_engine* e, const int32_t* sequences, int R, int L, int eos_token_id, int32_t* out_ids,
                    int32_t* out_lens, int L_out, void* stream) {
  CXRM_GUARD(e, e->impl->bridge_ids(sequences, R, L, eos_token_id, out_ids, out_lens, L_out, static_cast<cudaStream_t>(stream)));
}
int cxrm_test_sample(const float* logits, int R, int V, int top_k, float temperature, uint64_t seed, int step, int Tmax,
                     int32_t* out_tokens, float* out_logprob, void* stream) {
  try {
    sample_rows_test(logits, R, V, top_k, temperature, seed, step, Tmax, out_tokens, out_logprob, static_cast<cudaStream_t>(stream));
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_INVALID;
  }
}
int cxrm_scst_step_host(cxrm_engine* e, const float* pixels, int B, int N, const int32_t* prompt_ids, int P,
                        const cxrm_rollout_args* tmpl, const int32_t* label_ids, const int32_t* label_lens,
                        int L_label, int32_t* sequences, float* logprobs, float* reward, float* baseline,
                        float* advantage, int32_t* steps_out, void* stream) {
  if (!tmpl) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->scst_step_host(pixels, B, N, prompt_ids, P, *tmpl, label_ids, label_lens, L_label, sequences,
                                        logprobs, reward, baseline, advantage, steps_out, /*on_device=*/false,
                                        static_cast<cudaStream_t>(stream)));
}
int cxrm_scst_step_device(cxrm_engine* e, const float* pixels, int B, int N, const int32_t* prompt_ids, int P,
                          const cxrm_rollout_args* tmpl, const int32_t* label_ids, const int32_t* label_lens,
                          int L_label, int32_t* sequences, float* logprobs, float* reward, float* baseline,
                          float* advantage, int32_t* steps_out, void* stream) {
  if (!tmpl) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->scst_step_host(pixels, B, N, prompt_ids, P, *tmpl, label_ids, label_lens, L_label, sequences,
                                        logprobs, reward, baseline, advantage, steps_out, /*on_device=*/true,
                                        static_cast<cudaStream_t>(stream)));
}
int cxrm_train_step(cxrm_engine* e, const cxrm_train_args* a, int stage, void* stream) {
  if (!a) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->train_step(*a, stage, static_cast<cudaStream_t>(stream)));
}
int cxrm_train_stages(const cxrm_engine* e) { return (e && e->impl) ? e->impl->train_stages() : 0; }
int cxrm_grad_count(const cxrm_engine* e, int lora_only) { return (e && e->impl) ? e->impl->grad_count(lora_only != 0) : 0; }
int64_t cxrm_grad_total(const cxrm_engine* e, int lora_only) { return (e && e->impl) ? e->impl->grad_total(lora_only != 0) : 0; }
int cxrm_grad_info(const cxrm_engine* e, int lora_only, int i, char* name, size_t name_len, int64_t* offset, int64_t* numel,
                   int* stage) {
  if (!e || !e->impl || !name || name_len == 0) return CXRM_ERR_INVALID;
  std::string n;
  long long off = 0, ne = 0;
  int st = 0;
  if (!e->impl->grad_info(lora_only != 0, i, &n, &off, &ne, &st)) return CXRM_ERR_INVALID;
  std::strncpy(name, n.c_str(), name_len - 1);
  name[name_len - 1] = 0;
  if (offset) *offset = off;
  if (numel) *numel = ne;
  if (stage) *stage = st;
  return CXRM_OK;
}
int cxrm_last_phase_ms(cxrm_engine* e, float* out5) {
  if (!out5) return CXRM_ERR_INVALID;
  CXRM_GUARD(e, e->impl->last_phase_ms(out5));
}
int cxrm_set_profile(cxrm_engine* e, int on) { CXRM_GUARD(e, e->impl->set_profile(on != 0)); }
int cxrm_profile_report(cxrm_engine* e, char* buf, size_t len) {
  CXRM_GUARD(e, {
    const std::string r = e->impl->profile_report();
    if (!buf || len == 0) throw std::runtime_error("buffer");
    std::strncpy(buf, r.c_str(), len - 1);
    buf[len - 1] = 0;
  });
}

static unsigned long long* g_trace_buf = nullptr;
void cxrm_test_set_gemm_trace(unsigned long long* dev_buf) { g_trace_buf = dev_buf; }

int cxrm_test_gemm(int impl, int dtype, const void* A, const void* W, void* C, int M, int N, int K, const float* bias,
                   int act, const void* residual, int out_f32, void* stream) {
  try {
    GemmArgs g;
    g.A = A; g.lda = K; g.W = W; g.ldw = K; g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.act = act; g.residual = residual; g.ldr = N; g.out_f32 = out_f32; g.skip_flag = nullptr; g.c_head_stride = 0; g.trace = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (impl == 1) {
      if (dtype != CXRM_BF16 || gemm_tcgen05_supported(g) != 0) return CXRM_ERR_INVALID;
      g.trace = g_trace_buf;
      gemm_tcgen05(g, s);
    } else if (impl == 2) {
      if (dtype != CXRM_BF16 || gemm_skinny_supported(g) != 0) return CXRM_ERR_INVALID;
      gemm_tcgen05_skinny(g, nullptr, nullptr, s);
    } else if (dtype == CXRM_F32) {
      gemm_simt<float>(g, s);
    } else {
      gemm_simt<bf16>(g, s);
    }
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

void cxrm_test_set_pdl(int on) { g_pdl = on != 0; }


int cxrm_test_gemm_ln(const void* A, const void* W, void* out, int M, int N, int K, const float* bias, int act,
                      const void* residual, const float* gamma, const float* beta, float eps, float* partial_ws,
                      void* stream) {
  try {
    GemmArgs g;
    g.A = A; g.lda = K; g.W = W; g.ldw = K; g.C = nullptr; g.ldc = 0; g.M = M; g.N = N; g.K = K;
    g.bias = nullptr; g.act = 0; g.residual = nullptr; g.ldr = 0; g.out_f32 = 0; g.skip_flag = nullptr; g.c_head_stride = 0; g.trace = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (gemm_skinny_supported(g) != 0 || N > 1024 || N % 4 != 0) return CXRM_ERR_INVALID;
    if (partial_ws == nullptr) return CXRM_ERR_INVALID;
    int nsplit = 0;
    gemm_tcgen05_skinny(g, partial_ws, &nsplit, s);
    splitk_ln(partial_ws, nsplit, M, N, bias, act, residual, N, gamma, beta, eps, out, N, nullptr, s);
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

int cxrm_test_attention(int dtype, const void* q, const void* k, const void* v, void* o, int batch, int heads, int Lq,
                        int Lk, const uint8_t* key_mask, int causal, float scale, void* stream) {
  try {
    AttnArgs a{};
    const long long C = static_cast<long long>(heads) * 64;
    a.q = q; a.k = k; a.v = v; a.o = o;
    a.q_bs = Lq * C; a.q_hs = 64; a.q_ts = C;
    a.k_bs = Lk * C; a.k_hs = 64; a.k_ts = C;
    a.v_bs = Lk * C; a.v_hs = 64; a.v_ts = C;
    a.o_bs = Lq * C; a.o_hs = 64; a.o_ts = C;
    a.batch = batch; a.heads = heads; a.Lq = Lq; a.Lk = Lk;
    a.key_mask = key_mask; a.key_mask_ld = Lk; a.key_mask_per_q_batch = 1;
    a.causal = causal; a.q_pos_offset = Lk - Lq; a.scale = scale;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CXRM_F32)
      attention_simt<float>(a, s);
    else if (attention_tc5_supported(a) == 0)
      attention_tc5(a, s);
    else if (attention_mma_supported(a) == 0)
      attention_mma(a, s);
    else
      attention_simt<bf16>(a, s);
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

int cxrm_test_attention_packed(int dtype, const void* qkv, void* o, int n_seq, int heads, int Lmax, const int32_t* offsets,
                               const int32_t* lens, long long total, float scale, void* stream) {
  try {
    AttnArgs a{};
    const long long C = static_cast<long long>(heads) * 64;
    const size_t esz = dtype == CXRM_F32 ? 4 : 2;
    const char* base = static_cast<const char*>(qkv);
    a.q = base; a.k = base + C * esz; a.v = base + 2 * C * esz; a.o = o;
    a.q_hs = a.k_hs = a.v_hs = a.o_hs = 64;
    a.q_ts = a.k_ts = a.v_ts = 3 * C; a.o_ts = C;
    a.batch = n_seq; a.heads = heads; a.Lq = Lmax; a.Lk = Lmax;
    a.q_offset = offsets; a.Lq_per_batch = lens; a.kv_offset = offsets; a.Lk_per_batch = lens;
    a.total_q = a.total_kv = total;
    a.scale = scale;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CXRM_F32)
      attention_simt<float>(a, s);
    else if (attention_tc5_supported(a) == 0)
      attention_tc5(a, s);
    else if (attention_mma_supported(a) == 0)
      attention_mma(a, s);
    else
      attention_simt<bf16>(a, s);
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

int cxrm_test_layernorm(int dtype, const void* x, void* y, const float* gamma, const float* beta, long long rows, int C,
                        float eps, void* stream) {
  try {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CXRM_F32)
      layernorm<float>(static_cast<const float*>(x), C, static_cast<float*>(y), C, gamma, beta, rows, C, eps, s);
    else
      layernorm<bf16>(static_cast<const bf16*>(x), C, static_cast<bf16*>(y), C, gamma, beta, rows, C, eps, s);
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

int cxrm_test_ln_dwconv(int dtype, const void* x, void* q, void* k, void* v, float* stats, const float* gamma,
                        const float* beta, float eps, const float* w, const float* scale, const float* shift, int n_img,
                        int H, int W, int C, int cls, void* stream) {
  try {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CXRM_F32)
      ln_dwconv_qkv<float>(static_cast<const float*>(x), static_cast<float*>(q), static_cast<float*>(k),
                           static_cast<float*>(v), stats, gamma, beta, eps, w, scale, shift, n_img, H, W, C, cls, s);
    else
      ln_dwconv_qkv<bf16>(static_cast<const bf16*>(x), static_cast<bf16*>(q), static_cast<bf16*>(k), static_cast<bf16*>(v),
                          stats, gamma, beta, eps, w, scale, shift, n_img, H, W, C, cls, s);
    return CXRM_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return CXRM_ERR_CUDA;
  }
}

}  // extern "C"
