// Engine implementation: weight packing, CvT-21 encoder, cross-K/V prefill,
// KV-cached rollout (prefill + decode steps), teacher-forced forward, CXR-BERT
// reward, and the host-buffer SCST step.  Templated on the storage type
// (float = fp32 validation mode, bf16 = tensor-core mode).
#include "engine.cuh"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <type_traits>

namespace cxrm {

unsigned long long g_launch_count = 0;
bool g_pdl = false;

// ---- Arena --------------------------------------------------------------------
void Arena::init(size_t bytes) {
  release();
  CXRM_CUDA_CHECK(cudaMalloc(&base_, bytes));
  cap_ = bytes;
  off_ = 0;
}
void Arena::release() {
  if (base_) cudaFree(base_);
  base_ = nullptr;
  cap_ = off_ = 0;
}
void* Arena::alloc(size_t bytes) {
  const size_t a = (off_ + 255) & ~static_cast<size_t>(255);
  if (a + bytes > cap_)
    throw std::runtime_error("engine scratch arena exhausted: need " + std::to_string(a + bytes) + " of " +
                             std::to_string(cap_) + " bytes (raise the cxrm_config maxima)");
  off_ = a + bytes;
  return base_ + a;
}

namespace {

constexpr int DH = 768, DFF = 3072, NHEAD = 12;
constexpr int CVT_C[3] = {64, 192, 384};
constexpr int CVT_HEADS[3] = {1, 3, 6};
constexpr int CVT_K[3] = {7, 3, 3};
constexpr int CVT_S[3] = {4, 2, 2};
constexpr int CVT_P[3] = {2, 1, 1};
constexpr float LN_EPS_CVT = 1e-5f, BN_EPS = 1e-5f, LN_EPS_BERT = 1e-12f;
constexpr float LORA_SCALE = 32.0f / 8.0f;   // lora_alpha / r (modelling_longitudinal.py:165-166)

// ---- small device helpers local to the engine ------------------------------------
__global__ void valid_images_kernel(const float* pixels, uint8_t* valid, int n, long long img_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) valid[i] = pixels[i * img_stride] != 0.0f ? 1 : 0;   // pixel_values[:, :, 0, 0, 0] != 0.0
}
__global__ void expand_mask_kernel(const uint8_t* valid, uint8_t* mask, int n_img, int T) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long long>(n_img) * T) mask[i] = valid[i / T];
}
// per study: number of visible encoder tokens
__global__ void count_mask_kernel(const uint8_t* mask, int* len, int B, int S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int c = 0;
  if (mask)
    for (int j = 0; j < S; ++j) c += mask[static_cast<long long>(b) * S + j] ? 1 : 0;
  else
    c = S;
  len[b] = c;
}
// compact row list: idx[off[b] + k] = b*S + (k-th visible token of study b)
__global__ void compact_rows_kernel(const uint8_t* mask, const int* off, int* idx, int B, int S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int k = off[b];
  for (int j = 0; j < S; ++j)
    if (!mask || mask[static_cast<long long>(b) * S + j]) idx[k++] = b * S + j;
}
// Packed (variable-length) reward batch: offsets of the sequences and the packed token count (one block) ...
__global__ void seq_offsets_kernel(const int* __restrict__ lens, int n, int L, int* __restrict__ off, int* __restrict__ total) {
  if (threadIdx.x == 0) {
    int o = 0;
    for (int i = 0; i < n; ++i) {
      off[i] = o;
      o += max(0, min(lens[i], L));
    }
    *total = o;
  }
}
// ... and the packed ids / positions (one warp per sequence)
__global__ void __launch_bounds__(32) seq_pack_kernel(const int* __restrict__ ids, const int* __restrict__ lens,
                                                      const int* __restrict__ off, int L, int* __restrict__ pk_ids,
                                                      int* __restrict__ pk_pos, int* __restrict__ len_clamped) {
  const int i = blockIdx.x, len = max(0, min(lens[i], L));
  for (int j = threadIdx.x; j < len; j += 32) {
    pk_ids[off[i] + j] = ids[static_cast<long long>(i) * L + j];
    pk_pos[off[i] + j] = j;
  }
  if (threadIdx.x == 0) len_clamped[i] = len;
}

// Device-side id bridge standing in for split_and_decode_sections + re-tokenisation
// (reference modelling_longitudinal.py:413-457, tools/rewards/cxrbert.py:33-40).
__global__ void bridge_ids_kernel(const int* __restrict__ seq, int ld_seq, int L, int R, int bos, int sep, int eos,
                                  int n_special_vocab, const int* __restrict__ id_map, int cls_id, int sep_id,
                                  int* __restrict__ out_ids, int* __restrict__ out_len, int Lout) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int* row = seq + static_cast<long long>(r) * ld_seq;
  const int specials[3] = {bos, sep, eos};
  int lo[3], hi[3];
  int prev = 0;
  for (int j = 0; j < 3; ++j) {
    if (prev >= L) {
      lo[j] = hi[j] = 0;
      continue;
    }
    int col = 0;
    for (int c = 0; c < L; ++c)
      if (row[c] == specials[j]) {
        col = c;
        break;
      }
    if (col == 0) col = L;
    lo[j] = prev;
    hi[j] = col;
    prev = col;
  }
  int* out = out_ids + static_cast<long long>(r) * Lout;
  int n = 0;
  out[n++] = cls_id;
  for (int j = 1; j < 3; ++j)
    for (int c = lo[j]; c < hi[j]; ++c) {
      const int id = row[c];
      if (id < n_special_vocab) continue;     // skip_special_tokens=True
      if (n < Lout - 1) out[n++] = id_map[id];
    }
  out[n++] = sep_id;
  out_len[r] = n;
  for (int c = n; c < Lout; ++c) out[c] = 0;
}
__global__ void advantage_kernel(const float* r, const float* b, float* adv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) adv[i] = r[i] - b[i];
}

template <typename T>
class Engine : public EngineBase {
 public:
  Engine(const cxrm_config& c, int dev) : cfg(c), device(dev) { setup(); }
  ~Engine() override {
    for (void* p : owned) cudaFree(p);
    if (beam_scratch) cudaFree(beam_scratch);
    for (auto& kv : raw) cudaFree(kv.second.data);
    arena.release();
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (pf_stream) cudaStreamDestroy(pf_stream);
    for (cudaEvent_t e : ev_chunk) cudaEventDestroy(e);
    for (cudaEvent_t e : phase_ev)
      if (e) cudaEventDestroy(e);
    if (prefill_done_ev) cudaEventDestroy(prefill_done_ev);
  }

  // =========================================================================== weights
  void load_weight(const std::string& name, const float* data, const int64_t* shape, int ndim,
                   bool on_device) override {
    RawTensor t;
    t.shape.assign(shape, shape + ndim);
    const long long n = std::max<long long>(t.numel(), 1);
    auto it = raw.find(name);
    if (it != raw.end()) {
      cudaFree(it->second.data);
      raw.erase(it);
    }
    CXRM_CUDA_CHECK(cudaMalloc(&t.data, n * sizeof(float)));
    CXRM_CUDA_CHECK(cudaMemcpy(t.data, data, n * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    raw[name] = t;
    finalized = false;
  }

  const RawTensor& need(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = raw.find(name);
    if (it == raw.end()) throw WeightError("missing weight: " + name);
    consumed.insert(name);
    std::vector<int64_t> want(shape);
    if (it->second.shape != want) {
      std::string got;
      for (auto s : it->second.shape) got += std::to_string(s) + ",";
      throw WeightError("weight " + name + " has shape [" + got + "]");
    }
    return it->second;
  }
  bool has(const std::string& name) const { return raw.count(name) != 0; }
  // W[n_out, n_in] (fp32 staging) += (alpha / r) . B . A when the LoRA pair of `prefix` was loaded (both or neither)
  void merge_lora(const std::string& prefix, float* W, int n_out, int n_in) {
    const bool hasA = has(prefix + ".lora_A.weight"), hasB = has(prefix + ".lora_B.weight");
    if (!hasA && !hasB) return;
    if (hasA != hasB) throw WeightError("weight " + prefix + ": lora_A / lora_B must be loaded as a pair");
    const RawTensor& A = raw.at(prefix + ".lora_A.weight");
    const RawTensor& Bm = raw.at(prefix + ".lora_B.weight");
    consumed.insert(prefix + ".lora_A.weight");
    consumed.insert(prefix + ".lora_B.weight");
    if (!(A.shape.size() == 2 && A.shape[1] == n_in && Bm.shape.size() == 2 && Bm.shape[0] == n_out &&
          Bm.shape[1] == A.shape[0]))
      throw WeightError("weight " + prefix + ": LoRA shapes do not match [r, n_in] / [n_out, r]");
    lora_merge(W, A.data, Bm.data, n_out, n_in, static_cast<int>(A.shape[0]), LORA_SCALE, 0);
  }

  template <typename U> U* dalloc(long long n) {
    void* p = nullptr;
    CXRM_CUDA_CHECK(cudaMalloc(&p, std::max<long long>(n, 1) * sizeof(U)));
    owned.push_back(p);
    persistent_bytes += n * sizeof(U);
    return static_cast<U*>(p);
  }
  float* vecf(const std::string& name, int n) {
    const RawTensor& t = need(name, {n});
    float* d = dalloc<float>(n);
    CXRM_CUDA_CHECK(cudaMemcpyAsync(d, t.data, n * sizeof(float), cudaMemcpyDeviceToDevice, 0));
    return d;
  }
  struct Lin { T* w = nullptr; float* b = nullptr; int n_out = 0, n_in = 0; };
  struct LNp { float* g = nullptr; float* b = nullptr; };
  LNp lnp(const std::string& prefix, int n) { return LNp{vecf(prefix + ".weight", n), vecf(prefix + ".bias", n)}; }
  Lin lin(const std::string& prefix, int n_out, int n_in, bool bias = true) {
    Lin L;
    L.n_out = n_out;
    L.n_in = n_in;
    const RawTensor& w = need(prefix + ".weight", {n_out, n_in});
    L.w = dalloc<T>(static_cast<long long>(n_out) * n_in);
    pack_matrix<T>(w.data, L.w, n_out, n_in, n_in, 0);
    if (bias) L.b = vecf(prefix + ".bias", n_out);
    return L;
  }
  // rows of several [n_i, n_in] matrices stacked into one [sum n_i, n_in]; optional LoRA merge on each part
  Lin lin_cat(const std::vector<std::string>& prefixes, int n_each, int n_in) {
    Lin L;
    L.n_out = n_each * static_cast<int>(prefixes.size());
    L.n_in = n_in;
    L.w = dalloc<T>(static_cast<long long>(L.n_out) * n_in);
    L.b = dalloc<float>(L.n_out);
    float* tmp = nullptr;
    CXRM_CUDA_CHECK(cudaMalloc(&tmp, static_cast<size_t>(n_each) * n_in * sizeof(float)));
    for (size_t i = 0; i < prefixes.size(); ++i) {
      const std::string& p = prefixes[i];
      const RawTensor& w = need(p + ".weight", {n_each, n_in});
      const RawTensor& b = need(p + ".bias", {n_each});
      CXRM_CUDA_CHECK(cudaMemcpyAsync(tmp, w.data, static_cast<size_t>(n_each) * n_in * sizeof(float), cudaMemcpyDeviceToDevice, 0));
      merge_lora(p, tmp, n_each, n_in);
      pack_matrix<T>(tmp, L.w + static_cast<long long>(i) * n_each * n_in, n_each, n_in, n_in, 0);
      CXRM_CUDA_CHECK(cudaMemcpyAsync(L.b + i * n_each, b.data, n_each * sizeof(float), cudaMemcpyDeviceToDevice, 0));
    }
    CXRM_CUDA_CHECK(cudaStreamSynchronize(0));
    cudaFree(tmp);
    return L;
  }

  struct CvtLayerW { LNp ln1, ln2; float* dw; float* bn_scale; float* bn_shift; Lin q, k, v, o, fc1, fc2; };
  struct CvtStageW { Lin emb; LNp emb_ln; std::vector<CvtLayerW> layers; };
  struct BertLayerW { Lin qkv, o; LNp ln1; Lin cq, ckv, co; LNp ln2; Lin fc1, fc2; LNp ln3; };
  struct BertW { T* word = nullptr; T* pos = nullptr; T* type = nullptr; LNp emb_ln; std::vector<BertLayerW> layers; int vocab = 0; };

  T* table(const std::string& name, int rows, int cols) {
    const RawTensor& w = need(name, {rows, cols});
    T* d = dalloc<T>(static_cast<long long>(rows) * cols);
    pack_matrix<T>(w.data, d, rows, cols, cols, 0);
    return d;
  }

  void load_bert(BertW& bw, const std::string& pre, int layers, int vocab, bool cross) {
    bw.vocab = vocab;
    bw.word = table(pre + "embeddings.word_embeddings.weight", vocab, DH);
    bw.pos = table(pre + "embeddings.position_embeddings.weight", 512, DH);
    bw.type = table(pre + "embeddings.token_type_embeddings.weight", 2, DH);
    bw.emb_ln = lnp(pre + "embeddings.LayerNorm", DH);
    bw.layers.resize(layers);
    for (int l = 0; l < layers; ++l) {
      const std::string p = pre + "encoder.layer." + std::to_string(l) + ".";
      BertLayerW& w = bw.layers[l];
      w.qkv = lin_cat({p + "attention.self.query", p + "attention.self.key", p + "attention.self.value"}, DH, DH);
      w.o = lin(p + "attention.output.dense", DH, DH);
      w.ln1 = lnp(p + "attention.output.LayerNorm", DH);
      if (cross) {
        w.cq = lin(p + "crossattention.self.query", DH, DH);
        w.ckv = lin_cat({p + "crossattention.self.key", p + "crossattention.self.value"}, DH, DH);
        w.co = lin(p + "crossattention.output.dense", DH, DH);
        w.ln2 = lnp(p + "crossattention.output.LayerNorm", DH);
      }
      w.fc1 = lin(p + "intermediate.dense", DFF, DH);
      w.fc2 = lin(p + "output.dense", DH, DFF);
      w.ln3 = lnp(p + "output.LayerNorm", DH);
    }
  }

  void finalize_weights() override {
    CXRM_CUDA_CHECK(cudaSetDevice(device));
    // ---- encoder ----
    int cin = 3;
    for (int s = 0; s < 3; ++s) {
      const int C = CVT_C[s], k = CVT_K[s];
      const std::string p = "encoder.cvt.encoder.stages." + std::to_string(s) + ".";
      CvtStageW& st = stages[s];
      const RawTensor& w = need(p + "embedding.convolution_embeddings.projection.weight", {C, cin, k, k});
      const int K = cin * k * k;
      const int Kpad = (K + 7) / 8 * 8;
      st.emb.n_out = C;
      st.emb.n_in = Kpad;
      st.emb.w = dalloc<T>(static_cast<long long>(C) * Kpad);
      if (s == 0) {
        pack_matrix<T>(w.data, st.emb.w, C, K, Kpad, 0);          // K order (cin, ky, kx), zero padded
      } else {
        pack_conv_khwc<T>(w.data, st.emb.w, C, cin, k, 0);        // K order (ky, kx, cin)
      }
      st.emb.b = vecf(p + "embedding.convolution_embeddings.projection.bias", C);
      st.emb_ln = lnp(p + "embedding.convolution_embeddings.normalization", C);
      st.layers.resize(cfg.cvt_depth[s]);
      for (int i = 0; i < cfg.cvt_depth[s]; ++i) {
        const std::string q = p + "layers." + std::to_string(i) + ".";
        CvtLayerW& L = st.layers[i];
        L.ln1 = lnp(q + "layernorm_before", C);
        L.ln2 = lnp(q + "layernorm_after", C);
        L.dw = dalloc<float>(3LL * 9 * C);
        L.bn_scale = dalloc<float>(3LL * C);
        L.bn_shift = dalloc<float>(3LL * C);
        const char* names[3] = {"query", "key", "value"};
        for (int j = 0; j < 3; ++j) {
          const std::string cp = q + "attention.attention.convolution_projection_" + names[j] + ".convolution_projection.";
          pack_dw(need(cp + "convolution.weight", {C, 1, 3, 3}).data, L.dw + static_cast<long long>(j) * 9 * C, C, 0);
          bn_fold(need(cp + "normalization.weight", {C}).data, need(cp + "normalization.bias", {C}).data,
                  need(cp + "normalization.running_mean", {C}).data, need(cp + "normalization.running_var", {C}).data,
                  BN_EPS, L.bn_scale + j * C, L.bn_shift + j * C, C, 0);
        }
        L.q = lin(q + "attention.attention.projection_query", C, C);
        L.k = lin(q + "attention.attention.projection_key", C, C);
        L.v = lin(q + "attention.attention.projection_value", C, C);
        L.o = lin(q + "attention.output.dense", C, C);
        L.fc1 = lin(q + "intermediate.dense", 4 * C, C);
        L.fc2 = lin(q + "output.dense", C, 4 * C);
      }
      cin = C;
    }
    {
      const RawTensor& c = need("encoder.cvt.encoder.stages.2.cls_token", {1, 1, CVT_C[2]});
      cls_token = dalloc<float>(CVT_C[2]);
      CXRM_CUDA_CHECK(cudaMemcpyAsync(cls_token, c.data, CVT_C[2] * sizeof(float), cudaMemcpyDeviceToDevice, 0));
    }
    head_ln = lnp("encoder.projection_head.layer_norm", CVT_C[2]);
    head_proj = lin("encoder.projection_head.projection", DH, CVT_C[2], /*bias=*/false);
    // ---- decoder ----
    load_bert(dec, "decoder.bert.", cfg.dec_layers, cfg.vocab, true);
    dec_head_t = lin("decoder.cls.predictions.transform.dense", DH, DH);
    dec_head_ln = lnp("decoder.cls.predictions.transform.LayerNorm", DH);
    dec_vocab_bias = vecf("decoder.cls.predictions.bias", cfg.vocab);
    dec_lm.w = dec.word;            // tied LM head (SURVEY.md finding 3)
    dec_lm.b = dec_vocab_bias;
    dec_lm.n_out = cfg.vocab;
    dec_lm.n_in = DH;
    // ---- training: the LoRA factors themselves (the merged weights above serve the forward pass) ----
    lora_r.assign(cfg.dec_layers, 0);
    lora_A.assign(cfg.dec_layers, {nullptr, nullptr});
    lora_Bt.assign(cfg.dec_layers, {nullptr, nullptr});
    if (cfg.max_train_tokens > 0) {
      for (int l = 0; l < cfg.dec_layers; ++l) {
        const char* names[2] = {"query", "key"};
        for (int j = 0; j < 2; ++j) {
          const std::string p = "decoder.bert.encoder.layer." + std::to_string(l) + ".attention.self." + names[j];
          if (!has(p + ".lora_A.weight")) continue;
          const RawTensor& A = raw.at(p + ".lora_A.weight");
          const RawTensor& Bm = raw.at(p + ".lora_B.weight");
          const int r = static_cast<int>(A.shape[0]);
          CXRM_CHECK(r % 8 == 0 && r <= 64 && (lora_r[l] == 0 || lora_r[l] == r), "LoRA rank must be a multiple of 8 (<= 64), equal for query and key");
          lora_r[l] = r;
          lora_A[l][j] = dalloc<T>(1LL * r * DH);
          pack_matrix<T>(A.data, lora_A[l][j], r, DH, DH, 0);
          T* Bt = dalloc<T>(1LL * DH * r);
          pack_matrix<T>(Bm.data, Bt, DH, r, r, 0);                  // B [768, r]
          lora_Bt[l][j] = dalloc<T>(1LL * r * DH);
          transpose<T>(Bt, r, lora_Bt[l][j], DH, DH, r, 0);          // B^T [r, 768]
        }
      }
    }
    build_grad_layout();
    wt_cache.clear();
    // ---- reward model ----
    if (cfg.rwd_layers > 0 && has("reward.bert.embeddings.word_embeddings.weight")) {
      load_bert(rwd, "reward.bert.", cfg.rwd_layers, cfg.rwd_vocab, false);
      rp1 = lin("reward.cls_projection_head.dense_to_hidden", 128, DH);
      rp_ln = lnp("reward.cls_projection_head.LayerNorm", 128);
      rp2 = lin("reward.cls_projection_head.dense_to_output", 128, 128);
      have_reward = true;
    }
    CXRM_CUDA_CHECK(cudaDeviceSynchronize());
    // Every loaded tensor must have been used (cxrm.h: unknown weights are CXRM_ERR_WEIGHT).  A state_dict whose keys
    // were renamed only in part (peft's base_layer / lora_A.default naming) would otherwise lose its LoRA update
    // silently.  Tied / alias tensors of the reference's state_dict are accepted.
    std::string leftover;
    int n_left = 0;
    for (auto& kv : raw) {
      if (consumed.count(kv.first)) continue;
      const std::string& k = kv.first;
      if (k == "decoder.cls.predictions.decoder.weight" || k == "decoder.cls.predictions.decoder.bias" ||
          k.find("position_ids") != std::string::npos || k.find("num_batches_tracked") != std::string::npos ||
          (k.rfind("reward.", 0) == 0 && (cfg.rwd_layers == 0 || k.rfind("reward.cls.", 0) == 0 ||
                                         k.find("pooler") != std::string::npos)))
        continue;
      if (n_left++ < 4) leftover += (leftover.empty() ? "" : ", ") + k;
    }
    for (auto& kv : raw) cudaFree(kv.second.data);
    raw.clear();
    consumed.clear();
    if (n_left) throw WeightError("unknown weight: " + std::to_string(n_left) + " loaded tensor(s) match no parameter of the engine: " + leftover);
    finalized = true;
  }

  // =========================================================================== set-up
  void setup() {
    CXRM_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CXRM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    CXRM_CHECK(prop.major == 10, "this engine is built for sm_100a (B200) only; found sm_" +
                                     std::to_string(prop.major) + std::to_string(prop.minor));
    CXRM_CHECK(cfg.image_h % 16 == 0 && cfg.image_w % 16 == 0, "image size must be a multiple of 16");
    CXRM_CHECK(cfg.max_prompt + cfg.max_new_tokens <= 512, "prompt + new tokens must fit 512 positions");
    CXRM_CHECK(cfg.rwd_layers == 0 || (cfg.rwd_max_len >= 1 && cfg.rwd_max_len <= 512), "rwd_max_len must fit the 512 learned positions");
    if (cfg.enc_chunk <= 0) cfg.enc_chunk = 64;
    T2 = (cfg.image_h / 16) * (cfg.image_w / 16);
    Smax = cfg.max_images * T2;
    Rmax = 2 * cfg.max_studies;
    Lmax = (cfg.max_prompt + cfg.max_new_tokens + 15) & ~15;   // row stride of the caches / key mask (16-byte bulk copies)
    const long long B = cfg.max_studies;
    // persistent buffers
    memory = dalloc<T>(B * Smax * DH);
    mem_mask = dalloc<uint8_t>(B * Smax);
    mem_compact = dalloc<T>(B * Smax * DH);
    compact_idx = dalloc<int>(B * Smax);
    kv_off = dalloc<int>(B);
    kv_len = dalloc<int>(B);
    cross_kv = dalloc<T>(static_cast<long long>(cfg.dec_layers) * B * Smax * 2 * DH);
    self_k = dalloc<T>(static_cast<long long>(cfg.dec_layers) * Rmax * Lmax * DH);
    self_v = dalloc<T>(static_cast<long long>(cfg.dec_layers) * Rmax * Lmax * DH);
    logits = dalloc<float>(static_cast<long long>(Rmax) * cfg.vocab);
    valid_img = dalloc<uint8_t>(B * cfg.max_images);
    img_idx = dalloc<int>(B * cfg.max_images);
    // rollout state
    st.cur_token = dalloc<int>(Rmax);
    st.cur_len = dalloc<int>(Rmax);
    st.cur_type = dalloc<int>(Rmax);
    st.cur_pos = dalloc<int>(Rmax);
    st.n_valid = dalloc<int>(Rmax);
    st.seen = dalloc<unsigned>(Rmax);
    st.finished = dalloc<uint8_t>(Rmax);
    st.key_valid = dalloc<uint8_t>(static_cast<long long>(Rmax) * Lmax);
    st.seq = dalloc<int>(static_cast<long long>(Rmax) * Lmax);
    st.logprob = dalloc<float>(static_cast<long long>(Rmax) * cfg.max_new_tokens);
    st.margin = dalloc<float>(static_cast<long long>(Rmax) * cfg.max_new_tokens);
    st.topk_idx = dalloc<int>(static_cast<long long>(Rmax) * cfg.max_new_tokens * kTopKCap);
    st.topk_val = dalloc<float>(static_cast<long long>(Rmax) * cfg.max_new_tokens * kTopKCap);
    st.topk_cnt = dalloc<int>(static_cast<long long>(Rmax) * cfg.max_new_tokens);
    st.step = dalloc<int>(1);
    st.done = dalloc<int>(1);
    st.arrive = dalloc<unsigned>(1);
    st.seed = dalloc<unsigned long long>(1);
    pre_ids = dalloc<int>(static_cast<long long>(Rmax) * cfg.max_prompt);
    pre_types = dalloc<int>(static_cast<long long>(Rmax) * cfg.max_prompt);
    pre_pos = dalloc<int>(static_cast<long long>(Rmax) * cfg.max_prompt);
    pre_valid = dalloc<uint8_t>(static_cast<long long>(Rmax) * cfg.max_prompt);
    prompt_dev = dalloc<int>(B * cfg.max_prompt);
    // decode-attention work units (decode_attn.cu): partials + arrival tickets, zeroed once
    attn_ch = decode_attn_chunk(sizeof(T));
    cross_max_chunks = ceil_div(Smax, attn_ch);
    cross_max_units = cfg.max_studies * cross_max_chunks;
    self_max_chunks = ceil_div(Lmax, attn_ch);
    cross_ws = dalloc<float>(static_cast<long long>(decode_attn_ws_floats(Rmax, cross_max_chunks)));
    self_ws = dalloc<float>(static_cast<long long>(decode_attn_ws_floats(Rmax, self_max_chunks)));
    cross_tickets = dalloc<unsigned>(static_cast<long long>(cfg.max_studies) * NHEAD);
    self_tickets = dalloc<unsigned>(static_cast<long long>(Rmax) * NHEAD);
    CXRM_CUDA_CHECK(cudaMemset(cross_tickets, 0, sizeof(unsigned) * cfg.max_studies * NHEAD));
    CXRM_CUDA_CHECK(cudaMemset(self_tickets, 0, sizeof(unsigned) * Rmax * NHEAD));
    unit_tab = dalloc<int>(4LL * cross_max_units + 2LL * cfg.max_studies + 1);
    // the tensor-core attention units multiply p = 0 with whatever lies behind a short chunk: keep the caches finite
    CXRM_CUDA_CHECK(cudaMemset(cross_kv, 0, sizeof(T) * cfg.dec_layers * cross_layer_stride()));
    CXRM_CUDA_CHECK(cudaMemset(self_k, 0, sizeof(T) * cfg.dec_layers * self_layer_stride()));
    CXRM_CUDA_CHECK(cudaMemset(self_v, 0, sizeof(T) * cfg.dec_layers * self_layer_stride()));
    setup_attn_maps();
    skinny_ws = dalloc<float>(static_cast<long long>(gemm_skinny_partial_floats(DH)));
    chain_bar = dalloc<unsigned>(kChainBarWords);
    CXRM_CUDA_CHECK(cudaMemset(chain_bar, 0, kChainBarWords * sizeof(unsigned)));
    chain_ws = dalloc<float>(static_cast<long long>(kChainMaxSplit) * 64 * DH);
    chain_xf = dalloc<float>(static_cast<long long>(Rmax) * DH);
    chain_x1f = dalloc<float>(static_cast<long long>(Rmax) * DH);

    // scratch arena: max over the phases
    const long long e = sizeof(T);
    const long long H0 = cfg.image_h / 4, W0 = cfg.image_w / 4;
    const long long tok0 = H0 * W0;
    const long long nimg = cfg.enc_chunk;
    const long long big = nimg * tok0 * 64 + nimg * 384 * 2;            // mirrors encode_chunk()
    const long long enc_elems = 5 * big + 4 * (big / 4 + nimg * 384) + std::max<long long>(nimg * tok0 * 152, 4 * big);
    const long long enc_bytes = enc_elems * e + 8 * (nimg * tok0 + nimg) + (1 << 20);
    const long long dec_tok = static_cast<long long>(Rmax) * std::max(cfg.max_prompt, 1);
    const long long per_tok = (4LL * DH + 3 * DH + DFF) * e + 64;
    const long long dec_bytes = (dec_tok + 4LL * Rmax) * per_tok + (1 << 20);
    const long long rwd_tok = static_cast<long long>(std::max(cfg.rwd_max_seqs, 1)) * cfg.rwd_max_len;
    const long long rwd_bytes = cfg.rwd_layers > 0 ? rwd_tok * (per_tok + 16) + (1 << 22) : 0;
    long long train_bytes = 0;
    if (cfg.max_train_tokens > 0) {     // mirrors train_stage(stage 0)
      const long long Mx = cfg.max_train_tokens, V = cfg.vocab, NLd = cfg.dec_layers;
      const long long kvx = B * Smax;
      const long long elems = Mx * (2LL * DH + 15360LL * NLd + 7LL * DH + DFF + 3LL * DH + 2LL * V + DH) +
                              2 * std::max<long long>(Mx * DFF, kvx * 2 * DH) + kvx * 2 * DH;
      train_bytes = elems * e + Mx * (V * 4 + 2LL * NHEAD * 4 + 16) + (64 << 20);
    }
    arena.init(static_cast<size_t>(std::max({enc_bytes, dec_bytes, rwd_bytes, train_bytes})) + (8 << 20));
    CXRM_CUDA_CHECK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    CXRM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_a, cudaEventDisableTiming));
    CXRM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming));
    CXRM_CUDA_CHECK(cudaStreamCreateWithFlags(&pf_stream, cudaStreamNonBlocking));
    std::memset(&graph_key, 0, sizeof(graph_key));
  }
  void setup_attn_maps();
  size_t workspace_bytes() const override { return arena.capacity() + static_cast<size_t>(persistent_bytes); }
  void last_phase_ms(float* out5) const override {
    for (int i = 0; i < 5; ++i) out5[i] = phase_ms[i];
  }

  // =========================================================================== per-kernel-class profiler
  // When enabled (cxrm_set_profile), every kernel launch of the engine is bracketed by a CUDA event pair on
  // the launching stream and attributed to a class tag; graphs are bypassed so that each launch is visible.
  struct ProfRec { std::string tag; cudaEvent_t a, b; };
  template <class F> void PF(const char* tag, cudaStream_t s, F&& f) {
    if (!profiling) {
      f();
      return;
    }
    ProfRec r;
    r.tag = std::string(phase) + "." + tag;
    if (ev_pool.size() >= 2) {
      r.a = ev_pool.back(); ev_pool.pop_back();
      r.b = ev_pool.back(); ev_pool.pop_back();
    } else {
      CXRM_CUDA_CHECK(cudaEventCreate(&r.a));
      CXRM_CUDA_CHECK(cudaEventCreate(&r.b));
    }
    CXRM_CUDA_CHECK(cudaEventRecord(r.a, s));
    f();
    CXRM_CUDA_CHECK(cudaEventRecord(r.b, s));
    prof_recs.push_back(r);
  }
  void set_profile(bool on) override { profiling = on; }
  std::string profile_report() override {
    CXRM_CUDA_CHECK(cudaDeviceSynchronize());
    std::map<std::string, std::pair<double, long long>> agg;
    for (auto& r : prof_recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      auto& e = agg[r.tag];
      e.first += ms;
      e.second += 1;
      ev_pool.push_back(r.a);
      ev_pool.push_back(r.b);
    }
    prof_recs.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
      if (!first) out += ", ";
      first = false;
      out += "\"" + kv.first + "\": {\"ms\": " + std::to_string(kv.second.first) + ", \"n\": " +
             std::to_string(kv.second.second) + "}";
    }
    return out + "}";
  }

  // =========================================================================== GEMM dispatch
  void gemm(const T* A, int lda, const Lin& L, void* C, int ldc, long long M, int act, const T* residual, int ldr,
            bool out_f32, const int* skip, cudaStream_t s, const char* tag = "gemm", long long c_head_stride = 0) {
    GemmArgs g;
    g.c_head_stride = c_head_stride;
    g.trace = nullptr;
    g.A = A; g.lda = lda; g.W = L.w; g.ldw = L.n_in; g.C = C; g.ldc = ldc;
    g.M = static_cast<int>(M); g.N = L.n_out; g.K = L.n_in;
    g.bias = L.b; g.act = act; g.residual = residual; g.ldr = ldr; g.out_f32 = out_f32 ? 1 : 0; g.skip_flag = skip;
    std::string shaped;
    if (profiling) {   // per-shape attribution: gemm[N x K]
      shaped = std::string(tag) + "[" + std::to_string(g.N) + "x" + std::to_string(g.K) + "]";
      tag = shaped.c_str();
    }
    PF(tag, s, [&] { dispatch_gemm(g, s); });
  }
  void dispatch_gemm(const GemmArgs& g, cudaStream_t s);

  // out = LayerNorm(act(A.W^T + bias) + residual), eps 1e-12 (every LN that follows a decoder GEMM).
  // Decode steps in bf16 (M <= 64): skinny split-K GEMM into fp32 partials + one fused reduce/bias/residual/LN kernel;
  // otherwise GEMM with fused epilogue followed by the LayerNorm kernel.  `out` may alias `residual`.
  void gemm_ln(const T* A, int lda, const Lin& L, int act, const T* residual, int ldr, const LNp& ln, T* out, int ldo,
               long long M, const int* skip, cudaStream_t s) {
    if (use_skinny(M, L)) {
      GemmArgs g = make_args(A, lda, L, nullptr, 0, M, ACT_NONE, nullptr, 0, false, skip);
      int nsplit = 0;
      PF("gemm", s, [&] { gemm_tcgen05_skinny(g, skinny_ws, &nsplit, s); });
      PF("layernorm", s, [&] { splitk_ln(skinny_ws, nsplit, static_cast<int>(M), L.n_out, L.b, act, residual, ldr, ln.g, ln.b,
                                         LN_EPS_BERT, out, ldo, skip, s); });
      return;
    }
    T* tmp = out;
    gemm(A, lda, L, tmp, ldo, M, act, residual, ldr, false, skip, s);
    PF("layernorm", s, [&] { layernorm<T>(tmp, ldo, out, ldo, ln.g, ln.b, M, L.n_out, LN_EPS_BERT, s); });
  }
  // (A single-kernel version of the pair - 8-CTA cluster, LayerNorm statistics through distributed shared memory - was
  // built in round 1 and measured 17 us against 8 us per call: three cluster barriers; removed in round 2.)
  bool use_skinny(long long M, const Lin& L) const;
  bool chain_pdl() const;
  GemmArgs make_args(const T* A, int lda, const Lin& L, void* C, int ldc, long long M, int act, const T* residual,
                     int ldr, bool out_f32, const int* skip) const {
    GemmArgs g;
    g.c_head_stride = 0; g.trace = nullptr;
    g.A = A; g.lda = lda; g.W = L.w; g.ldw = L.n_in; g.C = C; g.ldc = ldc;
    g.M = static_cast<int>(M); g.N = L.n_out; g.K = L.n_in;
    g.bias = L.b; g.act = act; g.residual = residual; g.ldr = ldr; g.out_f32 = out_f32 ? 1 : 0; g.skip_flag = skip;
    return g;
  }

  // =========================================================================== encoder
  // n images (indices img_idx_dev into `pixels`) -> proj [n*T2, 768] in the arena
  T* encode_chunk(const float* pixels, const int* idx_dev, int n, cudaStream_t s) {
    phase = "enc";
    arena.reset();
    // every kernel of the chunk is a programmatic dependent launch of its predecessor: its CTAs are scheduled and set up
    // (barriers, TMEM, descriptors, resident weights) under the predecessor's tail and wait in griddepcontrol.wait
    struct EncPdl {
      explicit EncPdl(bool on) { g_pdl = on; }
      ~EncPdl() { g_pdl = false; }
    } enc_pdl(chain_pdl() && !profiling && std::getenv("CXRM_NO_ENC_PDL") == nullptr);
    int H = cfg.image_h, W = cfg.image_w;
    const long long tok0 = static_cast<long long>(H / 4) * (W / 4);
    const long long nt = static_cast<long long>(n) * tok0;   // largest token count (stage 1)
    // buffers sized for stage 1 (tokens*C is largest there: tok0*64 >= tok0/4*192 >= (tok0/16+1)*384 holds for tok0 >= 16)
    const long long big = nt * 64 + static_cast<long long>(n) * 384 * 2;
    T* x = arena.get<T>(big);
    T* x2 = arena.get<T>(big);
    T* y = arena.get<T>(big);
    T* q = arena.get<T>(big);
    T* qp = arena.get<T>(big);
    T* k = arena.get<T>(big / 4 + n * 384);
    T* v = arena.get<T>(big / 4 + n * 384);
    T* kp = arena.get<T>(big / 4 + n * 384);
    T* vp = arena.get<T>(big / 4 + n * 384);
    const long long hid_elems = std::max<long long>(nt * 152, 4 * big);
    T* hid = arena.get<T>(hid_elems);   // im2col buffer and MLP hidden share storage
    float* ln_stats = arena.get<float>(2 * (nt + n));   // (mean, rstd) per token of the largest stage
    T* prev = nullptr;                  // previous stage's tokens [n, Hp*Wp, Cp]
    int Hp = 0, Wp = 0;
    for (int s_ = 0; s_ < 3; ++s_) {
      const CvtStageW& sw = stages[s_];
      const int C = CVT_C[s_], ks = CVT_K[s_], sd = CVT_S[s_], pd = CVT_P[s_];
      int Ho, Wo;
      if (s_ == 0) {
        Ho = (H + 2 * pd - ks) / sd + 1;
        Wo = (W + 2 * pd - ks) / sd + 1;
        PF("im2col", s, [&] { im2col_pixels<T>(pixels, idx_dev, hid, n, H, W, ks, sd, pd, sw.emb.n_in, s); });
      } else {
        Ho = (Hp + 2 * pd - ks) / sd + 1;
        Wo = (Wp + 2 * pd - ks) / sd + 1;
        PF("im2col", s, [&] { im2col_tokens<T>(prev, hid, n, Hp, Wp, CVT_C[s_ - 1], ks, sd, pd, s); });
      }
      const long long rows = static_cast<long long>(n) * Ho * Wo;
      const int cls = (s_ == 2) ? 1 : 0;
      T* emb_out = cls ? x2 : x;
      gemm(hid, sw.emb.n_in, sw.emb, emb_out, C, rows, ACT_NONE, nullptr, 0, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(emb_out, C, emb_out, C, sw.emb_ln.g, sw.emb_ln.b, rows, C, LN_EPS_CVT, s); });
      if (cls) PF("cls", s, [&] { cat_cls<T>(x2, cls_token, x, n, Ho * Wo, C, s); });
      const int Tq = cls + Ho * Wo;
      const int Hk = (Ho + 2 - 3) / 2 + 1, Wk = (Wo + 2 - 3) / 2 + 1;
      const int Tk = cls + Hk * Wk;
      const long long rq = static_cast<long long>(n) * Tq, rk = static_cast<long long>(n) * Tk;
      for (const CvtLayerW& L : sw.layers) {
        // LayerNorm-before + convolutional projections in one pass over x (the normalised map is never stored)
        PF("ln_dwconv", s, [&] { ln_dwconv_qkv<T>(x, q, k, v, ln_stats, L.ln1.g, L.ln1.b, LN_EPS_CVT, L.dw, L.bn_scale, L.bn_shift, n, Ho, Wo,
                                                  C, cls, s); });
        gemm(q, C, L.q, qp, C, rq, ACT_NONE, nullptr, 0, false, nullptr, s);
        gemm(k, C, L.k, kp, C, rk, ACT_NONE, nullptr, 0, false, nullptr, s);
        gemm(v, C, L.v, vp, C, rk, ACT_NONE, nullptr, 0, false, nullptr, s);
        AttnArgs a{};
        a.q = qp; a.k = kp; a.v = vp; a.o = q;     // context reuses the q buffer
        a.q_bs = static_cast<long long>(Tq) * C; a.q_hs = 64; a.q_ts = C;
        a.k_bs = static_cast<long long>(Tk) * C; a.k_hs = 64; a.k_ts = C;
        a.v_bs = a.k_bs; a.v_hs = 64; a.v_ts = C;
        a.o_bs = a.q_bs; a.o_hs = 64; a.o_ts = C;
        a.batch = n; a.heads = CVT_HEADS[s_]; a.Lq = Tq; a.Lk = Tk;
        a.scale = 1.0f / sqrtf(static_cast<float>(C));   // embed_dim ** -0.5 (modeling_cvt.py:183)
        PF("attn", s, [&] { attention(a, s); });
        gemm(q, C, L.o, x2, C, rq, ACT_NONE, x, C, false, nullptr, s);          // + residual
        PF("layernorm", s, [&] { layernorm<T>(x2, C, y, C, L.ln2.g, L.ln2.b, rq, C, LN_EPS_CVT, s); });
        gemm(y, C, L.fc1, hid, 4 * C, rq, ACT_GELU, nullptr, 0, false, nullptr, s);
        gemm(hid, 4 * C, L.fc2, x, C, rq, ACT_NONE, x2, C, false, nullptr, s);  // + residual
      }
      if (cls) {
        PF("cls", s, [&] { drop_cls<T>(x, x2, n, Ho * Wo, C, s); });
        std::swap(x, x2);
      }
      // x now holds [n, Ho*Wo, C]; it becomes `prev`; continue in the other buffer
      prev = x;
      std::swap(x, x2);
      Hp = Ho;
      Wp = Wo;
    }
    // projection head: LN(eps 1e-12) -> Linear 384 -> 768 without bias (modelling_longitudinal.py:40-43)
    const long long rows = static_cast<long long>(n) * T2;
    PF("layernorm", s, [&] { layernorm<T>(prev, CVT_C[2], y, CVT_C[2], head_ln.g, head_ln.b, rows, CVT_C[2], LN_EPS_BERT, s); });
    T* proj = qp;
    gemm(y, CVT_C[2], head_proj, proj, DH, rows, ACT_NONE, nullptr, 0, false, nullptr, s);
    return proj;
  }

  void encode(const float* pixels, int B, int N, void* memory_out, uint8_t* mask_out, cudaStream_t s) override {
    CXRM_CHECK(finalized, "weights not finalized");
    CXRM_CHECK(B >= 1 && B <= cfg.max_studies && N >= 1 && N <= cfg.max_images, "B/N exceed the configured maxima");
    const int n_all = B * N;
    const long long img_stride = 3LL * cfg.image_h * cfg.image_w;
    valid_images_kernel<<<ceil_div(n_all, 128), 128, 0, s>>>(pixels, valid_img, n_all, img_stride);
    check_launch("valid_images");
    std::vector<uint8_t> valid(n_all);
    CXRM_CUDA_CHECK(cudaMemcpyAsync(valid.data(), valid_img, n_all, cudaMemcpyDeviceToHost, s));
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<int> idx;
    for (int i = 0; i < n_all; ++i)
      if (valid[i]) idx.push_back(i);
    encode_valid(pixels, B, N, idx, idx, nullptr, s);
    if (memory_out)
      CXRM_CUDA_CHECK(cudaMemcpyAsync(memory_out, memory, static_cast<size_t>(B) * enc_S * DH * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (mask_out)
      CXRM_CUDA_CHECK(cudaMemcpyAsync(mask_out, mem_mask, static_cast<size_t>(B) * enc_S, cudaMemcpyDeviceToDevice, s));
  }

  // images per encoder pass: the valid images split into equal passes of at most enc_chunk (a 4-image tail pass costs
  // the same ~240 launches as a full one)
  size_t balanced_chunk(size_t n_valid) const {
    if (n_valid == 0) return static_cast<size_t>(cfg.enc_chunk);
    const size_t passes = (n_valid + cfg.enc_chunk - 1) / cfg.enc_chunk;
    return (n_valid + passes - 1) / passes;
  }
  // Encode the valid images: the k-th one is image src[k] of `pixels` and lands in slot dst[k] (= b*N + n) of the
  // encoder memory [B, N*T2, 768]; valid_img (device flags per slot) must be set.  chunk_ready (nullable): event c
  // is awaited before chunk c is touched (host-buffer path: the chunk's pixels arrive on the copy stream meanwhile).
  void encode_valid(const float* pixels, int B, int N, const std::vector<int>& src, const std::vector<int>& dst,
                    const cudaEvent_t* chunk_ready, cudaStream_t s) {
    const int n_all = B * N;
    enc_B = B;
    enc_S = N * T2;
    fill_zero(memory, static_cast<size_t>(B) * enc_S * DH * sizeof(T), s);
    expand_mask_kernel<<<static_cast<unsigned>(ceil_div_ll(static_cast<long long>(n_all) * T2, 256)), 256, 0, s>>>(
        valid_img, mem_mask, n_all, T2);
    check_launch("expand_mask");
    if (dst.empty()) return;
    CXRM_CUDA_CHECK(cudaMemcpyAsync(img_idx, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    const size_t per = balanced_chunk(dst.size());
    for (size_t c0 = 0, c = 0; c0 < dst.size(); c0 += per, ++c) {
      const int n = static_cast<int>(std::min<size_t>(per, dst.size() - c0));
      if (chunk_ready) CXRM_CUDA_CHECK(cudaStreamWaitEvent(s, chunk_ready[c], 0));
      T* proj = encode_chunk(pixels, img_idx + c0, n, s);
      for (int i = 0; i < n;) {   // consecutive slots leave as one copy
        int j = i + 1;
        while (j < n && dst[c0 + j] == dst[c0 + j - 1] + 1) ++j;
        const long long d0 = dst[c0 + i];
        CXRM_CUDA_CHECK(cudaMemcpyAsync(memory + d0 * T2 * DH, proj + static_cast<long long>(i) * T2 * DH,
                                        static_cast<size_t>(j - i) * T2 * DH * sizeof(T), cudaMemcpyDeviceToDevice, s));
        i = j;
      }
    }
  }

  // =========================================================================== cross K/V
  void prefill_cross_kv(const void* mem_in, const uint8_t* mask_in, int B, int S, cudaStream_t s) override {
    CXRM_CHECK(finalized, "weights not finalized");
    const T* mem = static_cast<const T*>(mem_in);
    const uint8_t* mask = mask_in;
    if (!mem) {
      CXRM_CHECK(enc_B > 0, "cxrm_prefill_cross_kv(NULL) needs a preceding cxrm_encode");
      mem = memory;
      mask = mem_mask;
      B = enc_B;
      S = enc_S;
    }
    CXRM_CHECK(B >= 1 && B <= cfg.max_studies && S >= 1 && S <= Smax, "B/S exceed the configured maxima");
    phase = "xkv";
    count_mask_kernel<<<ceil_div(B, 64), 64, 0, s>>>(mask, kv_len, B, S);
    check_launch("count_mask");
    std::vector<int> len(B), off(B);
    CXRM_CUDA_CHECK(cudaMemcpyAsync(len.data(), kv_len, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    int total = 0, mx = 0;
    for (int b = 0; b < B; ++b) {
      off[b] = total;
      total += len[b];
      mx = std::max(mx, len[b]);
      CXRM_CHECK(len[b] > 0, "a study has no visible encoder token");
    }
    kv_total = total;
    kv_maxlen = mx;
    kv_B = B;
    h_kv_len = len;
    h_kv_off = off;
    CXRM_CUDA_CHECK(cudaMemcpyAsync(kv_off, off.data(), B * sizeof(int), cudaMemcpyHostToDevice, s));
    compact_rows_kernel<<<ceil_div(B, 64), 64, 0, s>>>(mask, kv_off, compact_idx, B, S);
    check_launch("compact_rows");
    PF("gather", s, [&] { gather_rows<T>(mem, compact_idx, mem_compact, total, DH, s); });
    for (int l = 0; l < cfg.dec_layers; ++l) {
      // head-major store: [k|v][head][token][64] (n / 64 = kv * 12 + head)
      T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
      gemm(mem_compact, DH, dec.layers[l].ckv, kvl, 2 * DH, total, ACT_NONE, nullptr, 0, false, nullptr, s, "gemm",
           cross_head_stride());
    }
    // uniform work units of the decode-step cross-attention: (study, chunk of attn_ch tokens)
    std::vector<int> tab(4 * static_cast<size_t>(cross_max_units) + 2 * cfg.max_studies + 1, 0);
    int* u_study = tab.data();
    int* u_j0 = u_study + cross_max_units;
    int* u_n = u_j0 + cross_max_units;
    int* u_chunk = u_n + cross_max_units;
    int* n_chunks = u_chunk + cross_max_units;
    int* first_unit = n_chunks + cfg.max_studies;
    int nu = 0;
    for (int b = 0; b < B; ++b) {
      const int nc = ceil_div(len[b], attn_ch);
      n_chunks[b] = nc;
      first_unit[b] = nu;
      for (int c = 0; c < nc; ++c, ++nu) {
        u_study[nu] = b;
        u_j0[nu] = off[b] + c * attn_ch;
        u_n[nu] = std::min(attn_ch, len[b] - c * attn_ch);
        u_chunk[nu] = c;
      }
    }
    first_unit[cfg.max_studies] = nu;
    CXRM_CUDA_CHECK(cudaMemcpyAsync(unit_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));   // `tab` is pageable host memory going out of scope
  }
  CrossUnits cross_units() const { return units_of(unit_tab); }
  CrossUnits units_of(const int* unit_tab) const {
    CrossUnits cu;
    cu.study = unit_tab;
    cu.j0 = unit_tab + cross_max_units;
    cu.n = unit_tab + 2 * cross_max_units;
    cu.chunk = unit_tab + 3 * cross_max_units;
    cu.n_chunks = unit_tab + 4 * cross_max_units;
    cu.first_unit = cu.n_chunks + cfg.max_studies;
    cu.n_units = cu.first_unit + cfg.max_studies;
    cu.max_units = cross_max_units;
    cu.max_chunks = cross_max_chunks;
    return cu;
  }
  long long cross_tok_cap() const { return static_cast<long long>(cfg.max_studies) * Smax; }
  long long cross_head_stride() const { return cross_tok_cap() * 64; }
  long long cross_layer_stride() const { return cross_tok_cap() * 2 * DH; }

  // =========================================================================== attention dispatch
  void attention(const AttnArgs& a, cudaStream_t s);

  // =========================================================================== decoder trunk
  // tokens M = R*qlen already embedded in x [M,768]; returns the buffer holding the output hidden states
  struct DecBufs { T* x; T* x1; T* qkv; T* ctx; T* hid; };
  DecBufs dec_bufs(long long M) {
    DecBufs b;
    b.x = arena.get<T>(M * DH);
    b.x1 = arena.get<T>(M * DH);
    b.qkv = arena.get<T>(M * 3 * DH);
    b.ctx = arena.get<T>(M * DH);
    b.hid = arena.get<T>(M * DFF);
    return b;
  }

  // full-sequence pass (prefill / teacher forcing): R rows of q tokens, causal + key_mask [R, ld_mask]
  // pk (nullable) + M_packed: the rows of b.x are the PACKED prompt tokens (kernels.h PromptPack) instead of the R x q grid
  T* decoder_full(DecBufs& b, int R, int q, int B, const uint8_t* key_mask, int ld_mask, bool store_cache,
                  cudaStream_t s, const PromptPack* pk = nullptr, long long M_packed = 0) {
    const long long M = pk ? M_packed : static_cast<long long>(R) * q;
    for (int l = 0; l < cfg.dec_layers; ++l) {
      const BertLayerW& w = dec.layers[l];
      gemm(b.x, DH, w.qkv, b.qkv, 3 * DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
      if (store_cache) {
        if (pk)
          PF("store_kv", s, [&] { prefill_store_kv<T>(b.qkv, self_k + l * self_layer_stride(), self_v + l * self_layer_stride(), pk->tok_slot,
                                                      pk->tok_cache, 1, static_cast<int>(M), Lmax, s, pk->tok_row); });
        else
          PF("store_kv", s, [&] { prefill_store_kv<T>(b.qkv, self_k + l * self_layer_stride(), self_v + l * self_layer_stride(), pre_pos, pre_valid,
                                                      R, q, Lmax, s); });
      }
      AttnArgs a{};
      a.q = b.qkv; a.k = b.qkv + DH; a.v = b.qkv + 2 * DH; a.o = b.ctx;
      a.q_bs = static_cast<long long>(q) * 3 * DH; a.q_hs = 64; a.q_ts = 3 * DH;
      a.k_bs = a.q_bs; a.k_hs = 64; a.k_ts = 3 * DH;
      a.v_bs = a.q_bs; a.v_hs = 64; a.v_ts = 3 * DH;
      a.o_bs = static_cast<long long>(q) * DH; a.o_hs = 64; a.o_ts = DH;
      a.batch = R; a.heads = NHEAD; a.Lq = q; a.Lk = q;
      a.key_mask = key_mask; a.key_mask_ld = ld_mask; a.key_mask_per_q_batch = 1;
      a.causal = 1; a.q_pos_offset = 0;
      a.scale = 0.125f;
      if (pk) {   // row r: queries / keys at packed tokens row_off[r] ..; every packed key is visible, the query-only
                  // last column (index row_lk[r]) sees all row_lk[r] keys through the causal rule
        a.q_offset = pk->row_off; a.Lq_per_batch = pk->row_lq;
        a.kv_offset = pk->row_off; a.Lk_per_batch = pk->row_lk;
        a.k_bs = a.v_bs = 0;
        a.key_mask = nullptr;
      }
      PF("attn", s, [&] { attention(a, s); });
      gemm(b.ctx, DH, w.o, b.x1, DH, M, ACT_NONE, b.x, DH, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(b.x1, DH, b.x1, DH, w.ln1.g, w.ln1.b, M, DH, LN_EPS_BERT, s); });
      // cross-attention over the ragged encoder K/V cache
      gemm(b.x1, DH, w.cq, b.qkv, DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
      const T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
      AttnArgs c{};
      c.q = b.qkv; c.k = kvl; c.v = kvl + NHEAD * cross_head_stride(); c.o = b.ctx;
      c.q_bs = static_cast<long long>(q) * DH; c.q_hs = 64; c.q_ts = DH;
      c.k_bs = 0; c.k_hs = cross_head_stride(); c.k_ts = 64;
      c.v_bs = 0; c.v_hs = cross_head_stride(); c.v_ts = 64;
      c.o_bs = c.q_bs; c.o_hs = 64; c.o_ts = DH;
      c.batch = R; c.heads = NHEAD; c.Lq = q; c.Lk = kv_maxlen;
      c.Lk_per_batch = kv_len; c.kv_offset = kv_off; c.kv_batch_mod = B;
      c.scale = 0.125f;
      if (pk) {
        c.q_offset = pk->row_off; c.Lq_per_batch = pk->row_lq;
      }
      PF("attn", s, [&] { attention(c, s); });
      gemm(b.ctx, DH, w.co, b.x, DH, M, ACT_NONE, b.x1, DH, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(b.x, DH, b.x, DH, w.ln2.g, w.ln2.b, M, DH, LN_EPS_BERT, s); });
      gemm(b.x, DH, w.fc1, b.hid, DFF, M, ACT_GELU, nullptr, 0, false, nullptr, s);
      gemm(b.hid, DFF, w.fc2, b.x1, DH, M, ACT_NONE, b.x, DH, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(b.x1, DH, b.x, DH, w.ln3.g, w.ln3.b, M, DH, LN_EPS_BERT, s); });
    }
    return b.x;
  }
  long long self_layer_stride() const { return static_cast<long long>(Rmax) * Lmax * DH; }

  // LM head on `rows` hidden rows -> fp32 logits (dense -> GELU -> LN -> tied decoder + bias)
  void lm_head(const T* hidden, long long rows, T* tmp, float* out, int ld_out, const int* skip, cudaStream_t s) {
    gemm_ln(hidden, DH, dec_head_t, ACT_GELU, nullptr, 0, dec_head_ln, tmp, DH, rows, skip, s);
    gemm(tmp, DH, dec_lm, out, ld_out, rows, ACT_NONE, nullptr, 0, true, skip, s);
  }

  // one decode step for R rows (all inputs come from the device-side rollout state)
  // Timing aid (never set in production): CXRM_ABLATE=<comma list of gemm,ln,self,cross,sample,embed> drops those
  // launches from the decode step, so that bench.py differences give each class's in-graph cost.
  static unsigned ablate_mask() {
    static int m = -1;
    if (m < 0) {
      m = 0;
      const char* e = std::getenv("CXRM_ABLATE");
      if (e) {
        const std::string v(e);
        const char* names[6] = {"gemm", "ln", "self", "cross", "sample", "embed"};
        for (int i = 0; i < 6; ++i)
          if (v.find(names[i]) != std::string::npos) m |= 1 << i;
      }
    }
    return static_cast<unsigned>(m);
  }

  // ---- decode step with every (split-K GEMM, reduce + LayerNorm) pair as ONE persistent two-phase launch
  // (decode_chain.cu): the GEMM phase leaves fp32 partials, a grid barrier replaces the kernel boundary, the second
  // phase reduces + normalises one row per CTA.  The projections that feed an attention kernel (QKV, cross-Q) and FFN-up
  // stay on the PDL-chained skinny kernels, so the attention kernels keep prefetching K/V under their predecessor.
  // A version that ran ALL GEMM / LayerNorm work between two attention kernels as one launch measured slower than the
  // PDL chain.  Per step 58 launches instead of 71.
  bool use_chain(int R) const {
    return std::is_same<T, bf16>::value && cfg.use_tensor_cores && R <= 64 && ablate_mask() == 0 &&
           decode_chain_available();
  }
  void ensure_chain(const DecBufs& b, T* head_tmp, int R) {
    if constexpr (std::is_same<T, bf16>::value) {
      if (chain_key_buf == b.x && chain_key_R == R && chain_key_head == head_tmp && !chain_launch.empty()) return;
      // fp32 residual stream beside the bf16 activations (CXRM_CHAIN_BF16_RES=1: bf16 residuals, pre-LN sum rounded to
      // bf16 as the stand-alone reduce + LayerNorm kernel and bf16 autocast do)
      static const bool f32res = std::getenv("CXRM_CHAIN_BF16_RES") == nullptr;
      std::vector<ChainPhase> ph;
      chain_launch.clear();
      auto pair = [&](const T* A, int Ktot, const Lin& L, int kslice, int act, const T* res, const float* res32, const LNp& ln,
                      T* out, float* out32) {
        const size_t first = ph.size();
        ChainPhase p;
        std::memset(&p, 0, sizeof(p));
        p.type = CH_GEMM;
        p.bn = 64; p.kslice = kslice; p.n_tiles = L.n_out / p.bn; p.nsplit = Ktot / kslice; p.epi = CE_PARTIAL; p.N = L.n_out;
        p.tmA = make_tensor_map_bf16_kblocks(A, R, Ktot, Ktot, 64, kslice / 64);
        p.tmB = make_tensor_map_bf16(L.w, L.n_out, L.n_in, L.n_in, p.bn, 64);
        p.partial = chain_ws;
        CXRM_CHECK(L.n_in == Ktot && L.n_out == DH && Ktot % kslice == 0 && kslice % 64 == 0 && p.nsplit <= kChainMaxSplit &&
                       p.bn * kslice * 2 <= 48 * 1024, "decode chain: GEMM phase shape");
        ph.push_back(p);
        ChainPhase q;
        std::memset(&q, 0, sizeof(q));
        q.type = CH_LN; q.nsplit = p.nsplit; q.bias = L.b; q.act = act; q.partial = chain_ws;
        const bool use32 = f32res && res32 != nullptr;
        q.residual = use32 ? nullptr : res; q.residual_f32 = use32 ? res32 : nullptr;
        q.round_pre = f32res ? 0 : 1;
        q.gamma = ln.g; q.beta = ln.b; q.eps = LN_EPS_BERT; q.out = out; q.ldo = DH; q.out_f32 = f32res ? out32 : nullptr;
        ph.push_back(q);
        const int ctas = std::max(p.n_tiles * p.nsplit, R);
        chain_launch.push_back({static_cast<int>(first), 2, ctas});
      };
      float* xf = chain_xf; float* x1f = chain_x1f;
      for (int l = 0; l < cfg.dec_layers; ++l) {
        const BertLayerW& w = dec.layers[l];
        // layer 0's input comes from the embedding kernel (bf16 only); later layers carry the fp32 copy written by LN3
        pair(b.ctx, DH, w.o, 192, ACT_NONE, b.x, l == 0 ? nullptr : xf, w.ln1, b.x1, x1f);
        pair(b.ctx, DH, w.co, 192, ACT_NONE, b.x1, x1f, w.ln2, b.x, xf);
        pair(b.hid, DFF, w.fc2, 256, ACT_NONE, b.x, xf, w.ln3, b.x, xf);
      }
      pair(b.x, DH, dec_head_t, 192, ACT_GELU, nullptr, nullptr, dec_head_ln, head_tmp, nullptr);
      static const bool want_trace = std::getenv("CXRM_CHAIN_TRACE") != nullptr;
      if (want_trace && !chain_trace) {
        const long long nt = static_cast<long long>(chain_launch.size()) * decode_chain_ctas() * kChainTraceSlots;
        chain_trace = dalloc<unsigned long long>(nt);
        CXRM_CUDA_CHECK(cudaMemset(chain_trace, 0, nt * sizeof(unsigned long long)));
      }
      chain_host = std::move(ph);
      chain_key_buf = b.x; chain_key_R = R; chain_key_head = head_tmp;
    }
  }
  void decode_step_chain(DecBufs& b, T* head_tmp, const RolloutParams& rp, const float* noise, cudaStream_t s) {
    phase = "decode";
    const int R = rp.R, B = rp.B;
    const int* skip = st.done;
    size_t k = 0;
    auto chain = [&]() {
      const auto fc = chain_launch[k];
      unsigned long long* tr = chain_trace ? chain_trace + k * static_cast<size_t>(decode_chain_ctas()) * kChainTraceSlots : nullptr;
      ++k;
      PF("gemm_ln", s, [&] { decode_chain(chain_host.data() + fc.first, fc.second, R, st, chain_bar, fc.ctas, s, tr); });
    };
    PF("embed_ln", s, [&] { embed_ln<T>(st.cur_token, st.cur_type, st.cur_pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, b.x, R,
                DH, LN_EPS_BERT, s); });
    struct PdlScope {
      explicit PdlScope(bool on) { g_pdl = on; }
      ~PdlScope() { g_pdl = false; }
    } pdl_scope(chain_pdl() && !profiling);
    for (int l = 0; l < cfg.dec_layers; ++l) {
      const BertLayerW& w = dec.layers[l];
      gemm(b.x, DH, w.qkv, b.qkv, 3 * DH, R, ACT_NONE, nullptr, 0, false, skip, s);
      PF("self_attn", s, [&] { decode_self_attention<T>(b.qkv, self_k + l * self_layer_stride(), self_v + l * self_layer_stride(), b.ctx, st, R,
                               rp.P, Lmax, self_ws, self_tickets, attn_maps_ptr, l, s); });
      chain();                                 // x1 = LN1(ctx . Wo + b + x)
      gemm(b.x1, DH, w.cq, b.qkv, DH, R, ACT_NONE, nullptr, 0, false, skip, s);
      const T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
      PF("cross_attn", s, [&] { decode_cross_attention<T>(b.qkv, DH, kvl, kvl + NHEAD * cross_head_stride(), cross_head_stride(), b.ctx,
                                cross_units(), st, R, B, cross_ws, cross_tickets, attn_maps_ptr, l, s); });
      chain();                                 // x = LN2(ctx . Wco + b + x1)
      gemm(b.x, DH, w.fc1, b.hid, DFF, R, ACT_GELU, nullptr, 0, false, skip, s);
      chain();                                 // x = LN3(hid . W2 + b + x)
    }
    chain();                                   // LM-head transform: LN(GELU(x . Wt + b))
    gemm(head_tmp, DH, dec_lm, logits, cfg.vocab, R, ACT_NONE, nullptr, 0, true, skip, s);
    PF("sample", s, [&] { sample_step(st, rp, logits, cfg.vocab, noise, s); });
  }

  void decode_step(DecBufs& b, T* head_tmp, const RolloutParams& rp, const float* noise, cudaStream_t s) {
    if (rp.beams == 0 && use_chain(rp.R)) {
      ensure_chain(b, head_tmp, rp.R);
      decode_step_chain(b, head_tmp, rp, noise, s);
      return;
    }
    phase = "decode";
    // beam search: every running beam is a virtual study (B == R) whose units point at its real study's encoder K/V
    const int R = rp.R, B = rp.B;
    const CrossUnits cunits = rp.beams ? units_of(unit_tab_beam) : cross_units();
    const int* skip = st.done;
    const unsigned abl = ablate_mask();
    const bool no_gemm = abl & 1, no_ln = abl & 2, no_self = abl & 4, no_cross = abl & 8, no_sample = abl & 16, no_embed = abl & 32;
    if (!no_embed)
      PF("embed_ln", s, [&] { embed_ln<T>(st.cur_token, st.cur_type, st.cur_pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, b.x, R,
                  DH, LN_EPS_BERT, s); });
    // every following kernel of the step is a programmatic dependent launch of its predecessor (common.cuh)
    struct PdlScope {
      explicit PdlScope(bool on) { g_pdl = on; }
      ~PdlScope() { g_pdl = false; }
    } pdl_scope(chain_pdl() && !profiling);
    auto G = [&](const T* A, int lda, const Lin& L, void* C, int ldc, int act) {
      if (!no_gemm) gemm(A, lda, L, C, ldc, R, act, nullptr, 0, false, skip, s);
    };
    auto GL = [&](const T* A, int lda, const Lin& L, int act, const T* res, const LNp& ln, T* out) {
      if (no_gemm && no_ln) return;
      if (no_gemm || no_ln) {   // ablation only: one half of the pair
        if (!no_gemm) gemm(A, lda, L, out, DH, R, act, res, DH, false, skip, s);
        if (!no_ln) PF("layernorm", s, [&] { layernorm<T>(out, DH, out, DH, ln.g, ln.b, R, DH, LN_EPS_BERT, s); });
        return;
      }
      gemm_ln(A, lda, L, act, res, DH, ln, out, DH, R, skip, s);
    };
    for (int l = 0; l < cfg.dec_layers; ++l) {
      const BertLayerW& w = dec.layers[l];
      G(b.x, DH, w.qkv, b.qkv, 3 * DH, ACT_NONE);
      if (!no_self)
        PF("self_attn", s, [&] { decode_self_attention<T>(b.qkv, self_k + l * self_layer_stride(), self_v + l * self_layer_stride(), b.ctx, st, R,
                                 rp.P, Lmax, self_ws, self_tickets, attn_maps_ptr, l, s); });
      GL(b.ctx, DH, w.o, ACT_NONE, b.x, w.ln1, b.x1);
      G(b.x1, DH, w.cq, b.qkv, DH, ACT_NONE);
      const T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
      // the grid covers cross_max_units so that the captured graph does not depend on the batch's image counts
      if (!no_cross)
        PF("cross_attn", s, [&] { decode_cross_attention<T>(b.qkv, DH, kvl, kvl + NHEAD * cross_head_stride(), cross_head_stride(), b.ctx,
                                  cunits, st, R, B, cross_ws, cross_tickets, attn_maps_ptr, l, s); });
      GL(b.ctx, DH, w.co, ACT_NONE, b.x1, w.ln2, b.x);
      G(b.x, DH, w.fc1, b.hid, DFF, ACT_GELU);
      GL(b.hid, DFF, w.fc2, ACT_NONE, b.x, w.ln3, b.x);
    }
    if (!no_gemm || !no_ln) {
      if (no_gemm || no_ln) {
        if (!no_ln) PF("layernorm", s, [&] { layernorm<T>(head_tmp, DH, head_tmp, DH, dec_head_ln.g, dec_head_ln.b, R, DH, LN_EPS_BERT, s); });
        if (!no_gemm) {
          gemm(b.x, DH, dec_head_t, head_tmp, DH, R, ACT_GELU, nullptr, 0, false, skip, s);
          gemm(head_tmp, DH, dec_lm, logits, cfg.vocab, R, ACT_NONE, nullptr, 0, true, skip, s);
        }
      } else {
        lm_head(b.x, R, head_tmp, logits, cfg.vocab, skip, s);
      }
    }
    if (rp.beams) {
      PF("beam_step", s, [&] { beam_step(st, rp, bs, logits, cfg.vocab, s); });
      PF("beam_reorder", s, [&] { beam_reorder_kv<T>(self_k, self_v, beam_scratch, st, bs, R, Lmax, rp.Tmax, cfg.dec_layers,
                                                      self_layer_stride(), s); });
    } else if (!no_sample) {
      PF("sample", s, [&] { sample_step(st, rp, logits, cfg.vocab, noise, s); });
    }
  }

  void rollout(const cxrm_rollout_args& a, cudaStream_t s_user) override {
    CXRM_CHECK(finalized, "weights not finalized");
    // stream capture is illegal on the legacy default stream: hop to the engine's own stream
    cudaStream_t s = s_user;
    const bool hop = cfg.use_cuda_graph && (s_user == nullptr || s_user == cudaStreamLegacy);
    if (hop) {
      s = side_stream;
      CXRM_CUDA_CHECK(cudaEventRecord(ev_a, s_user));
      CXRM_CUDA_CHECK(cudaStreamWaitEvent(s, ev_a, 0));
    }
    CXRM_CHECK(a.mode >= 1 && a.mode <= 3, "mode");
    CXRM_CHECK(kv_B == a.B && kv_total > 0, "cxrm_rollout needs cxrm_prefill_cross_kv for the same B");
    const int nm = a.mode == CXRM_BOTH ? 2 : 1;
    const int R = a.B * nm, P = a.P, Tn = a.max_new_tokens;
    CXRM_CHECK(R <= Rmax && R < 65536 && P >= 1 && P <= cfg.max_prompt && Tn >= 1 && Tn <= cfg.max_new_tokens, "rollout shape");
    CXRM_CHECK(a.n_special_sample <= kMaxSpecial && a.n_special_greedy <= kMaxSpecial, "too many special tokens");
    RolloutParams rp{};
    rp.R = R; rp.B = a.B; rp.P = P; rp.Lmax = Lmax; rp.Tmax = Tn; rp.V = cfg.vocab;
    auto set_blk = [&](int blk, bool greedy) {
      rp.mode_of_block[blk] = greedy ? 1 : 0;
      rp.n_special[blk] = greedy ? a.n_special_greedy : a.n_special_sample;
      for (int i = 0; i < rp.n_special[blk]; ++i) rp.special_ids[blk][i] = greedy ? a.special_greedy[i] : a.special_sample[i];
      for (int i = 0; i <= rp.n_special[blk]; ++i) rp.sections[blk][i] = greedy ? a.sections_greedy[i] : a.sections_sample[i];
    };
    if (a.mode == CXRM_BOTH) {
      set_blk(0, false);
      set_blk(1, true);
    } else {
      set_blk(0, a.mode == CXRM_GREEDY);
    }
    rp.mask_token_id = a.mask_token_id; rp.eos = a.eos_token_id; rp.pad = a.pad_token_id;
    rp.top_k = a.top_k; rp.temperature = a.temperature; rp.seed = a.seed;
    rp.want_margin = a.margins ? 1 : 0;
    // the kernels index logprob/topk buffers with Tmax = Tn.  Steps after every row finished are skipped inside the
    // kernels (`*done`), so the [R,Tn] outputs are cleared first: log-prob 0 where PAD (cxrm.h), no stale columns.
    {
      const size_t rt0 = static_cast<size_t>(R) * Tn;
      CXRM_CUDA_CHECK(cudaMemsetAsync(st.logprob, 0, rt0 * sizeof(float), s));
      CXRM_CUDA_CHECK(cudaMemsetAsync(st.margin, 0, rt0 * sizeof(float), s));
      CXRM_CUDA_CHECK(cudaMemsetAsync(st.topk_cnt, 0, rt0 * sizeof(int), s));
    }
    PF("init", s, [&] { rollout_init(st, rp, a.prompt_ids, pre_ids, pre_types, pre_pos, pre_valid, s); });

    arena.reset();
    const long long M = static_cast<long long>(R) * P;
    DecBufs db = dec_bufs(R);     // first: keeps the decode-step pointers (and the captured graph) independent of P
    T* last = arena.get<T>(static_cast<long long>(R) * DH);
    T* head_tmp = arena.get<T>(static_cast<long long>(R) * DH);
    // Prompt pass over the PACKED prompts: right-padded (masked) columns cost nothing - the benchmark's prompts fill 40 %
    // of their 64 x 247 grid - except each padded row's last column, which the reference still uses as the query of the
    // first new token (kernels.h PromptPack).  One small device -> host read (the packed token count) sizes the launches.
    static const bool packed_prefill = std::getenv("CXRM_NO_PACKED_PREFILL") == nullptr;
    if (packed_prefill) {
      PromptPack pk;
      pk.ids = arena.get<int>(M); pk.types = arena.get<int>(M); pk.pos = arena.get<int>(M);
      pk.tok_row = arena.get<int>(M); pk.tok_slot = arena.get<int>(M);
      pk.tok_cache = arena.get<uint8_t>(M);
      pk.row_off = arena.get<int>(R); pk.row_lq = arena.get<int>(R); pk.row_lk = arena.get<int>(R); pk.last_idx = arena.get<int>(R);
      pk.total = arena.get<int>(1);
      PF("pack", s, [&] { pack_prompt(pk, pre_ids, pre_types, pre_pos, pre_valid, R, P, s); });
      int total = 0;
      CXRM_CUDA_CHECK(cudaMemcpyAsync(&total, pk.total, sizeof(int), cudaMemcpyDeviceToHost, s));
      CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
      CXRM_CHECK(total >= R && total <= M, "packed prompt size");
      DecBufs pb = dec_bufs(total);
      PF("embed_ln", s, [&] { embed_ln<T>(pk.ids, pk.types, pk.pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, pb.x, total, DH,
                  LN_EPS_BERT, s); });
      T* hid = decoder_full(pb, R, P, a.B, nullptr, 0, /*store_cache=*/true, s, &pk, total);
      PF("take_last", s, [&] { gather_rows<T>(hid, pk.last_idx, last, R, DH, s); });
    } else {
      DecBufs pb = dec_bufs(M);
      PF("embed_ln", s, [&] { embed_ln<T>(pre_ids, pre_types, pre_pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, pb.x, M, DH,
                  LN_EPS_BERT, s); });
      T* hid = decoder_full(pb, R, P, a.B, pre_valid, P, /*store_cache=*/true, s);
      PF("take_last", s, [&] { take_last_token<T>(hid, last, R, P, DH, s); });
    }
    lm_head(last, R, head_tmp, logits, cfg.vocab, nullptr, s);
    PF("sample", s, [&] { sample_step(st, rp, logits, cfg.vocab, a.exp_noise, s); });

    if (!prefill_done_ev) CXRM_CUDA_CHECK(cudaEventCreate(&prefill_done_ev));
    CXRM_CUDA_CHECK(cudaEventRecord(prefill_done_ev, s));   // prompt pass + first token done; decode steps follow
    prefill_recorded = true;
    if (Tn > 1) {
      if (cfg.use_cuda_graph && !profiling) {
        run_decode_graph(db, head_tmp, rp, a.exp_noise, Tn - 1, s);
      } else {
        for (int t = 1; t < Tn; ++t) decode_step(db, head_tmp, rp, a.exp_noise, s);
      }
    }
    // outputs
    const int Lout = P + Tn;
    CXRM_CUDA_CHECK(cudaMemcpy2DAsync(a.sequences, Lout * sizeof(int), st.seq, Lmax * sizeof(int), Lout * sizeof(int), R,
                                      cudaMemcpyDeviceToDevice, s));
    const size_t rt = static_cast<size_t>(R) * Tn;
    if (a.logprobs) CXRM_CUDA_CHECK(cudaMemcpyAsync(a.logprobs, st.logprob, rt * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (a.margins) CXRM_CUDA_CHECK(cudaMemcpyAsync(a.margins, st.margin, rt * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (a.topk_idx) CXRM_CUDA_CHECK(cudaMemcpyAsync(a.topk_idx, st.topk_idx, rt * kTopKCap * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (a.topk_val) CXRM_CUDA_CHECK(cudaMemcpyAsync(a.topk_val, st.topk_val, rt * kTopKCap * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (a.topk_cnt) CXRM_CUDA_CHECK(cudaMemcpyAsync(a.topk_cnt, st.topk_cnt, rt * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (a.last_logits)
      CXRM_CUDA_CHECK(cudaMemcpyAsync(a.last_logits, logits, static_cast<size_t>(R) * cfg.vocab * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (chain_trace) {
      static int dumped = 0;
      if (dumped++ == 2) dump_chain_trace(s);   // third rollout of the process: warm
    }
    if (hop) {
      CXRM_CUDA_CHECK(cudaEventRecord(ev_b, s));
      CXRM_CUDA_CHECK(cudaStreamWaitEvent(s_user, ev_b, 0));
    }
    if (a.steps_out) {
      int steps = 0;
      CXRM_CUDA_CHECK(cudaMemcpyAsync(&steps, st.step, sizeof(int), cudaMemcpyDeviceToHost, s));
      CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
      *a.steps_out = steps;
    }
  }

  // The decode step reads everything that changes from device memory, so one captured
  // graph serves every step of every rollout with the same (R, B, buffers, parameters).
  void run_decode_graph(DecBufs& db, T* head_tmp, const RolloutParams& rp, const float* noise, int n_steps,
                        cudaStream_t s) {
    GraphKey key;
    std::memset(&key, 0, sizeof(key));
    key.R = rp.R; key.B = rp.B; key.P = rp.P; key.Tmax = rp.Tmax; key.top_k = rp.top_k;
    key.temperature = rp.temperature; key.noise = noise; key.buf = db.x;
    key.mask_id = rp.mask_token_id; key.eos = rp.eos; key.pad = rp.pad; key.want_margin = rp.want_margin;
    key.beams = rp.beams; key.length_penalty = rp.length_penalty; key.beam_scratch = rp.beams ? beam_scratch : nullptr;
    std::memcpy(key.special, rp.special_ids, sizeof(key.special));
    std::memcpy(key.sections, rp.sections, sizeof(key.sections));
    std::memcpy(key.nspecial, rp.n_special, sizeof(key.nspecial));
    std::memcpy(key.modes, rp.mode_of_block, sizeof(key.modes));
    if (!graph_exec || !graph_valid || std::memcmp(&key, &graph_key, sizeof(GraphKey)) != 0) {
      if (graph_exec) {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
      }
      // one eager step first: sets the kernels' function attributes outside the capture
      decode_step(db, head_tmp, rp, noise, s);
      --n_steps;
      cudaGraph_t graph = nullptr;
      const unsigned long long before = g_launch_count;
      CXRM_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      try {
        decode_step(db, head_tmp, rp, noise, s);
      } catch (...) {
        cudaStreamEndCapture(s, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      CXRM_CUDA_CHECK(cudaStreamEndCapture(s, &graph));
      graph_nodes = g_launch_count - before;
      g_launch_count = before;
      CXRM_CUDA_CHECK(cudaGraphInstantiate(&graph_exec, graph, 0));
      cudaGraphDestroy(graph);
      std::memcpy(&graph_key, &key, sizeof(GraphKey));
      graph_valid = true;
    }
    for (int t = 0; t < n_steps; ++t) CXRM_CUDA_CHECK(cudaGraphLaunch(graph_exec, s));
    g_launch_count += graph_nodes * static_cast<unsigned long long>(n_steps);
  }

  // =========================================================================== beam search
  // HF _beam_search over the KV-cached rollout (include/cxrm.h cxrm_rollout_beam; kernels and bookkeeping: beam.cu).
  void ensure_beam_buffers(int R, int Tn) {
    if (!unit_tab_beam) {
      unit_tab_beam = dalloc<int>(4LL * cross_max_units + 2LL * cfg.max_studies + 1);
      const long long nbm = static_cast<long long>(cfg.max_studies);
      bs.run_score = dalloc<float>(nbm);
      bs.fin_score = dalloc<float>(nbm);
      bs.fin_len = dalloc<int>(nbm);
      bs.is_fin = dalloc<uint8_t>(nbm);
      bs.can_improve = dalloc<uint8_t>(nbm);
      bs.fin_seq = dalloc<int>(nbm * kBeamMaxT);
      bs.src_row = dalloc<int>(nbm);
      bs.n_slots = dalloc<int>(4);
      bs.arrive = reinterpret_cast<unsigned*>(bs.n_slots + 1);
      bs.cnt_can = bs.n_slots + 2;
      bs.cnt_hit = bs.n_slots + 3;
    }
    const size_t need = beam_scratch_elems(R, Tn, cfg.dec_layers);
    if (need > beam_scratch_cap) {
      if (beam_scratch) CXRM_CUDA_CHECK(cudaFree(beam_scratch));
      beam_scratch = nullptr;
      CXRM_CUDA_CHECK(cudaMalloc(&beam_scratch, need * sizeof(T)));
      beam_scratch_cap = need;
      graph_valid = false;
    }
  }
  void rollout_beam(const cxrm_beam_args& a, cudaStream_t s_user) override {
    CXRM_CHECK(finalized, "weights not finalized");
    cudaStream_t s = s_user;
    const bool hop = cfg.use_cuda_graph && (s_user == nullptr || s_user == cudaStreamLegacy);
    if (hop) {
      s = side_stream;
      CXRM_CUDA_CHECK(cudaEventRecord(ev_a, s_user));
      CXRM_CUDA_CHECK(cudaStreamWaitEvent(s, ev_a, 0));
    }
    const int nb = a.num_beams, B = a.B, P = a.P, Tn = a.max_new_tokens;
    CXRM_CHECK(nb >= 2 && nb <= kMaxBeams, "num_beams must be 2..8 (num_beams == 1 is cxrm_rollout's greedy mode)");
    CXRM_CHECK(kv_B == B && kv_total > 0, "cxrm_rollout_beam needs cxrm_prefill_cross_kv for the same B");
    const int R = B * nb;
    CXRM_CHECK(R <= cfg.max_studies, "beam search: B * num_beams must not exceed the engine's max_studies");
    CXRM_CHECK(P >= 1 && P <= cfg.max_prompt && Tn >= 1 && Tn <= cfg.max_new_tokens && Tn <= kBeamMaxT, "rollout shape");
    CXRM_CHECK(a.n_special <= kMaxSpecial, "too many special tokens");
    CXRM_CHECK(a.sequences != nullptr && a.prompt_ids != nullptr, "sequences / prompt_ids");
    ensure_beam_buffers(R, Tn);
    RolloutParams rp{};
    rp.R = R; rp.B = R; rp.P = P; rp.Lmax = Lmax; rp.Tmax = Tn; rp.V = cfg.vocab;
    rp.mode_of_block[0] = 1;
    rp.n_special[0] = a.n_special;
    for (int i = 0; i < a.n_special; ++i) rp.special_ids[0][i] = a.special[i];
    for (int i = 0; i <= a.n_special; ++i) rp.sections[0][i] = a.sections[i];
    rp.mask_token_id = a.mask_token_id; rp.eos = a.eos_token_id; rp.pad = a.pad_token_id;
    rp.top_k = 0; rp.temperature = 1.0f; rp.seed = 0; rp.want_margin = 0;
    rp.beams = nb; rp.length_penalty = a.length_penalty;
    // virtual studies: beam j of study b is row j * B + b; its cross-attention units are those of study b
    {
      std::vector<int> tab(4 * static_cast<size_t>(cross_max_units) + 2 * cfg.max_studies + 1, 0);
      int* u_study = tab.data();
      int* u_j0 = u_study + cross_max_units;
      int* u_n = u_j0 + cross_max_units;
      int* u_chunk = u_n + cross_max_units;
      int* n_chunks = u_chunk + cross_max_units;
      int* first_unit = n_chunks + cfg.max_studies;
      int nu = 0;
      for (int v = 0; v < R; ++v) {
        const int b = v % B;
        const int nc = ceil_div(h_kv_len[b], attn_ch);
        n_chunks[v] = nc;
        first_unit[v] = nu;
        for (int c = 0; c < nc; ++c, ++nu) {
          CXRM_CHECK(nu < cross_max_units, "beam search: too many cross-attention units");
          u_study[nu] = v;
          u_j0[nu] = h_kv_off[b] + c * attn_ch;
          u_n[nu] = std::min(attn_ch, h_kv_len[b] - c * attn_ch);
          u_chunk[nu] = c;
        }
      }
      first_unit[cfg.max_studies] = nu;
      CXRM_CUDA_CHECK(cudaMemcpyAsync(unit_tab_beam, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));
      CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    for (int j = 0; j < nb; ++j)
      CXRM_CUDA_CHECK(cudaMemcpyAsync(prompt_dev + static_cast<long long>(j) * B * P, a.prompt_ids, sizeof(int) * B * P,
                                      cudaMemcpyDeviceToDevice, s));
    PF("init", s, [&] { rollout_init(st, rp, prompt_dev, pre_ids, pre_types, pre_pos, pre_valid, s); });
    beam_init(bs, B, nb, Tn, a.pad_token_id ? a.pad_token_id : a.eos_token_id, s);

    arena.reset();
    const long long M = static_cast<long long>(R) * P;
    DecBufs db = dec_bufs(R);
    T* last = arena.get<T>(static_cast<long long>(R) * DH);
    T* head_tmp = arena.get<T>(static_cast<long long>(R) * DH);
    DecBufs pb = dec_bufs(M);
    phase = "prefill";
    PF("embed_ln", s, [&] { embed_ln<T>(pre_ids, pre_types, pre_pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, pb.x, M, DH,
                LN_EPS_BERT, s); });
    T* hid = decoder_full(pb, R, P, B, pre_valid, P, /*store_cache=*/true, s);   // row r attends the encoder K/V of study r % B
    PF("take_last", s, [&] { take_last_token<T>(hid, last, R, P, DH, s); });
    lm_head(last, R, head_tmp, logits, cfg.vocab, nullptr, s);
    PF("beam_step", s, [&] { beam_step(st, rp, bs, logits, cfg.vocab, s); });   // step 0: no generated cache slot to move yet
    if (Tn > 1) {
      if (cfg.use_cuda_graph && !profiling) {
        run_decode_graph(db, head_tmp, rp, nullptr, Tn - 1, s);
      } else {
        for (int t = 1; t < Tn; ++t) decode_step(db, head_tmp, rp, nullptr, s);
      }
    }
    beam_finalize(st, bs, B, nb, P, Lmax, Tn, a.sequences, a.scores, a.lengths, s);
    if (hop) {
      CXRM_CUDA_CHECK(cudaEventRecord(ev_b, s));
      CXRM_CUDA_CHECK(cudaStreamWaitEvent(s_user, ev_b, 0));
    }
    if (a.steps_out) {
      int steps = 0;
      CXRM_CUDA_CHECK(cudaMemcpyAsync(&steps, st.step, sizeof(int), cudaMemcpyDeviceToHost, s));
      CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
      *a.steps_out = steps;
    }
  }

  // =========================================================================== teacher-forced forward
  void decoder_forward(const int* ids, const int* tt, const int* pos, const uint8_t* key_mask, int R, int L, int B,
                       bool last_only, float* logits_out, cudaStream_t s) override {
    CXRM_CHECK(finalized, "weights not finalized");
    CXRM_CHECK(kv_B == B && kv_total > 0, "cxrm_decoder_forward needs cxrm_prefill_cross_kv for the same B");
    CXRM_CHECK(R >= 1 && R % B == 0 && L >= 1 && L <= 512, "decoder_forward shape");
    phase = "fwd";
    arena.reset();
    const long long M = static_cast<long long>(R) * L;
    DecBufs b = dec_bufs(M);
    PF("embed_ln", s, [&] { embed_ln<T>(ids, tt, pos, dec.word, dec.type, dec.pos, dec.emb_ln.g, dec.emb_ln.b, b.x, M, DH, LN_EPS_BERT, s); });
    T* hid = decoder_full(b, R, L, B, key_mask, L, /*store_cache=*/false, s);
    if (last_only) {
      T* last = b.ctx;
      PF("take_last", s, [&] { take_last_token<T>(hid, last, R, L, DH, s); });
      lm_head(last, R, b.x1, logits_out, cfg.vocab, nullptr, s);
    } else {
      lm_head(hid, M, b.x1, logits_out, cfg.vocab, nullptr, s);
    }
  }


  // =========================================================================== teacher-forced forward + backward
  // (train.cu has the non-GEMM kernels and the derivation; include/cxrm.h cxrm_train_step the contract)
  struct GradSlot { std::string name; long long offset, numel; int stage; };
  int train_stages() const override { return cfg.dec_layers + 2; }
  const std::vector<GradSlot>& slots(bool lora_only) const { return grad_slots[lora_only ? 1 : 0]; }
  int grad_count(bool lora_only) const override { return static_cast<int>(slots(lora_only).size()); }
  long long grad_total(bool lora_only) const override {
    const auto& v = slots(lora_only);
    return v.empty() ? 0 : v.back().offset + v.back().numel;
  }
  bool grad_info(bool lora_only, int i, std::string* name, long long* offset, long long* numel, int* stage) const override {
    const auto& v = slots(lora_only);
    if (i < 0 || i >= static_cast<int>(v.size())) return false;
    *name = v[i].name; *offset = v[i].offset; *numel = v[i].numel; *stage = v[i].stage;
    return true;
  }
  void build_grad_layout() {
    for (int lo = 0; lo < 2; ++lo) {
      auto& v = grad_slots[lo];
      v.clear();
      long long off = 0;
      auto add = [&](const std::string& n, long long ne, int stage) {
        v.push_back({n, off, ne, stage});
        off += ne;
      };
      const int NL = cfg.dec_layers;
      if (!lo) {
        const std::string t = "decoder.cls.predictions.";
        add(t + "transform.dense.weight", 1LL * DH * DH, 0);
        add(t + "transform.dense.bias", DH, 0);
        add(t + "transform.LayerNorm.weight", DH, 0);
        add(t + "transform.LayerNorm.bias", DH, 0);
        add(t + "bias", cfg.vocab, 0);
      }
      for (int l = NL - 1; l >= 0; --l) {
        const int st = 1 + (NL - 1 - l);
        const std::string p = "decoder.bert.encoder.layer." + std::to_string(l) + ".";
        if (lo) {
          for (const char* nm : {"query", "key"}) {
            if (lora_r[l] <= 0) continue;
            add(p + "attention.self." + nm + ".lora_A.weight", 1LL * lora_r[l] * DH, st);
            add(p + "attention.self." + nm + ".lora_B.weight", 1LL * DH * lora_r[l], st);
          }
          continue;
        }
        for (const char* nm : {"query", "key", "value"}) add(p + "attention.self." + nm + ".weight", 1LL * DH * DH, st);   // contiguous = fused q|k|v
        for (const char* nm : {"query", "key", "value"}) add(p + "attention.self." + nm + ".bias", DH, st);
        add(p + "attention.output.dense.weight", 1LL * DH * DH, st);
        add(p + "attention.output.dense.bias", DH, st);
        add(p + "attention.output.LayerNorm.weight", DH, st);
        add(p + "attention.output.LayerNorm.bias", DH, st);
        add(p + "crossattention.self.query.weight", 1LL * DH * DH, st);
        add(p + "crossattention.self.query.bias", DH, st);
        for (const char* nm : {"key", "value"}) add(p + "crossattention.self." + nm + ".weight", 1LL * DH * DH, st);     // contiguous = fused k|v
        for (const char* nm : {"key", "value"}) add(p + "crossattention.self." + nm + ".bias", DH, st);
        add(p + "crossattention.output.dense.weight", 1LL * DH * DH, st);
        add(p + "crossattention.output.dense.bias", DH, st);
        add(p + "crossattention.output.LayerNorm.weight", DH, st);
        add(p + "crossattention.output.LayerNorm.bias", DH, st);
        add(p + "intermediate.dense.weight", 1LL * DFF * DH, st);
        add(p + "intermediate.dense.bias", DFF, st);
        add(p + "output.dense.weight", 1LL * DH * DFF, st);
        add(p + "output.dense.bias", DH, st);
        add(p + "output.LayerNorm.weight", DH, st);
        add(p + "output.LayerNorm.bias", DH, st);
      }
      if (!lo) {
        const std::string e = "decoder.bert.embeddings.";
        const int st = NL + 1;
        add(e + "position_embeddings.weight", 512LL * DH, st);
        add(e + "token_type_embeddings.weight", 2LL * DH, st);
        add(e + "LayerNorm.weight", DH, st);
        add(e + "LayerNorm.bias", DH, st);
        add(e + "word_embeddings.weight", 1LL * cfg.vocab * DH, st);   // tied: LM-head gradient + embedding gradient
      }
    }
  }
  long long slot_off(bool lo, const std::string& name) const {
    for (const auto& s : slots(lo))
      if (s.name == name) return s.offset;
    throw std::runtime_error("no gradient slot " + name);
  }
  T* transposed(const Lin& L) {     // W [n_out, n_in] -> W^T [n_in, n_out], cached (weights are constants between loads)
    auto it = wt_cache.find(L.w);
    if (it != wt_cache.end()) return it->second;
    T* t = dalloc<T>(1LL * L.n_out * L.n_in);
    transpose<T>(L.w, L.n_in, t, L.n_out, L.n_out, L.n_in, 0);
    CXRM_CUDA_CHECK(cudaStreamSynchronize(0));
    wt_cache[L.w] = t;
    return t;
  }

  struct Tape { T *x, *qkv, *ctx, *x1_pre, *x1, *cq, *cctx, *x2_pre, *x2, *h_pre, *hid, *x3_pre; };
  struct TrainState {
    cxrm_train_args a{};
    long long M = 0;
    int next_stage = 0;
    std::vector<Tape> tp;
    T *emb_pre = nullptr, *x0 = nullptr, *xL = nullptr, *t_pre = nullptr, *t_act = nullptr, *t_ln = nullptr;
    T *dx = nullptr, *d1 = nullptr, *d2 = nullptr, *dbig = nullptr, *dqkv = nullptr, *tA = nullptr, *tB = nullptr;
    T *dkv = nullptr;
    float *lse = nullptr, *Dd = nullptr, *row_loss = nullptr;
    float2* lnstats = nullptr;
    int* n_counted = nullptr;
  } tr;

  AttnArgs self_attn_args(const T* qkv, T* ctx, int R, int L, const uint8_t* key_mask) const {
    AttnArgs a{};
    a.q = qkv; a.k = qkv + DH; a.v = qkv + 2 * DH; a.o = ctx;
    a.q_bs = static_cast<long long>(L) * 3 * DH; a.q_hs = 64; a.q_ts = 3 * DH;
    a.k_bs = a.q_bs; a.k_hs = 64; a.k_ts = 3 * DH;
    a.v_bs = a.q_bs; a.v_hs = 64; a.v_ts = 3 * DH;
    a.o_bs = static_cast<long long>(L) * DH; a.o_hs = 64; a.o_ts = DH;
    a.batch = R; a.heads = NHEAD; a.Lq = L; a.Lk = L;
    a.key_mask = key_mask; a.key_mask_ld = L; a.key_mask_per_q_batch = 1;
    a.causal = 1; a.q_pos_offset = 0; a.scale = 0.125f;
    return a;
  }
  AttnArgs cross_attn_args(const T* q, T* ctx, const T* kvl, int R, int L, int B) const {
    AttnArgs c{};
    c.q = q; c.k = kvl; c.v = kvl + NHEAD * cross_head_stride(); c.o = ctx;
    c.q_bs = static_cast<long long>(L) * DH; c.q_hs = 64; c.q_ts = DH;
    c.k_bs = 0; c.k_hs = cross_head_stride(); c.k_ts = 64;
    c.v_bs = 0; c.v_hs = cross_head_stride(); c.v_ts = 64;
    c.o_bs = c.q_bs; c.o_hs = 64; c.o_ts = DH;
    c.batch = R; c.heads = NHEAD; c.Lq = L; c.Lk = kv_maxlen;
    c.Lk_per_batch = kv_len; c.kv_offset = kv_off; c.kv_batch_mod = B;
    c.scale = 0.125f;
    return c;
  }

  void train_step(const cxrm_train_args& a, int stage, cudaStream_t s) override {
    CXRM_CHECK(finalized, "weights not finalized");
    CXRM_CHECK(cfg.max_train_tokens > 0, "the engine was created without a training workspace (cxrm_config.max_train_tokens)");
    const int NL = cfg.dec_layers, NS = train_stages();
    CXRM_CHECK(stage >= -1 && stage < NS, "stage");
    if (stage == -1) {
      for (int st = 0; st < NS; ++st) train_stage(a, st, s);
      return;
    }
    train_stage(a, stage, s);
    (void)NL;
  }

  // rank-r LoRA contractions (N or M = r = 8): strict-FMA kernel in both modes, the shapes are far below a tensor-core tile
  void gemm_small(const T* A, int lda, const T* W, int n_out, int n_in, void* C, int ldc, long long M, bool out_f32, cudaStream_t s) {
    GemmArgs g;
    g.c_head_stride = 0; g.trace = nullptr;
    g.A = A; g.lda = lda; g.W = W; g.ldw = n_in; g.C = C; g.ldc = ldc; g.M = static_cast<int>(M); g.N = n_out; g.K = n_in;
    g.bias = nullptr; g.act = ACT_NONE; g.residual = nullptr; g.ldr = 0; g.out_f32 = out_f32 ? 1 : 0; g.skip_flag = nullptr;
    PF("lora", s, [&] { gemm_simt<T>(g, s); });
  }
  // dX[M, K] = dY[M, N] . W  (W [N, K]; wt = W^T [K, N])
  void dgrad(const T* dY, int ldy, const Lin& L, T* dX, long long M, const T* add_res, cudaStream_t s) {
    Lin t;
    t.w = transposed(L); t.b = nullptr; t.n_out = L.n_in; t.n_in = L.n_out;
    gemm(dY, ldy, t, dX, L.n_in, M, ACT_NONE, add_res, L.n_in, false, nullptr, s, "dgrad");
  }
  // dW[N, K] (fp32) = dY^T . X; db[N] = column sums of dY
  void wgrad(const T* dY, int ldy, int N, const T* X, int ldx, int K, long long M, float* dW, float* db, cudaStream_t s) {
    transpose<T>(dY, ldy, tr.tA, M, M, N, s);          // [N, M]
    transpose<T>(X, ldx, tr.tB, M, M, K, s);           // [K, M]
    Lin t;
    t.w = tr.tB; t.b = nullptr; t.n_out = K; t.n_in = static_cast<int>(M);
    gemm(tr.tA, static_cast<int>(M), t, dW, K, N, ACT_NONE, nullptr, 0, true, nullptr, s, "wgrad");
    if (db) colsum<T>(dY, ldy, M, N, db, false, s);
  }

  void train_stage(const cxrm_train_args& a, int stage, cudaStream_t s) {
    const int NL = cfg.dec_layers;
    const int R = a.R, L = a.L, B = kv_B;
    const long long M = static_cast<long long>(R) * L;
    const bool lo = a.lora_only != 0;
    float* G = a.grads;
    if (stage == 0) {
      CXRM_CHECK(kv_B > 0 && kv_total > 0 && R >= 1 && R % B == 0, "cxrm_train_step needs cxrm_prefill_cross_kv; rows must be a multiple of the studies");
      CXRM_CHECK(L >= 1 && L <= 512 && M % 8 == 0 && M <= cfg.max_train_tokens, "train batch: R * L must be a multiple of 8 and fit max_train_tokens");
      CXRM_CHECK(a.ids && a.token_type_ids && a.position_ids && a.key_mask && a.targets && a.loss_out && G, "train_step: null argument");
      CXRM_CHECK(a.loss_kind == 0 || (a.loss_kind == 1 && a.advantage), "loss_kind");
      phase = "train";
      tr = TrainState{};
      tr.a = a;
      tr.M = M;
      arena.reset();
      auto get = [&](long long n) { return arena.get<T>(n); };
      tr.emb_pre = get(M * DH); tr.x0 = get(M * DH);
      tr.tp.resize(NL);
      for (int l = 0; l < NL; ++l) {
        Tape& t = tr.tp[l];
        t.x = l == 0 ? tr.x0 : get(M * DH);
        t.qkv = get(M * 3 * DH); t.ctx = get(M * DH); t.x1_pre = get(M * DH); t.x1 = get(M * DH); t.cq = get(M * DH);
        t.cctx = get(M * DH); t.x2_pre = get(M * DH); t.x2 = get(M * DH); t.h_pre = get(M * DFF); t.hid = get(M * DFF);
        t.x3_pre = get(M * DH);
      }
      tr.xL = get(M * DH); tr.t_pre = get(M * DH); tr.t_act = get(M * DH); tr.t_ln = get(M * DH);
      tr.dx = get(M * DH); tr.d1 = get(M * DH); tr.d2 = get(M * DH); tr.dbig = get(M * DFF); tr.dqkv = get(M * 3 * DH);
      const long long tmax = std::max<long long>(M * DFF, lo ? 0 : static_cast<long long>(kv_total) * 2 * DH);
      tr.tA = get(tmax); tr.tB = get(tmax);
      tr.lse = arena.get<float>(static_cast<long long>(R) * NHEAD * L);
      tr.Dd = arena.get<float>(static_cast<long long>(R) * NHEAD * L);
      tr.row_loss = arena.get<float>(M);
      tr.lnstats = arena.get<float2>(M);
      tr.n_counted = arena.get<int>(1);
      if (!lo) tr.dkv = get(static_cast<long long>(kv_total) * 2 * DH);

      // ---------------- forward, keeping what the backward needs ----------------
      embed_sum<T>(a.ids, a.token_type_ids, a.position_ids, dec.word, dec.type, dec.pos, tr.emb_pre, M, DH, s);
      layernorm<T>(tr.emb_pre, DH, tr.x0, DH, dec.emb_ln.g, dec.emb_ln.b, M, DH, LN_EPS_BERT, s);
      for (int l = 0; l < NL; ++l) {
        const BertLayerW& w = dec.layers[l];
        Tape& t = tr.tp[l];
        T* x_next = l + 1 < NL ? tr.tp[l + 1].x : tr.xL;
        gemm(t.x, DH, w.qkv, t.qkv, 3 * DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
        attention(self_attn_args(t.qkv, t.ctx, R, L, a.key_mask), s);
        gemm(t.ctx, DH, w.o, t.x1_pre, DH, M, ACT_NONE, t.x, DH, false, nullptr, s);
        layernorm<T>(t.x1_pre, DH, t.x1, DH, w.ln1.g, w.ln1.b, M, DH, LN_EPS_BERT, s);
        gemm(t.x1, DH, w.cq, t.cq, DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
        const T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
        attention(cross_attn_args(t.cq, t.cctx, kvl, R, L, B), s);
        gemm(t.cctx, DH, w.co, t.x2_pre, DH, M, ACT_NONE, t.x1, DH, false, nullptr, s);
        layernorm<T>(t.x2_pre, DH, t.x2, DH, w.ln2.g, w.ln2.b, M, DH, LN_EPS_BERT, s);
        gemm(t.x2, DH, w.fc1, t.h_pre, DFF, M, ACT_NONE, nullptr, 0, false, nullptr, s);
        gelu_fwd<T>(t.h_pre, t.hid, M * DFF, s);
        gemm(t.hid, DFF, w.fc2, t.x3_pre, DH, M, ACT_NONE, t.x2, DH, false, nullptr, s);
        layernorm<T>(t.x3_pre, DH, x_next, DH, w.ln3.g, w.ln3.b, M, DH, LN_EPS_BERT, s);
      }
      gemm(tr.xL, DH, dec_head_t, tr.t_pre, DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
      gelu_fwd<T>(tr.t_pre, tr.t_act, M * DH, s);
      layernorm<T>(tr.t_act, DH, tr.t_ln, DH, dec_head_ln.g, dec_head_ln.b, M, DH, LN_EPS_BERT, s);
      // logits, loss and dlogits (the [M, V] buffers live only inside this stage: they reuse the arena tail)
      const long long V = cfg.vocab;
      float* lg = arena.get<float>(M * V);
      T* dz = arena.get<T>(M * V);
      gemm(tr.t_ln, DH, dec_lm, lg, static_cast<int>(V), M, ACT_NONE, nullptr, 0, true, nullptr, s);
      loss_head<T>(lg, M, static_cast<int>(V), a.targets, a.ignore_index, a.loss_kind, a.advantage, L, R, a.top_k,
                   a.temperature == 0.f ? 1.0f : a.temperature, dz, tr.row_loss, tr.n_counted, a.loss_out, s);
      // ---------------- LM head backward ----------------
      dgrad(dz, static_cast<int>(V), dec_lm, tr.d1, M, nullptr, s);                        // d t_ln = dz . E
      if (!lo) {
        T* dzt = arena.get<T>(M * V);                                                       // dE = dz^T . t_ln
        T* tlt = arena.get<T>(M * DH);
        transpose<T>(dz, V, dzt, M, M, static_cast<int>(V), s);
        transpose<T>(tr.t_ln, DH, tlt, M, M, DH, s);
        Lin t;
        t.w = tlt; t.b = nullptr; t.n_out = DH; t.n_in = static_cast<int>(M);
        gemm(dzt, static_cast<int>(M), t, G + slot_off(false, "decoder.bert.embeddings.word_embeddings.weight"), DH, V, ACT_NONE,
             nullptr, 0, true, nullptr, s, "wgrad");
        colsum<T>(dz, V, M, static_cast<int>(V), G + slot_off(false, "decoder.cls.predictions.bias"), false, s);
      }
      float* gG = lo ? nullptr : G + slot_off(false, "decoder.cls.predictions.transform.LayerNorm.weight");
      layernorm_bwd<T>(tr.t_act, tr.d1, dec_head_ln.g, LN_EPS_BERT, tr.d2, tr.lnstats, gG, gG ? gG + DH : nullptr, false, M, DH, s);
      gelu_bwd<T>(tr.t_pre, tr.d2, tr.d1, M * DH, s);                                       // d t_pre
      if (!lo)
        wgrad(tr.d1, DH, DH, tr.xL, DH, DH, M, G + slot_off(false, "decoder.cls.predictions.transform.dense.weight"),
              G + slot_off(false, "decoder.cls.predictions.transform.dense.bias"), s);
      dgrad(tr.d1, DH, dec_head_t, tr.dx, M, nullptr, s);                                   // dx = gradient of the trunk output
      tr.next_stage = 1;
      return;
    }
    CXRM_CHECK(stage == tr.next_stage && tr.M == M && tr.a.grads == a.grads, "cxrm_train_step stages must be called in order with the same arguments");
    if (stage <= NL) {
      const int l = NL - stage;
      const BertLayerW& w = dec.layers[l];
      Tape& t = tr.tp[l];
      const std::string p = "decoder.bert.encoder.layer." + std::to_string(l) + ".";
      auto g = [&](const std::string& n) { return lo ? nullptr : G + slot_off(false, p + n); };
      // LN3: x_{l+1} = LN(x3_pre)
      layernorm_bwd<T>(t.x3_pre, tr.dx, w.ln3.g, LN_EPS_BERT, tr.d1, tr.lnstats, g("output.LayerNorm.weight"), g("output.LayerNorm.bias"),
                       false, M, DH, s);                                                     // d1 = d x3_pre (also the residual branch to x2)
      if (!lo) wgrad(tr.d1, DH, DH, t.hid, DFF, DFF, M, g("output.dense.weight"), g("output.dense.bias"), s);
      dgrad(tr.d1, DH, w.fc2, tr.dbig, M, nullptr, s);                                      // d hid
      gelu_bwd<T>(t.h_pre, tr.dbig, tr.dbig, M * DFF, s);                                   // d h_pre
      if (!lo) wgrad(tr.dbig, DFF, DFF, t.x2, DH, DH, M, g("intermediate.dense.weight"), g("intermediate.dense.bias"), s);
      dgrad(tr.dbig, DFF, w.fc1, tr.d2, M, tr.d1, s);                                       // d x2 = d h_pre . W1 + d x3_pre
      // LN2: x2 = LN(x2_pre)
      layernorm_bwd<T>(t.x2_pre, tr.d2, w.ln2.g, LN_EPS_BERT, tr.d1, tr.lnstats, g("crossattention.output.LayerNorm.weight"),
                       g("crossattention.output.LayerNorm.bias"), false, M, DH, s);         // d1 = d x2_pre (residual branch to x1)
      if (!lo) wgrad(tr.d1, DH, DH, t.cctx, DH, DH, M, g("crossattention.output.dense.weight"), g("crossattention.output.dense.bias"), s);
      dgrad(tr.d1, DH, w.co, tr.d2, M, nullptr, s);                                         // d cctx
      const T* kvl = cross_kv + static_cast<long long>(l) * cross_layer_stride();
      AttnArgs ca = cross_attn_args(t.cq, t.cctx, kvl, R, L, B);
      if (!lo) {
        // d cq (packed [M, 768] in dqkv) and d(k | v) of the compact encoder tokens, token-major [kv_total, 1536], which
        // is the dY of the fused cross K|V projection: dW = d(k|v)^T . memory
        CXRM_CHECK(kv_total % 8 == 0, "cxrm_train_step (all parameters): the visible encoder-token count must be a multiple of 8");
        attention_bwd<T>(ca, tr.d2, tr.dqkv, tr.dkv, tr.dkv + DH, 0, 64, 2 * DH, tr.lse, tr.Dd, s);
        wgrad(tr.dkv, 2 * DH, 2 * DH, mem_compact, DH, DH, kv_total, g("crossattention.self.key.weight"),
              g("crossattention.self.key.bias"), s);
      } else {
        attention_bwd<T>(ca, tr.d2, tr.dqkv, nullptr, nullptr, 0, 0, 0, tr.lse, tr.Dd, s);
      }
      if (!lo) wgrad(tr.dqkv, DH, DH, t.x1, DH, DH, M, g("crossattention.self.query.weight"), g("crossattention.self.query.bias"), s);
      dgrad(tr.dqkv, DH, w.cq, tr.d2, M, tr.d1, s);                                         // d x1 = d cq . Wcq + d x2_pre
      // LN1: x1 = LN(x1_pre)
      layernorm_bwd<T>(t.x1_pre, tr.d2, w.ln1.g, LN_EPS_BERT, tr.d1, tr.lnstats, g("attention.output.LayerNorm.weight"),
                       g("attention.output.LayerNorm.bias"), false, M, DH, s);              // d1 = d x1_pre (residual branch to x)
      if (!lo) wgrad(tr.d1, DH, DH, t.ctx, DH, DH, M, g("attention.output.dense.weight"), g("attention.output.dense.bias"), s);
      dgrad(tr.d1, DH, w.o, tr.d2, M, nullptr, s);                                          // d ctx
      AttnArgs sa = self_attn_args(t.qkv, t.ctx, R, L, a.key_mask);
      attention_bwd<T>(sa, tr.d2, tr.dqkv, tr.dqkv + DH, tr.dqkv + 2 * DH, sa.k_bs, sa.k_hs, sa.k_ts, tr.lse, tr.Dd,
                       s);                                                                   // dq | dk | dv, token-major like qkv
      if (!lo) {
        wgrad(tr.dqkv, 3 * DH, 3 * DH, t.x, DH, DH, M, g("attention.self.query.weight"), g("attention.self.query.bias"), s);
      } else if (lora_r[l] > 0) {
        const int r = lora_r[l];
        const char* names[2] = {"query", "key"};
        for (int j = 0; j < 2; ++j) {
          const T* dY = tr.dqkv + j * DH;                      // [M, 768] slice, row pitch 2304
          const std::string lp = p + "attention.self." + names[j] + ".";
          float* dA = G + slot_off(true, lp + "lora_A.weight");
          float* dB = G + slot_off(true, lp + "lora_B.weight");
          T* P = tr.tB;                                        // X . A^T  [M, r]
          gemm_small(t.x, DH, lora_A[l][j], r, DH, P, r, M, false, s);
          T* Pt = tr.tB + M * r;                               // [r, M]
          transpose<T>(P, r, Pt, M, M, r, s);
          transpose<T>(dY, 3 * DH, tr.tA, M, M, DH, s);        // dY^T [768, M]
          gemm_small(tr.tA, static_cast<int>(M), Pt, r, static_cast<int>(M), dB, r, DH, true, s);            // dB = dY^T . P
          scale_f32(dB, LORA_SCALE, 1LL * DH * r, s);
          T* Q = tr.tB + 2 * M * r;                            // dY . B  [M, r]
          gemm_small(dY, 3 * DH, lora_Bt[l][j], r, DH, Q, r, M, false, s);
          T* Qt = tr.tB + 3 * M * r;                           // [r, M]
          transpose<T>(Q, r, Qt, M, M, r, s);
          transpose<T>(t.x, DH, tr.tA, M, M, DH, s);           // X^T [768, M]
          gemm_small(Qt, static_cast<int>(M), tr.tA, DH, static_cast<int>(M), dA, DH, r, true, s);           // dA = Q^T . X
          scale_f32(dA, LORA_SCALE, 1LL * r * DH, s);
        }
      }
      dgrad(tr.dqkv, 3 * DH, w.qkv, tr.dx, M, tr.d1, s);                                   // dx = dqkv . Wqkv + d x1_pre
      tr.next_stage = stage + 1;
      return;
    }
    // ---------------- embeddings ----------------
    if (!lo) {
      const std::string e = "decoder.bert.embeddings.";
      float* gG = G + slot_off(false, e + "LayerNorm.weight");
      layernorm_bwd<T>(tr.emb_pre, tr.dx, dec.emb_ln.g, LN_EPS_BERT, tr.d1, tr.lnstats, gG, gG + DH, false, M, DH, s);
      float* dpos = G + slot_off(false, e + "position_embeddings.weight");
      float* dtyp = G + slot_off(false, e + "token_type_embeddings.weight");
      CXRM_CUDA_CHECK(cudaMemsetAsync(dpos, 0, (512LL + 2) * DH * sizeof(float), s));      // position | token type are adjacent
      (void)dtyp;
      scatter_add_rows<T>(tr.d1, a.position_ids, dpos, M, DH, s);
      scatter_add_rows<T>(tr.d1, a.token_type_ids, G + slot_off(false, e + "token_type_embeddings.weight"), M, DH, s);
      scatter_add_rows<T>(tr.d1, a.ids, G + slot_off(false, e + "word_embeddings.weight"), M, DH, s);   // onto the LM-head part
    }
    tr.next_stage = 0;
  }

  // =========================================================================== reward model
  void reward_embed(const int* ids, const int* lens, int n, int L, float* emb_out, cudaStream_t s) override {
    CXRM_CHECK(finalized && have_reward, "reward model weights not loaded");
    CXRM_CHECK(n >= 1 && L >= 1 && L <= cfg.rwd_max_len && static_cast<long long>(n) * L <=
                   static_cast<long long>(std::max(cfg.rwd_max_seqs, 1)) * cfg.rwd_max_len, "reward batch too large");
    phase = "rwd";
    arena.reset();
    // Variable-length batch: the sequences are PACKED (tokens beyond a sequence's length are never computed: labels
    // are 32..256 tokens against 257-token generated reports, real reports are shorter still); attention runs per
    // sequence over its own tokens through the packed-query / ragged-key arguments, so no padding mask is needed.
    // One small device -> host read (the packed token count) sizes the launches.
    const long long Mmax = static_cast<long long>(n) * L;
    int* off = arena.get<int>(n);
    int* lenc = arena.get<int>(n);
    int* total_d = arena.get<int>(1);
    int* pk_ids = arena.get<int>(Mmax);
    int* pos = arena.get<int>(Mmax);
    seq_offsets_kernel<<<1, 32, 0, s>>>(lens, n, L, off, total_d);
    check_launch("seq_offsets");
    seq_pack_kernel<<<n, 32, 0, s>>>(ids, lens, off, L, pk_ids, pos, lenc);
    check_launch("seq_pack");
    int total = 0;
    CXRM_CUDA_CHECK(cudaMemcpyAsync(&total, total_d, sizeof(int), cudaMemcpyDeviceToHost, s));
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    CXRM_CHECK(total >= 1 && total <= Mmax, "reward batch: every sequence is empty");
    const long long M = total;
    DecBufs b = dec_bufs(M);
    PF("embed_ln", s, [&] { embed_ln<T>(pk_ids, nullptr, pos, rwd.word, rwd.type, rwd.pos, rwd.emb_ln.g, rwd.emb_ln.b, b.x, M, DH, LN_EPS_BERT, s); });
    for (int l = 0; l < cfg.rwd_layers; ++l) {
      const BertLayerW& w = rwd.layers[l];
      gemm(b.x, DH, w.qkv, b.qkv, 3 * DH, M, ACT_NONE, nullptr, 0, false, nullptr, s);
      AttnArgs a{};
      a.q = b.qkv; a.k = b.qkv + DH; a.v = b.qkv + 2 * DH; a.o = b.ctx;
      a.q_bs = 0; a.q_hs = 64; a.q_ts = 3 * DH;
      a.k_bs = 0; a.k_hs = 64; a.k_ts = 3 * DH;
      a.v_bs = 0; a.v_hs = 64; a.v_ts = 3 * DH;
      a.o_bs = 0; a.o_hs = 64; a.o_ts = DH;
      a.batch = n; a.heads = NHEAD; a.Lq = L; a.Lk = L;
      a.q_offset = off; a.Lq_per_batch = lenc;      // attention_mask from padding='longest' == (j < len)
      a.kv_offset = off; a.Lk_per_batch = lenc;
      a.total_q = a.total_kv = M;
      a.scale = 0.125f;
      PF("attn", s, [&] { attention(a, s); });
      gemm(b.ctx, DH, w.o, b.x1, DH, M, ACT_NONE, b.x, DH, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(b.x1, DH, b.x1, DH, w.ln1.g, w.ln1.b, M, DH, LN_EPS_BERT, s); });
      gemm(b.x1, DH, w.fc1, b.hid, DFF, M, ACT_GELU, nullptr, 0, false, nullptr, s);
      gemm(b.hid, DFF, w.fc2, b.x, DH, M, ACT_NONE, b.x1, DH, false, nullptr, s);
      PF("layernorm", s, [&] { layernorm<T>(b.x, DH, b.x, DH, w.ln3.g, w.ln3.b, M, DH, LN_EPS_BERT, s); });
    }
    // [CLS] rows (the first packed token of every sequence) -> 768 -> 128 GELU -> LN -> 128
    T* cls = b.qkv;
    PF("gather", s, [&] { gather_rows<T>(b.x, off, cls, n, DH, s); });
    T* h1 = b.ctx;
    gemm(cls, DH, rp1, h1, 128, n, ACT_GELU, nullptr, 0, false, nullptr, s);
    PF("layernorm", s, [&] { layernorm<T>(h1, 128, h1, 128, rp_ln.g, rp_ln.b, n, 128, LN_EPS_BERT, s); });
    gemm(h1, 128, rp2, emb_out, 128, n, ACT_NONE, nullptr, 0, true, nullptr, s);
  }

  // =========================================================================== host-buffer SCST step
  void set_id_map(const int* id_map_host, int n, int cls_id, int sep_id, int bos_id, int sep_dec_id,
                  int n_special) override {
    CXRM_CHECK(n == cfg.vocab, "id map must cover the decoder vocabulary");
    CXRM_CHECK(n_special >= 0 && n_special <= n, "n_special outside the decoder vocabulary");
    auto in_rwd = [&](int v) { return v >= 0 && v < cfg.rwd_vocab; };
    CXRM_CHECK(in_rwd(cls_id) && in_rwd(sep_id), "[CLS] / [SEP] id outside the reward vocabulary");
    for (int i = n_special; i < n; ++i)   // the embedding gather of the reward model does not bounds-check
      CXRM_CHECK(in_rwd(id_map_host[i]), "id_map[" + std::to_string(i) + "] outside the reward vocabulary");
    bridge_n_special = n_special;
    if (!id_map) id_map = dalloc<int>(n);
    CXRM_CUDA_CHECK(cudaMemcpy(id_map, id_map_host, n * sizeof(int), cudaMemcpyHostToDevice));
    bridge_cls = cls_id;
    bridge_sep = sep_id;
    bridge_bos = bos_id;
    bridge_sep_dec = sep_dec_id;
  }

  // generated ids [R, L] -> reward-model ids [R, Lout] ([CLS] findings impression [SEP], specials dropped) + lengths
  void bridge_ids(const int* seq, int R, int L, int eos, int* out_ids, int* out_lens, int Lout, cudaStream_t s) override {
    CXRM_CHECK(id_map != nullptr, "cxrm_bridge_ids needs cxrm_set_id_map");
    CXRM_CHECK(R >= 1 && L >= 1 && Lout >= 2, "bridge_ids shape");
    bridge_ids_kernel<<<ceil_div(R, 64), 64, 0, s>>>(seq, L, L, R, bridge_bos, bridge_sep_dec, eos, bridge_n_special, id_map,
                                                      bridge_cls, bridge_sep, out_ids, out_lens, Lout);
    check_launch("bridge_ids");
  }

  void scst_step_host(const float* pixels, int B, int N, const int* prompt_ids, int P, const cxrm_rollout_args& tmpl,
                      const int* label_ids, const int* label_lens, int L_label, int* sequences, float* logprobs,
                      float* reward, float* baseline, float* advantage, int* steps_out, bool on_device,
                      cudaStream_t s) override {
    CXRM_CHECK(finalized && have_reward && id_map, "scst_step_host needs all weights and cxrm_set_id_map");
    CXRM_CHECK(B <= cfg.max_studies && N <= cfg.max_images && P <= cfg.max_prompt, "scst_step_host shape");
    CXRM_CHECK(3 * B <= cfg.rwd_max_seqs && L_label <= cfg.rwd_max_len, "reward batch exceeds rwd_max_seqs");
    // timing aid: CXRM_PHASE_TIMES=1 prints the device time of each phase of the step (events on the step's stream)
    static const bool phase_times = std::getenv("CXRM_PHASE_TIMES") != nullptr;
    // device time of each phase of the step: five event records per step, read back by cxrm_last_phase_ms
    if (!phase_ev[0])
      for (cudaEvent_t& e : phase_ev) CXRM_CUDA_CHECK(cudaEventCreate(&e));
    cudaEvent_t* pe = phase_ev;
    int n_pe = 0;
    auto mark = [&]() { CXRM_CUDA_CHECK(cudaEventRecord(pe[n_pe++], s)); };
    mark();
    const int Tn = tmpl.max_new_tokens, R = 2 * B, Lseq = P + Tn;
    // longest possible reward input: [CLS] + generated words + [SEP], or the longest label
    const int Lr = std::min(cfg.rwd_max_len, std::max(Tn + 2, L_label));
    if (!h_pixels) {
      h_pixels = dalloc<float>(static_cast<long long>(cfg.max_studies) * cfg.max_images * 3 * cfg.image_h * cfg.image_w);
      h_seq = dalloc<int>(static_cast<long long>(Rmax) * Lmax);
      h_lp = dalloc<float>(static_cast<long long>(Rmax) * cfg.max_new_tokens);
      h_rids = dalloc<int>(3LL * cfg.max_studies * cfg.rwd_max_len);
      h_rlens = dalloc<int>(3LL * cfg.max_studies);
      h_emb = dalloc<float>(3LL * cfg.max_studies * 128);
      h_out = dalloc<float>(3LL * cfg.max_studies);
    }
    // cudaMemcpyDefault: the direction is inferred from the (unified) addresses, host or device
    CXRM_CUDA_CHECK(cudaMemcpyAsync(prompt_dev, prompt_ids, static_cast<size_t>(B) * P * sizeof(int), cudaMemcpyDefault, s));
    if (on_device) {
      encode(pixels, B, N, nullptr, nullptr, s);
    } else {
      // Host pixels: padding is detected on the host (pixel_values[b, n, 0, 0, 0] != 0, modelling_longitudinal.py:83),
      // only the valid images cross PCIe, compacted, one encoder chunk at a time on the copy stream, so that chunk
      // c + 1 is in flight while chunk c is being encoded.
      const int n_all = B * N;
      const long long img_stride = 3LL * cfg.image_h * cfg.image_w;
      std::vector<uint8_t> valid(n_all);
      std::vector<int> src, dst;
      for (int i = 0; i < n_all; ++i) {
        valid[i] = pixels[i * img_stride] != 0.0f ? 1 : 0;
        if (valid[i]) {
          src.push_back(static_cast<int>(dst.size()));
          dst.push_back(i);
        }
      }
      CXRM_CUDA_CHECK(cudaMemcpyAsync(valid_img, valid.data(), n_all, cudaMemcpyHostToDevice, s));
      const size_t per = balanced_chunk(dst.size());
      const size_t n_chunks = dst.empty() ? 0 : (dst.size() + per - 1) / per;
      while (ev_chunk.size() < n_chunks + 1) {
        cudaEvent_t e;
        CXRM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev_chunk.push_back(e);
      }
      // h_pixels may still be read by work queued on s (a previous step): the copy stream starts behind it
      CXRM_CUDA_CHECK(cudaEventRecord(ev_chunk[n_chunks], s));
      CXRM_CUDA_CHECK(cudaStreamWaitEvent(pf_stream, ev_chunk[n_chunks], 0));
      h2d_pixel_bytes = 0;
      for (size_t c = 0; c < n_chunks; ++c) {
        const size_t k0 = c * per, k1 = std::min(dst.size(), k0 + per);
        for (size_t k = k0; k < k1;) {   // consecutive source slots travel as one copy
          size_t j = k + 1;
          while (j < k1 && dst[j] == dst[j - 1] + 1) ++j;
          const size_t bytes = (j - k) * img_stride * sizeof(float);
          CXRM_CUDA_CHECK(cudaMemcpyAsync(h_pixels + static_cast<long long>(k) * img_stride, pixels + dst[k] * img_stride, bytes,
                                          cudaMemcpyHostToDevice, pf_stream));
          h2d_pixel_bytes += bytes;
          k = j;
        }
        CXRM_CUDA_CHECK(cudaEventRecord(ev_chunk[c], pf_stream));
      }
      CXRM_CHECK(B >= 1 && N >= 1, "scst_step_host shape");
      encode_valid(h_pixels, B, N, src, dst, ev_chunk.data(), s);
    }
    mark();
    prefill_cross_kv(nullptr, nullptr, 0, 0, s);
    mark();
    cxrm_rollout_args a = tmpl;
    a.mode = CXRM_BOTH; a.B = B; a.P = P; a.prompt_ids = prompt_dev;
    a.sequences = h_seq; a.logprobs = h_lp;
    a.margins = nullptr; a.topk_idx = nullptr; a.topk_val = nullptr; a.topk_cnt = nullptr; a.last_logits = nullptr;
    a.steps_out = nullptr;
    rollout(a, s);
    mark();
    // text bridge: sample rows [0,B), greedy rows [B,2B) -> reward ids rows [0,2B); labels -> rows [2B,3B)
    bridge_ids_kernel<<<ceil_div(R, 64), 64, 0, s>>>(h_seq, Lseq, Lseq, R, bridge_bos, bridge_sep_dec, a.eos_token_id,
                                                      bridge_n_special, id_map, bridge_cls, bridge_sep, h_rids, h_rlens, Lr);
    check_launch("bridge_ids");
    // labels are already reward-model ids, padded to L_label: copy into the [*, Lr] layout
    CXRM_CUDA_CHECK(cudaMemsetAsync(h_rids + static_cast<long long>(R) * Lr, 0, static_cast<size_t>(B) * Lr * sizeof(int), s));
    CXRM_CUDA_CHECK(cudaMemcpy2DAsync(h_rids + static_cast<long long>(R) * Lr, Lr * sizeof(int), label_ids,
                                      L_label * sizeof(int), L_label * sizeof(int), B, cudaMemcpyDefault, s));
    CXRM_CUDA_CHECK(cudaMemcpyAsync(h_rlens + R, label_lens, B * sizeof(int), cudaMemcpyDefault, s));
    reward_embed(h_rids, h_rlens, 3 * B, Lr, h_emb, s);
    cosine_rows(h_emb, h_emb + static_cast<long long>(R) * 128, h_out, B, 128, s);                               // sample vs label
    cosine_rows(h_emb + static_cast<long long>(B) * 128, h_emb + static_cast<long long>(R) * 128, h_out + B, B, 128, s);  // greedy vs label
    advantage_kernel<<<ceil_div(B, 64), 64, 0, s>>>(h_out, h_out + B, h_out + 2 * B, B);
    check_launch("advantage");
    CXRM_CUDA_CHECK(cudaMemcpyAsync(sequences, h_seq, static_cast<size_t>(R) * Lseq * sizeof(int), cudaMemcpyDefault, s));
    if (logprobs) CXRM_CUDA_CHECK(cudaMemcpyAsync(logprobs, h_lp, static_cast<size_t>(R) * Tn * sizeof(float), cudaMemcpyDefault, s));
    CXRM_CUDA_CHECK(cudaMemcpyAsync(reward, h_out, B * sizeof(float), cudaMemcpyDefault, s));
    CXRM_CUDA_CHECK(cudaMemcpyAsync(baseline, h_out + B, B * sizeof(float), cudaMemcpyDefault, s));
    CXRM_CUDA_CHECK(cudaMemcpyAsync(advantage, h_out + 2 * B, B * sizeof(float), cudaMemcpyDefault, s));
    if (steps_out) CXRM_CUDA_CHECK(cudaMemcpyAsync(steps_out, st.step, sizeof(int), cudaMemcpyDefault, s));
    mark();
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    {
      const char* names[4] = {"encode", "cross_kv", "rollout", "reward+copy"};
      std::string line = "[cxrm phases ms]";
      for (int i = 0; i + 1 < n_pe; ++i) {
        cudaEventElapsedTime(&phase_ms[i], pe[i], pe[i + 1]);
        line += std::string(" ") + names[i] + " " + std::to_string(phase_ms[i]);
      }
      // prompt pass (prefill + first token) inside the rollout: its end was recorded by rollout()
      phase_ms[4] = 0.f;
      if (prefill_done_ev && prefill_recorded) cudaEventElapsedTime(&phase_ms[4], pe[2], prefill_done_ev);
      if (phase_times) fprintf(stderr, "%s (prompt pass %.3f)\n", line.c_str(), phase_ms[4]);
    }
  }

 private:
  cxrm_config cfg;
  int device;
  std::map<std::string, RawTensor> raw;
  std::set<std::string> consumed;          // raw keys finalize_weights has used
  std::vector<void*> owned;
  long long persistent_bytes = 0;
  bool finalized = false, have_reward = false;
  Arena arena;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  bool profiling = false;
  const char* phase = "misc";
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> ev_pool;
  // weights
  CvtStageW stages[3];
  float* cls_token = nullptr;
  LNp head_ln;
  Lin head_proj;
  BertW dec, rwd;
  Lin dec_head_t, dec_lm, rp1, rp2;
  LNp dec_head_ln, rp_ln;
  float* dec_vocab_bias = nullptr;
  // geometry
  int T2 = 0, Smax = 0, Rmax = 0, Lmax = 0;
  // persistent activations / caches
  T* memory = nullptr; uint8_t* mem_mask = nullptr; T* mem_compact = nullptr; int* compact_idx = nullptr;
  int* kv_off = nullptr; int* kv_len = nullptr; T* cross_kv = nullptr; T* self_k = nullptr; T* self_v = nullptr;
  float* logits = nullptr; uint8_t* valid_img = nullptr; int* img_idx = nullptr;
  int enc_B = 0, enc_S = 0, kv_B = 0, kv_total = 0, kv_maxlen = 0;
  int attn_ch = 0, cross_max_chunks = 0, cross_max_units = 0, self_max_chunks = 0;
  float* cross_ws = nullptr; float* self_ws = nullptr;
  unsigned* cross_tickets = nullptr; unsigned* self_tickets = nullptr;
  int* unit_tab = nullptr;
  // beam search (beam.cu): virtual-study unit table, bookkeeping state, scratch of the cache reorder (allocated on first use)
  int* unit_tab_beam = nullptr;
  BeamState bs{};
  T* beam_scratch = nullptr;
  size_t beam_scratch_cap = 0;
  std::vector<int> h_kv_len, h_kv_off;    // host copies of kv_len / kv_off (cxrm_prefill_cross_kv)
  AttnMaps attn_maps{};
  const AttnMaps* attn_maps_ptr = nullptr;
  float* skinny_ws = nullptr;
  cudaStream_t pf_stream = nullptr;       // copy stream of the host-buffer step (pixel chunks travel under the encoding)
  cudaEvent_t phase_ev[5] = {};           // scst step: start, encoded, cross K/V, rollout, reward
  cudaEvent_t prefill_done_ev = nullptr;
  bool prefill_recorded = false;
  float phase_ms[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // encode, cross_kv, rollout, reward+copy, (prompt pass inside rollout)
  std::vector<cudaEvent_t> ev_chunk;      // host-buffer step: pixels of encoder chunk c have arrived
  size_t h2d_pixel_bytes = 0;             // pixel bytes copied by the last host-buffer step
  RolloutState st{};
  int* pre_ids = nullptr; int* pre_types = nullptr; int* pre_pos = nullptr; int* prompt_dev = nullptr;
  uint8_t* pre_valid = nullptr;           // [R, P] key mask of the prompt pass
  // host-step staging
  float* h_pixels = nullptr; int* h_seq = nullptr; float* h_lp = nullptr; int* h_rids = nullptr; int* h_rlens = nullptr;
  float* h_emb = nullptr; float* h_out = nullptr;
  int* id_map = nullptr; int bridge_cls = 0, bridge_sep = 0, bridge_bos = 1, bridge_sep_dec = 3, bridge_n_special = 12;
  // training
  std::vector<GradSlot> grad_slots[2];            // [0] every decoder parameter, [1] LoRA only
  std::map<const void*, T*> wt_cache;             // transposed weight copies for the dX GEMMs
  std::vector<int> lora_r;                        // LoRA rank per decoder layer (0: none)
  std::vector<std::array<T*, 2>> lora_A, lora_Bt; // per layer, (query, key): A [r, 768], B^T [r, 768]
  // persistent decode chain (decode_chain.cu): phase lists, (first, count) per launch of a step
  std::vector<ChainPhase> chain_host;
  unsigned* chain_bar = nullptr;
  unsigned long long* chain_trace = nullptr;                  // CXRM_CHAIN_TRACE: stamps of the last executed decode step
  // timeline of the last traced decode step: per launch and phase, when the slowest CTA started / ended it
  void dump_chain_trace(cudaStream_t s) {
    if (!chain_trace || chain_launch.empty()) return;
    CXRM_CUDA_CHECK(cudaStreamSynchronize(s));
    const int nc = decode_chain_ctas();
    std::vector<unsigned long long> h(chain_launch.size() * nc * kChainTraceSlots);
    CXRM_CUDA_CHECK(cudaMemcpy(h.data(), chain_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    auto at = [&](size_t k, int c, int slot) { return h[(k * nc + c) * kChainTraceSlots + slot]; };
    unsigned long long t00 = ~0ull;
    for (int c = 0; c < chain_launch[0].ctas; ++c) t00 = std::min(t00, at(0, c, 0));
    fprintf(stderr, "[cxrm chain trace] one decode step, us since the first chain CTA started (min..max over %d CTAs)\n", nc);
    for (size_t k = 0; k < chain_launch.size(); ++k) {
      auto mm = [&](int slot, double& lo, double& hi) {
        unsigned long long a = ~0ull, b = 0;
        for (int c = 0; c < chain_launch[k].ctas; ++c) {
          a = std::min(a, at(k, c, slot));
          b = std::max(b, at(k, c, slot));
        }
        lo = (a - t00) * 1e-3;
        hi = (b - t00) * 1e-3;
      };
      double lo, hi, lo1, hi1;
      mm(0, lo, hi);
      mm(1, lo1, hi1);
      fprintf(stderr, "  launch %2zu: entry %.1f..%.1f  dep-wait done %.1f..%.1f |", k, lo, hi, lo1, hi1);
      for (int i = 0; i < chain_launch[k].second; ++i) {
        double s0, s1, e0, e1;
        mm(2 + 8 * i, s0, s1);
        mm(2 + 8 * i + 7, e0, e1);
        fprintf(stderr, " p%d start %.1f..%.1f", i, s0, s1);
        if (at(k, 0, 2 + 8 * i + 3)) {   // GEMM sub-stamps of CTA 0 (it takes part in every GEMM phase), relative to its phase start
          const unsigned long long b = at(k, 0, 2 + 8 * i);
          auto rel = [&](int sub) { return (static_cast<double>(at(k, 0, 2 + 8 * i + sub)) - static_cast<double>(b)) * 1e-3; };
          fprintf(stderr, " [cta0 +: A-issued %.2f A-in %.2f mma-issued %.2f acc %.2f]", rel(5), rel(2), rel(3), rel(4));
        }
        fprintf(stderr, " end %.1f..%.1f |", e0, e1);
      }
      fprintf(stderr, "\n");
    }
  }
  float* chain_xf = nullptr; float* chain_x1f = nullptr;      // fp32 residual stream [Rmax, 768] x 2
  struct ChainSpan { int first, second, ctas; };
  float* chain_ws = nullptr;                                  // split-K partials [kChainMaxSplit][64][768]
  std::vector<ChainSpan> chain_launch;
  const void* chain_key_buf = nullptr; const void* chain_key_head = nullptr; int chain_key_R = 0;
  // CUDA graph of one decode step
  struct GraphKey {
    int R, B, P, Tmax, top_k; float temperature; const float* noise; const void* buf;
    int mask_id, eos, pad, want_margin;
    int beams; float length_penalty; const void* beam_scratch;
    int special[2][kMaxSpecial]; int sections[2][kMaxSpecial + 1]; int nspecial[2]; int modes[2];
  };
  GraphKey graph_key;
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_valid = false;
  unsigned long long graph_nodes = 0;
};

template <>
void Engine<float>::attention(const AttnArgs& a, cudaStream_t s) { attention_simt<float>(a, s); }
template <>
void Engine<bf16>::attention(const AttnArgs& a, cudaStream_t s) {
  if (cfg.use_tensor_cores && attention_tc5_supported(a) == 0)
    attention_tc5(a, s);
  else if (cfg.use_tensor_cores && attention_mma_supported(a) == 0)
    attention_mma(a, s);
  else
    attention_simt<bf16>(a, s);
}
template <>
void Engine<float>::setup_attn_maps() {}
template <>
void Engine<bf16>::setup_attn_maps() {
  const long long cross_rows = static_cast<long long>(cfg.dec_layers) * 2 * NHEAD * cross_tok_cap();
  const long long self_rows = static_cast<long long>(Rmax) * NHEAD * Lmax;
  CXRM_CHECK(cross_rows < (1LL << 31) && self_rows * cfg.dec_layers < (1LL << 31), "cache too large for 32-bit TMA row coordinates");
  attn_maps.cross = make_tensor_map_bf16(cross_kv, cross_rows, 64, 64, attn_ch, 64);   // one TMA box per K / V chunk
  attn_maps.self_k = make_tensor_map_bf16(self_k, self_rows * cfg.dec_layers, 64, 64, 64, 64);
  attn_maps.self_v = make_tensor_map_bf16(self_v, self_rows * cfg.dec_layers, 64, 64, 64, 64);
  attn_maps.self_rows_per_layer = static_cast<int>(self_rows);
  attn_maps_ptr = &attn_maps;
}
template <>
bool Engine<float>::chain_pdl() const { return false; }
template <>
bool Engine<bf16>::chain_pdl() const { return cfg.use_tensor_cores != 0; }
template <>
bool Engine<float>::use_skinny(long long, const Lin&) const { return false; }
template <>
bool Engine<bf16>::use_skinny(long long M, const Lin& L) const {
  return cfg.use_tensor_cores && M <= 64 && L.n_out <= 1024 && L.n_out % 4 == 0 && L.n_in % 8 == 0;
}
template <>
void Engine<float>::dispatch_gemm(const GemmArgs& g, cudaStream_t s) { gemm_simt<float>(g, s); }
template <>
void Engine<bf16>::dispatch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (cfg.use_tensor_cores && g.M <= 64 && gemm_skinny_supported(g) == 0)
    gemm_tcgen05_skinny(g, nullptr, nullptr, s);   // decode steps: latency-bound weight streaming
  else if (cfg.use_tensor_cores && gemm_tcgen05_supported(g) == 0)
    gemm_tcgen05(g, s);
  else
    gemm_simt<bf16>(g, s);
}

}  // namespace

EngineBase* make_engine(const cxrm_config& cfg, int device) {
  if (cfg.dtype == CXRM_F32) return new Engine<float>(cfg, device);
  if (cfg.dtype == CXRM_BF16) return new Engine<bf16>(cfg, device);
  throw std::runtime_error("unknown dtype");
}

}  // namespace cxrm
