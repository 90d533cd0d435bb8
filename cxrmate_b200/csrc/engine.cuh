// Engine: owns weights, caches and scratch; sequences the kernels of the path.
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cxrm.h"
#include "kernels.h"

namespace cxrm {

struct RawTensor {
  float* data = nullptr;          // device fp32 staging copy
  std::vector<int64_t> shape;
  long long numel() const {
    long long n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// bump allocator over one device arena; reset between phases
class Arena {
 public:
  void init(size_t bytes);
  void release();
  void reset() { off_ = 0; }
  void* alloc(size_t bytes);
  template <typename U> U* get(long long n) { return static_cast<U*>(alloc(static_cast<size_t>(n) * sizeof(U))); }
  size_t capacity() const { return cap_; }

 private:
  char* base_ = nullptr;
  size_t cap_ = 0, off_ = 0;
};

class EngineBase {
 public:
  virtual ~EngineBase() {}
  virtual void load_weight(const std::string& name, const float* data, const int64_t* shape, int ndim,
                           bool on_device) = 0;
  virtual void finalize_weights() = 0;
  virtual void encode(const float* pixels, int B, int N, void* memory_out, uint8_t* mask_out, cudaStream_t s) = 0;
  virtual void prefill_cross_kv(const void* memory, const uint8_t* mask, int B, int S, cudaStream_t s) = 0;
  virtual void rollout(const cxrm_rollout_args& a, cudaStream_t s) = 0;
  virtual void rollout_beam(const cxrm_beam_args& a, cudaStream_t s) = 0;
  virtual void decoder_forward(const int* ids, const int* tt, const int* pos, const uint8_t* key_mask, int R, int L,
                               int B, bool last_only, float* logits_out, cudaStream_t s) = 0;
  virtual void reward_embed(const int* ids, const int* lens, int n, int L, float* emb_out, cudaStream_t s) = 0;
  virtual void set_id_map(const int* id_map_host, int n, int cls_id, int sep_id, int bos_id, int sep_dec_id,
                          int n_special) = 0;
  virtual void scst_step_host(const float* pixels, int B, int N, const int* prompt_ids, int P,
                              const cxrm_rollout_args& tmpl, const int* label_ids, const int* label_lens, int L_label,
                              int* sequences, float* logprobs, float* reward, float* baseline, float* advantage,
                              int* steps_out, bool on_device, cudaStream_t s) = 0;
  virtual void bridge_ids(const int* seq, int R, int L, int eos, int* out_ids, int* out_lens, int Lout, cudaStream_t s) = 0;
  virtual size_t workspace_bytes() const = 0;
  virtual void last_phase_ms(float* out5) const = 0;
  virtual void train_step(const cxrm_train_args& a, int stage, cudaStream_t s) = 0;
  virtual int train_stages() const = 0;
  virtual int grad_count(bool lora_only) const = 0;
  virtual long long grad_total(bool lora_only) const = 0;
  virtual bool grad_info(bool lora_only, int i, std::string* name, long long* offset, long long* numel, int* stage) const = 0;
  virtual void set_profile(bool on) = 0;
  virtual std::string profile_report() = 0;
  std::string last_error;
  unsigned long long launches = 0;
};

EngineBase* make_engine(const cxrm_config& cfg, int device);

}  // namespace cxrm
