// Shared device/host helpers for the CXRMate SCST rollout engine (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <stdexcept>
#include <utility>
#include <string>

namespace cxrm {

using bf16 = __nv_bfloat16;

// typed failures, mapped to cxrm_status by the C ABI (capi.cu)
struct WeightError : std::runtime_error { using std::runtime_error::runtime_error; };   // CXRM_ERR_WEIGHT
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };     // CXRM_ERR_CUDA

#define CXRM_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      throw ::cxrm::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__));                         \
    }                                                                                              \
  } while (0)

#define CXRM_CHECK(cond, msg)                                                                       \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      throw std::runtime_error(std::string("check failed: ") + #cond + " - " + (msg) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__));                         \
    }                                                                                              \
  } while (0)

extern unsigned long long g_launch_count;   // kernels launched by this library (engine.cu)

inline void check_launch(const char* what) {
  ++g_launch_count;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw CudaError(std::string(what) + " launch failed: " + cudaGetErrorString(e));
}

// Programmatic dependent launch (PDL) for the decode-step kernel chain: while g_pdl is set, the chain's kernels are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so that kernel N+1 is scheduled as soon as every
// CTA of kernel N has executed pdl_launch_dependents(); its prologue (barrier init, TMEM allocation, descriptor
// prefetch, TMA loads of WEIGHT tiles - data that no kernel of the chain writes) overlaps kernel N, and it blocks in
// pdl_wait() until kernel N has completed and flushed.  Rules for a chain kernel: call pdl_launch_dependents() early,
// touch only chain-constant memory before pdl_wait(), and ALWAYS execute pdl_wait() (completion is transitive only
// through it).  Both instructions are no-ops in a kernel that was launched without the attribute.
extern bool g_pdl;   // engine.cu
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
  if (e != cudaSuccess) throw CudaError(std::string("cudaLaunchKernelEx failed: ") + cudaGetErrorString(e));
}
#endif

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- element conversion -------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors of T -------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void load_nc(const float* p) { raw = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void load_nc(const bf16* p) { raw = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---- reductions -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// exact-erf GELU (nn.GELU() default / ACT2FN["gelu"])
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }


// GELU for the bf16 tensor-core epilogues (the fp32 validation mode uses erff).  Default: 0.5 x (1 + tanh(x (a + b x^2 +
// c x^4))) with (a, b, c) fitted to the exact erf GELU on [-6, 6] (max |error| 2.5e-5, 19x closer than the textbook
// tanh form) and MUFU.TANH (relative error 2^-11 on the tanh): 6 FP ops + 1 MUFU against the 14 + 2 (RCP, EX2) of the
// Abramowitz-Stegun 7.1.26 erf below.  The result is rounded to bf16 (relative 2^-9) right after, which is 8x coarser
// than either error; the GELU epilogues are ALU-bound (ncu: tensor pipe 5-21 % on the MLP-up GEMMs of CvT) and run
// 1.3-1.65x faster with it: [294912 x 256 x 64] 99.5 -> 60.2 us, [73728 x 768 x 192] 78.0 -> 49.0, [18464 x 1536 x 384]
// 57.0 -> 45.0; the engine's bf16 error against the fp32 oracle is unchanged (tests/test_bf16_parity_gpu.py).
// -DCXRM_GELU_ERF (CXRM_DEFINES=CXRM_GELU_ERF) selects the erf formulation (|error| <= 1.5e-7).
__device__ __forceinline__ float gelu_fast(float x) {
#ifndef CXRM_GELU_ERF
  const float x2 = x * x;
  const float inner = x * fmaf(x2, fmaf(x2, -3.51516788e-4f, 3.70056460e-2f), 7.97507884e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#else
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));   // MUFU.RCP: 1 ulp-ish, far below bf16 rounding
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = 1.0f - p * t * __expf(-z * z);
  const float erf_x = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_x);
#endif
}

}  // namespace cxrm
