// Bandwidth-bound kernels of the path: LayerNorm, embedding gather + LayerNorm,
// im2col for the CvT convolutional token embeddings, the depth-wise
// convolutional q/k/v projections with folded BatchNorm, row gathers/scatters
// and the weight-packing helpers.  All math in fp32; storage type T.
#include <algorithm>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int kMaxPerLane = 24;  // supports C <= 768

// ---- LayerNorm: one warp per row -------------------------------------------
// Two-pass (mean, then centred variance) in registers, like
// torch.nn.functional.layer_norm (HF modeling_cvt.py:112,366; modeling_bert.py:62,292,348,479).
template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C,
                                 float eps) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    v[i] = (c < C) ? to_f(xr[c]) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  T* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    if (c < C) yr[c] = from_f<T>((v[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

template <typename T>
__global__ void embed_ln_kernel(const int* __restrict__ ids, const int* __restrict__ types, const int* __restrict__ pos,
                                const T* __restrict__ word, const T* __restrict__ type_emb,
                                const T* __restrict__ pos_emb, const float* __restrict__ gamma,
                                const float* __restrict__ beta, T* __restrict__ out, long long rows, int C, float eps) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  pdl_launch_dependents();   // first kernel of the decode-step chain: lets the QKV GEMM prefetch its weights
  if (row >= rows) return;
  const T* w = word + static_cast<long long>(ids[row]) * C;
  const T* t = type_emb + static_cast<long long>(types ? types[row] : 0) * C;
  const T* p = pos_emb + static_cast<long long>(pos[row]) * C;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    // (word + type) + position in fp32, same association order as BertEmbeddings.forward
    float e = 0.f;
    if (c < C) e = (to_f(w[c]) + to_f(t[c])) + to_f(p[c]);
    v[i] = e;
    s += e;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  T* yr = out + row * C;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * kWarp;
    if (c < C) yr[c] = from_f<T>((v[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

// Vectorised version (16-byte gathers, NV vectors per lane, C == 32 * NV * Vec16<T>::N): the decode step opens with
// this kernel, one warp per rollout row, so its latency is on every token's critical path.
template <typename T, int NV>
__global__ void __launch_bounds__(128) embed_ln_vec_kernel(const int* __restrict__ ids, const int* __restrict__ types,
                                                           const int* __restrict__ pos, const T* __restrict__ word,
                                                           const T* __restrict__ type_emb, const T* __restrict__ pos_emb,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           T* __restrict__ out, long long rows, int C, float eps) {
  constexpr int V = Vec16<T>::N;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  pdl_launch_dependents();
  if (row >= rows) return;
  const T* w = word + static_cast<long long>(ids[row]) * C;
  const T* t = type_emb + static_cast<long long>(types ? types[row] : 0) * C;
  const T* p = pos_emb + static_cast<long long>(pos[row]) * C;
  Vec16<T> wv[NV], tv[NV], pv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * kWarp) * V;
    wv[i].load(w + c);
    tv[i].load(t + c);
    pv[i].load(p + c);
  }
  float v[NV][V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a[V], b[V], c3[V];
    wv[i].unpack(a);
    tv[i].unpack(b);
    pv[i].unpack(c3);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[i][j] = (a[j] + b[j]) + c3[j];   // (word + type) + position, as BertEmbeddings.forward associates it
      s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < V; ++j) q += (v[i][j] - mean) * (v[i][j] - mean);
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * kWarp) * V;
    float o[V];
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = (v[i][j] - mean) * rstd * gamma[c + j] + beta[c + j];
    Vec16<T> ov;
    ov.pack(o);
    ov.store(out + row * C + c);
  }
}

// ---- im2col ------------------------------------------------------------------
__device__ __forceinline__ void store_pair2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair2(bf16* p, float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&v);
}
template <typename T>
__global__ void im2col_pixels_kernel(const float* __restrict__ pixels, const int* __restrict__ img_idx,
                                     T* __restrict__ out, int n_img, int H, int W, int Ho, int Wo, int ksz, int stride,
                                     int pad, int Kpad) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
  const int K = 3 * ksz * ksz;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kk = static_cast<int>(i % Kpad);
    const long long r = i / Kpad;
    const int ox = static_cast<int>(r % Wo);
    const int oy = static_cast<int>((r / Wo) % Ho);
    const int n = static_cast<int>(r / (static_cast<long long>(Wo) * Ho));
    float v = 0.f;
    if (kk < K) {
      const int kx = kk % ksz, ky = (kk / ksz) % ksz, c = kk / (ksz * ksz);
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const long long src = img_idx ? img_idx[n] : n;
        v = pixels[((src * 3 + c) * H + iy) * static_cast<long long>(W) + ix];
      }
    }
    out[i] = from_f<T>(v);
  }
}

// The same patch matrix, one block per (image, output row): the ksz input rows of the 3 channels are staged in shared
// memory with coalesced loads (the element-per-thread kernel above gathers 4-byte pixels with a div / mod chain per
// element: 350 us for a 32-image chunk whose 57 MB in + 90 MB out need ~25 us), then the row's Wo patches leave as
// consecutive 2-element stores.  Needs an even Kpad.
// (First version: one flat index per element, decomposed with runtime divisions by ksz / W / Kpad in both phases: ncu
// 400 us per 50-image pass for 88 MB in + 140 MB out, integer-division bound.  Now a warp owns whole rows: input rows as
// float4 per lane, output patches with the (channel, ky, kx) decomposition of a lane's <= 3 element pairs hoisted out
// of the loop over the row's patches.)
template <typename T>
__global__ void __launch_bounds__(256) im2col_pixels_rows_kernel(const float* __restrict__ pixels, const int* __restrict__ img_idx,
                                                                 T* __restrict__ out, int H, int W, int Ho, int Wo, int ksz,
                                                                 int stride, int pad, int Kpad) {
  extern __shared__ __align__(16) float rows_sm[];   // [3 * ksz][W + 8]: the row pad spreads a patch's ksz rows over the banks
  pdl_launch_dependents();
  pdl_wait();
  const int oy = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int lane = threadIdx.x % 32, wrp = threadIdx.x / 32;
  const long long src = img_idx ? img_idx[n] : n;
  const int K = 3 * ksz * ksz, W4 = W / 4, RS = W + 8;
  for (int r = wrp; r < 3 * ksz; r += 8) {            // r = c * ksz + ky
    const int c = r / ksz, ky = r - c * ksz;
    const int iy = oy * stride - pad + ky;
    const bool in = iy >= 0 && iy < H;
    const float4* g = reinterpret_cast<const float4*>(pixels + ((src * 3 + c) * H + (in ? iy : 0)) * static_cast<long long>(W));
    float4* d = reinterpret_cast<float4*>(rows_sm + r * RS);
    for (int i = lane; i < W4; i += 32) d[i] = in ? __ldg(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // this lane's element pairs kk0 = 2 * (lane + 32 j) of a patch: source offsets relative to the patch's first column
  const int half = Kpad / 2;
  int off[3][2], dx[3][2];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = 2 * (lane + 32 * j) + e;
      const int c = kk / (ksz * ksz), rem = kk - c * ksz * ksz, ky = rem / ksz, kx = rem - ky * ksz;
      dx[j][e] = kk < K ? kx - pad : -(1 << 20);       // padding columns of the patch: always out of range
      off[j][e] = (c * ksz + ky) * RS;
    }
  T* orow = out + (static_cast<long long>(n) * Ho + oy) * Wo * Kpad;
  for (int ox = wrp; ox < Wo; ox += 8) {
    const int x0 = ox * stride;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int pr = lane + 32 * j;
      if (pr < half) {
        const int ix0 = x0 + dx[j][0], ix1 = x0 + dx[j][1];
        const float v0 = (ix0 >= 0 && ix0 < W) ? rows_sm[off[j][0] + ix0] : 0.f;
        const float v1 = (ix1 >= 0 && ix1 < W) ? rows_sm[off[j][1] + ix1] : 0.f;
        store_pair2(orow + static_cast<long long>(ox) * Kpad + 2 * pr, v0, v1);
      }
    }
  }
}

// one thread per 16-byte channel vector
template <typename T>
__global__ void im2col_tokens_kernel(const T* __restrict__ in, T* __restrict__ out, int n_img, int H, int W, int C,
                                     int Ho, int Wo, int ksz, int stride, int pad) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = static_cast<long long>(n_img) * Ho * Wo * ksz * ksz * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * V;
    long long r = i / cv;
    const int tap = static_cast<int>(r % (ksz * ksz));
    r /= (ksz * ksz);
    const int ox = static_cast<int>(r % Wo);
    const int oy = static_cast<int>((r / Wo) % Ho);
    const int n = static_cast<int>(r / (static_cast<long long>(Wo) * Ho));
    const int ky = tap / ksz, kx = tap % ksz;
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    Vec16<T> v;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      v.load(in + ((static_cast<long long>(n) * H + iy) * W + ix) * C + c);
    } else {
      float z[V];
#pragma unroll
      for (int j = 0; j < V; ++j) z[j] = 0.f;
      v.pack(z);
    }
    v.store(out + i * V);
  }
}


// ---- LayerNorm, vectorised: LPR lanes per row, 16-byte loads, rows held in registers -------------------------
// Same two-pass fp32 arithmetic as layernorm_kernel.  STATS: write (mean, rstd) per row instead of the normalised
// row (consumed by ln_dwconv_qkv_kernel, which normalises on load).
constexpr int kLnMaxVec = 6;   // 16-byte vectors per lane
template <typename T, int LPR, bool STATS>
__global__ void __launch_bounds__(256) ln_rows_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy,
                                                      float2* __restrict__ stats, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, long long rows, int C, float eps) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  constexpr int V = Vec16<T>::N;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
  const int sub = threadIdx.x % LPR;
  if (row >= rows) return;   // LPR divides the warp size and rows are LPR-aligned in the warp: whole groups leave together
  const int cv = C / V;
  const T* xr = x + row * ldx;
  float v[kLnMaxVec][V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vi = sub + i * LPR;
    if (vi < cv) {
      Vec16<T> t;
      t.load(xr + vi * V);
      t.unpack(v[i]);
#pragma unroll
      for (int j = 0; j < V; ++j) s += v[i][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (sub + i * LPR < cv) {
#pragma unroll
      for (int j = 0; j < V; ++j) q += (v[i][j] - mean) * (v[i][j] - mean);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(kFull, q, o);
  const float rstd = rsqrtf(q / C + eps);
  if (STATS) {
    if (sub == 0) stats[row] = make_float2(mean, rstd);
    return;
  }
  T* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vi = sub + i * LPR;
    if (vi < cv) {
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = (v[i][j] - mean) * rstd * gamma[vi * V + j] + beta[vi * V + j];
      Vec16<T> t;
      t.pack(o);
      t.store(yr + vi * V);
    }
  }
}

// VW consecutive elements of T as one 8- or 16-byte access
template <typename T, int VW> struct VecW;
template <typename T> struct VecW<T, Vec16<T>::N> : Vec16<T> {};
template <> struct VecW<bf16, 4> {
  uint2 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint2*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = __uint_as_float(raw.x << 16); f[1] = __uint_as_float(raw.x & 0xffff0000u);
    f[2] = __uint_as_float(raw.y << 16); f[3] = __uint_as_float(raw.y & 0xffff0000u);
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    raw = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
};

// ---- CvT attention front end in one pass: LayerNorm (from per-token stats) -> depth-wise 3x3 (pad 1) + folded
// BatchNorm for q (stride 1) and k, v (stride 2), HF modeling_cvt.py:124-141,215-228,371-377.
// A thread owns one 16-byte channel vector of one image column and walks TY output rows with a rolling 3x3 window
// of NORMALISED values in registers: every input token is fetched 3 times per channel vector (its own column and
// both neighbours - L1 hits inside the block) instead of 9 + 9/4 times, the LayerNorm output is never written, and
// the stride-2 k/v outputs (window centre (2a, 2b) == the q window at even coordinates) come from the same registers.
// (bf16 runs with 4 channels per thread: the rolling window is 9 x V registers, and at V = 8 the kernel needed 189
// registers = one 256-thread CTA per SM, 141 us for a 38 MB map in ncu; 4 channels halve that.)
template <typename T, int TY, int V>
__global__ void __launch_bounds__(256, 3) ln_dwconv_qkv_kernel(const T* __restrict__ x, T* __restrict__ q,
                                                            T* __restrict__ k, T* __restrict__ v,
                                                            const float2* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ w,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int H, int W, int C, int cls,
                                                            int Hk, int Wk, int cols) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const int cv = C / V;
  const int tc = threadIdx.x % cv, tcol = threadIdx.x / cv;
  const int ox = blockIdx.x * cols + tcol;
  if (tcol >= cols || ox >= W) return;
  const int oy0 = blockIdx.y * TY, n = blockIdx.z, c = tc * V;
  const long long tok0 = static_cast<long long>(n) * (cls + H * W) + cls;
  float g[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = gamma[c + j];
    b[j] = beta[c + j];
  }
  float win[3][3][V];
  auto load_row = [&](int iy, float (&dst)[3][V]) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox - 1 + dx;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) {
#pragma unroll
        for (int j = 0; j < V; ++j) dst[dx][j] = 0.f;   // the convolution pads the NORMALISED map with zeros
      } else {
        const long long tok = tok0 + static_cast<long long>(iy) * W + ix;
        VecW<T, V> t;
        t.load(x + tok * C + c);
        float xf[V];
        t.unpack(xf);
        const float2 st = stats[tok];
#pragma unroll
        for (int j = 0; j < V; ++j) dst[dx][j] = (xf[j] - st.x) * st.y * g[j] + b[j];
      }
    }
  };
  load_row(oy0 - 1, win[0]);
  load_row(oy0, win[1]);
  // rolled on purpose: unrolled, the compiler keeps all 27 x V weights live across the rows (189 registers, spills)
#pragma unroll 1
  for (int t = 0; t < TY; ++t) {
    const int oy = oy0 + t;
    if (oy >= H) break;
    load_row(oy + 1, win[2]);
    const bool kv = ((oy | ox) & 1) == 0;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      if (o > 0 && !kv) break;
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wt = w + (static_cast<long long>(o) * 9 + ky * 3 + kx) * C + c;
          float wf[V];
#pragma unroll
          for (int j = 0; j < V; j += 4) {
            // volatile asm: re-read (L1 hit) every row instead of 27 x V loop-invariant registers (keeping the 9 q taps
            // in registers at 2 CTAs/SM measured slower: 67.6 vs 60.5 us per launch)
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(wf[j]), "=f"(wf[j + 1]), "=f"(wf[j + 2]), "=f"(wf[j + 3])
                         : "l"(wt + j));
          }
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = fmaf(win[ky][kx][j], wf[j], acc[j]);
        }
      float of[V];
#pragma unroll
      for (int j = 0; j < V; ++j) of[j] = fmaf(acc[j], scale[o * C + c + j], shift[o * C + c + j]);
      VecW<T, V> ov;
      ov.pack(of);
      if (o == 0) {
        ov.store(q + (tok0 + static_cast<long long>(oy) * W + ox) * C + c);
      } else {
        const long long orow = static_cast<long long>(n) * (cls + Hk * Wk) + cls + static_cast<long long>(oy >> 1) * Wk + (ox >> 1);
        ov.store((o == 1 ? k : v) + orow * C + c);
      }
    }
    // slide the window one row down
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        win[0][dx][j] = win[1][dx][j];
        win[1][dx][j] = win[2][dx][j];
      }
  }
}

// ---- the same front end, tiled through shared memory (round 2; C % 64 == 0) ---------------------------------------
// ncu of the register-window kernel above (stage 3, 18432 x 384): 41 us, ~100 executed instructions per output element
// (27 weight vectors re-read through L1 per output row, 64-bit address arithmetic per tap), issue slots 48 % busy, 61 %
// of the stalls on L1TEX scoreboards - instruction-bound at 5x the time its 35 MB of traffic needs.  Here a block owns
// a TH x 8 spatial tile and 64 channels: the NORMALISED tile + halo goes to shared memory once, a warp is one output
// column x 32 channel pairs with its weights and the folded BatchNorm constants in registers, and each output row costs
// 3 conflict-free shared loads for the rolling window.  (First version: register-staged loads, 27 x 2 weights live.)
template <typename T>
__device__ __forceinline__ float2 load_pair(const T* p);
template <>
__device__ __forceinline__ float2 load_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 load_pair<bf16>(const bf16* p) {
  const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
}
__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(bf16* p, float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&v);
}

// Second version: the tile arrives RAW through cp.async, the three convolutions run as two passes.
// ncu of the first version over an encoder pass: 100 us per launch at 330 GB/s of DRAM traffic, neither issue- nor
// bandwidth-bound: ~120 registers (27 x 2 weights) held it to 2 blocks per SM, and a block alternated between five
// serial batches of register-staged loads and its compute phase.  Here every 16-byte chunk of the tile + halo is
// requested at once with cp.async (no registers, zero-filled outside the image), tokens are
// normalised once in a sweep through registers, and q / (k | v) are separate passes with 9 weights live each (the k and v
// pass maps the 8 warps to the 4 even columns x {k, v}): 64 registers, 47 KB of shared memory, 4 blocks per SM.
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <typename T, int TH>
__global__ void __launch_bounds__(256, 4) ln_dwconv_async_kernel(const T* __restrict__ x, T* __restrict__ q, T* __restrict__ k,
                                                               T* __restrict__ v, const float2* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ w, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, int H, int W, int C, int cls,
                                                               int Hk, int Wk, int tiles_x) {
  constexpr int TW = 8, PW = TW + 2, PH = TH + 2, NP = PH * PW;
  constexpr int ROWB = 64 * static_cast<int>(sizeof(T));   // bytes of one position's 64 raw channels
  constexpr int CPR = ROWB / 16;                           // 16-byte chunks per position
  constexpr int PPW = (NP + 7) / 8;                        // positions per warp in the normalisation sweep
  // the raw tile lands at the START of the buffer the normalised fp32 tile will occupy
  __shared__ __align__(16) float2 tile[NP][32];
  __shared__ float2 sst[NP];                               // (mean, rstd); rstd = 0 marks a position outside the image
  unsigned char* raw = reinterpret_cast<unsigned char*>(&tile[0][0]);
  pdl_launch_dependents();
  const int cp = threadIdx.x % 32, wrp = threadIdx.x / 32;
  const int c0 = blockIdx.y * 64, c = c0 + 2 * cp;
  const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW, n = blockIdx.z;
  const long long tok0 = static_cast<long long>(n) * (cls + H * W) + cls;
  // constants of the layer: ahead of the dependency wait
  const float2 g = *reinterpret_cast<const float2*>(gamma + c), b = *reinterpret_cast<const float2*>(beta + c);
  float2 wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const float2*>(w + static_cast<long long>(t) * C + c);
  float2 sc = *reinterpret_cast<const float2*>(scale + c), sh = *reinterpret_cast<const float2*>(shift + c);
  pdl_wait();
  for (int i = threadIdx.x; i < NP * CPR; i += 256) {
    const int p = i / CPR, ch = i % CPR;
    const int iy = ty0 - 1 + p / PW, ix = tx0 - 1 + p % PW;
    const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
    const T* src = in ? x + (tok0 + static_cast<long long>(iy) * W + ix) * C + c0 + ch * (16 / static_cast<int>(sizeof(T))) : x;
    cp_async16_zfill(raw + p * ROWB + ch * 16, src, in);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int p = threadIdx.x; p < NP; p += 256) {
    const int iy = ty0 - 1 + p / PW, ix = tx0 - 1 + p % PW;
    const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
    sst[p] = in ? stats[tok0 + static_cast<long long>(iy) * W + ix] : make_float2(0.f, 0.f);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- normalise ONCE per position (zero outside the image: the convolution pads the NORMALISED map); the raw values
  // pass through registers because the fp32 tile overlays them ----
  {
    float2 xr[PPW];
#pragma unroll
    for (int i = 0; i < PPW; ++i) {
      const int p = wrp + 8 * i;
      xr[i] = p < NP ? load_pair<T>(reinterpret_cast<const T*>(raw + p * ROWB) + 2 * cp) : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PPW; ++i) {
      const int p = wrp + 8 * i;
      if (p < NP) {
        const float2 st = sst[p];
        const bool in = st.y != 0.f;
        tile[p][cp] = make_float2(in ? (xr[i].x - st.x) * st.y * g.x + b.x : 0.f, in ? (xr[i].y - st.x) * st.y * g.y + b.y : 0.f);
      }
    }
    __syncthreads();
  }
  auto ld = [&](int p) -> float2 { return tile[p][cp]; };
  // ---- q: stride 1, a warp = one tile column, rolling window down the column ----
  {
    const int col = wrp, ox = tx0 + col;
    if (ox < W) {
      float2 win[3][3];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        win[0][dx] = ld(0 * PW + col + dx);
        win[1][dx] = ld(1 * PW + col + dx);
      }
#pragma unroll
      for (int t = 0; t < TH; ++t) {
        const int oy = ty0 + t;
        if (oy >= H) break;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) win[2][dx] = ld((t + 2) * PW + col + dx);
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            a.x = fmaf(win[ky][kx].x, wt[ky * 3 + kx].x, a.x);
            a.y = fmaf(win[ky][kx].y, wt[ky * 3 + kx].y, a.y);
          }
        store_pair(q + (tok0 + static_cast<long long>(oy) * W + ox) * C + c, fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y));
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          win[0][dx] = win[1][dx];
          win[1][dx] = win[2][dx];
        }
      }
    }
  }
  // ---- k | v: stride 2 (window centres at the even coordinates; the tile origin is even), warp = (even column, k or v) ----
  {
    const int col = 2 * (wrp % 4), o = 1 + wrp / 4, ox = tx0 + col;
    if (ox >= W) return;
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const float2*>(w + static_cast<long long>(o * 9 + t) * C + c);
    sc = *reinterpret_cast<const float2*>(scale + o * C + c);
    sh = *reinterpret_cast<const float2*>(shift + o * C + c);
    T* out = o == 1 ? k : v;
    float2 top[3];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) top[dx] = ld(col + dx);
#pragma unroll
    for (int t = 0; t < TH; t += 2) {
      const int oy = ty0 + t;
      if (oy >= H) break;
      float2 mid[3], bot[3];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        mid[dx] = ld((t + 1) * PW + col + dx);
        bot[dx] = ld((t + 2) * PW + col + dx);
      }
      float2 a = make_float2(0.f, 0.f);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        a.x = fmaf(top[kx].x, wt[kx].x, a.x);
        a.y = fmaf(top[kx].y, wt[kx].y, a.y);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        a.x = fmaf(mid[kx].x, wt[3 + kx].x, a.x);
        a.y = fmaf(mid[kx].y, wt[3 + kx].y, a.y);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        a.x = fmaf(bot[kx].x, wt[6 + kx].x, a.x);
        a.y = fmaf(bot[kx].y, wt[6 + kx].y, a.y);
      }
      const long long orow = static_cast<long long>(n) * (cls + Hk * Wk) + cls + static_cast<long long>(oy >> 1) * Wk + (ox >> 1);
      store_pair(out + orow * C + c, fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y));
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) top[dx] = bot[dx];
    }
  }
}

// cls rows bypass the convolution: q = k = v = LayerNorm(x[cls]) (modeling_cvt.py:215-228)
template <typename T>
__global__ void ln_cls_kernel(const T* __restrict__ x, const float2* __restrict__ stats, const float* __restrict__ gamma,
                              const float* __restrict__ beta, T* __restrict__ q, T* __restrict__ k, T* __restrict__ v,
                              int n_img, int HWq, int HWk, int C) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * C) return;
  const int n = i / C, c = i % C;
  const long long row = static_cast<long long>(n) * (1 + HWq);
  const float2 st = stats[row];
  const T val = from_f<T>((to_f(x[row * C + c]) - st.x) * st.y * gamma[c] + beta[c]);
  q[row * C + c] = val;
  k[static_cast<long long>(n) * (1 + HWk) * C + c] = val;
  v[static_cast<long long>(n) * (1 + HWk) * C + c] = val;
}

template <typename T>
__global__ void cat_cls_kernel(const T* __restrict__ tokens, const float* __restrict__ cls_token, T* __restrict__ out,
                               int n_img, int HW, int C) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const long long total = static_cast<long long>(n_img) * (1 + HW) * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int t = static_cast<int>(r % (1 + HW));
    const long long n = r / (1 + HW);
    out[i] = (t == 0) ? from_f<T>(cls_token[c]) : tokens[(n * HW + (t - 1)) * C + c];
  }
}

template <typename T>
__global__ void drop_cls_kernel(const T* __restrict__ in, T* __restrict__ out, int n_img, int HW, int C) {
  pdl_launch_dependents();   // programmatic dependent launch (encoder chain): the next kernel may start its prologue
  pdl_wait();                // ... and nothing below touches global memory before the previous kernel has completed
  const long long total = static_cast<long long>(n_img) * HW * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int t = static_cast<int>(r % HW);
    const long long n = r / HW;
    out[i] = in[(n * (1 + HW) + 1 + t) * C + c];
  }
}

template <typename T, bool SCATTER>
__global__ void move_rows_kernel(const T* __restrict__ src, const int* __restrict__ idx, T* __restrict__ dst,
                                 long long n_rows, int C) {
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = n_rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * V;
    const long long r = i / cv;
    const int j = idx[r];
    Vec16<T> v;
    if (SCATTER) {
      if (j < 0) continue;
      v.load(src + r * C + c);
      v.store(dst + static_cast<long long>(j) * C + c);
    } else {
      if (j >= 0) {
        v.load(src + static_cast<long long>(j) * C + c);
      } else {
        float z[V];
#pragma unroll
        for (int q = 0; q < V; ++q) z[q] = 0.f;
        v.pack(z);
      }
      v.store(dst + r * C + c);
    }
  }
}

template <typename TS, typename TD>
__global__ void cast_copy_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = from_f<TD>(to_f(src[i]));
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = beta[c] - mean[c] * s;
}

__global__ void lora_merge_kernel(float* W, const float* A, const float* B, int n_out, int n_in, int r, float s) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n_out) * n_in) return;
  const int o = static_cast<int>(i / n_in), c = static_cast<int>(i % n_in);
  float d = 0.f;
  for (int j = 0; j < r; ++j) d = fmaf(B[o * r + j], A[j * n_in + c], d);
  W[i] += s * d;
}

template <typename T>
__global__ void pack_matrix_kernel(const float* src, T* dst, int rows, int cols_src, int cols_dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols_dst) return;
  const int r = static_cast<int>(i / cols_dst), c = static_cast<int>(i % cols_dst);
  dst[i] = from_f<T>(c < cols_src ? src[static_cast<long long>(r) * cols_src + c] : 0.f);
}

template <typename T>
__global__ void pack_conv_khwc_kernel(const float* src, T* dst, int Cout, int Cin, int ksz) {
  const long long total = static_cast<long long>(Cout) * Cin * ksz * ksz;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // dst index = ((o*ksz + ky)*ksz + kx)*Cin + c
  const int c = static_cast<int>(i % Cin);
  long long r = i / Cin;
  const int kx = static_cast<int>(r % ksz);
  r /= ksz;
  const int ky = static_cast<int>(r % ksz);
  const int o = static_cast<int>(r / ksz);
  dst[i] = from_f<T>(src[((static_cast<long long>(o) * Cin + c) * ksz + ky) * ksz + kx]);
}

__global__ void pack_dw_kernel(const float* src, float* dst, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  const int tap = i / C, c = i % C;
  dst[i] = src[c * 9 + tap];
}

inline unsigned grid_for(long long total, int block, long long cap = 148LL * 32) {
  long long g = ceil_div_ll(total, block);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace


// vectorised LayerNorm / row statistics; false when the shape or alignment needs the scalar kernel
template <typename T, bool STATS>
bool ln_rows_launch(const T* x, int ldx, T* y, int ldy, float2* stats, const float* gamma, const float* beta,
                    long long rows, int C, float eps, cudaStream_t stream) {
  constexpr int V = Vec16<T>::N;
  if (C % V != 0 || ldx % V != 0 || reinterpret_cast<uintptr_t>(x) % 16 != 0) return false;
  if (!STATS && (ldy % V != 0 || reinterpret_cast<uintptr_t>(y) % 16 != 0)) return false;
  const int cv = C / V;
  if (cv > 32 * kLnMaxVec) return false;
  // fewest lanes per row that still hold the row in kLnMaxVec vectors per lane: up to 6 independent 16-byte loads in
  // flight per thread (one load per thread left the kernel latency-bound: 1.5 TB/s in ncu)
  int lpr = 2;
  while (lpr * kLnMaxVec < cv) lpr *= 2;
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows * lpr, 256));
  switch (lpr) {
    case 2: launch_chain(ln_rows_kernel<T, 2, STATS>, dim3(grid), dim3(256), 0, stream, x, ldx, y, ldy, stats, gamma, beta, rows, C, eps); break;
    case 4: launch_chain(ln_rows_kernel<T, 4, STATS>, dim3(grid), dim3(256), 0, stream, x, ldx, y, ldy, stats, gamma, beta, rows, C, eps); break;
    case 8: launch_chain(ln_rows_kernel<T, 8, STATS>, dim3(grid), dim3(256), 0, stream, x, ldx, y, ldy, stats, gamma, beta, rows, C, eps); break;
    case 16: launch_chain(ln_rows_kernel<T, 16, STATS>, dim3(grid), dim3(256), 0, stream, x, ldx, y, ldy, stats, gamma, beta, rows, C, eps); break;
    default: launch_chain(ln_rows_kernel<T, 32, STATS>, dim3(grid), dim3(256), 0, stream, x, ldx, y, ldy, stats, gamma, beta, rows, C, eps); break;
  }
  check_launch(STATS ? "ln_stats" : "layernorm_vec");
  return true;
}

template <typename T>
void layernorm(const T* x, int ldx, T* y, int ldy, const float* gamma, const float* beta, long long rows, int C,
               float eps, cudaStream_t stream) {
  if (rows <= 0) return;
  CXRM_CHECK(C <= kMaxPerLane * kWarp, "layernorm supports C <= 768");
  if (ln_rows_launch<T, false>(x, ldx, y, ldy, nullptr, gamma, beta, rows, C, eps, stream)) return;
  const int block = 256;
  const long long grid = ceil_div_ll(rows * kWarp, block);
  launch_chain(layernorm_kernel<T>, dim3(static_cast<unsigned>(grid)), dim3(block), 0, stream, x, ldx, y, ldy, gamma, beta, rows, C, eps);
  check_launch("layernorm");
}

template <typename T>
void embed_ln(const int* ids, const int* types, const int* pos, const T* word, const T* type_emb, const T* pos_emb,
              const float* gamma, const float* beta, T* out, long long rows, int C, float eps, cudaStream_t stream) {
  if (rows <= 0) return;
  CXRM_CHECK(C <= kMaxPerLane * kWarp, "embed_ln supports C <= 768");
  constexpr int NV = 768 / (kWarp * Vec16<T>::N);   // the BERT width: 3 (bf16) / 6 (fp32) vectors per lane
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  if (C == 768 && al16(word) && al16(type_emb) && al16(pos_emb) && al16(out)) {
    embed_ln_vec_kernel<T, NV><<<static_cast<unsigned>(ceil_div_ll(rows * kWarp, 128)), 128, 0, stream>>>(
        ids, types, pos, word, type_emb, pos_emb, gamma, beta, out, rows, C, eps);
    check_launch("embed_ln");
    return;
  }
  const int block = 256;
  const long long grid = ceil_div_ll(rows * kWarp, block);
  embed_ln_kernel<T><<<static_cast<unsigned>(grid), block, 0, stream>>>(ids, types, pos, word, type_emb, pos_emb, gamma,
                                                                        beta, out, rows, C, eps);
  check_launch("embed_ln");
}

template <typename T>
void im2col_pixels(const float* pixels, const int* img_idx, T* out, int n_img, int H, int W, int ksz, int stride,
                   int pad, int Kpad, cudaStream_t stream) {
  const int Ho = (H + 2 * pad - ksz) / stride + 1, Wo = (W + 2 * pad - ksz) / stride + 1;
  const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
  if (total <= 0) return;
  const size_t smem = static_cast<size_t>(3) * ksz * (W + 8) * sizeof(float);
  if (Kpad % 2 == 0 && Kpad <= 192 && W % 4 == 0 && reinterpret_cast<uintptr_t>(pixels) % 16 == 0 && smem <= 48 * 1024 &&
      reinterpret_cast<uintptr_t>(out) % 8 == 0) {
    launch_chain(im2col_pixels_rows_kernel<T>, dim3(static_cast<unsigned>(n_img * Ho)), dim3(256), smem, stream, pixels, img_idx, out, H, W,
                 Ho, Wo, ksz, stride, pad, Kpad);
  } else {
    launch_chain(im2col_pixels_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, stream, pixels, img_idx, out, n_img, H, W, Ho, Wo, ksz,
                 stride, pad, Kpad);
  }
  check_launch("im2col_pixels");
}

template <typename T>
void im2col_tokens(const T* in, T* out, int n_img, int H, int W, int C, int ksz, int stride, int pad,
                   cudaStream_t stream) {
  const int Ho = (H + 2 * pad - ksz) / stride + 1, Wo = (W + 2 * pad - ksz) / stride + 1;
  CXRM_CHECK(C % Vec16<T>::N == 0, "im2col_tokens needs C multiple of the vector width");
  const long long total = static_cast<long long>(n_img) * Ho * Wo * ksz * ksz * (C / Vec16<T>::N);
  if (total <= 0) return;
  launch_chain(im2col_tokens_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, stream, in, out, n_img, H, W, C, Ho, Wo, ksz, stride, pad);
  check_launch("im2col_tokens");
}

template <typename T>
void ln_dwconv_qkv(const T* x, T* q, T* k, T* v, float* stats, const float* gamma, const float* beta, float eps,
                   const float* w, const float* scale, const float* shift, int n_img, int H, int W, int C, int cls,
                   cudaStream_t stream) {
  if (n_img <= 0) return;
  constexpr int V = 4, TY = 8;   // channels per thread (8- / 16-byte accesses for bf16 / fp32)
  const int cv = C / V;
  CXRM_CHECK(C % V == 0 && cv <= 256, "ln_dwconv_qkv: unsupported channel count");
  const int Hk = (H + 2 - 3) / 2 + 1, Wk = (W + 2 - 3) / 2 + 1;   // stride-2 window centres are the even coordinates
  float2* st = reinterpret_cast<float2*>(stats);
  const long long rows = static_cast<long long>(n_img) * (cls + H * W);
  const bool ok = ln_rows_launch<T, true>(x, C, nullptr, 0, st, nullptr, nullptr, rows, C, eps, stream);
  CXRM_CHECK(ok, "ln_dwconv_qkv: row statistics need 16-byte aligned rows");
  static const bool tiled = std::getenv("CXRM_NO_DWCONV_TILE") == nullptr;
  if (tiled && C % 64 == 0 && n_img <= 65535) {
    const int tiles_x = ceil_div(W, 8);
    const long long blocks16 = static_cast<long long>(tiles_x) * ceil_div(H, 16) * (C / 64) * n_img;
    if (blocks16 >= 148 * 8) {
      dim3 grid(tiles_x * ceil_div(H, 16), C / 64, n_img);
      launch_chain(ln_dwconv_async_kernel<T, 16>, grid, dim3(256), 0, stream, x, q, k, v, st, gamma, beta, w, scale, shift, H, W, C, cls,
                   Hk, Wk, tiles_x);
    } else {
      dim3 grid(tiles_x * ceil_div(H, 8), C / 64, n_img);
      launch_chain(ln_dwconv_async_kernel<T, 8>, grid, dim3(256), 0, stream, x, q, k, v, st, gamma, beta, w, scale, shift, H, W, C, cls,
                   Hk, Wk, tiles_x);
    }
    check_launch("ln_dwconv_async");
  } else {
    const int cols = std::min(W, 256 / cv);
    dim3 grid(ceil_div(W, cols), ceil_div(H, TY), n_img);
    CXRM_CHECK(grid.z <= 65535, "ln_dwconv_qkv: too many images per chunk");
    launch_chain(ln_dwconv_qkv_kernel<T, TY, V>, grid, dim3(cols * cv), 0, stream, x, q, k, v, st, gamma, beta, w, scale, shift, H, W, C, cls,
                 Hk, Wk, cols);
    check_launch("ln_dwconv_qkv");
  }
  if (cls) {
    launch_chain(ln_cls_kernel<T>, dim3(ceil_div(n_img * C, 256)), dim3(256), 0, stream, x, st, gamma, beta, q, k, v, n_img, H * W, Hk * Wk, C);
    check_launch("ln_cls");
  }
}

template <typename T>
void cat_cls(const T* tokens, const float* cls_token, T* out, int n_img, int HW, int C, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_img) * (1 + HW) * C;
  if (total <= 0) return;
  launch_chain(cat_cls_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, stream, tokens, cls_token, out, n_img, HW, C);
  check_launch("cat_cls");
}

template <typename T>
void drop_cls(const T* in, T* out, int n_img, int HW, int C, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_img) * HW * C;
  if (total <= 0) return;
  launch_chain(drop_cls_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, stream, in, out, n_img, HW, C);
  check_launch("drop_cls");
}

template <typename T>
void gather_rows(const T* src, const int* idx, T* dst, long long n_rows, int C, cudaStream_t stream) {
  if (n_rows <= 0) return;
  CXRM_CHECK(C % Vec16<T>::N == 0, "gather_rows needs C multiple of the vector width");
  move_rows_kernel<T, false><<<grid_for(n_rows * (C / Vec16<T>::N), 256), 256, 0, stream>>>(src, idx, dst, n_rows, C);
  check_launch("gather_rows");
}

template <typename T>
void scatter_rows(const T* src, const int* idx, T* dst, long long n_rows, int C, cudaStream_t stream) {
  if (n_rows <= 0) return;
  CXRM_CHECK(C % Vec16<T>::N == 0, "scatter_rows needs C multiple of the vector width");
  move_rows_kernel<T, true><<<grid_for(n_rows * (C / Vec16<T>::N), 256), 256, 0, stream>>>(src, idx, dst, n_rows, C);
  check_launch("scatter_rows");
}

template <typename TS, typename TD>
void cast_copy(const TS* src, TD* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return;
  cast_copy_kernel<TS, TD><<<grid_for(n, 256), 256, 0, stream>>>(src, dst, n);
  check_launch("cast_copy");
}

void fill_zero(void* p, size_t bytes, cudaStream_t stream) { CXRM_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream)); }

void bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
             float* shift, int C, cudaStream_t stream) {
  bn_fold_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(gamma, beta, mean, var, eps, scale, shift, C);
  check_launch("bn_fold");
}

void lora_merge(float* W, const float* A, const float* B, int n_out, int n_in, int r, float s, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_out) * n_in;
  lora_merge_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(W, A, B, n_out, n_in, r, s);
  check_launch("lora_merge");
}

template <typename T>
void pack_matrix(const float* src, T* dst, int rows, int cols_src, int cols_dst, cudaStream_t stream) {
  const long long total = static_cast<long long>(rows) * cols_dst;
  if (total <= 0) return;
  pack_matrix_kernel<T><<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(src, dst, rows, cols_src,
                                                                                          cols_dst);
  check_launch("pack_matrix");
}

template <typename T>
void pack_conv_khwc(const float* src, T* dst, int Cout, int Cin, int ksz, cudaStream_t stream) {
  const long long total = static_cast<long long>(Cout) * Cin * ksz * ksz;
  pack_conv_khwc_kernel<T><<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(src, dst, Cout, Cin, ksz);
  check_launch("pack_conv_khwc");
}

void pack_dw(const float* src, float* dst, int C, cudaStream_t stream) {
  pack_dw_kernel<<<ceil_div(9 * C, 256), 256, 0, stream>>>(src, dst, C);
  check_launch("pack_dw");
}

#define INST(T)                                                                                                      \
  template void layernorm<T>(const T*, int, T*, int, const float*, const float*, long long, int, float, cudaStream_t); \
  template void embed_ln<T>(const int*, const int*, const int*, const T*, const T*, const T*, const float*,           \
                            const float*, T*, long long, int, float, cudaStream_t);                                   \
  template void im2col_pixels<T>(const float*, const int*, T*, int, int, int, int, int, int, int, cudaStream_t);      \
  template void im2col_tokens<T>(const T*, T*, int, int, int, int, int, int, int, cudaStream_t);                      \
  template void ln_dwconv_qkv<T>(const T*, T*, T*, T*, float*, const float*, const float*, float, const float*,        \
                                 const float*, const float*, int, int, int, int, int, cudaStream_t);                  \
  template void cat_cls<T>(const T*, const float*, T*, int, int, int, cudaStream_t);                                  \
  template void drop_cls<T>(const T*, T*, int, int, int, cudaStream_t);                                               \
  template void gather_rows<T>(const T*, const int*, T*, long long, int, cudaStream_t);                               \
  template void scatter_rows<T>(const T*, const int*, T*, long long, int, cudaStream_t);                              \
  template void pack_matrix<T>(const float*, T*, int, int, int, cudaStream_t);                                        \
  template void pack_conv_khwc<T>(const float*, T*, int, int, int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST
template void cast_copy<float, float>(const float*, float*, long long, cudaStream_t);
template void cast_copy<float, bf16>(const float*, bf16*, long long, cudaStream_t);
template void cast_copy<bf16, float>(const bf16*, float*, long long, cudaStream_t);
template void cast_copy<bf16, bf16>(const bf16*, bf16*, long long, cudaStream_t);

}  // namespace cxrm
