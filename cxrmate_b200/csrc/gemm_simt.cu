// Strict-fp32 FMA GEMM with fused epilogue: the arithmetic of the fp32 validation
// mode (no TF32, no tensor cores, k accumulated in ascending order per output),
// and the debug path for bf16 storage.
//
// C[M,N] = epi(A[M,K] . W[N,K]^T)  -- the nn.Linear contraction used by every
// dense layer on the path (HF modeling_cvt.py:232-234,258,304,315;
// modeling_bert.py:172-174,291,334,347,476,487).
#include "kernels.h"

namespace cxrm {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256, TM = 8, TN = 4;

template <typename T>
__device__ __forceinline__ void load8(const T* p, bool ok, float* f);
template <>
__device__ __forceinline__ void load8<float>(const float* p, bool ok, float* f) {
  if (ok) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
  }
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, bool ok, float* f) {
  if (ok) {
    Vec16<bf16> v;
    v.load(p);
    v.unpack(f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  if (g.skip_flag && *g.skip_flag) return;
  const T* __restrict__ A = static_cast<const T*>(g.A);
  const T* __restrict__ W = static_cast<const T*>(g.W);
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid / 16, tx = tid % 16;

  // loader mapping: A tile 128 rows x 16 k -> thread loads 8 k of one row; W tile 64 rows x 16 k -> 4 k of one row
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int w_row = tid >> 2, w_k = (tid & 3) * 4;
  const bool a_row_ok = (m0 + a_row) < g.M;
  const bool w_row_ok = (n0 + w_row) < g.N;
  const T* a_ptr = A + (m0 + a_row) * static_cast<long long>(g.lda) + a_k;
  const T* w_ptr = W + static_cast<long long>(n0 + w_row) * g.ldw + w_k;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[8], rw[8];
  auto fetch = [&](int k0) {
    load8<T>(a_ptr + k0, a_row_ok && (k0 + a_k) < g.K, ra);
    // W: 4 elements (K % 8 == 0 so a 4-chunk is either fully inside or fully outside)
    if (w_row_ok && (k0 + w_k) < g.K) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rw[i] = to_f(w_ptr[k0 + i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) rw[i] = 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_k + i][a_row] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[w_k + i][w_row] = rw[i];
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const T* __restrict__ R = static_cast<const T*>(g.residual);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.act == ACT_GELU) v = gelu_erf(v);
      if (R) v += to_f(R[m * g.ldr + n]);
      if (g.c_head_stride > 0)
        static_cast<T*>(g.C)[static_cast<long long>(n / 64) * g.c_head_stride + m * 64 + (n % 64)] = from_f<T>(v);
      else if (g.out_f32)
        static_cast<float*>(g.C)[m * g.ldc + n] = v;
      else
        static_cast<T*>(g.C)[m * g.ldc + n] = from_f<T>(v);
    }
  }
}

}  // namespace

template <typename T>
void gemm_simt(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return;
  CXRM_CHECK(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm_simt needs K, lda, ldw multiples of 8");
  dim3 grid(static_cast<unsigned>(ceil_div_ll(g.M, BM)), static_cast<unsigned>(ceil_div(g.N, BN)));
  CXRM_CHECK(grid.y <= 65535, "gemm_simt N too large");
  gemm_simt_kernel<T><<<grid, NT, 0, stream>>>(g);
  check_launch("gemm_simt");
}

template void gemm_simt<float>(const GemmArgs&, cudaStream_t);
template void gemm_simt<bf16>(const GemmArgs&, cudaStream_t);

}  // namespace cxrm
