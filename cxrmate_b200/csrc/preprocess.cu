// GPU image preprocessing of the reference's validation / test pipeline:
//   transforms.Compose([Resize(384), CenterCrop([384, 384]), ToTensor(), Normalize(mean, std)])
// (modules/lightning_modules/single.py:248-262, multi.py:89-103) applied to `Image.open(path).convert('RGB')`
// (data/dicom_id.py:91-92).  The resize is Pillow's antialiased bilinear resampler (third-party, Pillow
// src/libImaging/Resample.c): a separable triangle filter whose support grows with the shrink factor, evaluated in
// 22-bit fixed point with a uint8 intermediate, horizontal pass first.  Both passes are reproduced bit for bit; only
// the rows and columns that survive the centre crop are computed.
//
// HBM-bound byte work: the source image (7-23 MB for a chest X-ray) is read once by the horizontal pass (each thread
// walks <= ~17 contiguous taps, neighbouring threads share them through L1), the [rows x 384] intermediate stays in L2.
#include <cmath>
#include <vector>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct AxisCoeffs {
  int ksize = 0;
  std::vector<int> bounds;   // [out][2]: first tap, number of taps
  std::vector<int> kk;       // [out][ksize] fixed-point weights
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter, whole-image box
AxisCoeffs make_coeffs(int in_size, int out_size) {
  AxisCoeffs c;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  c.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  c.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  c.kk.assign(static_cast<size_t>(out_size) * c.ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(c.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = xmax; x < c.ksize; ++x) w[x] = 0.0;
    for (int x = 0; x < c.ksize; ++x) {
      const double v = w[x] * (1 << kPrecisionBits);
      c.kk[static_cast<size_t>(xx) * c.ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    c.bounds[2 * xx] = xmin;
    c.bounds[2 * xx + 1] = xmax;
  }
  return c;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[y - y_first][ox][c] = horizontal pass of source row y at output column left + ox
template <int C>
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, long long pitch, int y_first, int n_rows, int left, int size,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                uint8_t* __restrict__ tmp) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int ry = blockIdx.y;
  if (ox >= size || ry >= n_rows) return;
  const int xx = left + ox;
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* k = kk + static_cast<long long>(xx) * ksize;
  const uint8_t* row = src + static_cast<long long>(y_first + ry) * pitch + static_cast<long long>(x0) * C;
  int acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
  for (int x = 0; x < n; ++x) {
    const int wgt = k[x];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(row[x * C + c]) * wgt;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) tmp[(static_cast<long long>(ry) * size + ox) * C + c] = clip8(acc[c]);
}

// out[c][oy][ox] = ((vertical pass at output row top + oy) / 255 - mean[c]) / std[c]; a grey image fills all 3 channels
template <int C>
__global__ void resize_v_norm_kernel(const uint8_t* __restrict__ tmp, int y_first, int top, int size, const int* __restrict__ bounds,
                                     const int* __restrict__ kk, int ksize, float m0, float m1, float m2, float s0, float s1,
                                     float s2, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= size) return;
  const int yy = top + oy;
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int* k = kk + static_cast<long long>(yy) * ksize;
  int acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < n; ++y) {
    const uint8_t* p = tmp + (static_cast<long long>(y0 - y_first + y) * size + ox) * C;
    const int wgt = k[y];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(p[c]) * wgt;
  }
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn(static_cast<float>(clip8(acc[C == 1 ? 0 : c])), 255.0f);      // ToTensor: uint8 / 255
    out[(static_cast<long long>(c) * size + oy) * size + ox] = __fdiv_rn(__fsub_rn(v, mean[c]), sd[c]);   // Normalize
  }
}

}  // namespace

void preprocess_image(const uint8_t* img_dev, int H, int W, int channels, long long pitch, int size, const float* mean,
                      const float* stdv, float* out, cudaStream_t stream) {
  CXRM_CHECK(img_dev && out && H >= 1 && W >= 1 && (channels == 1 || channels == 3) && size >= 1, "preprocess_image arguments");
  CXRM_CHECK(pitch >= static_cast<long long>(W) * channels, "preprocess_image: row pitch smaller than a row");
  // torchvision Resize(int): the shorter edge becomes `size`, the other int(size * long / short)
  const int nh = W <= H ? static_cast<int>(static_cast<long long>(size) * H / W) : size;
  const int nw = W <= H ? size : static_cast<int>(static_cast<long long>(size) * W / H);
  // torchvision center_crop: int(round((n - size) / 2.0)) with Python's round-half-even
  auto off = [&](int n) { return static_cast<int>(std::nearbyint((n - size) / 2.0)); };
  const int top = off(nh), left = off(nw);
  const AxisCoeffs cx = make_coeffs(W, nw), cy = make_coeffs(H, nh);
  const int y_first = cy.bounds[2 * top];
  const int y_last = cy.bounds[2 * (top + size - 1)] + cy.bounds[2 * (top + size - 1) + 1];
  const int n_rows = y_last - y_first;
  const size_t nbx = cx.bounds.size(), nkx = cx.kk.size(), nby = cy.bounds.size(), nky = cy.kk.size();
  int* tab = nullptr;
  uint8_t* tmp = nullptr;
  CXRM_CUDA_CHECK(cudaMallocAsync(&tab, (nbx + nkx + nby + nky) * sizeof(int), stream));
  CXRM_CUDA_CHECK(cudaMallocAsync(&tmp, static_cast<size_t>(n_rows) * size * channels, stream));
  std::vector<int> host;
  host.reserve(nbx + nkx + nby + nky);
  host.insert(host.end(), cx.bounds.begin(), cx.bounds.end());
  host.insert(host.end(), cx.kk.begin(), cx.kk.end());
  host.insert(host.end(), cy.bounds.begin(), cy.bounds.end());
  host.insert(host.end(), cy.kk.begin(), cy.kk.end());
  CXRM_CUDA_CHECK(cudaMemcpyAsync(tab, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  CXRM_CUDA_CHECK(cudaStreamSynchronize(stream));   // `host` is pageable memory going out of scope
  const int* bx = tab;
  const int* kx = tab + nbx;
  const int* by = kx + nkx;
  const int* ky = by + nby;
  const dim3 blk(128), g1(ceil_div(size, 128), n_rows), g2(ceil_div(size, 128), size);
  if (channels == 1) {
    resize_h_kernel<1><<<g1, blk, 0, stream>>>(img_dev, pitch, y_first, n_rows, left, size, bx, kx, cx.ksize, tmp);
    check_launch("resize_h");
    resize_v_norm_kernel<1><<<g2, blk, 0, stream>>>(tmp, y_first, top, size, by, ky, cy.ksize, mean[0], mean[1], mean[2], stdv[0],
                                                   stdv[1], stdv[2], out);
  } else {
    resize_h_kernel<3><<<g1, blk, 0, stream>>>(img_dev, pitch, y_first, n_rows, left, size, bx, kx, cx.ksize, tmp);
    check_launch("resize_h");
    resize_v_norm_kernel<3><<<g2, blk, 0, stream>>>(tmp, y_first, top, size, by, ky, cy.ksize, mean[0], mean[1], mean[2], stdv[0],
                                                   stdv[1], stdv[2], out);
  }
  check_launch("resize_v_norm");
  CXRM_CUDA_CHECK(cudaFreeAsync(tmp, stream));
  CXRM_CUDA_CHECK(cudaFreeAsync(tab, stream));
}

}  // namespace cxrm
