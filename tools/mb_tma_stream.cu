// Microbenchmark: how fast can one B200 stream a K/V-cache-like byte range into shared memory?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_tma tools/mb_tma_stream.cu -lcuda && /tmp/mb_tma
// Variants: 2-D TMA boxes (rows x 128 B, 128B swizzle) through an S-stage mbarrier ring with G CTAs per SM,
// 1-D cp.async.bulk of the same bytes, and plain 16-byte loads.  Each for an L2-resident range (second pass over
// 48 MB), and cold ranges of 157 MB and 1 GB (L2 flushed by a 512 MB memset in between).
// Answers what bounds decode_cross_persist_kernel (profiles/README notes): the memory system or the ring structure.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}

// one thread per CTA drives an S-stage ring; tiles are assigned round-robin over the grid (tile t -> CTA t % grid)
template <int MODE>   // 0: 2-D TMA box, 1: 1-D bulk copy
__global__ void ring_kernel(const __grid_constant__ CUtensorMap tm, const char* base, int box_rows, int stages, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = box_rows * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + stages * tile_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long issued = 0, done = 0;
    const long long first = blockIdx.x, step = gridDim.x;
    const long long mine = first < n_tiles ? (n_tiles - first + step - 1) / step : 0;
    auto issue = [&](long long k) {
      const int s = static_cast<int>(k % stages);
      const long long t = first + k * step;
      mbar_expect(&full[s], tile_bytes);
      if (MODE == 0) tma2d(tiles + s * tile_bytes, &tm, 0, static_cast<int>(t * box_rows), &full[s]);
      else bulk1d(tiles + s * tile_bytes, base + t * tile_bytes, tile_bytes, &full[s]);
    };
    for (; issued < mine && issued < stages; ++issued) issue(issued);
    for (; done < mine; ++done) {
      mbar_wait(&full[done % stages], (done / stages) & 1);
      if (issued < mine) issue(issued++);   // the stage just drained is refilled at once
    }
  }
}

__global__ void ldg_kernel(const uint4* __restrict__ p, long long n, unsigned* sink) {
  unsigned acc = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n; i += stride) { uint4 v = __ldg(p + i); acc ^= v.x ^ v.y; }
  if (acc == 0x12345678u) *sink = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t GB = 1ull << 30;
  char* buf; char* flush; unsigned* sink;
  CK(cudaMalloc(&buf, GB)); CK(cudaMalloc(&flush, GB / 2)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, GB));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));

  auto timeit = [&](const char* name, size_t bytes, bool warm, auto&& launch) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      if (!warm) CK(cudaMemsetAsync(flush, rep, GB / 2));
      else launch();   // untimed pass that leaves the range in L2
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-58s %7.1f MB %s  %8.1f us  %7.0f GB/s\n", name, bytes / 1e6, warm ? "L2-warm" : "cold   ", best * 1e3, bytes / (best * 1e-3) / 1e9);
  };

  struct Range { size_t bytes; bool warm; };
  const Range ranges[] = {{48ull << 20, true}, {157ull * 1000 * 1000, false}, {GB, false}};
  for (const Range& r : ranges) {
    const size_t bytes = r.bytes / (256 * 128) * (256 * 128);
    {
      char nm[128]; snprintf(nm, sizeof nm, "ldg.128 x8 unrolled, 148x8 CTAs x 256");
      timeit(nm, bytes, r.warm, [&] { ldg_kernel<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, sink); });
    }
    for (int mode = 0; mode < 2; ++mode)
      for (int box : {64, 192, 256})
        for (int cfg = 0; cfg < 5; ++cfg) {
          const int per_sm[5] = {1, 2, 2, 4, 1}, stg[5] = {2, 2, 4, 2, 4};
          const int tile = box * 128, stages = stg[cfg], g = per_sm[cfg];
          const size_t smem = static_cast<size_t>(stages) * tile + 1024 + 256;
          if (smem * g > 225 * 1024) continue;
          CUtensorMap tm;
          const cuuint64_t dims[2] = {64, bytes / 128}; const cuuint64_t strides[1] = {128};
          const cuuint32_t bx[2] = {64, static_cast<cuuint32_t>(box)}; const cuuint32_t es[2] = {1, 1};
          CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
          const long long n_tiles = bytes / tile;
          char nm[128]; snprintf(nm, sizeof nm, "%s box %3d rows (%2d KB), %d CTA/SM x %d stages", mode == 0 ? "TMA-2D " : "bulk-1D", box, tile / 1024, g, stages);
          // pad the dynamic smem so that exactly g CTAs fit an SM
          const size_t pad = std::max(smem, static_cast<size_t>(224 * 1024 / g - 1024));
          if (mode == 0) timeit(nm, bytes, r.warm, [&] { ring_kernel<0><<<148 * g, 32, pad>>>(tm, buf, box, stages, n_tiles); });
          else timeit(nm, bytes, r.warm, [&] { ring_kernel<1><<<148 * g, 32, pad>>>(tm, buf, box, stages, n_tiles); });
          CK(cudaGetLastError());
        }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
