// Microbenchmark: sustained tcgen05.mma rate of ONE CTA per SM with the operand layout the GEMMs use (bf16, K-major,
// 128-byte swizzle, 64-wide k-blocks = 4 MMAs of K = 16), operands resident in shared memory, no TMA, no epilogue.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_mma tools/mb_mma.cu && /tmp/mb_mma
// Prints clocks per MMA and the fraction of the tensor-pipe floor (128 * N / 256 clocks per 128 x N x 16 MMA) for
// N = 64 .. 256, for 1 .. 4 distinct k-block stages cycled through, committing every k-block or once.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int BN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int kblocks, int stages, int commit_each, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < stages * STAGE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x / 32 == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (threadIdx.x / 32 == 1) {
    constexpr uint32_t idesc = make_idesc(BN);
    const uint32_t base = smem_u32(tiles);
    long long t0 = clock64();
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % stages;
      const uint64_t da = make_desc(base + s * STAGE), db = make_desc(base + s * STAGE + A_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        if (commit_each) umma_commit(&bar[1]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar[0]);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar[0], 0);
    long long t2 = clock64();
    if (threadIdx.x % 32 == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x / 32 == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

template <int BN>
void run(int grid, long long* dout) {
  for (int stages : {1, 4})
    for (int commit_each : {0, 1}) {
      const int kblocks = 256;
      const size_t smem = static_cast<size_t>(stages) * (128 * 128 + BN * 128) + 2048;
      CK(cudaFuncSetAttribute(mma_rate_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      mma_rate_kernel<BN><<<grid, 128, smem>>>(kblocks, stages, commit_each, dout);
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
      const double per = static_cast<double>(h[1]) / (kblocks * 4), floor_ = 128.0 * BN / 256.0;
      printf("N=%3d grid=%3d stages=%d commit_each=%d: issue %.1f clk/MMA, complete %.1f clk/MMA (floor %.0f -> %.0f%% of the tensor pipe)\n",
             BN, grid, stages, commit_each, static_cast<double>(h[0]) / (kblocks * 4), per, floor_, 100.0 * floor_ / per);
    }
}

int main() {
  long long* dout;
  CK(cudaMalloc(&dout, 16));
  for (int grid : {1, 148}) {
    run<64>(grid, dout);
    run<128>(grid, dout);
    run<192>(grid, dout);
    run<256>(grid, dout);
  }
  return 0;
}
