// bf16 GEMM on the 5th-generation tensor cores of sm_100a:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) -> shared memory ring
//   -> tcgen05.mma (one elected thread, operands from shared-memory descriptors)
//   -> fp32 accumulator in TMEM -> tcgen05.ld -> fused epilogue
//   (bias, exact GELU, residual add, bf16 or fp32 store).
//
// C[M,N] = epi(A[M,K] . W[N,K]^T).  Both operands are K-major, which is the
// natural layout of nn.Linear (x[M,K], weight[N,K]); out-of-range rows/columns
// and the K tail are zero-filled by TMA.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer, warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "kernels.h"

namespace cxrm {

namespace {

constexpr int BM = 128, BK = 64, STAGES = 4, NTHREADS = 192;
constexpr int A_BYTES = BM * BK * 2;   // 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64) = 2)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) |
         (static_cast<uint32_t>(BM >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, GemmArgs g,
                                                           int vec_ok) {
  using C_ = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  if (g.skip_flag && *g.skip_flag) return;

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * C_::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], C_::STAGE_BYTES);
        uint8_t* a_dst = tiles + s * C_::STAGE_BYTES;
        tma_load_2d(a_dst, &tmA, kb * BK, m0, &full[s]);
        tma_load_2d(a_dst + A_BYTES, &tmB, kb * BK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * C_::STAGE_BYTES);
        const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // +32 bytes per 16-element K step inside the 128-byte swizzle atom: start-address field += 2
          umma(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
               (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ---- epilogue: 4 warps x 32 TMEM lanes ---------------------------------------
    const int quarter = warp % 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const long long m = static_cast<long long>(m0) + quarter * 32 + lane;
    const bool row_ok = m < g.M;
    const bf16* __restrict__ R = static_cast<const bf16*>(g.residual);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= g.N) break;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), r);
      if (!row_ok) continue;
      const int nb = n0 + c0;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool full_chunk = (nb + 32 <= g.N);
      if (g.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (full_chunk || nb + j < g.N) ? g.bias[nb + j] : 0.f;
      }
      if (g.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      }
      if (g.c_head_stride > 0) {
        // head-major K/V cache store: 32 consecutive columns never straddle a 64-wide head
        bf16* cp = static_cast<bf16*>(g.C) + static_cast<long long>(nb / 64) * g.c_head_stride + m * 64 + (nb % 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          Vec16<bf16> ov;
          ov.pack(v + 8 * q);
          ov.store(cp + 8 * q);
        }
      } else if (full_chunk && vec_ok) {
        if (R) {
          const bf16* rp = R + m * g.ldr + nb;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec16<bf16> rv;
            rv.load(rp + 8 * q);
            float rf[8];
            rv.unpack(rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * q + j] += rf[j];
          }
        }
        if (g.out_f32) {
          float* cp = static_cast<float*>(g.C) + m * g.ldc + nb;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          bf16* cp = static_cast<bf16*>(g.C) + m * g.ldc + nb;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec16<bf16> ov;
            ov.pack(v + 8 * q);
            ov.store(cp + 8 * q);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nb + j;
          if (n >= g.N) continue;
          float o = v[j];
          if (R) o += to_f(R[m * g.ldr + n]);
          if (g.out_f32)
            static_cast<float*>(g.C)[m * g.ldc + n] = o;
          else
            static_cast<bf16*>(g.C)[m * g.ldc + n] = from_f<bf16>(o);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::TMEM_COLS)
                 : "memory");
  }
}

// ---- host side -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [rows, cols] with row pitch ld elements; box = [64 cols, box_rows rows]; 128-byte swizzle
CUtensorMap make_map(const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed (CUresult " + std::to_string(static_cast<int>(r)) + ")");
  return m;
}

int pick_bn(int M, int N) {
  if (M <= 256) {
    // weight-streaming regime (decode steps): many narrow tiles so that enough CTAs pull weights concurrently
    if (N <= 32 || N >= 512) return 32;
  }
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N % 256 != 0 && N % 192 == 0 && N <= 768) return 192;
  return 256;
}

template <int BN>
void launch(const GemmArgs& g, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    configured = true;
  }
  const CUtensorMap ta = make_map(g.A, g.M, g.K, g.lda, BM);
  const CUtensorMap tb = make_map(g.W, g.N, g.K, g.ldw, BN);
  const int esz = g.out_f32 ? 4 : 2;
  int vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((static_cast<long long>(g.ldc) * esz) % 16 == 0);
  if (g.residual)
    vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(g.residual) % 16 == 0) && (g.ldr % 8 == 0);
  dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN));
  gemm_tc_kernel<BN><<<grid, NTHREADS, Cfg<BN>::SMEM, stream>>>(ta, tb, g, vec_ok);
  check_launch("gemm_tcgen05");
}

}  // namespace

int gemm_tcgen05_supported(const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 1;
  if (g.K % 8 != 0 || g.lda % 8 != 0 || g.ldw % 8 != 0) return 2;
  if (reinterpret_cast<uintptr_t>(g.A) % 16 != 0 || reinterpret_cast<uintptr_t>(g.W) % 16 != 0) return 3;
  if (ceil_div(g.N, 32) > 65535) return 4;
  if (g.c_head_stride > 0 && (g.N % 64 != 0 || g.out_f32 || g.residual || reinterpret_cast<uintptr_t>(g.C) % 16 != 0 ||
                              g.c_head_stride % 8 != 0))
    return 5;
  return 0;
}

void gemm_tcgen05(const GemmArgs& g, cudaStream_t stream) {
  CXRM_CHECK(gemm_tcgen05_supported(g) == 0, "shape/alignment not supported by the tcgen05 GEMM");
  switch (pick_bn(g.M, g.N)) {
    case 32: launch<32>(g, stream); break;
    case 64: launch<64>(g, stream); break;
    case 128: launch<128>(g, stream); break;
    case 192: launch<192>(g, stream); break;
    default: launch<256>(g, stream); break;
  }
}

}  // namespace cxrm
