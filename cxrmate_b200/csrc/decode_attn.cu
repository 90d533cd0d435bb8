// One-token (decode-step) attention over the K/V caches: the HBM-bound core of the
// rollout (18,432 B per cached token per row, cross K/V
// counted once per study).
//
// Both caches are HEAD-MAJOR: for one layer, head h of key j lives at
// base + (h * tokens + j) * 64, so the keys a (study|row, head) pair attends to
// are ONE contiguous byte range.  The work is cut into uniform units
// (group, head, chunk of <= CH keys); a unit
//   1. pulls its K and V chunk into shared memory with two bulk async copies
//      (cp.async.bulk -> UBLKCP, completion on an mbarrier): the whole chunk is
//      in flight from one thread, independent of occupancy and registers;
//   2. scores = q.K^T * 1/8 for the NQ query rows that share the chunk (the
//      sample and greedy rows of a study share the encoder K/V: one read);
//   3. chunk-local softmax statistics and sum_j p_j V_j;
//   4. writes (max, sum, out[64]) partials; the LAST unit of a (group, head) to
//      arrive (atomic ticket) merges the partials and stores the context row, so
//      no second kernel is needed.
// Uniform units remove the load imbalance of one-block-per-study (1..5 images).
//
// Reference semantics: HF modeling_bert.py:143-207 (self, cache append, key padding
// mask as finfo.min == skipped keys) and :210-284 (cross, encoder mask: masked
// tokens were dropped from the cache at cxrm_prefill_cross_kv).
#include "kernels.h"

namespace cxrm {

namespace {

constexpr int H = 768, HD = 64, NH = 12;
constexpr int NT = 128;

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
// 1-D bulk async copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename T, int NQ>
struct UnitSmem {
  static constexpr int VN = Vec16<T>::N;
  static constexpr int LPK = HD / VN;   // lanes per key
  static constexpr int NG = NT / LPK;   // key groups per block
};

// Shared-memory carve-up (dynamic): K[CH][64] T | V[CH][64] T | sc[NQ][CH] f32 | 2 mbarriers.
// The cross-group reduction buffer red[NG][NQ][64] f32 aliases the K region (dead after the score pass).
template <typename T, int NQ>
__host__ __device__ constexpr size_t unit_smem_bytes(int CH) {
  return static_cast<size_t>(2) * CH * HD * sizeof(T) + static_cast<size_t>(NQ) * CH * sizeof(float) + 64;
}

// The part shared by self- and cross-attention.  On entry K/V chunk copies have been ISSUED on bars[0]/bars[1]
// (n keys); `fix_last` optionally overrides key n-1 (the token being appended) after the copies land.
//   qf     : this thread's slice of the NQ query rows (dims sub*VN .. +VN), fp32
//   valid  : nullable per-key validity bytes for this chunk (self-attention key padding mask)
//   part   : partial slots of this unit: part[i * part_row_stride] -> (m, l, o[64]) of query row i
template <typename T, int NQ>
__device__ __forceinline__ void unit_core(const T* Ks, const T* Vs, float* sc, float* red, uint64_t* bars, int n, int CH,
                                          const float (&qf)[NQ][Vec16<T>::N], const uint8_t* __restrict__ valid,
                                          float* __restrict__ part, long long part_row_stride) {
  constexpr int VN = Vec16<T>::N, LPK = HD / VN, NG = NT / LPK;
  __shared__ float sh_red[NQ][NT / kWarp];
  const int tid = threadIdx.x, g = tid / LPK, sub = tid % LPK;

  // ---- scores -------------------------------------------------------------------
  mbar_wait(&bars[0], 0);
  for (int jb = 0; jb < n; jb += NG) {   // warp-uniform trip count (group shuffles need every lane)
    const int j = jb + g;
    float kf[VN];
    if (j < n) {
      Vec16<T> kv;
      kv.load(Ks + static_cast<long long>(j) * HD + sub * VN);
      kv.unpack(kf);
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) kf[e] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < VN; ++e) d = fmaf(qf[i][e], kf[e], d);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
      if (j < n && sub == 0) sc[i * CH + j] = (!valid || valid[j]) ? d * 0.125f : -INFINITY;
    }
  }
  __syncthreads();
  // ---- chunk-local softmax statistics ----------------------------------------------
  float mx[NQ], sum[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float m = -INFINITY;
    for (int j = tid; j < n; j += NT) m = fmaxf(m, sc[i * CH + j]);
    m = warp_max(m);
    if (tid % kWarp == 0) sh_red[i][tid / kWarp] = m;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NT / kWarp; ++w) m = fmaxf(m, sh_red[i][w]);
    mx[i] = m;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float sm = 0.f;
    for (int j = tid; j < n; j += NT) {
      const float s = sc[i * CH + j];
      const float e = (s == -INFINITY) ? 0.f : expf(s - mx[i]);
      sc[i * CH + j] = e;
      sm += e;
    }
    sm = warp_sum(sm);
    if (tid % kWarp == 0) sh_red[i][tid / kWarp] = sm;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    float sm = 0.f;
#pragma unroll
    for (int w = 0; w < NT / kWarp; ++w) sm += sh_red[i][w];
    sum[i] = sm;
  }
  // ---- weighted sum of V -----------------------------------------------------------
  mbar_wait(&bars[1], 0);
  float acc[NQ][VN];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[i][e] = 0.f;
  for (int j = g; j < n; j += NG) {
    Vec16<T> vv;
    vv.load(Vs + static_cast<long long>(j) * HD + sub * VN);
    float vf[VN];
    vv.unpack(vf);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float pj = sc[i * CH + j];
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[i][e] = fmaf(pj, vf[e], acc[i][e]);
    }
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) red[(g * NQ + i) * HD + sub * VN + e] = acc[i][e];
  __syncthreads();
  if (tid < NQ * HD) {
    const int i = tid / HD, d = tid % HD;
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) o += red[(gg * NQ + i) * HD + d];
    float* w = part + i * part_row_stride;
    w[2 + d] = o;
    if (d == 0) {
      float m_i = mx[0], s_i = sum[0];
#pragma unroll
      for (int q = 1; q < NQ; ++q)
        if (i == q) {
          m_i = mx[q];
          s_i = sum[q];
        }
      w[0] = m_i;
      w[1] = s_i;
    }
  }
}

// merge `nchunk` partials (stride HD+2 floats) of one (row, head) and store the context row
template <typename T>
__device__ __forceinline__ void merge_partials(const float* __restrict__ w, int nchunk, T* __restrict__ dst, int d) {
  float m = -INFINITY;
  for (int c = 0; c < nchunk; ++c) m = fmaxf(m, __ldcg(w + c * (HD + 2)));
  float l = 0.f, o = 0.f;
  for (int c = 0; c < nchunk; ++c) {
    const float mc = __ldcg(w + c * (HD + 2));
    if (mc == -INFINITY) continue;
    const float a = expf(mc - m);
    l += a * __ldcg(w + c * (HD + 2) + 1);
    o += a * __ldcg(w + c * (HD + 2) + 2 + d);
  }
  dst[d] = from_f<T>(l > 0.f ? o / l : 0.f);
}

// =============================================================================
// cross-attention: grid (max_units, NH); unit table built at cxrm_prefill_cross_kv
// =============================================================================
template <typename T, int NQ>
__global__ void __launch_bounds__(NT) decode_cross_units_kernel(const T* __restrict__ q, int ldq,
                                                                const T* __restrict__ kc, const T* __restrict__ vc,
                                                                long long head_stride, T* __restrict__ ctx,
                                                                CrossUnits cu, RolloutState st, int B, int CH,
                                                                float* __restrict__ ws, unsigned* __restrict__ tickets) {
  constexpr int VN = Vec16<T>::N, LPK = HD / VN, NG = NT / LPK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw);
  T* Vs = Ks + static_cast<size_t>(CH) * HD;
  float* sc = reinterpret_cast<float*>(Vs + static_cast<size_t>(CH) * HD);
  float* red = reinterpret_cast<float*>(Ks);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sc + NQ * CH);
  static_assert(NG * NQ * HD * sizeof(float) <= 96 * HD * sizeof(T), "reduction buffer must fit the K chunk");
  __shared__ int sh_last;

  if (*st.done) return;
  const int u = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  if (u >= *cu.n_units) return;
  const int b = cu.study[u], j0 = cu.j0[u], n = cu.n[u], c = cu.chunk[u];
  bool all_fin = true;
#pragma unroll
  for (int i = 0; i < NQ; ++i) all_fin = all_fin && st.finished[b + i * B];
  if (all_fin) return;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = static_cast<uint32_t>(n) * HD * sizeof(T);
    const long long off = static_cast<long long>(h) * head_stride + static_cast<long long>(j0) * HD;
    mbar_expect_tx(&bars[0], bytes);
    bulk_g2s(Ks, kc + off, bytes, &bars[0]);
    mbar_expect_tx(&bars[1], bytes);
    bulk_g2s(Vs, vc + off, bytes, &bars[1]);
  }
  const int sub = tid % LPK;
  float qf[NQ][VN];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    Vec16<T> qv;
    qv.load(q + static_cast<long long>(b + i * B) * ldq + h * HD + sub * VN);
    qv.unpack(qf[i]);
  }
  __syncthreads();   // barrier inits visible to every waiter

  const int maxc = cu.max_chunks;
  float* part = ws + ((static_cast<long long>(b) * NH + h) * maxc + c) * (HD + 2);
  const long long row_stride = static_cast<long long>(B) * NH * maxc * (HD + 2);   // query row i = b + i*B
  unit_core<T, NQ>(Ks, Vs, sc, red, bars, n, CH, qf, nullptr, part, row_stride);

  // ---- last unit of (study, head) merges -----------------------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned nchunk = static_cast<unsigned>(cu.n_chunks[b]);
    const unsigned prev = atomicAdd(&tickets[b * NH + h], 1u);
    sh_last = (prev == nchunk - 1) ? 1 : 0;
    if (sh_last) tickets[b * NH + h] = 0;   // ready for the next launch
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    if (tid < NQ * HD) {
      const int i = tid / HD, d = tid % HD;
      const int r = b + i * B;
      const float* w = ws + i * row_stride + (static_cast<long long>(b) * NH + h) * maxc * (HD + 2);
      merge_partials<T>(w, cu.n_chunks[b], ctx + static_cast<long long>(r) * H + h * HD, d);
    }
  }
}

// =============================================================================
// self-attention: grid (R * max_chunks, NH); appends the new token's K/V
// =============================================================================
template <typename T>
__global__ void __launch_bounds__(NT) decode_self_units_kernel(const T* __restrict__ qkv, T* __restrict__ kcache,
                                                               T* __restrict__ vcache, T* __restrict__ ctx,
                                                               RolloutState st, int Lmax, int CH, int max_chunks,
                                                               float* __restrict__ ws,
                                                               unsigned* __restrict__ tickets) {
  constexpr int VN = Vec16<T>::N, LPK = HD / VN, NG = NT / LPK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw);
  T* Vs = Ks + static_cast<size_t>(CH) * HD;
  float* sc = reinterpret_cast<float*>(Vs + static_cast<size_t>(CH) * HD);
  float* red = reinterpret_cast<float*>(Ks);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sc + CH);
  __shared__ int sh_last;

  if (*st.done) return;
  const int r = blockIdx.x / max_chunks, c = blockIdx.x % max_chunks, h = blockIdx.y, tid = threadIdx.x;
  if (st.finished[r]) return;
  const int L = st.cur_len[r];   // slot of the token being fed; keys 0..L
  const int ntot = L + 1;
  const int nchunk = ceil_div(ntot, CH);
  if (c >= nchunk) return;
  const int j0 = c * CH, n = min(CH, ntot - j0);
  const bool has_new = (c == nchunk - 1);          // this chunk ends with the token being appended
  const int n_cached = has_new ? n - 1 : n;

  const long long base = (static_cast<long long>(r) * NH + h) * Lmax * HD;   // [row][head][Lmax][64]
  T* kbase = kcache + base;
  T* vbase = vcache + base;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = static_cast<uint32_t>(n_cached) * HD * sizeof(T);
    if (bytes > 0) {
      mbar_expect_tx(&bars[0], bytes);
      bulk_g2s(Ks, kbase + static_cast<long long>(j0) * HD, bytes, &bars[0]);
      mbar_expect_tx(&bars[1], bytes);
      bulk_g2s(Vs, vbase + static_cast<long long>(j0) * HD, bytes, &bars[1]);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    }
  }
  const T* qrow = qkv + static_cast<long long>(r) * 3 * H + h * HD;
  if (has_new) {
    // the new token's key / value come from the QKV projection: into the chunk and into the cache
    if (tid < HD) {
      const T kvn = qrow[H + tid];
      Ks[static_cast<long long>(n - 1) * HD + tid] = kvn;
      kbase[static_cast<long long>(L) * HD + tid] = kvn;
    } else {
      const int d = tid - HD;
      const T vvn = qrow[2 * H + d];
      Vs[static_cast<long long>(n - 1) * HD + d] = vvn;
      vbase[static_cast<long long>(L) * HD + d] = vvn;
    }
  }
  const int sub = tid % LPK;
  float qf[1][VN];
  {
    Vec16<T> qv;
    qv.load(qrow + sub * VN);
    qv.unpack(qf[0]);
  }
  __syncthreads();

  float* part = ws + ((static_cast<long long>(r) * NH + h) * max_chunks + c) * (HD + 2);
  unit_core<T, 1>(Ks, Vs, sc, red, bars, n, CH, qf, st.key_valid + static_cast<long long>(r) * Lmax + j0, part, 0);

  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(&tickets[r * NH + h], 1u);
    sh_last = (prev == static_cast<unsigned>(nchunk) - 1) ? 1 : 0;
    if (sh_last) tickets[r * NH + h] = 0;
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    if (tid < HD)
      merge_partials<T>(ws + (static_cast<long long>(r) * NH + h) * max_chunks * (HD + 2), nchunk,
                        ctx + static_cast<long long>(r) * H + h * HD, tid);
  }
}

// qkv [R*P, 3*768] -> head-major caches [R][12][Lmax][64], columns [0,P)
template <typename T>
__global__ void prefill_store_kv_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                                        int R, int P, int Lmax) {
  constexpr int VN = Vec16<T>::N;
  const int cv = H / VN;
  const long long total = static_cast<long long>(R) * P * cv * 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * VN;
    long long t = i / cv;
    const int which = static_cast<int>(t % 2);
    t /= 2;
    const int pcol = static_cast<int>(t % P);
    const long long r = t / P;
    Vec16<T> v;
    v.load(qkv + (r * P + pcol) * 3 * H + (1 + which) * H + c);
    const int h = c / HD, d = c % HD;
    v.store((which ? vcache : kcache) + ((r * NH + h) * Lmax + pcol) * HD + d);
  }
}

}  // namespace

int decode_attn_chunk(size_t elem_size) { return elem_size == 2 ? 192 : 96; }

size_t decode_attn_ws_floats(int rows, int max_chunks) {
  return static_cast<size_t>(rows) * NH * max_chunks * (HD + 2);
}

template <typename T>
void decode_self_attention(const T* qkv, T* kcache, T* vcache, T* ctx, const RolloutState& st, int R, int Lmax,
                           float* ws, unsigned* tickets, cudaStream_t stream) {
  const int CH = decode_attn_chunk(sizeof(T));
  const int max_chunks = ceil_div(Lmax, CH);
  const size_t smem = unit_smem_bytes<T, 1>(CH);
  static size_t configured = 0;
  if (smem > configured) {
    CXRM_CUDA_CHECK(cudaFuncSetAttribute(decode_self_units_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    configured = smem;
  }
  decode_self_units_kernel<T><<<dim3(R * max_chunks, NH), NT, smem, stream>>>(qkv, kcache, vcache, ctx, st, Lmax, CH,
                                                                             max_chunks, ws, tickets);
  check_launch("decode_self_attention");
}

template <typename T>
void decode_cross_attention(const T* q, int ldq, const T* kc, const T* vc, long long head_stride, T* ctx,
                            const CrossUnits& cu, const RolloutState& st, int R, int B, float* ws, unsigned* tickets,
                            cudaStream_t stream) {
  const int nq = R / B;
  CXRM_CHECK(nq == 1 || nq == 2, "decode_cross_attention: 1 or 2 rows per study");
  const int CH = decode_attn_chunk(sizeof(T));
  auto launch = [&](auto kern, size_t smem) {
    static size_t configured = 0;
    if (smem > configured) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured = smem;
    }
    kern<<<dim3(cu.max_units, NH), NT, smem, stream>>>(q, ldq, kc, vc, head_stride, ctx, cu, st, B, CH, ws, tickets);
  };
  if (nq == 1)
    launch(decode_cross_units_kernel<T, 1>, unit_smem_bytes<T, 1>(CH));
  else
    launch(decode_cross_units_kernel<T, 2>, unit_smem_bytes<T, 2>(CH));
  check_launch("decode_cross_attention");
}

template <typename T>
void prefill_store_kv(const T* qkv, T* kcache, T* vcache, int R, int P, int Lmax, cudaStream_t stream) {
  const long long total = static_cast<long long>(R) * P * (H / Vec16<T>::N) * 2;
  if (total <= 0) return;
  long long grid = ceil_div_ll(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  prefill_store_kv_kernel<T><<<static_cast<unsigned>(grid), 256, 0, stream>>>(qkv, kcache, vcache, R, P, Lmax);
  check_launch("prefill_store_kv");
}

#define INST(T)                                                                                                      \
  template void decode_self_attention<T>(const T*, T*, T*, T*, const RolloutState&, int, int, float*, unsigned*,      \
                                         cudaStream_t);                                                               \
  template void decode_cross_attention<T>(const T*, int, const T*, const T*, long long, T*, const CrossUnits&,        \
                                          const RolloutState&, int, int, float*, unsigned*, cudaStream_t);            \
  template void prefill_store_kv<T>(const T*, T*, T*, int, int, int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace cxrm
