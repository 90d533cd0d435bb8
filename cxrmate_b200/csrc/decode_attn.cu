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
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

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


// =====================================================================================================
// bf16 path: the same units, but the two small contractions of a unit (scores = q.K^T, out = p.V) run on the
// tensor cores through mma.sync.m16n8k16 with the NQ (1-2) query rows padded to the 16-row tile.  This is not
// about FLOPs (the kernel is HBM-bound): the SIMT formulation needs ~5200 warp-instructions per 48 KiB unit
// (unpack + shuffle reductions dominate) and ncu showed it issue-bound at 2.7 TB/s; the MMA formulation
// needs ~450.  K/V chunks arrive by 2-D TMA (cp.async.bulk.tensor) with the 128-byte swizzle so that the
// ldmatrix reads of 8 keys x 16 B are bank-conflict free.
// =====================================================================================================
#ifndef CXRM_CHB
#define CXRM_CHB 192
#endif
#ifndef CXRM_PSTAGES
#define CXRM_PSTAGES 2
#endif
constexpr int CHB = CXRM_CHB;       // keys per unit (bf16): 192 or 128 (4 warps x a multiple of 16 keys)
constexpr int WKEYS = CHB / 4;      // keys per warp
constexpr int WNT = WKEYS / 8;      // 8-key score tiles per warp
constexpr int WKK = WKEYS / 16;     // 16-key steps of p.V per warp
static_assert(CHB % 64 == 0 && CHB <= 256, "CHB: whole 64-row TMA boxes of the self cache, one TMA box of the cross cache");
constexpr int SELF_BOX = 64;        // rows per TMA box of the self cache (a unit loads only the boxes it needs)

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of (key row r, 16-byte chunk c) inside a 128B-swizzled tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

// ---- persistent formulation ------------------------------------------------------------------------
// One CTA per SM for the whole launch.  Warp 4 is the TMA producer: it walks the CTA's contiguous range of
// work items and keeps a 4-stage ring of (K chunk, V chunk) tiles = 192 KiB per SM requested ahead of the
// four consumer warps, so that HBM requests never drain while a tile is being reduced (the one-unit-per-CTA
// version left the memory pipe idle during every unit's epilogue and CTA relaunch: 3.5 TB/s in ncu).
// Consumers keep flash-style running (max, sum, out) state in registers across consecutive chunks of the same
// (group, head) and only flush when the group changes, so most groups are finished by a single CTA without
// touching the global partial buffer; groups that straddle CTAs use the ticket merge.
constexpr int PSTAGES = CXRM_PSTAGES;           // per CTA; two CTAs per SM -> 4 x 48 KiB requested ahead per SM (CHB 192, 2 stages)
constexpr int TILE_BYTES = CHB * 128;           // one K or V chunk
constexpr int ST_Q = 2 * TILE_BYTES;            // stage layout: K | V | q rows (2 x 128 B) | key mask (<= 192 B)
constexpr int ST_MASK = ST_Q + 256;
constexpr int STAGE_BYTES = ST_MASK + 256 + 512;   // 50176 = 49 KiB: keeps every stage 1024-byte aligned
static_assert(STAGE_BYTES % 1024 == 0, "stage alignment (128B swizzle)");
constexpr int NCONS = 128;                      // consumer threads (4 warps); warp 4 = producer
constexpr int PNT = NCONS + 32;
constexpr int MAXI = 128;                       // work items per CTA (metadata staged in shared memory)
constexpr size_t kPersistSmem = static_cast<size_t>(PSTAGES) * STAGE_BYTES + 4 * 2 * (HD + 2) * sizeof(float) +
                                5 * MAXI * sizeof(int) + MAXI + 256 + 1024;

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// running flash state of one consumer warp: rows g < NQ live in lanes 4g..4g+3
struct WarpAcc {
  float m, l;
  float o[8][2];
  __device__ __forceinline__ void reset() {
    m = -INFINITY;
    l = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = 0.f;
  }
};

// One consumer warp folds keys [w*48, w*48+48) of the staged chunk (n valid keys) into its running state.
template <int NQ>
__device__ __forceinline__ void warp_tile_update(uint32_t Ks, uint32_t Vs, int n, const uint32_t (&qa0)[4],
                                                 const uint32_t (&qa2)[4], const uint8_t* __restrict__ valid,
                                                 WarpAcc& acc) {
  const int lane = threadIdx.x % kWarp, w = threadIdx.x / kWarp;
  const int g = lane >> 2, t = lane & 3;
  const int k0 = w * WKEYS;
  if (k0 >= n) return;   // warp-uniform
  // ---- scores: 6 n-tiles of 8 keys ----
  float sc[WNT][4];
  const int mi = lane >> 3, mr = lane & 7;   // ldmatrix: this lane addresses row mr of matrix mi
#pragma unroll
  for (int nt = 0; nt < WNT; ++nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[nt][j] = 0.f;
    const int key = k0 + nt * 8 + mr;
    uint32_t b[4];
    ldsm_x4(Ks + swz(key, mi), b[0], b[1], b[2], b[3]);          // dims 0..31
    mma_bf16(sc[nt], qa0[0], 0u, qa2[0], 0u, b[0], b[1]);
    mma_bf16(sc[nt], qa0[1], 0u, qa2[1], 0u, b[2], b[3]);
    ldsm_x4(Ks + swz(key, 4 + mi), b[0], b[1], b[2], b[3]);      // dims 32..63
    mma_bf16(sc[nt], qa0[2], 0u, qa2[2], 0u, b[0], b[1]);
    mma_bf16(sc[nt], qa0[3], 0u, qa2[3], 0u, b[2], b[3]);
  }
  // ---- online softmax (row g: sc[nt][0..1] = keys k0 + nt*8 + 2t, +1) ----
  float mt = -INFINITY;
  if (n < k0 + WKEYS || valid) {   // warp-uniform: a short chunk or a key-padding mask
#pragma unroll
    for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = k0 + nt * 8 + 2 * t + j;
        const bool vis = key < n && (!valid || valid[key]);
        sc[nt][j] = vis ? sc[nt][j] * 0.125f : -INFINITY;
        mt = fmaxf(mt, sc[nt][j]);
      }
  } else {
#pragma unroll
    for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[nt][j] *= 0.125f;
        mt = fmaxf(mt, sc[nt][j]);
      }
  }
  mt = fmaxf(mt, __shfl_xor_sync(kFull, mt, 1));
  mt = fmaxf(mt, __shfl_xor_sync(kFull, mt, 2));
  const float m_new = fmaxf(acc.m, mt);
  // __expf = one FMUL + MUFU.EX2 (the probabilities are rounded to bf16 right below); __expf(-inf) = 0
  const float scale = (acc.m == -INFINITY) ? 0.f : __expf(acc.m - m_new);   // m_new == -inf only if acc.m == -inf too
  float ls = 0.f;
  uint32_t pa[WNT];
#pragma unroll
  for (int nt = 0; nt < WNT; ++nt) {
    const float p0 = (sc[nt][0] == -INFINITY) ? 0.f : __expf(sc[nt][0] - m_new);
    const float p1 = (sc[nt][1] == -INFINITY) ? 0.f : __expf(sc[nt][1] - m_new);
    ls += p0 + p1;
    pa[nt] = (g < NQ) ? pack_bf16(p0, p1) : 0u;
  }
  ls += __shfl_xor_sync(kFull, ls, 1);
  ls += __shfl_xor_sync(kFull, ls, 2);
  acc.l = acc.l * scale + ls;
  acc.m = m_new;
  // ---- out = out * scale + p.V : 3 k-steps of 16 keys x 8 n-tiles of 8 dims ----
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    o[nt][0] = acc.o[nt][0] * scale;
    o[nt][1] = acc.o[nt][1] * scale;
    o[nt][2] = o[nt][3] = 0.f;
  }
#pragma unroll
  for (int kk = 0; kk < WKK; ++kk) {
    if (k0 + kk * 16 >= n) break;   // warp-uniform
    const int key = k0 + kk * 16 + (mi & 1) * 8 + mr;
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(Vs + swz(key, 2 * dp + (mi >> 1)), b[0], b[1], b[2], b[3]);
      mma_bf16(o[2 * dp], pa[2 * kk], 0u, pa[2 * kk + 1], 0u, b[0], b[1]);
      mma_bf16(o[2 * dp + 1], pa[2 * kk], 0u, pa[2 * kk + 1], 0u, b[2], b[3]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    acc.o[nt][0] = o[nt][0];
    acc.o[nt][1] = o[nt][1];
  }
}

// q rows staged in shared memory (row i at qs + i*128 bytes) -> A fragments of the 16-row tile (rows >= NQ zero)
template <int NQ>
__device__ __forceinline__ void load_q_frags(const unsigned char* qs, uint32_t (&qa0)[4], uint32_t (&qa2)[4]) {
  const int lane = threadIdx.x % kWarp, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    qa0[ks] = 0u;
    qa2[ks] = 0u;
  }
  if (g < NQ) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa0[ks] = *reinterpret_cast<const uint32_t*>(qs + g * 128 + (ks * 16 + 2 * t) * 2);
      qa2[ks] = *reinterpret_cast<const uint32_t*>(qs + g * 128 + (ks * 16 + 8 + 2 * t) * 2);
    }
  }
}

// Consumers (128 threads): finish a (group, head): merge the 4 warps' states and store the context rows, directly
// when this CTA saw every chunk of the group (n_parts == 1), else through the global partials + ticket.
//   rows[i]: global row index of query row i; slot(i) = ws + ((rows[i]*NH + h)*maxp + part)*(HD+2)
template <int NQ>
__device__ __forceinline__ void flush_group(const WarpAcc& acc, float* wpart, const int (&rows)[NQ], int h, int part,
                                            int n_parts, int maxp, float* __restrict__ ws, unsigned* ticket,
                                            bf16* __restrict__ ctx, int* sh_last) {
  const int tid = threadIdx.x, lane = tid % kWarp, w = tid / kWarp, g = lane >> 2, t = lane & 3;
  if (g < NQ) {
    float* mine = wpart + (w * NQ + g) * (HD + 2);
    if (t == 0) {
      mine[0] = acc.m;
      mine[1] = acc.l;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mine[2 + nt * 8 + 2 * t] = acc.o[nt][0];
      mine[2 + nt * 8 + 2 * t + 1] = acc.o[nt][1];
    }
  }
  cons_sync();
  if (tid < NQ * HD) {
    const int i = tid / HD, d = tid % HD;
    float m = -INFINITY;
#pragma unroll
    for (int x = 0; x < 4; ++x) m = fmaxf(m, wpart[(x * NQ + i) * (HD + 2)]);
    float l = 0.f, o = 0.f;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float mw = wpart[(x * NQ + i) * (HD + 2)];
      if (mw == -INFINITY) continue;
      const float a = expf(mw - m);
      l += a * wpart[(x * NQ + i) * (HD + 2) + 1];
      o += a * wpart[(x * NQ + i) * (HD + 2) + 2 + d];
    }
    if (n_parts == 1) {
      ctx[static_cast<long long>(rows[i]) * H + h * HD + d] = from_f<bf16>(l > 0.f ? o / l : 0.f);
    } else {
      float* dst = ws + ((static_cast<long long>(rows[i]) * NH + h) * maxp + part) * (HD + 2);
      dst[2 + d] = o;
      if (d == 0) {
        dst[0] = m;
        dst[1] = l;
      }
    }
  }
  if (n_parts > 1) {
    // release / acquire through ONE thread: the partial stores of all 128 consumers are ordered before thread 0's
    // fence by the barrier, the fence makes them visible before the ticket; the last arriver fences again before the
    // barrier that lets the others read the partials (through L2: __ldcg in merge_partials).  128 threads fencing
    // (MEMBAR + L1 invalidate each) showed up as the membar stalls of this kernel in ncu.
    cons_sync();
    if (tid == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(ticket, 1u);
      const int last = (prev == static_cast<unsigned>(n_parts) - 1) ? 1 : 0;
      if (last) {
        *ticket = 0;   // ready for the next launch
        __threadfence();
      }
      *sh_last = last;
    }
    cons_sync();
    if (*sh_last) {
      if (tid < NQ * HD) {
        const int i = tid / HD, d = tid % HD;
        merge_partials<bf16>(ws + (static_cast<long long>(rows[i]) * NH + h) * maxp * (HD + 2), n_parts,
                             ctx + static_cast<long long>(rows[i]) * H + h * HD, d);
      }
    }
  }
  cons_sync();   // wpart / sh_last reusable
}

// ---- cross-attention: items = (head, unit) with the unit fastest, so a CTA's range walks the chunks of a study in order.
// The encoder K/V cache and the unit table are constants of the whole rollout, so the first PSTAGES tiles of the
// CTA are requested BEFORE the programmatic-dependency wait: they stream in while the preceding Q-projection GEMM
// (a latency-bound kernel on a few dozen SMs) is still running.  Only q and the finished flags wait.
template <int NQ>
__global__ void __launch_bounds__(PNT, 2) decode_cross_persist_kernel(const __grid_constant__ CUtensorMap tm_kv,
                                                                      int row_base_k, int row_base_v, int tok_cap,
                                                                      const bf16* __restrict__ q, int ldq,
                                                                      bf16* __restrict__ ctx, CrossUnits cu,
                                                                      RolloutState st, int B, float* __restrict__ ws,
                                                                      unsigned* __restrict__ tickets, int dbg_nocompute) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* wpart = reinterpret_cast<float*>(tiles + PSTAGES * STAGE_BYTES);
  int* m_study = reinterpret_cast<int*>(wpart + 4 * 2 * (HD + 2));
  int* m_j0 = m_study + MAXI;
  int* m_n = m_j0 + MAXI;
  int* m_first = m_n + MAXI;
  int* m_nch = m_first + MAXI;
  uint64_t* full = reinterpret_cast<uint64_t*>(m_nch + MAXI);
  uint64_t* empty = full + PSTAGES;
  uint8_t* m_fin = reinterpret_cast<uint8_t*>(empty + PSTAGES);   // [MAXI] every row of the item's study finished
  __shared__ int sh_last;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid / kWarp;
  const int n_units = *cu.n_units;
  const int n_items = n_units * NH;
  const int per = ceil_div(n_items, static_cast<int>(gridDim.x));   // <= MAXI by the launcher's grid size
  const int lo = blockIdx.x * per, hi = min(n_items, lo + per);
  if (lo >= hi) {
    pdl_wait();
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int x = tid; x < hi - lo; x += PNT) {
    const int u = (lo + x) % n_units;
    const int b = cu.study[u];
    m_study[x] = b;
    m_j0[x] = cu.j0[u];
    m_n[x] = cu.n[u];
    m_first[x] = cu.first_unit[b];
    m_nch[x] = cu.n_chunks[b];
  }
  __syncthreads();
  const int npre = min(PSTAGES, hi - lo);   // items whose K/V are requested ahead of the dependency wait
  if (tid == NCONS) {
    for (int x = 0; x < npre; ++x) {
      const int h = (lo + x) / n_units;
      unsigned char* dst = tiles + x * STAGE_BYTES;
      mbar_expect_tx(&full[x], 2 * TILE_BYTES + NQ * 128);
      tma_load_2d(dst, &tm_kv, 0, row_base_k + h * tok_cap + m_j0[x], &full[x]);
      tma_load_2d(dst + TILE_BYTES, &tm_kv, 0, row_base_v + h * tok_cap + m_j0[x], &full[x]);
    }
  }
  // the finished flags are written by the PREVIOUS step's last kernel and a step opens with a fully serialised launch:
  // they are stable for the whole step and readable ahead of the dependency wait (as in the self-attention kernel)
  const bool done = *st.done != 0;
  for (int x = tid; x < hi - lo; x += PNT) {
    bool fin = true;
#pragma unroll
    for (int i = 0; i < NQ; ++i) fin = fin && st.finished[m_study[x] + i * B];
    m_fin[x] = (fin || done) ? 1 : 0;
  }
  __syncthreads();
  pdl_wait();

  if (warp == 4) {
    // ===== producer =====
    if (tid % kWarp == 0) {
      // q rows of the prefetched items (also when the study is finished: the stage's barrier expects them)
      for (int x = 0; x < npre; ++x) {
        const int h = (lo + x) / n_units, b = m_study[x];
#pragma unroll
        for (int i = 0; i < NQ; ++i)
          bulk_g2s(tiles + x * STAGE_BYTES + ST_Q + i * 128, q + static_cast<long long>(b + i * B) * ldq + h * HD, 128, &full[x]);
      }
      int k = npre;
      for (int it = lo + npre; it < hi; ++it) {
        if (m_fin[it - lo]) continue;
        const int b = m_study[it - lo];
        const int h = it / n_units;
        const int s = k % PSTAGES;
        mbar_wait(&empty[s], ((k / PSTAGES) & 1) ^ 1);
        unsigned char* dst = tiles + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], 2 * TILE_BYTES + NQ * 128);
        tma_load_2d(dst, &tm_kv, 0, row_base_k + h * tok_cap + m_j0[it - lo], &full[s]);
        tma_load_2d(dst + TILE_BYTES, &tm_kv, 0, row_base_v + h * tok_cap + m_j0[it - lo], &full[s]);
#pragma unroll
        for (int i = 0; i < NQ; ++i)
          bulk_g2s(dst + ST_Q + i * 128, q + static_cast<long long>(b + i * B) * ldq + h * HD, 128, &full[s]);
        ++k;
      }
    }
    return;
  }

  // ===== consumers =====
  WarpAcc acc;
  acc.reset();
  uint32_t qa0[4], qa2[4];
  int cur_b = -1, cur_h = -1, cur_x = 0, k = 0;
  auto finish = [&]() {
    if (cur_b < 0) return;
    const int first_item = cur_h * n_units + m_first[cur_x];
    const int last_item = first_item + m_nch[cur_x] - 1;
    const int p0 = first_item / per, n_parts = last_item / per - p0 + 1;
    int rows[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) rows[i] = cur_b + i * B;
    flush_group<NQ>(acc, wpart, rows, cur_h, static_cast<int>(blockIdx.x) - p0, n_parts, cu.max_chunks, ws,
                    &tickets[cur_b * NH + cur_h], ctx, &sh_last);
    acc.reset();
  };
  for (int it = lo; it < hi; ++it) {
    const int x = it - lo;
    const bool fin = m_fin[x] != 0;
    if (fin && x >= npre) continue;          // never requested
    const int s = k % PSTAGES;
    mbar_wait(&full[s], (k / PSTAGES) & 1);
    if (!fin) {
      const int b = m_study[x], h = it / n_units;
      if (b != cur_b || h != cur_h) {
        finish();
        cur_b = b;
        cur_h = h;
        cur_x = x;
      }
      const unsigned char* stage = tiles + s * STAGE_BYTES;
      load_q_frags<NQ>(stage + ST_Q, qa0, qa2);
      const uint32_t base = smem_u32(stage);
      if (!dbg_nocompute) warp_tile_update<NQ>(base, base + TILE_BYTES, m_n[x], qa0, qa2, nullptr, acc);
    }
    __syncwarp();
    if (tid % kWarp == 0) mbar_arrive(&empty[s]);   // a prefetched tile of a finished study is simply released
    ++k;
  }
  finish();
}

// ---- self-attention: items = (head, row, chunk); every live row has the same cache length P + step
__global__ void __launch_bounds__(PNT, 2) decode_self_persist_kernel(const __grid_constant__ CUtensorMap tm_k,
                                                                     const __grid_constant__ CUtensorMap tm_v,
                                                                     int layer_row0, const bf16* __restrict__ qkv,
                                                                     bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                                                                     bf16* __restrict__ ctx, RolloutState st, int R,
                                                                     int P, int Lmax, int max_chunks,
                                                                     float* __restrict__ ws,
                                                                     unsigned* __restrict__ tickets, int dbg_nocompute) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* wpart = reinterpret_cast<float*>(tiles + PSTAGES * STAGE_BYTES);
  int* m_n = reinterpret_cast<int*>(wpart + 4 * 2 * (HD + 2));   // keys in the chunk; 0: skip
  int* m_nch = m_n + MAXI;                                       // chunks of the row
  int* m_len = m_nch + MAXI;                                     // cache slot of the token being fed
  uint64_t* full = reinterpret_cast<uint64_t*>(m_n + 5 * MAXI);
  uint64_t* empty = full + PSTAGES;
  __shared__ int sh_last;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid / kWarp, lane = tid % kWarp;
  // The rollout state (step, done, cur_len, finished, key_valid) is written by the sampling kernel of the PREVIOUS
  // step, and a step opens with a plain (fully serialised) launch: it is readable ahead of the dependency wait, and
  // so are all cache rows except the slot of the token being fed.  Only q/k/v of the QKV projection wait.
  const bool done = *st.done != 0;
  // a live row's token sits at slot (visible prompt tokens) + step <= P + step: chunks of the longest possible row
  const int mc = ceil_div(P + *st.step + 1, CHB);
  const int n_items = NH * R * mc;
  // whole (row, head) groups per CTA (a multiple of mc items): no group straddles two CTAs, so the cross-CTA ticket
  // merge (global partials, fences, atomics) never runs in this kernel; a few CTAs at the end of the grid stay idle
  const int per = ceil_div(ceil_div(n_items, static_cast<int>(gridDim.x)), mc) * mc;   // <= MAXI by the launcher's grid size
  const int lo = blockIdx.x * per, hi = min(n_items, lo + per);
  if (done || lo >= hi) {
    pdl_wait();
    return;
  }

  if (tid == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // item -> (h, r, c) = (it / (mc*R), (it / mc) % R, it % mc); dead when the row is finished or the chunk is beyond its cache
  for (int x = tid; x < hi - lo; x += PNT) {
    const int it = lo + x, c = it % mc, r = (it / mc) % R;
    const int ntot = st.cur_len[r] + 1;
    const int nchunk = ceil_div(ntot, CHB);
    const bool live = !st.finished[r] && c < nchunk;
    m_n[x] = live ? min(CHB, ntot - c * CHB) : 0;
    m_nch[x] = nchunk;
    m_len[x] = ntot - 1;
  }
  __syncthreads();

  if (warp == 4) {
    // ===== producer (whole warp: the append of the new token's K/V is a 32-lane copy) =====
    // K/V and mask of one item -> stage s (the barrier also expects the 128 q bytes, sent after the dependency wait)
    auto request_kv = [&](int it, int s) {
      const int n = m_n[it - lo];
      const int c = it % mc, r = (it / mc) % R, h = it / (mc * R);
      unsigned char* dst = tiles + s * STAGE_BYTES;
      const int nbox = ceil_div(n, SELF_BOX);
      const int row0 = layer_row0 + (r * NH + h) * Lmax + c * CHB;
      const int mbytes = min(CHB, Lmax - c * CHB);   // Lmax % 16 == 0
      mbar_expect_tx(&full[s], 2 * nbox * SELF_BOX * 128 + 128 + mbytes);
      for (int x = 0; x < nbox; ++x) {
        tma_load_2d(dst + x * SELF_BOX * 128, &tm_k, 0, row0 + x * SELF_BOX, &full[s]);
        tma_load_2d(dst + TILE_BYTES + x * SELF_BOX * 128, &tm_v, 0, row0 + x * SELF_BOX, &full[s]);
      }
      bulk_g2s(dst + ST_MASK, st.key_valid + static_cast<long long>(r) * Lmax + c * CHB, mbytes, &full[s]);
    };
    // ahead of the dependency wait: the K/V + mask tiles of the leading live items stream in while the QKV projection
    // is still running - ALL of them: the row of the token being fed is stale in the tile that holds it, and the
    // consumer warp that owns that key overwrites it in shared memory from the projection's output (patch_new_token)
    int it0 = lo, npre = 0, pre_it[PSTAGES];
#pragma unroll
    for (int x = 0; x < PSTAGES; ++x) pre_it[x] = 0;
    if (lane == 0) {
      for (; it0 < hi && npre < PSTAGES; ++it0) {
        if (m_n[it0 - lo] == 0) continue;
        request_kv(it0, npre);
#pragma unroll
        for (int x = 0; x < PSTAGES; ++x)
          if (x == npre) pre_it[x] = it0;
        ++npre;
      }
    }
    it0 = __shfl_sync(kFull, it0, 0);
    npre = __shfl_sync(kFull, npre, 0);
    pdl_wait();
    if (lane == 0) {
#pragma unroll
      for (int x = 0; x < PSTAGES; ++x)
        if (x < npre) {
          const int r = (pre_it[x] / mc) % R, h = pre_it[x] / (mc * R);
          bulk_g2s(tiles + x * STAGE_BYTES + ST_Q, qkv + static_cast<long long>(r) * 3 * H + h * HD, 128, &full[x]);
        }
    }
    int k = npre;
    for (int it = lo; it < hi; ++it) {
      const int n = m_n[it - lo];
      if (n == 0) continue;
      const int c = it % mc, r = (it / mc) % R, h = it / (mc * R);
      const bf16* qrow = qkv + static_cast<long long>(r) * 3 * H + h * HD;
      if (c == m_nch[it - lo] - 1) {
        // the fed token's K/V go into the cache for the FOLLOWING steps (plain stores: this step reads them from the
        // projection's output, the next step's kernels start after this one has completed)
        const long long base = (static_cast<long long>(r) * NH + h) * Lmax * HD;
        const int L = m_len[it - lo];
        const uint32_t kx = *reinterpret_cast<const uint32_t*>(qrow + H + 2 * lane);
        const uint32_t vx = *reinterpret_cast<const uint32_t*>(qrow + 2 * H + 2 * lane);
        *reinterpret_cast<uint32_t*>(kcache + base + static_cast<long long>(L) * HD + 2 * lane) = kx;
        *reinterpret_cast<uint32_t*>(vcache + base + static_cast<long long>(L) * HD + 2 * lane) = vx;
      }
      if (it < it0) continue;   // requested ahead of the wait
      if (lane == 0) {
        const int s = k % PSTAGES;
        mbar_wait(&empty[s], ((k / PSTAGES) & 1) ^ 1);
        request_kv(it, s);
        bulk_g2s(tiles + s * STAGE_BYTES + ST_Q, qrow, 128, &full[s]);
      }
      ++k;
    }
    return;
  }
  pdl_wait();   // consumers: every thread of a chain kernel passes the dependency wait

  // ===== consumers =====
  WarpAcc acc;
  acc.reset();
  uint32_t qa0[4], qa2[4];
  int cur_r = -1, cur_h = -1, cur_nchunk = 0, k = 0;
  auto finish = [&]() {
    if (cur_r < 0) return;
    const int first_item = (cur_h * R + cur_r) * mc;
    const int last_item = first_item + cur_nchunk - 1;
    const int p0 = first_item / per, n_parts = last_item / per - p0 + 1;
    const int rows[1] = {cur_r};
    flush_group<1>(acc, wpart, rows, cur_h, static_cast<int>(blockIdx.x) - p0, n_parts, max_chunks, ws,
                   &tickets[cur_r * NH + cur_h], ctx, &sh_last);
    acc.reset();
  };
  for (int it = lo; it < hi; ++it) {
    const int n = m_n[it - lo];
    if (n == 0) continue;
    const int r = (it / mc) % R, h = it / (mc * R);
    if (r != cur_r || h != cur_h) {
      finish();
      cur_r = r;
      cur_h = h;
      cur_nchunk = m_nch[it - lo];
    }
    const int s = k % PSTAGES;
    // the chunk that ends with the token being fed: the warp that owns its key (the chunk's last one) takes K/V of that
    // token straight from the QKV projection - requested before the tile wait - and overwrites the stale row of the
    // tile (128-byte-swizzled: 16-byte chunk c of row r sits at c ^ (r & 7)).  No cache write -> fence -> TMA read-back
    // on the critical path, and the tile itself no longer depends on the projection.
    const bool patch = (it % mc == m_nch[it - lo] - 1) && warp == (n - 1) / WKEYS;
    uint32_t kx = 0, vx = 0;
    if (patch) {
      const bf16* qrow = qkv + static_cast<long long>(r) * 3 * H + h * HD;
      kx = *reinterpret_cast<const uint32_t*>(qrow + H + 2 * lane);
      vx = *reinterpret_cast<const uint32_t*>(qrow + 2 * H + 2 * lane);
    }
    mbar_wait(&full[s], (k / PSTAGES) & 1);
    unsigned char* stage = tiles + s * STAGE_BYTES;
    if (patch) {
      const int row = n - 1;
      const int off = row * 128 + ((((lane >> 2) ^ (row & 7)) << 4) | ((lane & 3) << 2));
      *reinterpret_cast<uint32_t*>(stage + off) = kx;
      *reinterpret_cast<uint32_t*>(stage + TILE_BYTES + off) = vx;
      __syncwarp();
    }
    load_q_frags<1>(stage + ST_Q, qa0, qa2);
    const uint32_t base = smem_u32(stage);
    if (!dbg_nocompute) warp_tile_update<1>(base, base + TILE_BYTES, n, qa0, qa2, stage + ST_MASK, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    ++k;
  }
  finish();
}

// qkv [R*P, 3*768] -> head-major caches [R][12][Lmax][64], columns [0,P)
template <typename T>
__global__ void prefill_store_kv_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                                        const int* __restrict__ slot, const uint8_t* __restrict__ valid, int R, int P,
                                        int Lmax, const int* __restrict__ row_of) {
  constexpr int VN = Vec16<T>::N;
  const int cv = H / VN;
  const long long total = static_cast<long long>(R) * P * cv * 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * VN;
    long long t = i / cv;
    const int which = static_cast<int>(t % 2);
    t /= 2;
    if (!valid[t]) continue;
    const long long r = row_of ? row_of[t] : t / P;       // (token index t = r * P + column in the padded layout)
    Vec16<T> v;
    v.load(qkv + t * 3 * H + (1 + which) * H + c);
    const int h = c / HD, d = c % HD;
    v.store((which ? vcache : kcache) + ((r * NH + h) * Lmax + slot[t]) * HD + d);
  }
}

// timing experiment only: CXRM_ATTN_NOCOMPUTE=1 keeps the TMA pipeline but skips the MMAs (memory-side ceiling)
int dbg_nocompute() {
  static int v = -1;
  if (v < 0) v = std::getenv("CXRM_ATTN_NOCOMPUTE") ? 1 : 0;
  return v;
}

// experiment knob (tools/exp_two_stream.py): persistent decode-attention CTAs per SM (default 2)
int attn_ctas_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = std::getenv("CXRM_ATTN_CPS");
    v = e ? std::max(1, std::atoi(e)) : 2;
  }
  return v;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    CXRM_CUDA_CHECK(cudaGetDevice(&dev));
    CXRM_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

}  // namespace

int decode_attn_chunk(size_t elem_size) { return elem_size == 2 ? CHB : 96; }

size_t decode_attn_ws_floats(int rows, int max_chunks) {
  return static_cast<size_t>(rows) * NH * max_chunks * (HD + 2);
}

template <typename T>
void decode_self_attention(const T* qkv, T* kcache, T* vcache, T* ctx, const RolloutState& st, int R, int P, int Lmax,
                           float* ws, unsigned* tickets, const AttnMaps* maps, int layer, cudaStream_t stream) {
  const int CH = decode_attn_chunk(sizeof(T));
  const int max_chunks = ceil_div(Lmax, CH);
  if constexpr (std::is_same<T, bf16>::value) {
    CXRM_CHECK(maps != nullptr, "bf16 decode attention needs the cache tensor maps");
    static bool configured = false;
    if (!configured) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(decode_self_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kPersistSmem)));
      configured = true;
    }
    const int layer_row0 = layer * maps->self_rows_per_layer;
    // the kernel rounds its per-CTA item count up to whole (row, head) groups: leave max_chunks of head room in MAXI
    CXRM_CHECK(max_chunks < MAXI / 2, "self-attention cache too long for the per-CTA work list");
    const int grid = std::max(attn_ctas_per_sm() * num_sms(), ceil_div(NH * R * max_chunks, MAXI - max_chunks));
    launch_chain(decode_self_persist_kernel, dim3(grid), dim3(PNT), kPersistSmem, stream, maps->self_k, maps->self_v,
                 layer_row0, qkv, kcache, vcache, ctx, st, R, P, Lmax, max_chunks, ws, tickets, dbg_nocompute());
  } else {
    const size_t smem = unit_smem_bytes<T, 1>(CH);
    static size_t configured = 0;
    if (smem > configured) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(decode_self_units_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
      configured = smem;
    }
    decode_self_units_kernel<T><<<dim3(R * max_chunks, NH), NT, smem, stream>>>(qkv, kcache, vcache, ctx, st, Lmax, CH,
                                                                               max_chunks, ws, tickets);
  }
  check_launch("decode_self_attention");
}

template <typename T>
void decode_cross_attention(const T* q, int ldq, const T* kc, const T* vc, long long head_stride, T* ctx,
                            const CrossUnits& cu, const RolloutState& st, int R, int B, float* ws, unsigned* tickets,
                            const AttnMaps* maps, int layer, cudaStream_t stream) {
  const int nq = R / B;
  CXRM_CHECK(nq == 1 || nq == 2, "decode_cross_attention: 1 or 2 rows per study");
  const int CH = decode_attn_chunk(sizeof(T));
  if constexpr (std::is_same<T, bf16>::value) {
    CXRM_CHECK(maps != nullptr, "bf16 decode attention needs the cache tensor maps");
    const int tok_cap = static_cast<int>(head_stride / HD);
    const int row_k = layer * 2 * NH * tok_cap, row_v = row_k + NH * tok_cap;
    // (one flag for both instantiations: they share a function-pointer type, so a generic lambda would share it too)
    static bool configured = false;
    if (!configured) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(decode_cross_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kPersistSmem)));
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(decode_cross_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kPersistSmem)));
      configured = true;
    }
    auto launch = [&](auto kern) {
      const int grid = std::max(attn_ctas_per_sm() * num_sms(), ceil_div(NH * cu.max_units, MAXI));
      launch_chain(kern, dim3(grid), dim3(PNT), kPersistSmem, stream, maps->cross, row_k, row_v, tok_cap, q, ldq, ctx, cu,
                   st, B, ws, tickets, dbg_nocompute());
    };
    if (nq == 1)
      launch(decode_cross_persist_kernel<1>);
    else
      launch(decode_cross_persist_kernel<2>);
  } else {
    auto launch = [&](auto kern, size_t smem) {
      CXRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kern<<<dim3(cu.max_units, NH), NT, smem, stream>>>(q, ldq, kc, vc, head_stride, ctx, cu, st, B, CH, ws, tickets);
    };
    if (nq == 1)
      launch(decode_cross_units_kernel<T, 1>, unit_smem_bytes<T, 1>(CH));
    else
      launch(decode_cross_units_kernel<T, 2>, unit_smem_bytes<T, 2>(CH));
  }
  check_launch("decode_cross_attention");
}

template <typename T>
void prefill_store_kv(const T* qkv, T* kcache, T* vcache, const int* slot, const uint8_t* valid, int R, int P, int Lmax,
                      cudaStream_t stream, const int* row_of) {
  const long long total = static_cast<long long>(R) * P * (H / Vec16<T>::N) * 2;
  if (total <= 0) return;
  long long grid = ceil_div_ll(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  prefill_store_kv_kernel<T><<<static_cast<unsigned>(grid), 256, 0, stream>>>(qkv, kcache, vcache, slot, valid, R, P, Lmax, row_of);
  check_launch("prefill_store_kv");
}

#define INST(T)                                                                                                      \
  template void decode_self_attention<T>(const T*, T*, T*, T*, const RolloutState&, int, int, int, float*, unsigned*, \
                                         const AttnMaps*, int, cudaStream_t);                                         \
  template void decode_cross_attention<T>(const T*, int, const T*, const T*, long long, T*, const CrossUnits&,        \
                                          const RolloutState&, int, int, float*, unsigned*, const AttnMaps*, int,     \
                                          cudaStream_t);                                                              \
  template void prefill_store_kv<T>(const T*, T*, T*, const int*, const uint8_t*, int, int, int, cudaStream_t, const int*);
INST(float)
INST(bf16)
#undef INST

}  // namespace cxrm
