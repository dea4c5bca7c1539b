// Causal / segment / window masked multi-head attention for sm_100a, forward and backward, on
// tcgen05 tensor cores with TMEM accumulators and TMA-fed, swizzled shared-memory tiles.
//
// Replaces model_tiny_gpt.py:103-131 (both the SDPA and the manual branch) together with the mask
// of :273-295, which is never materialised: the kernels take per-token segment STARTS (int32) and a
// window, turn the mask into a per-row interval jlo(i) <= j <= i, and skip whole KV tiles that the
// causal / window / segment structure rules out.  GQA reads kv head h/(H/Hk) directly (:94-96).
//
// Layout: packed qkv bf16 [B*T, W], W=(H+2Hk)*hd, column blocks q | k | v.  One 3-D TMA map
// {W, T, B} with box {AW, 128, 1} serves Q, K, V, dO ... (AW = swizzle-atom width in elements).
// Every tile lives in smem as rows of 2*AW bytes with the TMA swizzle; the same bytes are read as a
// K-major operand (reduction over hd) or an MN-major operand (reduction over rows) by choosing the
// UMMA descriptor, so no transposes are ever made.
//
// TMEM lane == query row, so softmax statistics are per-thread scalars (no shuffles); 256 threads
// split each row's columns in two halves.
#include <string.h>

#include "common.cuh"

namespace cgpt {
namespace {

constexpr int BQ = 128;   // query rows per tile
constexpr int BKV = 128;  // kv rows per tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int HD>
struct HeadCfg {
  static_assert(HD % 16 == 0 && HD <= 128, "head_dim must be a multiple of 16, <= 128");
  static constexpr int AW = (HD % 64 == 0) ? 64 : ((HD % 32 == 0) ? 32 : 16);  // atom width (elements)
  static constexpr int NA = HD / AW;                                            // atoms per row
  static constexpr int ROWB = AW * 2;                                           // bytes per smem row
  static constexpr int SBO = 8 * ROWB;                                          // 8-row group pitch
  static constexpr uint32_t LAYOUT = AW == 64 ? kLayoutSW128 : (AW == 32 ? kLayoutSW64 : kLayoutSW32);
  static constexpr int SWZ = ROWB;                                              // TMA swizzle bytes
  static constexpr int ATOM_BYTES = 128 * ROWB;                                 // one atom of a 128-row tile
  static constexpr int TILE_BYTES = NA * ATOM_BYTES;                            // 128 x HD bf16
  // K-major view of a [128 x HD] tile, k-step ks (16 elements of hd)
  __device__ static uint64_t kmajor(uint32_t base, int ks) {
    const int e = ks * 16;
    return umma_smem_desc(base + (e / AW) * ATOM_BYTES + (e % AW) * 2, 16, SBO, LAYOUT);
  }
  // MN-major view (hd is the M/N dim, rows are the reduction), k-step ks = 16 rows
  __device__ static uint64_t mnmajor(uint32_t base, int ks) {
    return umma_smem_desc(base + ks * 16 * ROWB, ATOM_BYTES, SBO, LAYOUT);
  }
};

// P / dS tile: [128 rows x 128 cols] bf16, two SW128 atoms of 64 columns (16 KB each).
constexpr int kPTileBytes = 2 * 128 * 128;
__device__ __forceinline__ uint64_t ptile_kmajor(uint32_t base, int ks) {  // reduction over the 128 columns
  return umma_smem_desc(base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, kLayoutSW128);
}
__device__ __forceinline__ uint64_t ptile_mnmajor(uint32_t base, int ks) {  // reduction over the 128 rows
  return umma_smem_desc(base + ks * 2048, 16384, 1024, kLayoutSW128);
}
// thread `row` stores 8 consecutive bf16 (one 16-byte chunk `chunk` in 0..15) of its row
__device__ __forceinline__ void ptile_store(uint8_t* base, int row, int chunk, uint4 v) {
  const int atom = chunk >> 3, c = chunk & 7;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(base) + atom * 16384 + row * 128 +
                                                                  ((c ^ (row & 7)) << 4)),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <int HD>
__device__ __forceinline__ void tma_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int col, int row, int b) {
  using C = HeadCfg<HD>;
#pragma unroll
  for (int a = 0; a < C::NA; ++a) tma_load_3d(dst + a * C::ATOM_BYTES, tm, bar, col + a * C::AW, row, b);
}

__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx: 1 MUFU, flushes denormal results to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// column j (absolute) is visible from row i iff jlo <= j <= i  <=>  (unsigned)(j - jlo) <= (unsigned)(i - jlo)
__device__ __forceinline__ bool visible(int j, int jlo, unsigned span) { return static_cast<unsigned>(j - jlo) <= span; }

// first position p in [0,T) with a[p] > key (a non-decreasing); T if none
__device__ __forceinline__ int upper_bound_i32(const int32_t* a, int T, int key) {
  int lo = 0, hi = T;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The mask of model_tiny_gpt.py:273-295 as a per-row interval: because segment ids are non-decreasing,
//   j<=i && i-j<window && seg[i]==seg[j]   <=>   jlo(i) <= j <= i,
//   jlo(i) = max(seg_start[i], i-window+1), seg_start[i] = last position <= i holding <SEP> (or 0).
__device__ __forceinline__ int row_jlo(const int32_t* seg_start_b, int i, int T, int window) {
  if (i >= T) return 0x3fffffff;  // rows past the end see nothing
  int lo = seg_start_b ? seg_start_b[i] : 0;
  if (window > 0) lo = max(lo, i - window + 1);
  return lo;
}

// ===================================================================================== forward
template <int HD>
struct FwdSmem {
  using C = HeadCfg<HD>;
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + C::TILE_BYTES;         // 2 buffers
  static constexpr int kV = kK + 2 * C::TILE_BYTES;     // 2 buffers
  static constexpr int kP = kV + 2 * C::TILE_BYTES;
  static constexpr int kBar = kP + kPTileBytes;
  static constexpr int kTotal = kBar + 64;
  static constexpr int kMaxPerCta2 = 115712;            // (228 KB - 2 x 1 KB reserved) / 2
  static constexpr bool kTwoCtas = kTotal <= kMaxPerCta2;
  // slack for aligning the base up to 1024 B; trimmed when it would cost the second resident CTA
  static constexpr int kDynamic = kTwoCtas ? (kTotal + 1024 <= kMaxPerCta2 ? kTotal + 1024 : kMaxPerCta2) : kTotal + 1024;
};

// 256 threads: thread t owns query row (t & 127) and the column half (t >> 7) of every tile, so the
// softmax work per thread is 64 scores per KV tile.  Thread 0 also drives TMA and the MMAs.
template <int HD>
__global__ void __launch_bounds__(256, FwdSmem<HD>::kTwoCtas ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm, const int32_t* __restrict__ seg_start,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, int Hk, int window,
                float scale_log2, int smem_bytes, const DropoutCfg drop_in) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = FwdSmem<HD>;
  constexpr int TMEM_COLS = 256;
  constexpr int HH = HD / 2;  // output columns per thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) + S::kTotal > smem_bytes) __trap();
  uint8_t* sQ = smem + S::kQ;
  uint8_t* sP = smem + S::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* q_bar = bars;
  uint64_t* kv_bar = bars + 1;  // [2]
  uint64_t* mma_bar = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heaviest (most KV tiles) first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (H / Hk);
  const int q0 = qb * BQ;
  const int qcol = h * HD, kcol = (H + kvh) * HD, vcol = (H + Hk + kvh) * HD;
  const int32_t* ssb = seg_start ? seg_start + (size_t)b * T : nullptr;

  if (tid == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_bar, 1);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;          // 128 columns: scores
  const uint32_t tO = tmem_base + 128;    // HD columns: P·V of the current tile
  const int kv_lo = row_jlo(ssb, q0, T, window) / BKV;  // jlo is non-decreasing in i
  const int kv_hi = qb;
  const int nblk = kv_hi - kv_lo + 1;

  if (tid == 0) {
    mbar_expect_tx(q_bar, C::TILE_BYTES);
    tma_tile<HD>(sQ, &tm, q_bar, qcol, q0, b);
    mbar_expect_tx(&kv_bar[0], 2 * C::TILE_BYTES);
    tma_tile<HD>(smem + S::kK, &tm, &kv_bar[0], kcol, kv_lo * BKV, b);
    tma_tile<HD>(smem + S::kV, &tm, &kv_bar[0], vcol, kv_lo * BKV, b);
  }

  const int i = q0 + row;  // my query row
  const int jlo = row_jlo(ssb, i, T, window);
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  // exchange slot: written by (row, half), read by (row, 1-half); lives in the part of the P tile that
  // only the READER overwrites later (atom 1-half, row `row`), see the sync structure below.
  float* xchg_wr = reinterpret_cast<float*>(sP + (1 - half) * 16384 + row * 128);
  float* xchg_rd = reinterpret_cast<float*>(sP + half * 16384 + row * 128);
  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[HH];
#pragma unroll
  for (int c = 0; c < HH; ++c) o_acc[c] = 0.f;
  uint32_t mma_phase = 0;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);

  const bool leader = (warp == 0) && elect_one();
  const uint32_t sQ_u = smem_u32(sQ), sP_u = smem_u32(sP), sKV_u = smem_u32(smem + S::kK);

  for (int it = 0; it < nblk; ++it) {
    const int buf = it & 1;
    const int kv0 = (kv_lo + it) * BKV;
    uint8_t* sK = smem + S::kK + buf * C::TILE_BYTES;
    uint8_t* sV = smem + S::kV + buf * C::TILE_BYTES;
    if (warp == 0) {  // warp-uniform control flow (descriptors stay in uniform registers); one lane issues
      if (it + 1 < nblk && leader) {  // prefetch next KV tile (its buffer was released by the last PV wait)
        const int nb = buf ^ 1;
        mbar_expect_tx(&kv_bar[nb], 2 * C::TILE_BYTES);
        tma_tile<HD>(smem + S::kK + nb * C::TILE_BYTES, &tm, &kv_bar[nb], kcol, kv0 + BKV, b);
        tma_tile<HD>(smem + S::kV + nb * C::TILE_BYTES, &tm, &kv_bar[nb], vcol, kv0 + BKV, b);
      }
      if (it == 0) mbar_wait(q_bar, 0);
      mbar_wait(&kv_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tS, C::kmajor(sQ_u, ks), C::kmajor(sKV_u + buf * C::TILE_BYTES, ks), idesc_s, ks > 0);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();

    // per-score tests only for rows that do not see the whole kv tile (diagonal, segment start, window, ragged end)
    const bool need_mask = (i >= T) || (kv0 + BKV - 1 > i) || (kv0 < jlo);
    // pass 1: row maximum over my 64 columns (raw scores: the scale is positive).  Rows past the end
    // (jlo > i) have span = 0xffffffff... guarded by row_valid below.
    const unsigned span = static_cast<unsigned>(i - jlo);
    const bool row_valid = (i < T);
    float mraw = -INFINITY;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c4 = half * 2 + cc;
      uint32_t r[32];
      tmem_ld32(tS + lane_base + c4 * 32, r);
      tmem_ld_wait();
      if (need_mask) {
        const int jb = kv0 + c4 * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          mraw = fmaxf(mraw, (row_valid && visible(jb + j, jlo, span)) ? __uint_as_float(r[j]) : -INFINITY);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(r[j]));
      }
    }
    const float mloc = mraw * scale_log2;
    *xchg_wr = mloc;
    __syncthreads();
    const float m_new = fmaxf(m_run, fmaxf(mloc, *xchg_rd));
    const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = fast_exp2(m_run - m_safe);
    // pass 2: probabilities -> bf16 P tile in smem
    float lsum = 0.f;
    const float neg_m = -m_safe;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c4 = half * 2 + cc;
      uint32_t r[32];
      tmem_ld32(tS + lane_base + c4 * 32, r);
      tmem_ld_wait();
      float p[32];
      if (need_mask) {
        const int jb = kv0 + c4 * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pv = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
          p[j] = (row_valid && visible(jb + j, jlo, span)) ? pv : 0.f;
          lsum += p[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
          lsum += p[j];
        }
      }
      if (drop.thresh) {  // dropout on the (still unnormalised) probabilities; the row sum stays undropped (:104,129)
        const uint32_t keep = attn_keep_mask32(drop, b * H + h, i, kv0 + c4 * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) p[j] = ((keep >> j) & 1u) ? p[j] * drop.inv_keep16 : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
        v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
        v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
        v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
        ptile_store(sP, row, c4 * 4 + q, v);
      }
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_bf16(tO, ptile_kmajor(sP_u, ks), C::mnmajor(sKV_u + (S::kV - S::kK) + buf * C::TILE_BYTES, ks), idesc_o, ks > 0);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < HH; c0 += 8) {
      uint32_t r[8];
      tmem_ld8(tO + lane_base + half * HH + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) o_acc[c0 + j] = o_acc[c0 + j] * alpha + __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();  // nobody may lag a full mbarrier phase behind thread 0
  }

  // combine the two halves' row sums (the P tile is free: the last PV MMA has completed)
  *xchg_wr = l_run;
  __syncthreads();
  const float l_tot = l_run + *xchg_rd;
  if (i < T) {
    const float inv = 1.f / l_tot;
    __nv_bfloat16* o = out + ((size_t)b * T + i) * (size_t)(H * HD) + h * HD + half * HH;
#pragma unroll
    for (int c = 0; c < HH; c += 8) {
      uint4 v;
      v.x = pack_bf16(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
      v.y = pack_bf16(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
      v.z = pack_bf16(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
      v.w = pack_bf16(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
      *reinterpret_cast<uint4*>(o + c) = v;
    }
    if (half == 0) lse[((size_t)b * H + h) * T + i] = (m_run + log2f(l_tot)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ===================================================================================== backward
// delta[b,h,i] = sum_c dO[i,c] * O[i,c]   (one warp per (token, head))
// ---------------------------------------------------------------- work order shared by the persistent kernels
// Items are (tile index k of nt, head, batch).  Under a causal mask tile k costs k+1 (forward: kv tiles per query
// tile) or nt-k (backward: query tiles per kv tile) units, so items are dealt in PAIRS (k, nt-1-k) of one head —
// equal work — round-robin and head-major: the ~148 pairs in flight belong to a few dozen heads, whose K/V, Q/dO and
// dQ rows therefore stay in L2.  The cursor walks one CTA's pairs without divisions in the loop.
struct PairCursor {
  int pi, bh, kk0, bhm, head, b, sub, k;  // k = tile index of the current item
  bool valid;
  int G, nt, npk, nheads, n_pairs, qG, rG, qGm, qH, rH;
  __device__ __forceinline__ int rot() const {  // pair slot rotated by the head index: a CTA's pairs cycle through all k
    const int kk = kk0 + bhm;
    return kk >= npk ? kk - npk : kk;
  }
  __device__ __forceinline__ void begin(int cta, int grid, int n_tiles, int heads, int batch, bool high_first) {
    G = grid; nt = n_tiles; npk = (n_tiles + 1) / 2; nheads = heads; n_pairs = npk * heads * batch;
    qG = G / npk; rG = G - qG * npk; qGm = qG % npk; qH = qG / heads; rH = qG - qH * heads;
    pi = cta;
    valid = pi < n_pairs;
    bh = pi / npk;
    kk0 = pi - bh * npk;
    bhm = bh % npk;
    b = bh / heads;
    head = bh - b * heads;
    sub = 0;
    flip = high_first;
    k = flip ? nt - 1 - rot() : rot();
  }
  bool flip;  // false: the low index of a pair first; true: the high index first
  __device__ __forceinline__ void next() {
    if (sub == 0) {  // second half of the pair: the mirrored tile (absent for the middle tile of an odd count)
      sub = 1;
      const int kk = rot();
      if (nt - 1 - kk != kk) {
        k = flip ? kk : nt - 1 - kk;
        return;
      }
    }
    sub = 0;
    pi += G;
    if (pi >= n_pairs) {
      valid = false;
      return;
    }
    kk0 += rG;
    const int carry = kk0 >= npk;
    if (carry) kk0 -= npk;
    bh += qG + carry;
    bhm += qGm + carry;
    while (bhm >= npk) bhm -= npk;
    head += rH + carry;
    b += qH;
    while (head >= nheads) {
      head -= nheads;
      ++b;
    }
    k = flip ? nt - 1 - rot() : rot();
  }
};

// ===================================================================================== forward, hd <= 64
// Persistent, warp-specialised, one continuous (query tile, kv tile) stream:
//   warp 9 : TMA producer — Q per item (double-buffered), (K, V) tiles through a kStages ring.
//   warp 8 : MMA issuer (warp-uniform control flow, elected lane) — S(t) = Q K^T into the S buffer t&1, then
//            O_t = P(t-1) V into the O buffer (t-1)&1: the scores of the next tile are computed while the softmax
//            warps still work on the current one.
//   warps 0..7 : softmax (thread = query row x column half): two passes over S in TMEM, P (bf16) into the P buffer
//            t&1, then fold O_{t-1} (TMEM) into the register accumulator with the rescale factor of that tile.
//            The output leaves through TMA stores staged in the warp's own rows of the P buffer.
// TMEM (512 cols): S0 128 | S1 128 | O0 hd | O1 hd.
template <int HD>
struct FwdWsSmem {
  using C = HeadCfg<HD>;
  static constexpr int kStages = 3;
  static constexpr int kQ = 0;                                   // 2 buffers
  static constexpr int kK = kQ + 2 * C::TILE_BYTES;              // kStages
  static constexpr int kV = kK + kStages * C::TILE_BYTES;        // kStages
  static constexpr int kP = kV + kStages * C::TILE_BYTES;        // 2 buffers
  static constexpr int kXchg = kP + 2 * kPTileBytes;             // 3 x 256 floats: row max (2 buffers) / row sum exchange
  static constexpr int kBar = kXchg + 3 * 256 * 4;
  static constexpr int kTotal = kBar + 256;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(HD <= 64 && kDynamic <= 232448, "warp-specialised forward: shared memory");
};

template <int HD>
__global__ void __launch_bounds__(320, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_out,
                   const int32_t* __restrict__ seg_start, float* __restrict__ lse, int Bsz, int T, int H, int Hk,
                   int window, float scale_log2, const DropoutCfg drop_in, int smem_bytes) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = FwdWsSmem<HD>;
  constexpr int NS = S::kStages;
  constexpr int HH = HD / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) + S::kTotal > smem_bytes) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2] every S MMA of the item has retired
  uint64_t* kv_full = bars + 4;       // [NS]
  uint64_t* kv_empty = bars + 4 + NS; // [NS] the P·V MMAs that read the stage have retired
  uint64_t* s_bar = bars + 4 + 2 * NS;   // [2]
  uint64_t* p_bar = s_bar + 2;           // [2] P of a tile is in smem (and its S buffer has been consumed)
  uint64_t* o_bar = p_bar + 2;           // [2] P·V of a tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_bar + 2);
  float* xchg = reinterpret_cast<float*>(smem + S::kXchg);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rep = H / Hk;
  const int nqb = (T + BQ - 1) / BQ;

  if (tid == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_bar[i], 1);
      mbar_init(&p_bar[i], 8);
      mbar_init(&o_bar[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tO0 = tmem_base + 256;  // S buffer b at tS0 + 128 b, O buffer b at tO0 + HD b

  // kv-tile range of an item: [row_jlo(first row) / BKV, qb]
  auto kv_lo_of = [&](int qb, int b) { return row_jlo(seg_start ? seg_start + (size_t)b * T : nullptr, qb * BQ, T, window) / BKV; };

  if (warp == 9) {
    // ================================================================= TMA producer
    const bool leader = elect_one();
    PairCursor cur;
    cur.begin(blockIdx.x, gridDim.x, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, h = cur.head, b = cur.b, kv_lo = nx_lo;
      const int kvh = h / rep;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);  // fetched one item ahead
      const int qbuf = n_it & 1;
      mbar_wait(&q_empty[qbuf], ((n_it >> 1) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(&q_full[qbuf], C::TILE_BYTES);
        tma_tile<HD>(smem + S::kQ + qbuf * C::TILE_BYTES, &tm, &q_full[qbuf], h * HD, qb * BQ, b);
      }
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const int st = t % NS;
        mbar_wait(&kv_empty[st], ((t / NS) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[st], 2 * C::TILE_BYTES);
          tma_tile<HD>(smem + S::kK + st * C::TILE_BYTES, &tm, &kv_full[st], (H + kvh) * HD, kvb * BKV, b);
          tma_tile<HD>(smem + S::kV + st * C::TILE_BYTES, &tm, &kv_full[st], (H + Hk + kvh) * HD, kvb * BKV, b);
        }
      }
    }
  } else if (warp == 8) {
    // ================================================================= MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);
    const uint32_t sQ_u = smem_u32(smem + S::kQ), sK_u = smem_u32(smem + S::kK), sV_u = smem_u32(smem + S::kV);
    const uint32_t sP_u = smem_u32(smem + S::kP);
    auto issue_pv = [&](uint32_t tt) {  // O buffer tt&1 = P(tt) V(tt); releases the (K, V) stage of the tile
      const uint32_t pb = tt & 1, st = tt % NS;
      mbar_wait(&p_bar[pb], (tt >> 1) & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_bf16(tO0 + pb * HD, ptile_kmajor(sP_u + pb * kPTileBytes, ks), C::mnmajor(sV_u + st * C::TILE_BYTES, ks),
                    idesc_o, ks > 0);
        umma_commit(&o_bar[pb]);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    };
    PairCursor cur;
    cur.begin(blockIdx.x, gridDim.x, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, kv_lo = nx_lo;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);
      const int qbuf = n_it & 1;
      mbar_wait(&q_full[qbuf], (n_it >> 1) & 1);
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const uint32_t st = t % NS, sb = t & 1;
        mbar_wait(&kv_full[st], (t / NS) & 1);
        // the S buffer t&1 was consumed by the softmax of tile t-2: awaited by issue_pv(t-2) one iteration ago
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_bf16(tS0 + sb * 128, C::kmajor(sQ_u + qbuf * C::TILE_BYTES, ks), C::kmajor(sK_u + st * C::TILE_BYTES, ks),
                      idesc_s, ks > 0);
          umma_commit(&s_bar[sb]);
          if (kvb == qb) umma_commit(&q_empty[qbuf]);  // last S MMA of the item: its Q buffer is free after this
        }
        __syncwarp();
        if (t >= 1) issue_pv(t - 1);
      }
    }
    if (t >= 1) {
      issue_pv(t - 1);
      mbar_wait(&o_bar[(t - 1) & 1], ((t - 1) >> 1) & 1);  // every MMA has retired before the CTA tears down
    }
  } else {
    // ================================================================= softmax warps
    const int row = tid & 127, half = tid >> 7;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    PairCursor cur;
    cur.begin(blockIdx.x, gridDim.x, nqb, H, Bsz, true);
    // own row's segment start and the tile's first row's, fetched one item ahead
    int nx_ss = 0, nx_ss0 = 0;
    auto prefetch_item = [&](const PairCursor& c) {
      const int i0 = c.k * BQ;
      nx_ss0 = seg_start ? seg_start[(size_t)c.b * T + i0] : 0;
      nx_ss = (seg_start && i0 + row < T) ? seg_start[(size_t)c.b * T + i0 + row] : 0;
    };
    if (cur.valid) prefetch_item(cur);
    uint32_t t = 0;
    while (cur.valid) {
      const int qb = cur.k, h = cur.head, b = cur.b;
      const int q0 = qb * BQ, i = q0 + row;
      const bool row_valid = i < T;
      int jlo = row_valid ? nx_ss : 0x3fffffff, jlo0 = nx_ss0;
      if (window > 0) {
        jlo = row_valid ? max(jlo, i - window + 1) : jlo;
        jlo0 = max(jlo0, q0 - window + 1);
      }
      const int kv_lo = jlo0 / BKV;
      cur.next();
      if (cur.valid) prefetch_item(cur);
      const unsigned span = static_cast<unsigned>(i - jlo);
      float m_run = -INFINITY, l_run = 0.f, alpha_pend = 1.f;
      float o_acc[HH];
#pragma unroll
      for (int c = 0; c < HH; ++c) o_acc[c] = 0.f;
      // O_{tt} (TMEM) folded into the accumulator with the rescale factor of tile tt
      auto fold = [&](uint32_t tt, float alpha) {
        const uint32_t ob = tt & 1;
        mbar_wait(&o_bar[ob], (tt >> 1) & 1);
        tc_fence_after();
        uint32_t r[HH];
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) tmem_ld8(tO0 + ob * HD + lane_base + half * HH + c0, *reinterpret_cast<uint32_t(*)[8]>(&r[c0]));
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HH; ++c) o_acc[c] = fmaf(o_acc[c], alpha, __uint_as_float(r[c]));
        tc_fence_before();
      };
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const uint32_t sb = t & 1;
        const int kv0 = kvb * BKV;
        uint8_t* sP = smem + S::kP + sb * kPTileBytes;
        float* xw = xchg + sb * 256 + (1 - half) * 128 + row;
        float* xr = xchg + sb * 256 + half * 128 + row;
        // bulk stores staged in this warp's rows of the P buffers (the previous item's output) have read their smem
        bulk_wait_read0();
        __syncwarp();
        mbar_wait(&s_bar[sb], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t tS = tS0 + sb * 128;
        // A row sees the column interval [jlo, i].  Per 32-column chunk the whole WARP (32 consecutive rows) votes:
        // chunk visible to every row -> no per-score tests; to no row -> skipped (no TMEM read, no exp); mixed -> tested.
        // On a diagonal tile only the 32x32 blocks on the diagonal are mixed.
        unsigned cls[2];  // 0 = all visible, 1 = none, 2 = mixed
        float mraw = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c4 = half * 2 + cc;
          const int jb = kv0 + c4 * 32;
          const bool full = row_valid && jb >= jlo && jb + 31 <= i;
          const bool none = !row_valid || jb > i || jb + 31 < jlo;
          cls[cc] = __all_sync(0xffffffffu, full) ? 0u : (__all_sync(0xffffffffu, none) ? 1u : 2u);
          if (cls[cc] == 1u) continue;
          uint32_t r[32];
          tmem_ld32(tS + lane_base + c4 * 32, r);
          tmem_ld_wait();
          if (cls[cc] == 2u) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              mraw = fmaxf(mraw, (row_valid && visible(jb + j, jlo, span)) ? __uint_as_float(r[j]) : -INFINITY);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(r[j]));
          }
        }
        const float mloc = mraw * scale_log2;
        *xw = mloc;
        named_bar_sync(1 + (warp & 3), 64);  // only the two warps that share these 32 rows meet
        const float m_new = fmaxf(m_run, fmaxf(mloc, *xr));
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2(m_run - m_safe);
        float lsum = 0.f;
        const float neg_m = -m_safe;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c4 = half * 2 + cc;
          if (cls[cc] == 1u) {  // nothing visible: the P chunk is zero
#pragma unroll
            for (int q = 0; q < 4; ++q) ptile_store(sP, row, c4 * 4 + q, make_uint4(0u, 0u, 0u, 0u));
            continue;
          }
          uint32_t r[32];
          tmem_ld32(tS + lane_base + c4 * 32, r);
          tmem_ld_wait();
          float p[32];
          if (cls[cc] == 2u) {
            const int jb = kv0 + c4 * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pv = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
              p[j] = (row_valid && visible(jb + j, jlo, span)) ? pv : 0.f;
              lsum += p[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              p[j] = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
              lsum += p[j];
            }
          }
          if (drop.thresh) {  // dropout on the (still unnormalised) probabilities; the row sum stays undropped (:104,129)
            const uint32_t keep = attn_keep_mask32(drop, b * H + h, i, kv0 + c4 * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = ((keep >> j) & 1u) ? p[j] * drop.inv_keep16 : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
            v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
            v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
            v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
            ptile_store(sP, row, c4 * 4 + q, v);
          }
        }
        l_run = l_run * alpha + lsum;
        m_run = m_new;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_bar[sb]);
        // while the tensor cores work on this tile's P·V: fold the previous tile of the item
        if (kvb > kv_lo) fold(t - 1, alpha_pend);
        alpha_pend = alpha;
      }
      fold(t - 1, alpha_pend);  // last tile of the item (also: its P buffer is free now)
      // combine the two halves' row sums
      {
        float* xw = xchg + 2 * 256 + (1 - half) * 128 + row;  // third buffer: never aliases a row-max exchange
        float* xr = xchg + 2 * 256 + half * 128 + row;
        *xw = l_run;
        named_bar_sync(1 + (warp & 3), 64);
        const float l_tot = l_run + *xr;
        const float inv = 1.f / l_tot;
        // output rows -> bf16, staged in this warp's rows of the last tile's P buffer, one TMA store per warp
        constexpr int ROWB = HH * 2;
        uint8_t* stg = smem + S::kP + ((t - 1) & 1) * kPTileBytes + half * 16384 + (warp & 3) * 4096;
        const int sw = ROWB == 64 ? ((lane >> 1) & 3) : (ROWB == 32 ? ((lane >> 2) & 1) : 0);
        const uint32_t rowp = smem_u32(stg + lane * ROWB);
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          const uint32_t off = static_cast<uint32_t>(((c0 >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + off),
                       "r"(pack_bf16(o_acc[c0 + 0] * inv, o_acc[c0 + 1] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 2] * inv, o_acc[c0 + 3] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 4] * inv, o_acc[c0 + 5] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 6] * inv, o_acc[c0 + 7] * inv))
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tm_out, smem_u32(stg), h * HD + half * HH, q0 + (warp & 3) * 32, b);
          bulk_commit();
        }
        if (half == 0 && row_valid) lse[((size_t)b * H + h) * T + i] = (m_run + log2f(l_tot)) * kLn2;
      }
    }
    bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ===================================================================================== forward, hd <= 64, two streams
// Two independent copies of the warp-specialised pipeline in ONE CTA (kernels that use tcgen05 get one CTA per SM):
// stream s in {0, 1} owns softmax warps 4s..4s+3 (thread = one full query row: no row-max exchange, no named
// barriers), MMA warp 8+s, TMA warp 10+s, 256 TMEM columns (S 128 | O0 | O1), its own Q / P buffers, 2-stage
// (K, V) ring and mbarriers, and walks the work items of virtual CTA 2*blockIdx.x+s.  Each scheduler therefore
// holds two softmax warps that do NOT move in lockstep: while one stream waits for its MMAs (P·V of tile t, then S
// of tile t+1) the other one computes.  12 warps x 168 registers, 2 x 113 KB of shared memory.
template <int HD>
struct FwdW2Smem {
  using C = HeadCfg<HD>;
  static constexpr int kStages = 2;
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + C::TILE_BYTES;
  static constexpr int kV = kK + kStages * C::TILE_BYTES;
  static constexpr int kP = kV + kStages * C::TILE_BYTES;
  static constexpr int kBar = kP + kPTileBytes;
  static constexpr int kStream = (kBar + 128 + 1023) / 1024 * 1024;  // bytes per stream
  static constexpr int kTotal = 2 * kStream;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(HD <= 64 && kDynamic <= 232448, "two-stream forward: shared memory");
};

template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_fwd_w2_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_out,
                   const int32_t* __restrict__ seg_start, float* __restrict__ lse, int Bsz, int T, int H, int Hk,
                   int window, float scale_log2, const DropoutCfg drop_in, int smem_bytes) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = FwdW2Smem<HD>;
  constexpr int NS = S::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem0 - smem_raw) + S::kTotal > smem_bytes) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strm = warp < 8 ? (warp >> 2) : (warp & 1);  // softmax warps 0-3 | 4-7, MMA warps 8 | 9, TMA warps 10 | 11
  uint8_t* smem = smem0 + strm * S::kStream;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* q_full = bars;             // Q of the item has landed
  uint64_t* q_empty = bars + 1;        // every S MMA of the item has retired
  uint64_t* kv_full = bars + 2;        // [NS]
  uint64_t* kv_empty = bars + 2 + NS;  // [NS] the P·V MMA that read the stage has retired
  uint64_t* s_bar = bars + 2 + 2 * NS; // S of a tile is in TMEM (and the previous P·V has retired: P is free)
  uint64_t* p_bar = s_bar + 1;         // P of a tile is in smem, its S has been consumed
  uint64_t* o_bar = p_bar + 1;         // [2] P·V of a tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem0 + S::kBar) + 32;  // shared by both streams (behind stream 0's barriers)

  const int rep = H / Hk;
  const int vcta = blockIdx.x * 2 + strm, vgrid = gridDim.x * 2;
  const int nqb = (T + BQ - 1) / BQ;

  if (tid == 0 || tid == 128) {  // one thread of each stream initialises that stream's barriers
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm_out);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, 4);
    mbar_init(&o_bar[0], 1);
    mbar_init(&o_bar[1], 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base + strm * 256, tO0 = tS + 128;  // O buffer b at tO0 + 64 b

  auto kv_lo_of = [&](int qb, int b) { return row_jlo(seg_start ? seg_start + (size_t)b * T : nullptr, qb * BQ, T, window) / BKV; };

  if (warp >= 10) {
    // ================================================================= TMA producer
    const bool leader = elect_one();
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, h = cur.head, b = cur.b, kv_lo = nx_lo;
      const int kvh = h / rep;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);
      mbar_wait(q_empty, (n_it & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(q_full, C::TILE_BYTES);
        tma_tile<HD>(smem + S::kQ, &tm, q_full, h * HD, qb * BQ, b);
      }
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const int st = t % NS;
        mbar_wait(&kv_empty[st], ((t / NS) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[st], 2 * C::TILE_BYTES);
          tma_tile<HD>(smem + S::kK + st * C::TILE_BYTES, &tm, &kv_full[st], (H + kvh) * HD, kvb * BKV, b);
          tma_tile<HD>(smem + S::kV + st * C::TILE_BYTES, &tm, &kv_full[st], (H + Hk + kvh) * HD, kvb * BKV, b);
        }
      }
    }
  } else if (warp >= 8) {
    // ================================================================= MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);
    const uint32_t sQ_u = smem_u32(smem + S::kQ), sK_u = smem_u32(smem + S::kK), sV_u = smem_u32(smem + S::kV);
    const uint32_t sP_u = smem_u32(smem + S::kP);
    auto issue_pv = [&](uint32_t tt) {  // O buffer tt&1 = P(tt) V(tt); releases the (K, V) stage of the tile
      const uint32_t st = tt % NS;
      mbar_wait(p_bar, tt & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_bf16(tO0 + (tt & 1) * 64, ptile_kmajor(sP_u, ks), C::mnmajor(sV_u + st * C::TILE_BYTES, ks), idesc_o, ks > 0);
        umma_commit(&o_bar[tt & 1]);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    };
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, kv_lo = nx_lo;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        // P·V of the previous tile first: the commit behind S(t) then also covers it, so the softmax warps may
        // overwrite the single P buffer as soon as they see S(t)
        if (t >= 1) issue_pv(t - 1);
        if (kvb == kv_lo) mbar_wait(q_full, n_it & 1);
        const uint32_t st = t % NS;
        mbar_wait(&kv_full[st], (t / NS) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_bf16(tS, C::kmajor(sQ_u, ks), C::kmajor(sK_u + st * C::TILE_BYTES, ks), idesc_s, ks > 0);
          umma_commit(s_bar);
          if (kvb == qb) umma_commit(q_empty);  // last S MMA of the item: the Q buffer is free after this
        }
        __syncwarp();
      }
    }
    if (t >= 1) {
      issue_pv(t - 1);
      mbar_wait(&o_bar[(t - 1) & 1], ((t - 1) >> 1) & 1);  // every MMA has retired before the CTA tears down
    }
  } else {
    // ================================================================= softmax warps: thread = query row
    const int row = tid & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_ss = 0, nx_ss0 = 0;
    auto prefetch_item = [&](const PairCursor& c) {
      const int i0 = c.k * BQ;
      nx_ss0 = seg_start ? seg_start[(size_t)c.b * T + i0] : 0;
      nx_ss = (seg_start && i0 + row < T) ? seg_start[(size_t)c.b * T + i0 + row] : 0;
    };
    if (cur.valid) prefetch_item(cur);
    uint32_t t = 0;
    uint8_t* sP = smem + S::kP;
    while (cur.valid) {
      const int qb = cur.k, h = cur.head, b = cur.b;
      const int q0 = qb * BQ, i = q0 + row;
      const bool row_valid = i < T;
      int jlo = row_valid ? nx_ss : 0x3fffffff, jlo0 = nx_ss0;
      if (window > 0) {
        jlo = row_valid ? max(jlo, i - window + 1) : jlo;
        jlo0 = max(jlo0, q0 - window + 1);
      }
      const int kv_lo = jlo0 / BKV;
      cur.next();
      if (cur.valid) prefetch_item(cur);
      const unsigned span = static_cast<unsigned>(i - jlo);
      float m_run = -INFINITY, l_run = 0.f, alpha_pend = 1.f;
      float o_acc[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) o_acc[c] = 0.f;
      auto fold = [&](uint32_t tt, float alpha) {  // O_{tt} (TMEM) folded in with the rescale factor of tile tt
        const uint32_t ob = tt & 1;
        mbar_wait(&o_bar[ob], (tt >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tO0 + ob * 64 + lane_base + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) o_acc[c0 + c] = fmaf(o_acc[c0 + c], alpha, __uint_as_float(r[c]));
        }
        tc_fence_before();
      };
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const int kv0 = kvb * BKV;
        mbar_wait(s_bar, t & 1);
        tc_fence_after();
        // per 32-column chunk the warp (32 consecutive rows) votes: visible to every row / to none / mixed
        unsigned cls_bits = 0;  // 2 bits per chunk: 0 = all visible, 1 = none, 2 = mixed (kept in a register)
        float mraw = -INFINITY;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          const int jb = kv0 + c4 * 32;
          const bool full = row_valid && jb >= jlo && jb + 31 <= i;
          const bool none = !row_valid || jb > i || jb + 31 < jlo;
          const unsigned cl = __all_sync(0xffffffffu, full) ? 0u : (__all_sync(0xffffffffu, none) ? 1u : 2u);
          cls_bits |= cl << (2 * c4);
          if (cl == 1u) continue;
          uint32_t r[32];
          tmem_ld32(tS + lane_base + c4 * 32, r);
          tmem_ld_wait();
          if (cl == 2u) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              mraw = fmaxf(mraw, (row_valid && visible(jb + j, jlo, span)) ? __uint_as_float(r[j]) : -INFINITY);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(r[j]));
          }
        }
        const float m_new = fmaxf(m_run, mraw * scale_log2);
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2(m_run - m_safe);
        const float neg_m = -m_safe;
        float ls0 = 0.f, ls1 = 0.f;
        // the previous item's output was staged in this warp's rows of the P buffer: its bulk store has read them
        bulk_wait_read0();
        __syncwarp();
        // pass 2 (exponentials -> P)
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          const unsigned cl = (cls_bits >> (2 * c4)) & 3u;
          uint32_t r[32];
          if (cl != 1u) {
            tmem_ld32(tS + lane_base + c4 * 32, r);
            tmem_ld_wait();
          }
          if (cl == 1u) {  // nothing visible: the P chunk is zero
#pragma unroll
            for (int q = 0; q < 4; ++q) ptile_store(sP, row, c4 * 4 + q, make_uint4(0u, 0u, 0u, 0u));
            continue;
          }
          float p[32];
          if (cl == 2u) {
            const int jb = kv0 + c4 * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pv = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
              p[j] = (row_valid && visible(jb + j, jlo, span)) ? pv : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            ls0 += p[j];
            ls1 += p[j + 1];
          }
          if (drop.thresh) {  // dropout on the (still unnormalised) probabilities; the row sum stays undropped (:104,129)
            const uint32_t keep = attn_keep_mask32(drop, b * H + h, i, kv0 + c4 * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = ((keep >> j) & 1u) ? p[j] * drop.inv_keep16 : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
            v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
            v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
            v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
            ptile_store(sP, row, c4 * 4 + q, v);
          }
        }
        const float lsum = ls0 + ls1;
        l_run = l_run * alpha + lsum;
        m_run = m_new;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_bar);
        if (kvb > kv_lo) fold(t - 1, alpha_pend);  // the previous tile's P·V retired long ago
        alpha_pend = alpha;
      }
      fold(t - 1, alpha_pend);  // last tile of the item (its P·V has retired: the P buffer is free)
      {
        const float inv = 1.f / l_run;
        // output rows -> bf16, staged in this warp's rows of the P buffer (first atom), one TMA store per warp
        constexpr int ROWB = HD * 2;
        uint8_t* stg = smem + S::kP + (warp & 3) * 4096;
        const int sw = ROWB == 128 ? (lane & 7) : (ROWB == 64 ? ((lane >> 1) & 3) : (ROWB == 32 ? ((lane >> 2) & 1) : 0));
        const uint32_t rowp = smem_u32(stg + lane * ROWB);
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 8) {
          const uint32_t off = static_cast<uint32_t>(((c0 >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + off),
                       "r"(pack_bf16(o_acc[c0 + 0] * inv, o_acc[c0 + 1] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 2] * inv, o_acc[c0 + 3] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 4] * inv, o_acc[c0 + 5] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 6] * inv, o_acc[c0 + 7] * inv))
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tm_out, smem_u32(stg), h * HD, q0 + (warp & 3) * 32, b);
          bulk_commit();
        }
        if (row_valid) lse[((size_t)b * H + h) * T + i] = (m_run + log2f(l_run)) * kLn2;
      }
    }
    bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Two-stream forward with O accumulated in TMEM and a single-pass softmax (lazy reference maximum).
template <int HD, bool DROP>
__global__ void __launch_bounds__(384, 1)
attn_fwd_w3_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_out,
                   const int32_t* __restrict__ seg_start, float* __restrict__ lse, int Bsz, int T, int H, int Hk,
                   int window, float scale_log2, const DropoutCfg drop_in, int smem_bytes) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = FwdW2Smem<HD>;
  constexpr int NS = S::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem0 - smem_raw) + S::kTotal > smem_bytes) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strm = warp < 8 ? (warp >> 2) : (warp & 1);  // softmax warps 0-3 | 4-7, MMA warps 8 | 9, TMA warps 10 | 11
  uint8_t* smem = smem0 + strm * S::kStream;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* q_full = bars;             // Q of the item has landed
  uint64_t* q_empty = bars + 1;        // every S MMA of the item has retired
  uint64_t* kv_full = bars + 2;        // [NS]
  uint64_t* kv_empty = bars + 2 + NS;  // [NS] the P·V MMA that read the stage has retired
  uint64_t* s_bar = bars + 2 + 2 * NS; // S of a tile is in TMEM (and the previous P·V has retired: P is free)
  uint64_t* p_bar = s_bar + 1;         // P of a tile is in smem, its S has been consumed
  uint64_t* o_bar = p_bar + 1;         // the LAST P·V of an item has retired: O (TMEM) is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem0 + S::kBar) + 32;  // shared by both streams (behind stream 0's barriers)

  const int rep = H / Hk;
  const int vcta = blockIdx.x * 2 + strm, vgrid = gridDim.x * 2;
  const int nqb = (T + BQ - 1) / BQ;

  if (tid == 0 || tid == 128) {  // one thread of each stream initialises that stream's barriers
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm_out);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, 4);
    mbar_init(o_bar, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail; its results are read from here on
  const uint32_t tS = tmem_base + strm * 256, tO = tS + 128;  // S: 128 columns, O: HD columns (accumulated over an item)

  auto kv_lo_of = [&](int qb, int b) { return row_jlo(seg_start ? seg_start + (size_t)b * T : nullptr, qb * BQ, T, window) / BKV; };

  if (warp >= 10) {
    // ================================================================= TMA producer
    const bool leader = elect_one();
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, h = cur.head, b = cur.b, kv_lo = nx_lo;
      const int kvh = h / rep;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);
      mbar_wait(q_empty, (n_it & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(q_full, C::TILE_BYTES);
        tma_tile<HD>(smem + S::kQ, &tm, q_full, h * HD, qb * BQ, b);
      }
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const int st = t % NS;
        mbar_wait(&kv_empty[st], ((t / NS) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&kv_full[st], 2 * C::TILE_BYTES);
          tma_tile<HD>(smem + S::kK + st * C::TILE_BYTES, &tm, &kv_full[st], (H + kvh) * HD, kvb * BKV, b);
          tma_tile<HD>(smem + S::kV + st * C::TILE_BYTES, &tm, &kv_full[st], (H + Hk + kvh) * HD, kvb * BKV, b);
        }
      }
    }
  } else if (warp >= 8) {
    // ================================================================= MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, false, true);
    const uint32_t sQ_u = smem_u32(smem + S::kQ), sK_u = smem_u32(smem + S::kK), sV_u = smem_u32(smem + S::kV);
    const uint32_t sP_u = smem_u32(smem + S::kP);
    // O (TMEM) += P(tt) V(tt); `first` = first tile of its item (overwrite), `last` = last tile (publishes O)
    auto issue_pv = [&](uint32_t tt, bool first, bool last) {
      const uint32_t st = tt % NS;
      mbar_wait(p_bar, tt & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_bf16(tO, ptile_kmajor(sP_u, ks), C::mnmajor(sV_u + st * C::TILE_BYTES, ks), idesc_o, ks > 0 || !first);
        if (last) umma_commit(o_bar);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    };
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_lo = cur.valid ? kv_lo_of(cur.k, cur.b) : 0;
    uint32_t t = 0;
    bool pv_first = false, pv_last = false;  // flags of tile t-1
    for (int n_it = 0; cur.valid; ++n_it) {
      const int qb = cur.k, kv_lo = nx_lo;
      cur.next();
      if (cur.valid) nx_lo = kv_lo_of(cur.k, cur.b);
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        // P·V of the previous tile first: the commit behind S(t) then also covers it, so the softmax warps may
        // overwrite the single P buffer (and rescale O) as soon as they see S(t)
        if (t >= 1) issue_pv(t - 1, pv_first, pv_last);
        pv_first = kvb == kv_lo;
        pv_last = kvb == qb;
        if (kvb == kv_lo) mbar_wait(q_full, n_it & 1);
        const uint32_t st = t % NS;
        mbar_wait(&kv_full[st], (t / NS) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_bf16(tS, C::kmajor(sQ_u, ks), C::kmajor(sK_u + st * C::TILE_BYTES, ks), idesc_s, ks > 0);
          umma_commit(s_bar);
          if (kvb == qb) umma_commit(q_empty);  // last S MMA of the item: the Q buffer is free after this
        }
        __syncwarp();
      }
    }
    if (t >= 1) {
      issue_pv(t - 1, pv_first, pv_last);
      mbar_wait(&kv_empty[(t - 1) % NS], ((t - 1) / NS) & 1);  // every MMA has retired before the CTA tears down
    }
  } else {
    // ================================================================= softmax warps: thread = query row
    const int row = tid & 127;
    const bool sleader = elect_one();  // issues this warp's output stores (uniform-register UTMASTG, no R2UR loop)
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    PairCursor cur;
    cur.begin(vcta, vgrid, nqb, H, Bsz, true);
    int nx_ss = 0, nx_ss0 = 0;
    auto prefetch_item = [&](const PairCursor& c) {
      const int i0 = c.k * BQ;
      nx_ss0 = seg_start ? seg_start[(size_t)c.b * T + i0] : 0;
      nx_ss = (seg_start && i0 + row < T) ? seg_start[(size_t)c.b * T + i0 + row] : 0;
    };
    if (cur.valid) prefetch_item(cur);
    uint32_t t = 0;
    int n_item = 0;
    uint8_t* sP = smem + S::kP;
    constexpr float kLagSum = 1099511627776.f;  // 2^40: the row sum of a tile whose maximum overshoots m_ref by ~2^33..2^40
    while (cur.valid) {
      const int qb = cur.k, h = cur.head, b = cur.b;
      const int q0 = qb * BQ, i = q0 + row;
      const bool row_valid = i < T;
      int jlo = row_valid ? nx_ss : 0x3fffffff, jlo0 = nx_ss0;
      if (window > 0) {
        jlo = row_valid ? max(jlo, i - window + 1) : jlo;
        jlo0 = max(jlo0, q0 - window + 1);
      }
      const int kv_lo = jlo0 / BKV;
      cur.next();
      if (cur.valid) prefetch_item(cur);
      const unsigned span = static_cast<unsigned>(i - jlo);
      // Softmax is shift invariant: any reference r with |r - rowmax| < ~100 (log2 units) gives the same P (bf16 keeps
      // the fp32 exponent range) and the same O / l up to an exact power of two.  m_ref is the exact maximum of the
      // row's FIRST visible tile and is only moved when a later tile overshoots it by 2^kLag (checked on the row sum):
      // one pass over S per tile, and O never leaves TMEM (the P·V MMAs accumulate into it; no per-tile fold).
      float m_ref = -INFINITY, l_run = 0.f;
      for (int kvb = kv_lo; kvb <= qb; ++kvb, ++t) {
        const int kv0 = kvb * BKV;
        // per 32-column chunk the warp (32 consecutive rows) votes: visible to every row / to none / mixed
        unsigned cls_bits = 0;  // 2 bits per chunk: 0 = all visible, 1 = none, 2 = mixed
        bool sees = false;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int jb = kv0 + c4 * 32;
          const bool full = row_valid && jb >= jlo && jb + 31 <= i;
          const bool none = !row_valid || jb > i || jb + 31 < jlo;
          sees |= !none;
          cls_bits |= (__all_sync(0xffffffffu, full) ? 0u : (__all_sync(0xffffffffu, none) ? 1u : 2u)) << (2 * c4);
        }
        mbar_wait(s_bar, t & 1);  // S(t) is in TMEM; P·V(t-1) has retired (P is free, O is quiescent)
        tc_fence_after();
        // the previous item's output was staged in this warp's rows of the P buffer: its bulk store has read them
        bulk_wait_read0();
        __syncwarp();
        bool need_max = __any_sync(0xffffffffu, sees && m_ref == -INFINITY);
        float lsum = 0.f;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
          if (need_max) {  // exact masked maximum of the tile -> new reference; l and O follow by 2^(old - new)
            float mraw = -INFINITY;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
              const unsigned cl = (cls_bits >> (2 * c4)) & 3u;
              if (cl == 1u) continue;
              const int jb = kv0 + c4 * 32;
              uint32_t r[32];
              tmem_ld32(tS + lane_base + c4 * 32, r);
              tmem_ld_wait();
              if (cl == 2u) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  mraw = fmaxf(mraw, (row_valid && visible(jb + j, jlo, span)) ? __uint_as_float(r[j]) : -INFINITY);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(r[j]));
              }
            }
            const float m_new = fmaxf(m_ref, mraw * scale_log2);
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = fast_exp2(m_ref - m_safe);  // 0 for a row that had seen nothing yet
            l_run *= alpha;
            if (kvb > kv_lo && __any_sync(0xffffffffu, alpha != 1.f)) {  // O *= alpha (rows of this warp only)
#pragma unroll
              for (int c0 = 0; c0 < HD; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tO + lane_base + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * alpha);
                tmem_st16(tO + lane_base + c0, r);
              }
              tmem_st_wait();
            }
            m_ref = m_new;
          }
          const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;  // rows that see nothing: every p is masked to 0
          float ls0 = 0.f, ls1 = 0.f;
          // exponentials -> P, chunk c4+1 in flight from TMEM while chunk c4 is processed
          uint32_t ra[32], rb[32];
          auto chunk = [&](int c4, const uint32_t (&r)[32]) {
            const unsigned cl = (cls_bits >> (2 * c4)) & 3u;
            if (cl == 1u) {  // nothing visible: the P chunk is zero
#pragma unroll
              for (int q = 0; q < 4; ++q) ptile_store(sP, row, c4 * 4 + q, make_uint4(0u, 0u, 0u, 0u));
              return;
            }
            float p[32];
            if (cl == 2u) {
              const int jb = kv0 + c4 * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float pv = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
                p[j] = (row_valid && visible(jb + j, jlo, span)) ? pv : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) p[j] = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2, neg_m));
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              ls0 += p[j];
              ls1 += p[j + 1];
            }
            if constexpr (DROP) {  // dropout on the (still unnormalised) probabilities; the row sum stays undropped (:104,129)
              const uint32_t keep = attn_keep_mask32(drop, b * H + h, i, kv0 + c4 * 32);
#pragma unroll
              for (int j = 0; j < 32; ++j) p[j] = ((keep >> j) & 1u) ? p[j] * drop.inv_keep16 : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 v;
              v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
              v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
              v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
              v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
              ptile_store(sP, row, c4 * 4 + q, v);
            }
          };
          auto load = [&](int c4, uint32_t (&r)[32]) {
            if (((cls_bits >> (2 * c4)) & 3u) != 1u) tmem_ld32(tS + lane_base + c4 * 32, r);
          };
          load(0, ra);
          tmem_ld_wait();
          load(1, rb);
          chunk(0, ra);
          tmem_ld_wait();
          load(2, ra);
          chunk(1, rb);
          tmem_ld_wait();
          load(3, rb);
          chunk(2, ra);
          tmem_ld_wait();
          chunk(3, rb);
          lsum = ls0 + ls1;
          // a tile that overshoots the reference by more than 2^kLag (or produced inf / nan): move the reference, redo
          need_max = __any_sync(0xffffffffu, !(lsum <= kLagSum));
          if (!need_max) break;
        }
        l_run += lsum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_bar);
      }
      // last tile of the item: its P·V (and with it the whole accumulation) has retired
      mbar_wait(o_bar, n_item & 1);
      ++n_item;
      tc_fence_after();
      float o_acc[HD];
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tO + lane_base + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) o_acc[c0 + c] = __uint_as_float(r[c]);
      }
      tc_fence_before();
      const float m_run = m_ref;
      {
        const float inv = 1.f / l_run;
        // output rows -> bf16, staged in this warp's rows of the P buffer (first atom), one TMA store per warp
        constexpr int ROWB = HD * 2;
        uint8_t* stg = smem + S::kP + (warp & 3) * 4096;
        const int sw = ROWB == 128 ? (lane & 7) : (ROWB == 64 ? ((lane >> 1) & 3) : (ROWB == 32 ? ((lane >> 2) & 1) : 0));
        const uint32_t rowp = smem_u32(stg + lane * ROWB);
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 8) {
          const uint32_t off = static_cast<uint32_t>(((c0 >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + off),
                       "r"(pack_bf16(o_acc[c0 + 0] * inv, o_acc[c0 + 1] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 2] * inv, o_acc[c0 + 3] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 4] * inv, o_acc[c0 + 5] * inv)),
                       "r"(pack_bf16(o_acc[c0 + 6] * inv, o_acc[c0 + 7] * inv))
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (sleader) {
          tma_store_3d(&tm_out, smem_u32(stg), h * HD, q0 + (warp & 3) * 32, b);
          bulk_commit();
        }
        if (row_valid) lse[((size_t)b * H + h) * T + i] = (m_run + log2f(l_run)) * kLn2;
      }
    }
    bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// last query tile that can see each (batch, kv tile): the warp-specialised backward reads this table instead of
// searching seg_start itself.  Side job of the first threads of the delta kernels.
__device__ __forceinline__ int qhi_ctr_index(int B, int T) { return (B * ((T + BKV - 1) / BKV) + 63) / 64 * 64; }
__device__ __forceinline__ void fill_qhi_tab(int e, int B, int T, const int32_t* __restrict__ seg_start, int window,
                                             int* __restrict__ qhi_tab, int ctr_init) {
  const int nkb = (T + BKV - 1) / BKV;
  if (e == 0) qhi_tab[qhi_ctr_index(B, T)] = ctr_init;  // work counter of the backward kernel (first unclaimed pair)
  if (e >= B * nkb) return;
  const int b = e / nkb, kvb = e - b * nkb;
  int hi_pos = T - 1;
  const int kv_last = min(T - 1, kvb * BKV + BKV - 1);
  if (window > 0) hi_pos = min(hi_pos, kv_last + window - 1);
  if (seg_start) hi_pos = min(hi_pos, upper_bound_i32(seg_start + (size_t)b * T, T, kv_last) - 1);
  qhi_tab[e] = min(hi_pos / BQ, (T + BQ - 1) / BQ - 1);
}

// delta[b,h,t] = sum_c out[b,t,h,c] * dout[b,t,h,c].  G = hd/8 threads per (token, head), 16-byte loads: the
// whole warp reads 512 contiguous bytes of each tensor.
template <int G>
__global__ void attn_delta_vec_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                      float* __restrict__ delta, int B, int T, int H,
                                      const int32_t* __restrict__ seg_start, int window, int* __restrict__ qhi_tab,
                                      int ctr_init, float4* __restrict__ dq_zero) {
  pdl_wait();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte chunk index
  if (qhi_tab && e < (long long)B * ((T + BKV - 1) / BKV)) fill_qhi_tab((int)e, B, T, seg_start, window, qhi_tab, ctr_init);
  const long long total = (long long)B * T * H * G;
  // the fp32 dQ workspace has 8 floats per chunk of O: it is cleared here (32 contiguous bytes per thread) instead of
  // by a separate memset pass in front of the backward kernel
  if (dq_zero && e < total) {
    dq_zero[2 * e] = make_float4(0.f, 0.f, 0.f, 0.f);
    dq_zero[2 * e + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
  if (e < total) {
    const uint4 a = reinterpret_cast<const uint4*>(o)[e];
    const uint4 g = reinterpret_cast<const uint4*>(dout)[e];
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 x = unpack_bf16(av[q]), y = unpack_bf16(gv[q]);
      s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (e < total && (threadIdx.x & (G - 1)) == 0) {
    const long long w = e / G;  // (token, head)
    const long long tok = w / H;
    const int h = (int)(w - tok * H);
    const long long bb = tok / T, t = tok - bb * T;
    delta[(bb * H + h) * T + t] = s;
  }
}

// generic head sizes (48, 96): one warp per (token, head)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int B, int T, int H, int hd,
                                  const int32_t* __restrict__ seg_start, int window, int* __restrict__ qhi_tab,
                                  int ctr_init) {
  if (qhi_tab) fill_qhi_tab(blockIdx.x * blockDim.x + threadIdx.x, B, T, seg_start, window, qhi_tab, ctr_init);
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * T * H) return;
  const long long tok = w / H;
  const int h = (int)(w - tok * H);
  const __nv_bfloat16* po = o + tok * (long long)(H * hd) + h * hd;
  const __nv_bfloat16* pd = dout + tok * (long long)(H * hd) + h * hd;
  float s = 0.f;
  for (int c = lane * 2; c < hd; c += 64) {
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(po + c));
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(pd + c));
    s += a.x * g.x + a.y * g.y;
  }
  s = warp_sum(s);
  if (lane == 0) {
    const long long bb = tok / T, t = tok - bb * T;
    delta[(bb * H + h) * T + t] = s;
  }
}

template <int HD>
struct BwdSmem {
  using C = HeadCfg<HD>;
  static constexpr int NQB = HD > 96 ? 1 : 2;          // (Q, dO) buffers
  // HD <= 64: dQ has its own TMEM columns and smem staging, so the S/dP MMAs of the next query tile are issued
  // right behind the gradient MMAs of the current one and overlap with the dQ write-out.
  static constexpr bool kPipelined = HD <= 64;
  static constexpr int kK = 0;
  static constexpr int kV = kK + C::TILE_BYTES;
  static constexpr int kQ = kV + C::TILE_BYTES;
  static constexpr int kdO = kQ + NQB * C::TILE_BYTES;
  static constexpr int kP = kdO + NQB * C::TILE_BYTES;
  static constexpr int kdS = kP + kPTileBytes;
  // dQ leaves through TMA tensor reduce-adds when a half row (HD/2 floats) is 64 or 128 bytes (or 2 x 128):
  // each warp stages its 32 rows as swizzled {HD/2 or 32 floats, 32 rows} boxes; otherwise one small bulk
  // reduce per thread from padded rows.
  static constexpr bool kTmaDq = (HD == 32 || HD == 64 || HD == 128);
  static constexpr int kdQBoxCols = HD == 32 ? 16 : 32;         // floats per box row
  static constexpr int kdQBoxes = (HD / 2) / kdQBoxCols;        // boxes per half row (1, or 2 for HD = 128)
  static constexpr int kdQRow = kTmaDq ? HD * 4 : HD * 4 + (HD <= 96 ? 16 : 0);
  static constexpr int kdQ = kdS + kPTileBytes;                 // dedicated staging when pipelined, else aliases P+dS
  static constexpr int kBar = kdQ + (kPipelined ? 128 * kdQRow : 0);
  static constexpr int kTotal = kBar + 64;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kPipelined || 128 * kdQRow <= 2 * kPTileBytes, "dQ staging must fit in the P+dS region");
  static_assert(kDynamic <= 232448, "exceeds the shared memory of one CTA");
};

// Persistent: one CTA per SM walks work items (kv tile, kv head, batch), heaviest first.  Per item it loops
// over the query heads of the GQA group and the query tiles that can see the kv tile; dK/dV accumulate in TMEM
// over that loop; each dQ tile is reduce-added into an fp32 workspace [B*H, T, hd] by TMA.  The K/V and first
// (Q, dO) tiles of the NEXT item are requested as soon as the last MMAs of the current one retire, so their
// latency hides behind the dK/dV write-out.  TMEM is allocated and the barriers are initialised once.
// 256 threads: thread t owns row (t & 127) of every tile and the column half (t >> 7).
template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const int32_t* __restrict__ seg_start,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                float* __restrict__ dq_ws, int Bsz, int T, int H, int Hk, int window, float scale,
                const DropoutCfg drop_in) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = BwdSmem<HD>;
  constexpr int TMEM_COLS = 512;
  constexpr int HH = HD / 2;
  constexpr int NQB = S::NQB;
  constexpr bool kPipe = S::kPipelined;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + S::kK;
  uint8_t* sV = smem + S::kV;
  uint8_t* sP = smem + S::kP;
  uint8_t* sdS = smem + S::kdS;
  uint8_t* sdQ = kPipe ? smem + S::kdQ : sP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* kv_bar = bars;
  uint64_t* q_bar = bars + 1;  // [2]
  uint64_t* s_bar = bars + 3;  // S and dP of a query tile are in TMEM
  uint64_t* g_bar = bars + 4;  // dV, dK, dQ MMAs of a query tile have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  int* s_qhi = reinterpret_cast<int*>(bars + 6);  // [2]: query-tile upper bound of the current / next item

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  const int rep = H / Hk;
  const int W = (H + 2 * Hk) * HD;
  const int nqb_total = (T + BQ - 1) / BQ;
  const int per_kvb = Hk * Bsz;
  const int n_items = nqb_total * per_kvb;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_bar, 1);
    mbar_init(&q_bar[0], 1);
    mbar_init(&q_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(g_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;              // 128 cols: scores
  const uint32_t tdP = tmem_base + 128;       // 128 cols
  const uint32_t tdV = tmem_base + 256;       // HD cols, accumulates over an item
  const uint32_t tdK = tmem_base + 256 + HD;  // HD cols, accumulates over an item
  const uint32_t tdQ = kPipe ? tmem_base + 256 + 2 * HD : tS;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);   // S = Q Kᵀ, dP = dO Vᵀ
  constexpr uint32_t idesc_kv = umma_idesc_bf16(128, HD, true, true);     // dV = Pᵀ dO, dK = dSᵀ Q
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, false, true);     // dQ = dS K
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float scale_log2 = scale * kLog2e;

  // item -> (kv tile, kv head, batch); kv tile 0 sees the most query tiles, so items are ordered heaviest first
  auto decode = [&](int item, int& kvb, int& kvh, int& b) {
    kvb = item / per_kvb;
    const int r = item - kvb * per_kvb;
    kvh = r % Hk;
    b = r / Hk;
  };
  // last query tile that can see kv tile kvb of batch b (thread 0: binary search over the segment starts)
  auto last_q_tile = [&](int kvb, int b) {
    int hi_pos = T - 1;
    const int kv_last = min(T - 1, kvb * BKV + BKV - 1);
    if (window > 0) hi_pos = min(hi_pos, kv_last + window - 1);
    if (seg_start) hi_pos = min(hi_pos, upper_bound_i32(seg_start + (size_t)b * T, T, kv_last) - 1);
    return min(hi_pos / BQ, nqb_total - 1);
  };
  // thread 0: request K, V and the first (Q, dO) of an item.  `gq` = running count of (Q, dO) tile loads.
  auto request_item = [&](int kvb, int kvh, int b, int gq) {
    mbar_expect_tx(kv_bar, 2 * C::TILE_BYTES);
    tma_tile<HD>(sK, &tm_qkv, kv_bar, (H + kvh) * HD, kvb * BKV, b);
    tma_tile<HD>(sV, &tm_qkv, kv_bar, (H + Hk + kvh) * HD, kvb * BKV, b);
    const int nb = gq % NQB;
    mbar_expect_tx(&q_bar[nb], 2 * C::TILE_BYTES);
    tma_tile<HD>(smem + S::kQ + nb * C::TILE_BYTES, &tm_qkv, &q_bar[nb], (kvh * rep) * HD, kvb * BQ, b);
    tma_tile<HD>(smem + S::kdO + nb * C::TILE_BYTES, &tm_do, &q_bar[nb], (kvh * rep) * HD, kvb * BQ, b);
  };

  uint32_t s_phase = 0, g_phase = 0, kv_phase = 0;
  int gq0 = 0;   // (Q, dO) tiles consumed before the current item
  int slot = 0;  // which s_qhi entry belongs to the current item
  if (tid == 0 && (int)blockIdx.x < n_items) {
    int kvb, kvh, b;
    decode(blockIdx.x, kvb, kvh, b);
    s_qhi[0] = last_q_tile(kvb, b);
    request_item(kvb, kvh, b, 0);
  }
  __syncthreads();

  for (int item = blockIdx.x; item < n_items; item += gridDim.x, slot ^= 1) {
    int kvb, kvh, b;
    decode(item, kvb, kvh, b);
    const int kv0 = kvb * BKV;
    const int kcol = (H + kvh) * HD, vcol = (H + Hk + kvh) * HD;
    const int32_t* ssb = seg_start ? seg_start + (size_t)b * T : nullptr;
    const int qb_lo = kvb, qb_hi = s_qhi[slot];
    const int nq = qb_hi - qb_lo + 1;
    const int niter = nq * rep;  // >= 1: a kv tile always sees its own diagonal query tile
    const int next_item = item + gridDim.x;

    auto load_q = [&](int it2) {
      const int nb = (gq0 + it2) % NQB;
      const int nh = kvh * rep + it2 / nq, nq0 = (qb_lo + it2 % nq) * BQ;
      mbar_expect_tx(&q_bar[nb], 2 * C::TILE_BYTES);
      tma_tile<HD>(smem + S::kQ + nb * C::TILE_BYTES, &tm_qkv, &q_bar[nb], nh * HD, nq0, b);
      tma_tile<HD>(smem + S::kdO + nb * C::TILE_BYTES, &tm_do, &q_bar[nb], nh * HD, nq0, b);
    };
    // S and dP of query tile it2 (thread 0 only)
    auto issue_scores = [&](int it2) {
      const int g = gq0 + it2, nb = g % NQB;
      const uint32_t q_s = smem_u32(smem + S::kQ + nb * C::TILE_BYTES), do_s = smem_u32(smem + S::kdO + nb * C::TILE_BYTES);
      mbar_wait(&q_bar[nb], (g / NQB) & 1);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tS, C::kmajor(q_s, ks), C::kmajor(smem_u32(sK), ks), idesc_s, ks > 0);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tdP, C::kmajor(do_s, ks), C::kmajor(smem_u32(sV), ks), idesc_s, ks > 0);
      umma_commit(s_bar);
    };

    int n_kvb = 0, n_kvh = 0, n_b = 0;
    if (tid == 0) {
      mbar_wait(kv_bar, kv_phase);
      issue_scores(0);
      if (next_item < n_items) {  // look ahead: the next item's tile range (global loads, off the critical path)
        decode(next_item, n_kvb, n_kvh, n_b);
        s_qhi[slot ^ 1] = last_q_tile(n_kvb, n_b);
      }
    }
    kv_phase ^= 1;

    // per-row statistics of the first tile (later tiles are fetched one tile ahead)
    // (raw loads only: the first dependent use is one tile later, so no warp stalls on them)
    float nx_lse = 0.f, nx_dl = 0.f;
    int nx_ss = 0;
    {
      const int i0 = qb_lo * BQ + row;
      const size_t st0 = ((size_t)b * H + kvh * rep) * T + (i0 < T ? i0 : 0);
      nx_lse = i0 < T ? lse[st0] : 0.f;
      nx_dl = i0 < T ? delta[st0] : 0.f;
      nx_ss = (ssb && i0 < T) ? ssb[i0] : 0;
    }

    for (int it = 0; it < niter; ++it) {
      const int buf = (gq0 + it) % NQB;
      const int hq = kvh * rep + it / nq;     // query head
      const int qb = qb_lo + it % nq;
      const int q0 = qb * BQ;
      uint8_t* sQ = smem + S::kQ + buf * C::TILE_BYTES;
      uint8_t* sdO = smem + S::kdO + buf * C::TILE_BYTES;
      // with two (Q, dO) buffers the other one was released when the previous tile's gradient MMAs retired
      if (tid == 0 && NQB == 2 && it + 1 < niter) load_q(it + 1);
      const int i = q0 + row;
      const bool row_ok = i < T;
      const float lse2 = nx_lse * kLog2e, dl = nx_dl;
      int jlo = row_ok ? nx_ss : 0x3fffffff;
      if (window > 0) jlo = max(jlo, i - window + 1);
      if (it + 1 < niter) {  // next tile's statistics: the loads complete while this tile is processed
        const int nh = kvh * rep + (it + 1) / nq, ni = (qb_lo + (it + 1) % nq) * BQ + row;
        const size_t st2 = ((size_t)b * H + nh) * T + (ni < T ? ni : 0);
        nx_lse = ni < T ? lse[st2] : 0.f;
        nx_dl = ni < T ? delta[st2] : 0.f;
        nx_ss = (ssb && ni < T) ? ssb[ni] : 0;
      }
      // per-score tests only for rows that do not see the whole kv tile (diagonal, segment start, window, ragged end)
      const bool need_mask = !row_ok || (kv0 + BKV - 1 > i) || (kv0 < jlo);
      mbar_wait(s_bar, s_phase);
      s_phase ^= 1;
      tc_fence_after();

#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c4 = half * 2 + cc;
        uint32_t rs[32], rp[32];
        tmem_ld32(tS + lane_base + c4 * 32, rs);
        tmem_ld32(tdP + lane_base + c4 * 32, rp);
        tmem_ld_wait();
        float p[32], ds[32];
        if (need_mask) {
          const int jb = kv0 + c4 * 32;
          const unsigned span = static_cast<unsigned>(i - jlo);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pv = fast_exp2(fmaf(__uint_as_float(rs[j]), scale_log2, -lse2));
            p[j] = (row_ok && visible(jb + j, jlo, span)) ? pv : 0.f;
            ds[j] = p[j] * (__uint_as_float(rp[j]) - dl);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            p[j] = fast_exp2(fmaf(__uint_as_float(rs[j]), scale_log2, -lse2));
            ds[j] = p[j] * (__uint_as_float(rp[j]) - dl);
          }
        }
        if (drop.thresh) {  // P feeds dV as dropout(P); dS = P * (dP * mask/(1-p) - delta)
          const uint32_t keep = attn_keep_mask32(drop, b * H + hq, i, kv0 + c4 * 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float mk = ((keep >> j) & 1u) ? drop.inv_keep16 : 0.f;
              const float pj = p[j];
              ds[j] = pj * (__uint_as_float(rp[j]) * mk - dl);
              p[j] = pj * mk;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v, w;
          v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
          v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
          v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
          v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
          w.x = pack_bf16(ds[q * 8 + 0], ds[q * 8 + 1]);
          w.y = pack_bf16(ds[q * 8 + 2], ds[q * 8 + 3]);
          w.z = pack_bf16(ds[q * 8 + 4], ds[q * 8 + 5]);
          w.w = pack_bf16(ds[q * 8 + 6], ds[q * 8 + 7]);
          ptile_store(sP, row, c4 * 4 + q, v);
          ptile_store(sdS, row, c4 * 4 + q, w);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BQ / 16; ++ks)  // dV[kv,hd] += Pᵀ[kv,q] dO[q,hd]
          umma_bf16(tdV, ptile_mnmajor(smem_u32(sP), ks), C::mnmajor(smem_u32(sdO), ks), idesc_kv, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < BQ / 16; ++ks)  // dK[kv,hd] += dSᵀ[kv,q] Q[q,hd]
          umma_bf16(tdK, ptile_mnmajor(smem_u32(sdS), ks), C::mnmajor(smem_u32(sQ), ks), idesc_kv, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)  // dQ[q,hd] = dS[q,kv] K[kv,hd]
          umma_bf16(tdQ, ptile_kmajor(smem_u32(sdS), ks), C::mnmajor(smem_u32(sK), ks), idesc_q, ks > 0);
        umma_commit(g_bar);
        // pipelined: the next tile's scores queue up right behind (their TMEM columns and operands are free)
        if (kPipe && it + 1 < niter) issue_scores(it + 1);
      }
      mbar_wait(g_bar, g_phase);
      g_phase ^= 1;
      tc_fence_after();
      if (tid == 0) {
        if (NQB == 1 && it + 1 < niter) load_q(it + 1);  // single buffer: free only now
        // last tile of the item: K, V, Q, dO buffers are all free -> request the next item's tiles now
        if (it + 1 == niter && next_item < n_items) request_item(n_kvb, n_kvh, n_b, gq0 + niter);
      }
      // dQ tile -> smem -> reduce-add into the fp32 workspace [B*H, T, hd].  The read-completion of the previous
      // tile's bulk operations is only awaited right before the staging is rewritten.
      if constexpr (S::kTmaDq) {
        // per warp: rows 32*(warp&3).., this thread's half; swizzled {kdQBoxCols, 32} boxes, 1-2 TMA tensor
        // reduce-adds per warp (rows beyond T are clipped by the 3-D map)
        constexpr int BC = S::kdQBoxCols, NB = S::kdQBoxes, BOX_BYTES = 32 * BC * 4, ROWB = BC * 4;
        uint8_t* wbase = sdQ + (warp * NB) * BOX_BYTES;  // warp w owns boxes [w*NB, w*NB+NB)
        const int lr = tid & 31;
        if (lr == 0) bulk_wait_read0();
        __syncwarp();
        uint32_t rq[HH];
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {  // issue all loads, wait once
          uint32_t (&r8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rq[c0]);
          tmem_ld8(tdQ + lane_base + half * HH + c0, r8);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          const int bx = c0 / BC, cc = (c0 % BC) * 4;  // box, byte column inside the box row
          const uint32_t rowp = smem_u32(wbase + bx * BOX_BYTES + lr * ROWB);
          const int sw = (ROWB == 128) ? (lr & 7) : ((lr >> 1) & 3);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (((cc >> 4) ^ sw) << 4)), "r"(rq[c0]),
                       "r"(rq[c0 + 1]), "r"(rq[c0 + 2]), "r"(rq[c0 + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((((cc >> 4) + 1) ^ sw) << 4)),
                       "r"(rq[c0 + 4]), "r"(rq[c0 + 5]), "r"(rq[c0 + 6]), "r"(rq[c0 + 7])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lr == 0) {
#pragma unroll
          for (int bx = 0; bx < NB; ++bx) {
            asm volatile(
                "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&tm_dq)),
                "r"(smem_u32(wbase + bx * BOX_BYTES)), "r"(half * HH + bx * BC), "r"(q0 + (warp & 3) * 32), "r"(b * H + hq)
                : "memory");
          }
          bulk_commit();
          if (!kPipe) bulk_wait_read0();  // staging aliases the P/dS tiles
        }
        if (!kPipe) __syncwarp();
      } else {
        uint8_t* myrow = sdQ + row * S::kdQRow + half * (HH * 4);
        bulk_wait_read0();
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          uint32_t r[8];
          tmem_ld8(tdQ + lane_base + half * HH + c0, r);
          tmem_ld_wait();
          *reinterpret_cast<uint4*>(myrow + c0 * 4) = make_uint4(r[0], r[1], r[2], r[3]);
          *reinterpret_cast<uint4*>(myrow + c0 * 4 + 16) = make_uint4(r[4], r[5], r[6], r[7]);
        }
        fence_proxy_async_smem();
        if (row_ok) {
          float* g = dq_ws + (((size_t)b * H + hq) * T + i) * HD + half * HH;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g),
                       "r"(smem_u32(myrow)), "r"(HH * 4)
                       : "memory");
        }
        bulk_commit();
        if (!kPipe) bulk_wait_read0();  // staging aliases the P/dS tiles
      }
      if (!kPipe) {
        tc_fence_before();
        __syncthreads();  // P/dS region and the S/dQ TMEM columns are free again
        if (tid == 0 && it + 1 < niter) {
          tc_fence_after();
          issue_scores(it + 1);
        }
      }
    }
    gq0 += niter;

    // dK (scaled) and dV -> bf16 into the k / v column blocks of dqkv (TMEM lane = kv row).  The next item's
    // tiles are already in flight; its first gradient MMA (which overwrites these columns) is only issued
    // after the next block-wide barrier.
    {
      const int j = kv0 + row;
      __nv_bfloat16* gk = dqkv + ((size_t)b * T + min(j, T - 1)) * W + kcol + half * HH;
      __nv_bfloat16* gv = dqkv + ((size_t)b * T + min(j, T - 1)) * W + vcol + half * HH;
      uint32_t rka[HH], rva[HH];
#pragma unroll
      for (int c0 = 0; c0 < HH; c0 += 8) {  // all TMEM loads in flight, one wait
        tmem_ld8(tdK + lane_base + half * HH + c0, *reinterpret_cast<uint32_t(*)[8]>(&rka[c0]));
        tmem_ld8(tdV + lane_base + half * HH + c0, *reinterpret_cast<uint32_t(*)[8]>(&rva[c0]));
      }
      tmem_ld_wait();
      if (j < T) {
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          const uint32_t* rk = &rka[c0];
          const uint32_t* rv = &rva[c0];
          uint4 a, c;
          a.x = pack_bf16(__uint_as_float(rk[0]) * scale, __uint_as_float(rk[1]) * scale);
          a.y = pack_bf16(__uint_as_float(rk[2]) * scale, __uint_as_float(rk[3]) * scale);
          a.z = pack_bf16(__uint_as_float(rk[4]) * scale, __uint_as_float(rk[5]) * scale);
          a.w = pack_bf16(__uint_as_float(rk[6]) * scale, __uint_as_float(rk[7]) * scale);
          c.x = pack_bf16(__uint_as_float(rv[0]), __uint_as_float(rv[1]));
          c.y = pack_bf16(__uint_as_float(rv[2]), __uint_as_float(rv[3]));
          c.z = pack_bf16(__uint_as_float(rv[4]), __uint_as_float(rv[5]));
          c.w = pack_bf16(__uint_as_float(rv[6]), __uint_as_float(rv[7]));
          *reinterpret_cast<uint4*>(gk + c0) = a;
          *reinterpret_cast<uint4*>(gv + c0) = c;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // s_qhi[slot^1] visible; dK/dV columns drained before the next item's MMAs
    tc_fence_after();
  }
  bulk_wait_read0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ===================================================================================== backward, warp-specialised
// hd <= 64.  Same work decomposition as attn_bwd_kernel (persistent CTA, items = (kv tile, kv head, batch)), but
// the tensor-core work is driven by a dedicated warp and the (item, query tile) sequence is ONE continuous tile
// stream: with segment masks most items see only 1-3 query tiles, so nothing may drain at an item boundary.
//
//   warp 8 : warp-uniform control flow, one elected lane issues.  Runs one tile ahead with the loads and the S/dP
//            MMAs — across item boundaries too (K/V are double-buffered, the next item's K/V are requested as soon
//            as the previous item's last gradient MMAs have retired) — then issues the gradient MMAs (dV, dK, dQ)
//            of the current tile, which therefore run while the math warps already work on the next one.
//   warps 0..7 : S,dP (TMEM) -> P (single buffer, released by a commit right behind the dV MMAs), dS (double-
//            buffered) as bf16 in swizzled smem; then the dQ write-out of the PREVIOUS tile (TMEM dQ is
//            double-buffered) through per-warp TMA reduce-adds staged in the warp's own rows of the dS buffer that
//            tile has released; at the end of an item dK/dV leave through TMA stores staged in the P buffer.
//
// TMEM (512 cols): S 128 | dP 128 | dV hd | dK hd | dQ 2*hd.   smem: 2x(K, V), 2x(Q, dO), P, 2x dS.
template <int HD>
struct BwdWsSmem {
  using C = HeadCfg<HD>;
  static constexpr int kK = 0;                            // 2 buffers
  static constexpr int kV = kK + 2 * C::TILE_BYTES;       // 2 buffers
  static constexpr int kQ = kV + 2 * C::TILE_BYTES;       // 2 buffers
  static constexpr int kdO = kQ + 2 * C::TILE_BYTES;      // 2 buffers
  static constexpr int kP = kdO + 2 * C::TILE_BYTES;      // 1 buffer
  static constexpr int kdS = kP + kPTileBytes;            // 2 buffers
  static constexpr int kBar = kdS + 2 * kPTileBytes;
  static constexpr int kTotal = kBar + 128 + 8 * 8 + 8 * 16;  // + work ring: 8 mbarriers, 8 pair descriptors
  static constexpr int kDynamic = (kTotal + 1024 <= 232448) ? kTotal + 1024 : 232448;
  static constexpr bool kTmaDq = (HD == 32 || HD == 64);
  static constexpr int kdQBoxCols = HD == 32 ? 16 : 32;
  static_assert(HD <= 64, "warp-specialised backward needs 256 + 4*hd TMEM columns");
  static_assert(kTotal <= 232448, "exceeds the shared memory of one CTA");
};

template <int HD>
__global__ void __launch_bounds__(288, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dkv,
                   const int32_t* __restrict__ seg_start,
                   int* __restrict__ qhi_tab, const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dq_ws, float* __restrict__ dqkv_colsum,
                   int Bsz, int T, int H, int Hk, int window, float scale, const DropoutCfg drop_in, int smem_bytes) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  using C = HeadCfg<HD>;
  using S = BwdWsSmem<HD>;
  constexpr int TMEM_COLS = 512;
  constexpr int HH = HD / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) + S::kTotal > smem_bytes) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* kv_bar = bars;          // [2] K, V of an item have landed (buffer = item ordinal & 1)
  uint64_t* q_bar = bars + 2;       // [2] Q, dO of a tile have landed
  uint64_t* s_bar = bars + 4;       // S, dP of a tile are in TMEM
  uint64_t* p_bar = bars + 5;       // [2] P, dS of a tile are in smem (and S/dP TMEM has been consumed)
  uint64_t* g_bar = bars + 7;       // [2] dV, dK, dQ MMAs of a tile have retired
  uint64_t* dkv_bar = bars + 9;     // the math warps have drained dK/dV of an item from TMEM
  uint64_t* pfree_bar = bars + 10;  // the dV MMAs of a tile have retired: the P buffer may be rewritten
  uint64_t* sfree_bar = bars + 11;  // the math warps hold S, dP of a tile in registers: the TMEM tiles may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* sched_bar = bars + 16;                                 // [8] a pair descriptor has been published
  int4* sched = reinterpret_cast<int4*>(bars + 24);                // [8] {pair slot kk, kv head, batch, valid}

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rep = H / Hk;
  const int W = (H + 2 * Hk) * HD;
  const int nqb_total = (T + BQ - 1) / BQ;
  // Work order.  An item is one kv tile of one (batch, kv head).  Items are paired (kv tile k with kv tile
  // nkb-1-k of the same head: k+1 plus nkb-k query tiles, i.e. equal work under a causal mask) and the pairs are
  // dealt round-robin, head-major: a CTA runs the two halves of its pair back to back, and the ~148 pairs in
  // flight belong to a few dozen heads, so their dQ rows (fp32 reduce-add targets) and Q/dO tiles stay in L2
  // instead of being fetched from HBM once per kv tile.
  const int nkb = nqb_total, npk = (nkb + 1) / 2;
  const int n_pairs = npk * Hk * Bsz;
  const int G = gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    tma_prefetch_desc(&tm_dkv);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(pfree_bar, 1);
    mbar_init(sfree_bar, 8);
    for (int i = 0; i < 8; ++i) mbar_init(&sched_bar[i], 1);
    mbar_init(&q_bar[0], 1);
    mbar_init(&q_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(&p_bar[0], 8);
    mbar_init(&p_bar[1], 8);
    mbar_init(&g_bar[0], 1);
    mbar_init(&g_bar[1], 1);
    mbar_init(dkv_bar, 8);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 256 + HD;
  const uint32_t tdQ0 = tmem_base + 256 + 2 * HD;  // + buf * HD
  pdl_wait();  // the prologue above overlapped the previous kernel's tail; its results are read from here on

  // Work distribution.  The first pair of a CTA is its block index; every further pair is claimed from a global
  // counter (initialised to the grid size by the delta kernel), so CTAs that drew short items (segment masks make
  // the work per kv tile vary a lot) simply claim more.  Warp 8 claims one pair ahead, decodes it and publishes
  // the descriptor through an 8-slot ring in shared memory; the math warps follow the same sequence from the ring.
  struct Cursor {
    int kk, kvh, b, sub, kvb;
    bool valid;
  };
  int* work_ctr = qhi_tab + qhi_ctr_index(Bsz, T);
  auto first_half = [&](Cursor& c) {
    c.sub = 0;
    c.kvb = c.kk;
  };
  // second half of the pair: the mirrored kv tile (absent for the middle tile of an odd count)
  auto second_half = [&](Cursor& c) {
    if (c.sub == 0 && nkb - 1 - c.kk != c.kk) {
      c.sub = 1;
      c.kvb = nkb - 1 - c.kk;
      return true;
    }
    return false;
  };
  auto ring_read = [&](uint32_t n) {  // pair number n of this CTA (consumers)
    const uint32_t slot = n & 7;
    mbar_wait(&sched_bar[slot], (n >> 3) & 1);
    const int4 d = sched[slot];
    Cursor c;
    c.kk = d.x;
    c.kvh = d.y;
    c.b = d.z;
    c.valid = d.w != 0;
    first_half(c);
    return c;
  };

  if (warp == 8) {
    // ================================================================= TMA + MMA warp
    // All 32 lanes run the control flow (so addresses and descriptors stay warp-uniform, i.e. in uniform registers);
    // only the elected lane issues TMA / tcgen05 instructions and barrier transactions.
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, HD, true, true);
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, false, true);
    const uint32_t sK_u = smem_u32(smem + S::kK), sV_u = smem_u32(smem + S::kV);  // + kvbuf * TILE_BYTES
    int g = 0;  // global tile counter: buffer = g & 1, barrier phase = (g >> 1) & 1
    int n_it = 0;
    // last query tile of an item (table written by the delta kernel); fetched one item ahead
    auto item_qhi = [&](const Cursor& c) { return qhi_tab[c.b * nqb_total + c.kvb]; };
    auto load_kv = [&](int kvb, int kvh, int b, int ordinal) {
      const int kb = ordinal & 1;
      if (leader) {
        mbar_expect_tx(&kv_bar[kb], 2 * C::TILE_BYTES);
        tma_tile<HD>(smem + S::kK + kb * C::TILE_BYTES, &tm_qkv, &kv_bar[kb], (H + kvh) * HD, kvb * BKV, b);
        tma_tile<HD>(smem + S::kV + kb * C::TILE_BYTES, &tm_qkv, &kv_bar[kb], (H + Hk + kvh) * HD, kvb * BKV, b);
      }
    };
    auto load_q = [&](int gg, int head, int qb, int b) {
      const int nb = gg & 1;
      if (leader) {
        mbar_expect_tx(&q_bar[nb], 2 * C::TILE_BYTES);
        tma_tile<HD>(smem + S::kQ + nb * C::TILE_BYTES, &tm_qkv, &q_bar[nb], head * HD, qb * BQ, b);
        tma_tile<HD>(smem + S::kdO + nb * C::TILE_BYTES, &tm_do, &q_bar[nb], head * HD, qb * BQ, b);
      }
    };
    auto issue_scores = [&](int gg, int ordinal) {  // S = Q K^T, dP = dO V^T of tile gg with the K/V of item `ordinal`
      const int nb = gg & 1, kb = ordinal & 1;
      const uint32_t q_s = smem_u32(smem + S::kQ + nb * C::TILE_BYTES), do_s = smem_u32(smem + S::kdO + nb * C::TILE_BYTES);
      const uint32_t k_s = sK_u + kb * C::TILE_BYTES, v_s = sV_u + kb * C::TILE_BYTES;
      mbar_wait(&q_bar[nb], (gg >> 1) & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tS, C::kmajor(q_s, ks), C::kmajor(k_s, ks), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tdP, C::kmajor(do_s, ks), C::kmajor(v_s, ks), idesc_s, ks > 0);
        umma_commit(s_bar);
      }
      __syncwarp();
    };
    // publisher side of the work ring
    uint32_t n_pub = 0;  // pairs published so far
    auto publish = [&](int pi) {  // decode pair index pi (>= n_pairs: end of work), publish, return its first item
      Cursor c;
      c.valid = pi < n_pairs;
      const int bh = c.valid ? pi / npk : 0;
      const int kk0 = pi - bh * npk;
      c.kk = (kk0 + bh) % npk;  // pair slot rotated by the head index (consecutive pairs of a head differ in weight)
      c.b = bh / Hk;
      c.kvh = bh - c.b * Hk;
      first_half(c);
      const uint32_t slot = n_pub & 7;
      if (lane == 0) {
        sched[slot] = make_int4(c.kk, c.kvh, c.b, c.valid ? 1 : 0);
        mbar_arrive(&sched_bar[slot]);  // release: the descriptor is visible to whoever observes the phase
      }
      ++n_pub;
      return c;
    };
    auto claim = [&]() {  // next unclaimed pair; the value is first used one pair later, so the latency is hidden
      int v = 0;
      if (lane == 0) v = atomicAdd(work_ctr, 1);
      return v;
    };
    int claimed = 0;       // lane 0: pair index claimed for the pair after the current one
    bool more = true;      // false once the end-of-work descriptor has been published
    Cursor cur = publish(blockIdx.x);
    more = cur.valid;
    if (more) claimed = claim();
    auto cursor_next = [&](Cursor c) {
      if (second_half(c)) return c;
      if (!more) {
        c.valid = false;
        return c;
      }
      c = publish(__shfl_sync(0xffffffffu, claimed, 0));
      more = c.valid;
      if (more) claimed = claim();
      return c;
    };
    int nx_qhi = cur.valid ? item_qhi(cur) : 0;
    if (cur.valid) {  // prologue of the stream: the first item's K/V and its first tile
      load_kv(cur.kvb, cur.kvh, cur.b, 0);
      load_q(0, cur.kvh * rep, cur.kvb, cur.b);
      mbar_wait(&kv_bar[0], 0);
      issue_scores(0, 0);
    }
    for (; cur.valid; ++n_it) {
      const int kvb = cur.kvb, kvh = cur.kvh, b = cur.b;
      const Cursor nx = cursor_next(cur);
      const int nxt = nx.valid ? 0 : -1;
      const int qb_lo = kvb, qb_hi = nx_qhi;
      const int nq = qb_hi - qb_lo + 1, niter = nq * rep;
      const int nkvb = nx.kvb, nkvh = nx.kvh, nb_ = nx.b;
      if (nx.valid) nx_qhi = item_qhi(nx);
      const int kb = n_it & 1;
      int ld_h = kvh * rep, ld_q = qb_lo;  // (head, query tile) of tile `it`, advanced incrementally
      for (int it = 0; it < niter; ++it, ++g) {
        const int buf = g & 1;
        const bool last = it + 1 == niter;
        const bool has_next = !last || nxt >= 0;
        if (++ld_q > qb_hi) {
          ld_q = qb_lo;
          ++ld_h;
        }
        if (has_next) {
          // tile g-1 has retired: the (Q, dO) buffer of tile g+1 is free, and at the first tile of an item so are
          // the K/V of the previous item, whose buffer the NEXT item's K/V go into
          if (g >= 1) mbar_wait(&g_bar[buf ^ 1], ((g - 1) >> 1) & 1);
          if (it == 0 && nxt >= 0) load_kv(nkvb, nkvh, nb_, n_it + 1);
          if (!last)
            load_q(g + 1, ld_h, ld_q, b);
          else
            load_q(g + 1, nkvh * rep, nkvb, nb_);
        }
        // Within an item the next tile's scores are issued as soon as the math warps have pulled S / dP of this tile
        // into registers (sfree_bar: a few hundred cycles into their work), so S(g+1), dP(g+1) are computed WHILE P, dS
        // of tile g are being made.  After an item's LAST tile the math warps first need the gradients (dQ write-out,
        // dK/dV epilogue), so there the scores of the next item follow the gradient MMAs.
        if (has_next && !last) {
          mbar_wait(sfree_bar, g & 1);
          tc_fence_after();
          issue_scores(g + 1, n_it);
        }
        mbar_wait(&p_bar[buf], (g >> 1) & 1);  // P, dS of tile g are in smem
        tc_fence_after();
        if (it == 0 && n_it > 0) {  // dK/dV accumulators still hold the previous item until the math warps drain them
          mbar_wait(dkv_bar, (n_it - 1) & 1);
          tc_fence_after();
        }
        const uint32_t sP = smem_u32(smem + S::kP), sdS = smem_u32(smem + S::kdS + buf * kPTileBytes);
        const uint32_t sQ = smem_u32(smem + S::kQ + buf * C::TILE_BYTES), sdO = smem_u32(smem + S::kdO + buf * C::TILE_BYTES);
        const uint32_t k_s = sK_u + kb * C::TILE_BYTES;
        const uint32_t acc = it > 0;
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < BQ / 16; ++ks)  // dV[kv,hd] += P^T[kv,q] dO[q,hd]
            umma_bf16(tdV, ptile_mnmajor(sP, ks), C::mnmajor(sdO, ks), idesc_kv, acc || (ks > 0));
          umma_commit(pfree_bar);  // the single P buffer is free as soon as these have retired
#pragma unroll
          for (int ks = 0; ks < BQ / 16; ++ks)  // dK[kv,hd] += dS^T[kv,q] Q[q,hd]
            umma_bf16(tdK, ptile_mnmajor(sdS, ks), C::mnmajor(sQ, ks), idesc_kv, acc || (ks > 0));
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks)  // dQ[q,hd] = dS[q,kv] K[kv,hd]
            umma_bf16(tdQ0 + buf * HD, ptile_kmajor(sdS, ks), C::mnmajor(k_s, ks), idesc_q, ks > 0);
          umma_commit(&g_bar[buf]);
        }
        __syncwarp();
        if (has_next && last) {
          mbar_wait(&kv_bar[kb ^ 1], ((n_it + 1) >> 1) & 1);
          issue_scores(g + 1, n_it + 1);
        }
      }
      cur = nx;
    }
    // every MMA has retired before the CTA tears down its TMEM / smem
    if (g >= 1) mbar_wait(&g_bar[(g - 1) & 1], ((g - 1) >> 1) & 1);
  } else {
    // ================================================================= math warps
    const int row = tid & 127, half = tid >> 7;
    // the lane that issues (and commits) this warp's bulk tensor copies: behind an elected predicate ptxas issues
    // UTMAREDG / UTMASTG straight from uniform registers instead of an R2UR ... BRA.U.ANY loop per instruction
    const bool mleader = elect_one();
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float scale_log2 = scale * kLog2e;
    int g = 0, n_it = 0;
    uint32_t s_phase = 0;

    // dQ of tile (gg) of head hq / query tile q0 -> reduce-add into the workspace; staging = P buffer (gg & 1)
    auto dq_phase = [&](int gg, int hq, int q0, int b) {
      const int buf = gg & 1;
      mbar_wait(&g_bar[buf], (gg >> 1) & 1);
      tc_fence_after();
      // staging = the 4 KB of the dS tile that THIS warp writes (atom `half`, rows 32*(warp&3)..): no other warp's
      // dS stores can touch it, so only the issuing thread's own read-completion wait orders its reuse
      uint8_t* stage = smem + S::kdS + buf * kPTileBytes + half * 16384 + (warp & 3) * 4096;
      const uint32_t tq = tdQ0 + buf * HD + lane_base + half * HH;
      if constexpr (S::kTmaDq) {
        constexpr int BC = S::kdQBoxCols, BOX_BYTES = 32 * BC * 4, ROWB = BC * 4;
        static_assert(BOX_BYTES <= 4096, "per-warp dQ staging");
        uint8_t* wbase = stage;
        uint32_t rq[HH];
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) tmem_ld8(tq + c0, *reinterpret_cast<uint32_t(*)[8]>(&rq[c0]));
        tmem_ld_wait();
        const int sw = (ROWB == 128) ? (lane & 7) : ((lane >> 1) & 3);
        const uint32_t rowp = smem_u32(wbase + lane * ROWB);
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 4)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (((c0 >> 2) ^ sw) << 4)), "r"(rq[c0]),
                       "r"(rq[c0 + 1]), "r"(rq[c0 + 2]), "r"(rq[c0 + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (mleader) {
          asm volatile(
              "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                  reinterpret_cast<uint64_t>(&tm_dq)),
              "r"(smem_u32(wbase)), "r"(half * HH), "r"(q0 + (warp & 3) * 32), "r"(b * H + hq)
              : "memory");
          bulk_commit();  // read completion is awaited at the top of the next tile (the staging aliases a P buffer)
        }
      } else {
        static_assert(32 * HH * 4 <= 4096 && (HH * 4) % 16 == 0, "per-warp dQ staging");
        uint8_t* myrow = stage + lane * (HH * 4);
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          uint32_t r[8];
          tmem_ld8(tq + c0, r);
          tmem_ld_wait();
          *reinterpret_cast<uint4*>(myrow + c0 * 4) = make_uint4(r[0], r[1], r[2], r[3]);
          *reinterpret_cast<uint4*>(myrow + c0 * 4 + 16) = make_uint4(r[4], r[5], r[6], r[7]);
        }
        fence_proxy_async_smem();
        const int i = q0 + row;
        if (i < T) {
          float* gp = dq_ws + (((size_t)b * H + hq) * T + i) * HD + half * HH;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gp),
                       "r"(smem_u32(myrow)), "r"(HH * 4)
                       : "memory");
        }
        bulk_commit();
      }
      tc_fence_before();
    };

    // row statistics of the first tile of an item and its query-tile range, fetched one item ahead
    float nx_lse = 0.f, nx_dl = 0.f;
    int nx_ss = 0, nx_qhi = 0;
    auto prefetch_item = [&](const Cursor& c) {
      const int kvb = c.kvb, kvh = c.kvh, b = c.b;
      const int i0 = kvb * BQ + row;
      const size_t st0 = ((size_t)b * H + kvh * rep) * T + (i0 < T ? i0 : 0);
      nx_qhi = qhi_tab[b * nqb_total + kvb];
      nx_lse = i0 < T ? lse[st0] : 0.f;
      nx_dl = i0 < T ? delta[st0] : 0.f;
      nx_ss = (seg_start && i0 < T) ? seg_start[(size_t)b * T + i0] : 0;
    };
    uint32_t n_read = 0;  // pairs read from the ring so far
    auto cursor_next = [&](Cursor c) {
      if (second_half(c)) return c;
      return ring_read(n_read++);
    };
    Cursor cur = ring_read(n_read++);
    if (cur.valid) prefetch_item(cur);

    for (; cur.valid; ++n_it) {
      const int kvb = cur.kvb, kvh = cur.kvh, b = cur.b;
      const Cursor nx = cursor_next(cur);
      const int kv0 = kvb * BKV;
      const int kcol = (H + kvh) * HD, vcol = (H + Hk + kvh) * HD;
      const int32_t* ssb = seg_start ? seg_start + (size_t)b * T : nullptr;
      const int qb_lo = kvb, qb_hi = nx_qhi;
      const int nq = qb_hi - qb_lo + 1, niter = nq * rep;
      int prev_hq = 0, prev_q0 = 0;
      int hq = kvh * rep, qb = qb_lo;  // (head, query tile) of tile `it`, advanced incrementally
      for (int it = 0; it < niter; ++it, ++g) {
        const int buf = g & 1;
        const int q0 = qb * BQ;
        const int i = q0 + row;
        const bool row_ok = i < T;
        const float lse2 = nx_lse * kLog2e, dl = nx_dl;
        int jlo = row_ok ? nx_ss : 0x3fffffff;
        if (window > 0) jlo = max(jlo, i - window + 1);
        if (it + 1 < niter) {
          const int nh = qb < qb_hi ? hq : hq + 1, ni = (qb < qb_hi ? qb + 1 : qb_lo) * BQ + row;
          const size_t st2 = ((size_t)b * H + nh) * T + (ni < T ? ni : 0);
          nx_lse = ni < T ? lse[st2] : 0.f;
          nx_dl = ni < T ? delta[st2] : 0.f;
          nx_ss = (ssb && ni < T) ? ssb[ni] : 0;
        } else if (nx.valid) {
          prefetch_item(nx);
        }
        // the dQ write-out issued during the previous tile was staged in the part of the dS buffer this warp is
        // about to rewrite (and dK/dV of the previous item in its part of P): the issuing threads wait until their
        // bulk copies have read the smem
        bulk_wait_read0();
        __syncwarp();
        mbar_wait(s_bar, s_phase);
        s_phase ^= 1;
        tc_fence_after();
        uint8_t* sP = smem + S::kP;
        uint8_t* sdS = smem + S::kdS + buf * kPTileBytes;
        auto release_sdp = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sfree_bar);
        };
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c4 = half * 2 + cc;
          // the warp (32 consecutive rows) classifies the 32-column chunk: visible to every row -> no per-score
          // tests; to no row -> zeros, nothing read or exponentiated; mixed (the diagonal 32x32 blocks) -> tested
          const int jb = kv0 + c4 * 32;
          const bool full = row_ok && jb >= jlo && jb + 31 <= i;
          const bool none = !row_ok || jb > i || jb + 31 < jlo;
          const bool w_full = __all_sync(0xffffffffu, full), w_none = __all_sync(0xffffffffu, none);
          if (w_none) {
            if (cc == 0 && g >= 1) mbar_wait(pfree_bar, (g - 1) & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ptile_store(sP, row, c4 * 4 + q, make_uint4(0u, 0u, 0u, 0u));
              ptile_store(sdS, row, c4 * 4 + q, make_uint4(0u, 0u, 0u, 0u));
            }
            if (cc == 1) release_sdp();
            continue;
          }
          uint32_t rs[32], rp[32];
          tmem_ld32(tS + lane_base + c4 * 32, rs);
          tmem_ld32(tdP + lane_base + c4 * 32, rp);
          tmem_ld_wait();
          if (cc == 1) release_sdp();  // S, dP of the tile are consumed: the next tile's scores may be computed
          float p[32], ds[32];
          if (!w_full) {
            const unsigned span = static_cast<unsigned>(i - jlo);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pv = fast_exp2(fmaf(__uint_as_float(rs[j]), scale_log2, -lse2));
              p[j] = (row_ok && visible(jb + j, jlo, span)) ? pv : 0.f;
              ds[j] = p[j] * (__uint_as_float(rp[j]) - dl);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              p[j] = fast_exp2(fmaf(__uint_as_float(rs[j]), scale_log2, -lse2));
              ds[j] = p[j] * (__uint_as_float(rp[j]) - dl);
            }
          }
          if (drop.thresh) {
            const uint4 sd = attn_dropout_seed(drop, b * H + hq, i, (kv0 + c4 * 32) >> 5);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float mk[8];
              attn_keep_scale8(attn_dropout_octet(sd, q), drop.thresh16, drop.inv_keep16, mk);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int j = q * 8 + e;
                const float pj = p[j];
                ds[j] = pj * (__uint_as_float(rp[j]) * mk[e] - dl);
                p[j] = pj * mk[e];
              }
            }
          }
          // the single P buffer: the dV MMAs of the previous tile must have retired (they were issued right behind
          // this tile's S/dP MMAs, so this rarely waits)
          if (cc == 0 && g >= 1) mbar_wait(pfree_bar, (g - 1) & 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v, w;
            v.x = pack_bf16(p[q * 8 + 0], p[q * 8 + 1]);
            v.y = pack_bf16(p[q * 8 + 2], p[q * 8 + 3]);
            v.z = pack_bf16(p[q * 8 + 4], p[q * 8 + 5]);
            v.w = pack_bf16(p[q * 8 + 6], p[q * 8 + 7]);
            w.x = pack_bf16(ds[q * 8 + 0], ds[q * 8 + 1]);
            w.y = pack_bf16(ds[q * 8 + 2], ds[q * 8 + 3]);
            w.z = pack_bf16(ds[q * 8 + 4], ds[q * 8 + 5]);
            w.w = pack_bf16(ds[q * 8 + 6], ds[q * 8 + 7]);
            ptile_store(sP, row, c4 * 4 + q, v);
            ptile_store(sdS, row, c4 * 4 + q, w);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_bar[buf]);
        // while the tensor cores chew on this tile's gradients: write out dQ of the previous tile
        if (it >= 1) dq_phase(g - 1, prev_hq, prev_q0, b);
        prev_hq = hq;
        prev_q0 = q0;
        if (++qb > qb_hi) {
          qb = qb_lo;
          ++hq;
        }
      }
      dq_phase(g - 1, prev_hq, prev_q0, b);  // also waits for the last gradient MMAs of the item

      // dK (scaled) and dV -> bf16 into the k / v column blocks of dqkv (TMEM lane = kv row).  Staged in this warp's
      // own 4 KB of the P buffer (free: every gradient MMA of the item has retired, and the dQ reduce-adds in flight
      // were staged in dS) and written by two TMA stores, which clip the rows past T.
      {
        constexpr int ROWB = HH * 2;  // bytes per staged row: 64 / 32 -> TMA swizzle of that span, else none
        uint8_t* stg = smem + S::kP + half * 16384 + (warp & 3) * 4096;
        uint32_t rka[HH], rva[HH];
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          tmem_ld8(tdK + lane_base + half * HH + c0, *reinterpret_cast<uint32_t(*)[8]>(&rka[c0]));
          tmem_ld8(tdV + lane_base + half * HH + c0, *reinterpret_cast<uint32_t(*)[8]>(&rva[c0]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_bar);  // the accumulators may be overwritten by the next item
        const int sw = ROWB == 64 ? ((lane >> 1) & 3) : (ROWB == 32 ? ((lane >> 2) & 1) : 0);
        const uint32_t rowk = smem_u32(stg + lane * ROWB), rowv = rowk + 32 * ROWB;
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 8) {
          const uint32_t* rk = &rka[c0];
          const uint32_t* rv = &rva[c0];
          const uint32_t off = static_cast<uint32_t>(((c0 >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowk + off),
                       "r"(pack_bf16(__uint_as_float(rk[0]) * scale, __uint_as_float(rk[1]) * scale)),
                       "r"(pack_bf16(__uint_as_float(rk[2]) * scale, __uint_as_float(rk[3]) * scale)),
                       "r"(pack_bf16(__uint_as_float(rk[4]) * scale, __uint_as_float(rk[5]) * scale)),
                       "r"(pack_bf16(__uint_as_float(rk[6]) * scale, __uint_as_float(rk[7]) * scale))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowv + off),
                       "r"(pack_bf16(__uint_as_float(rv[0]), __uint_as_float(rv[1]))),
                       "r"(pack_bf16(__uint_as_float(rv[2]), __uint_as_float(rv[3]))),
                       "r"(pack_bf16(__uint_as_float(rv[4]), __uint_as_float(rv[5]))),
                       "r"(pack_bf16(__uint_as_float(rv[6]), __uint_as_float(rv[7])))
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (mleader) {  // read completion is awaited (by this lane) before the warp next writes its dS rows
          tma_store_3d(&tm_dkv, smem_u32(stg), kcol + half * HH, kv0 + (warp & 3) * 32, b);
          tma_store_3d(&tm_dkv, smem_u32(stg) + 32 * ROWB, vcol + half * HH, kv0 + (warp & 3) * 32, b);
          bulk_commit();
        }
        if (dqkv_colsum != nullptr && lane < HH) {
          // column sums of the staged (bf16-rounded) dK / dV piece, lane = column: the key / value bias gradients
          // without re-reading dqkv from HBM.  kv rows past T hold exact zeros (their P and dS columns are masked).
          float sk = 0.f, sv = 0.f;
          const uint32_t cb = smem_u32(stg) + (lane & 7) * 2;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int swr = ROWB == 64 ? ((rr >> 1) & 3) : (ROWB == 32 ? ((rr >> 2) & 1) : 0);
            const uint32_t a = cb + rr * ROWB + (((lane >> 3) ^ swr) << 4);
            uint16_t hk, hv;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hk) : "r"(a));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(a + 32 * ROWB));
            sk += __uint_as_float(static_cast<uint32_t>(hk) << 16);
            sv += __uint_as_float(static_cast<uint32_t>(hv) << 16);
          }
          atomicAdd(dqkv_colsum + kcol + half * HH + lane, sk);
          atomicAdd(dqkv_colsum + vcol + half * HH + lane, sv);
        }
      }
      cur = nx;
    }
    bulk_wait_read0();  // every thread that issued reduce-adds drains its own groups before the smem goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dq (fp32 [B,H,T,hd]) * scale -> bf16 into the q column block of dqkv; optionally also the column sums of the
// (bf16-rounded) result = the query bias gradient.  CTA = up to 256 rows of one (batch, head); a thread owns 8
// columns (two 16-byte loads, one 16-byte store).
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ dq_ws, __nv_bfloat16* __restrict__ dqkv, int T, int H, int hd, int W,
                       float scale, float* __restrict__ colsum) {
  __shared__ float red[256][8];
  pdl_wait();
  const int hd8 = hd >> 3;
  const int rpi = 256 / hd8;  // rows per iteration
  const int tid = threadIdx.x;
  const int r = tid / hd8, c8 = tid - r * hd8;
  const bool active = r < rpi;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int t0 = blockIdx.x * 256, t1 = min(T, t0 + 256);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (active) {
    for (int t = t0 + r; t < t1; t += rpi) {
      const float4* src = reinterpret_cast<const float4*>(dq_ws + ((size_t)bh * T + t) * hd + c8 * 8);
      const float4 v = src[0], w = src[1];
      const uint4 o = make_uint4(pack_bf16(v.x * scale, v.y * scale), pack_bf16(v.z * scale, v.w * scale),
                                 pack_bf16(w.x * scale, w.y * scale), pack_bf16(w.z * scale, w.w * scale));
      *reinterpret_cast<uint4*>(dqkv + ((size_t)b * T + t) * W + h * hd + c8 * 8) = o;
      if (colsum != nullptr) {
        const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16(ov[q]);
          acc[2 * q] += f.x;
          acc[2 * q + 1] += f.y;
        }
      }
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[tid][j] = active ? acc[j] : 0.f;
  __syncthreads();
  if (tid < hd) {
    const int cc8 = tid >> 3, j = tid & 7;
    float tsum = 0.f;
    for (int rr = 0; rr < rpi; ++rr) tsum += red[rr * hd8 + cc8][j];
    atomicAdd(colsum + h * hd + tid, tsum);
  }
}

// ===================================================================================== dense probabilities
// Introspection path only (use_sdpa=False stores last_attn, model_tiny_gpt.py:116-128): one warp per
// (b, h, i) row, SIMT dot products.  O(T^2 hd) and not tuned: never on the training path.
__global__ void attn_probs_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ seg,
                                  float* __restrict__ att, int B, int T, int H, int Hk, int hd, int window,
                                  float scale, const DropoutCfg drop_in) {
  const DropoutCfg drop = resolve_dropout(drop_in);
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * H * T) return;
  const int i = (int)(w % T);
  const int h = (int)((w / T) % H);
  const int b = (int)(w / ((long long)T * H));
  const int W = (H + 2 * Hk) * hd;
  const int kvh = h / (H / Hk);
  const __nv_bfloat16* q = qkv + ((size_t)b * T + i) * W + h * hd;
  const int jlo = row_jlo(seg ? seg + (size_t)b * T : nullptr, i, T, window);
  float* row = att + (size_t)w * T;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) {
    float s = -INFINITY;
    if (j >= jlo && j <= i) {
      const __nv_bfloat16* k = qkv + ((size_t)b * T + j) * W + (H + kvh) * hd;
      float acc = 0.f;
      for (int c = 0; c < hd; c += 2) {
        const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(q + c));
        const float2 kk = unpack_bf16(*reinterpret_cast<const uint32_t*>(k + c));
        acc += a.x * kk.x + a.y * kk.y;
      }
      s = acc * scale;
    }
    row[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float se = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float s = row[j];
    const float p = (s == -INFINITY) ? 0.f : expf(s - mx);
    row[j] = p;
    se += p;
  }
  se = warp_sum(se);
  const float inv = 1.f / se;
  for (int j = lane; j < T; j += 32) {
    float v = row[j] * inv;
    if (drop.thresh) {
      v = attn_keep(drop, b * H + h, i, j) ? v * drop.inv_keep16 : 0.f;
    }
    row[j] = v;
  }
}

inline size_t delta_floats(int B, int T, int H) { return ((size_t)B * H * T + 63) / 64 * 64; }
// range table [B, kv tiles] rounded up to 64 entries, + 64 entries for the work counter
inline size_t qhi_floats(int B, int T) { return ((size_t)B * ((T + BKV - 1) / BKV) + 63) / 64 * 64 + 64; }

int make_qkv_tmap(CUtensorMap* tm, const void* base, int B, int T, int W, int aw) {
  const uint64_t dims[3] = {(uint64_t)W, (uint64_t)T, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)W * 2, (uint64_t)T * W * 2};
  const uint32_t box[3] = {(uint32_t)aw, 128u, 1u};
  return make_tmap_bf16(tm, base, 3, dims, str, box, aw * 2);
}

template <int HD>
int launch_fwd(const void* qkv, const int32_t* seg, void* out, float* lse, int B, int T, int H, int Hk, int window,
               float scale, const DropoutCfg& drop, cudaStream_t st) {
  CUtensorMap tm;
  int rc = make_qkv_tmap(&tm, qkv, B, T, (H + 2 * Hk) * HD, HeadCfg<HD>::AW);
  if (rc) return rc;
  if constexpr (HD <= 64) {
    using SW = FwdWsSmem<HD>;
    CUtensorMap to;  // output stores: [hd/2 columns x 32 rows] boxes of out [B, T, H*hd]
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)H * HD * 2, (uint64_t)T * H * HD * 2};
    const uint32_t box[3] = {(uint32_t)(HD / 2), 32u, 1u};
    rc = make_tmap_bf16(&to, out, 3, dims, str, box, HD);  // swizzle span = row bytes (64 / 32) or none
    if (rc) return rc;
    const int n_pairs = (((T + BQ - 1) / BQ + 1) / 2) * H * B;  // the kernels deal query tiles in balanced pairs
    static const int variant = [] {  // CGPT_ATTN_FWD=ws selects the one-CTA-per-SM kernel (probe / fallback)
      const char* e = getenv("CGPT_ATTN_FWD");
      if (e && e[0] == 'w' && e[1] == 's') return 0;
      return (e && e[0] == 'w' && e[1] == '2') ? 1 : 2;  // w2: two-pass / register-folded O; default: attn_fwd_w3_kernel
    }();
    if (variant >= 1) {
      using S2 = FwdW2Smem<HD>;
      CUtensorMap to2;  // output stores: [hd columns x 32 rows] boxes
      const uint32_t box2[3] = {(uint32_t)HD, 32u, 1u};
      rc = make_tmap_bf16(&to2, out, 3, dims, str, box2, 2 * HD);  // swizzle span = row bytes (128 / 64 / 32) or none
      if (rc) return rc;
      auto k2 = variant == 2 ? (drop.thresh ? attn_fwd_w3_kernel<HD, true> : attn_fwd_w3_kernel<HD, false>)
                             : attn_fwd_w2_kernel<HD>;
      static bool configured2 = false;
      if (!configured2) {
        CGPT_CHECK(cudaFuncSetAttribute(attn_fwd_w2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::kDynamic));
        CGPT_CHECK(cudaFuncSetAttribute(attn_fwd_w3_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::kDynamic));
        CGPT_CHECK(cudaFuncSetAttribute(attn_fwd_w3_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::kDynamic));
        configured2 = true;
      }
      const int half_pairs = (n_pairs + 1) / 2;  // two item streams per CTA
      const int g2 = half_pairs < num_sms() ? half_pairs : num_sms();
      if (variant == 2)  // only the w3 kernels carry the pdl_wait() a programmatic launch needs
        CGPT_CHECK(launch_pdl(k2, dim3(g2), dim3(384), S2::kDynamic, st, 1, (long long)B * T, tm, to2, seg, lse, B, T, H, Hk, window,
                              scale * kLog2e, drop, (int)S2::kDynamic));
      else
        k2<<<g2, 384, S2::kDynamic, st>>>(tm, to2, seg, lse, B, T, H, Hk, window, scale * kLog2e, drop, S2::kDynamic);
      count_launch();
      CGPT_LAUNCH_CHECK();
      return 0;
    }
    auto kern = attn_fwd_ws_kernel<HD>;
    static bool configured = false;
    if (!configured) {
      CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SW::kDynamic));
      configured = true;
    }
    const int grid = n_pairs < num_sms() ? n_pairs : num_sms();
    kern<<<grid, 320, SW::kDynamic, st>>>(tm, to, seg, lse, B, T, H, Hk, window, scale * kLog2e, drop, SW::kDynamic);
  } else {
    using S = FwdSmem<HD>;
    auto kern = attn_fwd_kernel<HD>;
    static bool configured = false;
    if (!configured) {
      CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kDynamic));
      configured = true;
    }
    dim3 grid((T + BQ - 1) / BQ, H, B);
    kern<<<grid, 256, S::kDynamic, st>>>(tm, seg, reinterpret_cast<__nv_bfloat16*>(out), lse, T, H, Hk, window,
                                         scale * kLog2e, S::kDynamic, drop);
  }
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

template <int HD>
int launch_bwd(const void* qkv, const int32_t* seg, const void* out, const void* dout, const float* lse, void* dqkv,
               void* ws, float* colsum, int B, int T, int H, int Hk, int window, float scale, const DropoutCfg& drop,
               cudaStream_t st) {
  using S = BwdSmem<HD>;
  const int W = (H + 2 * Hk) * HD;
  CUtensorMap tq, td;
  int rc = make_qkv_tmap(&tq, qkv, B, T, W, HeadCfg<HD>::AW);
  if (rc) return rc;
  rc = make_qkv_tmap(&td, dout, B, T, H * HD, HeadCfg<HD>::AW);
  if (rc) return rc;
  float* delta = reinterpret_cast<float*>(ws);
  int* qhi_tab = reinterpret_cast<int*>(delta + delta_floats(B, T, H));
  float* dq_ws = delta + delta_floats(B, T, H) + qhi_floats(B, T);  // 256-byte aligned: bulk reduce-adds need 16
  CUtensorMap tdq;
  memset(&tdq, 0, sizeof(tdq));
  if (S::kTmaDq) {
    const uint64_t dims[3] = {(uint64_t)HD, (uint64_t)T, (uint64_t)B * H};
    const uint64_t str[2] = {(uint64_t)HD * 4, (uint64_t)T * HD * 4};
    const uint32_t box[3] = {(uint32_t)S::kdQBoxCols, 32u, 1u};
    rc = make_tmap_f32(&tdq, dq_ws, 3, dims, str, box, S::kdQBoxCols * 4);
    if (rc) return rc;
  }
  CUtensorMap tdkv;  // dK / dV stores of the warp-specialised kernel: [hd/2 columns x 32 rows] boxes of dqkv
  memset(&tdkv, 0, sizeof(tdkv));
  if constexpr (HD <= 64) {
    const uint64_t dims[3] = {(uint64_t)W, (uint64_t)T, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)W * 2, (uint64_t)T * W * 2};
    const uint32_t box[3] = {(uint32_t)(HD / 2), 32u, 1u};
    rc = make_tmap_bf16(&tdkv, dqkv, 3, dims, str, box, HD);  // swizzle span = row bytes (64 / 32) or none
    if (rc) return rc;
  }
  constexpr bool kDeltaVec = (HD / 8 == 2 || HD / 8 == 4 || HD / 8 == 8 || HD / 8 == 16);
  if (!kDeltaVec) CGPT_CHECK(cudaMemsetAsync(dq_ws, 0, (size_t)B * H * T * HD * sizeof(float), st));
  // grid of the warp-specialised kernel = initial value of its work counter (pairs 0..grid-1 are implicit)
  const int ws_pairs = (((T + BKV - 1) / BKV + 1) / 2) * Hk * B;
  const int ws_grid = ws_pairs < num_sms() ? ws_pairs : num_sms();
  {
    const __nv_bfloat16* po = reinterpret_cast<const __nv_bfloat16*>(out);
    const __nv_bfloat16* pd = reinterpret_cast<const __nv_bfloat16*>(dout);
    constexpr int G = HD / 8;
    if constexpr (G == 2 || G == 4 || G == 8 || G == 16) {
      const long long chunks = (long long)B * T * H * G;  // >= B * kv tiles, so the table job fits too
      CGPT_CHECK(launch_pdl(attn_delta_vec_kernel<G>, dim3((unsigned)((chunks + 255) / 256)), dim3(256), 0, st, 1, (long long)B * T, po, pd,
                            delta, B, T, H, seg, window, qhi_tab, ws_grid, reinterpret_cast<float4*>(dq_ws)));
    } else {
      const long long warps = (long long)B * T * H;
      attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(po, pd, delta, B, T, H, HD, seg, window, qhi_tab,
                                                                     ws_grid);
    }
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  const int n_items = ((T + BKV - 1) / BKV) * Hk * B;
  int grid = n_items < num_sms() ? n_items : num_sms();
  if constexpr (HD <= 64) {
    using SW = BwdWsSmem<HD>;
    grid = ws_grid;  // the kernel deals kv tiles in balanced pairs
    auto kern = attn_bwd_ws_kernel<HD>;
    static bool configured = false;
    if (!configured) {
      CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SW::kDynamic));
      configured = true;
    }
    CGPT_CHECK(launch_pdl(kern, dim3(grid), dim3(288), SW::kDynamic, st, 1, (long long)B * T, tq, td, tdq, tdkv, seg, qhi_tab, lse, delta,
                          reinterpret_cast<__nv_bfloat16*>(dqkv), dq_ws, colsum, B, T, H, Hk, window, scale, drop,
                          (int)SW::kDynamic));
  } else {
    auto kern = attn_bwd_kernel<HD>;
    static bool configured = false;
    if (!configured) {
      CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kDynamic));
      configured = true;
    }
    kern<<<grid, 256, S::kDynamic, st>>>(tq, td, tdq, seg, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv), dq_ws, B, T,
                                         H, Hk, window, scale, drop);
  }
  count_launch();
  CGPT_LAUNCH_CHECK();
  {
    // hd <= 64: the k / v column sums came out of the main kernel's epilogue, the q part comes out of this one
    float* qsum = (HD <= 64) ? colsum : nullptr;
    CGPT_CHECK(launch_pdl(attn_dq_convert_kernel, dim3((T + 255) / 256, B * H), dim3(256), 0, st, 1, (long long)B * T, dq_ws,
                          reinterpret_cast<__nv_bfloat16*>(dqkv), T, H, HD, W, scale, qsum));
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  if (colsum != nullptr && HD > 64) return cgpt_colsum_bf16(dqkv, W, colsum, B * T, W, reinterpret_cast<cgpt_stream_t>(st));
  return 0;
}

int check_attn_args(const char* who, int B, int T, int H, int Hk, int hd) {
  CGPT_REQUIRE(B > 0 && T > 0 && H > 0 && Hk > 0, "%s: bad sizes B=%d T=%d H=%d Hk=%d", who, B, T, H, Hk);
  CGPT_REQUIRE(H % Hk == 0, "%s: n_head=%d must be divisible by n_kv_head=%d", who, H, Hk);
  CGPT_REQUIRE(hd == 16 || hd == 32 || hd == 48 || hd == 64 || hd == 96 || hd == 128,
               "%s: head_dim=%d not supported (16, 32, 48, 64, 96, 128)", who, hd);
  return 0;
}

}  // namespace
}  // namespace cgpt

using namespace cgpt;

#define DISPATCH_HD(hd, CALL)        \
  switch (hd) {                      \
    case 16: return CALL(16);        \
    case 32: return CALL(32);        \
    case 48: return CALL(48);        \
    case 64: return CALL(64);        \
    case 96: return CALL(96);        \
    default: return CALL(128);       \
  }

extern "C" {

int cgpt_attn_fwd(const void* qkv, const int32_t* seg, void* out, float* lse, int B, int T, int H, int Hk, int hd,
                  int window, float scale, float dropout_p, uint64_t seed, uint64_t offset, cgpt_stream_t stream) {
  CGPT_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_fwd: dropout_p=%f must be in [0,1)", dropout_p);
  const DropoutCfg drop = make_dropout(dropout_p, seed, offset);
  CGPT_REQUIRE(qkv && out && lse, "attn_fwd: null pointer");
  int rc = check_attn_args("attn_fwd", B, T, H, Hk, hd);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CALL(HD) launch_fwd<HD>(qkv, seg, out, lse, B, T, H, Hk, window, scale, drop, st)
  DISPATCH_HD(hd, CALL)
#undef CALL
}

int64_t cgpt_attn_bwd_workspace(int B, int T, int H, int Hk, int hd) {
  (void)Hk;
  return (int64_t)sizeof(float) *
         ((int64_t)cgpt::delta_floats(B, T, H) + (int64_t)cgpt::qhi_floats(B, T) + (int64_t)B * H * T * hd);
}

int cgpt_attn_bwd(const void* qkv, const int32_t* seg, const void* out, const void* dout, const float* lse, void* dqkv,
                  void* ws, int B, int T, int H, int Hk, int hd, int window, float scale, float dropout_p, uint64_t seed,
                  uint64_t offset, cgpt_stream_t stream) {
  return cgpt_attn_bwd_colsum(qkv, seg, out, dout, lse, dqkv, ws, nullptr, B, T, H, Hk, hd, window, scale, dropout_p, seed,
                              offset, stream);
}

int cgpt_attn_bwd_colsum(const void* qkv, const int32_t* seg, const void* out, const void* dout, const float* lse,
                         void* dqkv, void* ws, float* dqkv_colsum, int B, int T, int H, int Hk, int hd, int window,
                         float scale, float dropout_p, uint64_t seed, uint64_t offset, cgpt_stream_t stream) {
  CGPT_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_bwd: dropout_p=%f must be in [0,1)", dropout_p);
  const DropoutCfg drop = make_dropout(dropout_p, seed, offset);
  CGPT_REQUIRE(qkv && out && dout && lse && dqkv && ws, "attn_bwd: null pointer");
  CGPT_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "attn_bwd: workspace must be 256-byte aligned");
  int rc = check_attn_args("attn_bwd", B, T, H, Hk, hd);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CALL(HD) launch_bwd<HD>(qkv, seg, out, dout, lse, dqkv, ws, dqkv_colsum, B, T, H, Hk, window, scale, drop, st)
  DISPATCH_HD(hd, CALL)
#undef CALL
}

int cgpt_attn_probs(const void* qkv, const int32_t* seg, float* att, int B, int T, int H, int Hk, int hd, int window,
                    float scale, float dropout_p, uint64_t seed, uint64_t offset, cgpt_stream_t stream) {
  const DropoutCfg drop = make_dropout(dropout_p, seed, offset);
  CGPT_REQUIRE(qkv && att, "attn_probs: null pointer");
  CGPT_REQUIRE(B > 0 && T > 0 && H > 0 && Hk > 0 && H % Hk == 0 && hd % 2 == 0, "attn_probs: bad sizes");
  const long long warps = (long long)B * H * T;
  attn_probs_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), seg, att, B, T, H, Hk, hd, window, scale, drop);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
