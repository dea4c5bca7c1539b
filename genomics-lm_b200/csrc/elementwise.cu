// Memory-bound kernels of the codon-GPT step: integer scans, embedding gather / scatter-add,
// LayerNorm fwd/bwd, casts, column sums, RoPE, SwiGLU gate, AdamW.  All HBM-bound: vectorised,
// coalesced accesses, warp-shuffle reductions, grids sized in multiples of the SM count.
#include <stdlib.h>

#include "common.cuh"

namespace cgpt {
namespace {

// ============================================================ integer scans (one warp per row)
__global__ void segment_ids_kernel(const int64_t* __restrict__ idx, int32_t* __restrict__ seg, int B, int T,
                                   int sep_id) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int64_t* r = idx + (size_t)row * T;
  int32_t* o = seg + (size_t)row * T;
  int running = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool f = (t < T) && (r[t] == (int64_t)sep_id);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int incl = __popc(m & (0xffffffffu >> (31 - lane)));
    if (t < T) o[t] = running + incl;
    running += __popc(m);
  }
}

__global__ void segment_starts_kernel(const int64_t* __restrict__ idx, int32_t* __restrict__ start, int B, int T,
                                      int sep_id) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int64_t* r = idx + (size_t)row * T;
  int32_t* o = start + (size_t)row * T;
  int carry = 0;  // no separator yet: the segment starts at 0
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool f = (t < T) && (r[t] == (int64_t)sep_id);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const unsigned at_or_before = m & (0xffffffffu >> (31 - lane));
    const int mine = at_or_before ? (t0 + 31 - __clz(at_or_before)) : carry;
    if (t < T) o[t] = mine;
    if (m) carry = t0 + 31 - __clz(m);
  }
}

struct IdSet {
  int64_t v[8];
  int n;
};

__global__ void next_in_set_kernel(const int64_t* __restrict__ yb, int32_t* __restrict__ next, int B, int T,
                                   IdSet ids) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int64_t* r = yb + (size_t)row * T;
  int32_t* o = next + (size_t)row * T;
  int carry = T;  // sentinel: nothing at or after
  for (int t0 = ((T - 1) / 32) * 32; t0 >= 0; t0 -= 32) {
    const int t = t0 + lane;
    bool f = false;
    if (t < T) {
      const int64_t tok = r[t];
      for (int i = 0; i < ids.n; ++i) f |= (tok == ids.v[i]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const unsigned at_or_after = m >> lane;
    const int mine = at_or_after ? (t + __ffs(at_or_after) - 1) : carry;
    if (t < T) o[t] = mine;
    if (m) carry = t0 + __ffs(m) - 1;
  }
}

struct EdgeSet {
  int64_t v[8];
  int n;
};

__global__ void termination_labels_kernel(const int64_t* __restrict__ yb, const int32_t* __restrict__ next_stop,
                                          int64_t* __restrict__ labels, int B, int T, EdgeSet e, int64_t ignore) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T) return;
  const int t = (int)(i % T);
  int64_t lab;
  if (yb[i] == 0) {
    lab = ignore;
  } else if (next_stop[i] >= T) {
    lab = e.n;
  } else {
    const int64_t dist = next_stop[i] - t;
    int c = 0;
    for (int k = 0; k < e.n; ++k) c += (dist > e.v[k]);
    lab = c;
  }
  labels[i] = lab;
}

// ============================================================ embedding
__global__ void embed_fwd_kernel(const int64_t* __restrict__ idx, const float4* __restrict__ tok,
                                 const float4* __restrict__ pos, float4* __restrict__ x, size_t n4, int T, int d4,
                                 int vocab) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / d4;
    const int c = (int)(i - m * d4);
    const int64_t v = idx[m];
    if (v < 0 || v >= vocab) {  // F.embedding raises for such an id; here the row is NaN: loud downstream, no OOB read
      const float q = __int_as_float(0x7fc00000);
      x[i] = make_float4(q, q, q, q);
      continue;
    }
    float4 a = __ldg(tok + (size_t)v * d4 + c);
    if (pos) {
      const float4 p = __ldg(pos + (size_t)(m % T) * d4 + c);
      a.x += p.x;
      a.y += p.y;
      a.z += p.z;
      a.w += p.w;
    }
    x[i] = a;
  }
}

// Token-embedding gradient: V (~68) rows receive M updates.  Each CTA owns a token range and a
// column slice, accumulates a private [V, cols] table in shared memory with plain adds (threads own
// columns, tokens are walked serially -> no atomics, no contention), then flushes with one atomic per
// table element.
__global__ void embed_bwd_tok_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dx,
                                     float* __restrict__ dtok, int M, int d, int vocab, int cols_per_cta,
                                     int toks_per_cta) {
  extern __shared__ float table[];  // [vocab][cols_per_cta]
  const int c0 = blockIdx.y * cols_per_cta;
  const int ncols = min(cols_per_cta, d - c0);
  const int m0 = blockIdx.x * toks_per_cta;
  const int m1 = min(M, m0 + toks_per_cta);
  for (int i = threadIdx.x; i < vocab * cols_per_cta; i += blockDim.x) table[i] = 0.f;
  __syncthreads();
  for (int m = m0; m < m1; ++m) {
    int64_t v = idx[m];
    v = v < 0 ? 0 : (v >= vocab ? vocab - 1 : v);
    const float* src = dx + (size_t)m * d + c0;
    float* dst = table + (size_t)v * cols_per_cta;
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) dst[c] += src[c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < vocab * cols_per_cta; i += blockDim.x) {
    const int v = i / cols_per_cta, c = i - v * cols_per_cta;
    const float val = table[i];
    if (c < ncols && val != 0.f) atomicAdd(dtok + (size_t)v * d + c0 + c, val);
  }
}

// Small vocabularies (the codon table: ~68 rows): a CTA owns 32 columns and a token range; each of its 8 warps
// keeps a PRIVATE [vocab, 32] table in shared memory (plain read-modify-write: a lane only ever touches its own
// column, so no atomics and no contention however skewed the token histogram is) and walks every 8th token with
// kUnroll tokens in flight, so the loop is bound by HBM rather than by one dependent load chain per token.
constexpr int kEmbTokWarps = 8, kEmbTokUnroll = 8;
__global__ void __launch_bounds__(kEmbTokWarps * 32)
embed_bwd_tok_small_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dx, float* __restrict__ dtok,
                           int M, int d, int vocab, int toks_per_cta) {
  extern __shared__ float table[];  // [warps][vocab][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const bool col_ok = c < d;
  const int m0 = blockIdx.x * toks_per_cta;
  const int m1 = min(M, m0 + toks_per_cta);
  for (int i = threadIdx.x; i < kEmbTokWarps * vocab * 32; i += blockDim.x) table[i] = 0.f;
  __syncthreads();
  float* mine = table + (size_t)warp * vocab * 32 + lane;
  for (int m = m0 + warp; m < m1; m += kEmbTokWarps * kEmbTokUnroll) {
    int v[kEmbTokUnroll];
    float x[kEmbTokUnroll];
#pragma unroll
    for (int u = 0; u < kEmbTokUnroll; ++u) {
      const int mm = m + u * kEmbTokWarps;
      int64_t t = mm < m1 ? idx[mm] : 0;
      v[u] = static_cast<int>(t < 0 ? 0 : (t >= vocab ? vocab - 1 : t));
      x[u] = (mm < m1 && col_ok) ? __ldg(dx + (size_t)mm * d + c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kEmbTokUnroll; ++u) mine[v[u] * 32] += x[u];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < vocab * 32; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kEmbTokWarps; ++w) acc += table[(size_t)w * vocab * 32 + i];
    const int row = i >> 5, cc = blockIdx.y * 32 + (i & 31);
    if (cc < d && acc != 0.f) atomicAdd(dtok + (size_t)row * d + cc, acc);
  }
}

__global__ void embed_bwd_pos_kernel(const float4* __restrict__ dx, float4* __restrict__ dpos, int B, int T,
                                     int d4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)T * d4) return;
  float4 acc = dpos[i];
  for (int b = 0; b < B; ++b) {
    const float4 v = __ldg(dx + (size_t)b * T * d4 + i);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  dpos[i] = acc;
}

// ============================================================ LayerNorm (one warp per row, d <= 1024, d % 4 == 0)
// VPT = float4 per lane (compile time): lane l owns columns 4*(l + 32k), k < VPT, for every row it visits,
// so gamma/beta and the dgamma/dbeta partial sums live in registers across the whole row loop.
constexpr int kLnMaxVec = 8;

template <int VPT>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ yb, float* __restrict__ yf, float* __restrict__ mean,
                     float* __restrict__ rstd, int M, int d, float eps, int rev) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = d >> 2;
  float4 g[VPT], bt[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int c = lane + 32 * k;
    g[k] = c < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[k] = c < nvec ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_d = 1.f / d;
  for (int rowi = blockIdx.x * wpb + (threadIdx.x >> 5); rowi < M; rowi += gridDim.x * wpb) {
    const int row = rev ? M - 1 - rowi : rowi;  // rev: last rows first (the producer's newest lines are still in L2)
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    float4 v[VPT];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      v[k] = c < nvec ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      if (lane + 32 * k < nvec) {
        const float a = v[k].x - mu, b = v[k].y - mu, e = v[k].z - mu, f = v[k].w - mu;
        q += (a * a + b * b) + (e * e + f * f);
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        float4 o;
        o.x = (v[k].x - mu) * rs * g[k].x + bt[k].x;
        o.y = (v[k].y - mu) * rs * g[k].y + bt[k].y;
        o.z = (v[k].z - mu) * rs * g[k].z + bt[k].z;
        o.w = (v[k].w - mu) * rs * g[k].w + bt[k].w;
        if (yf) reinterpret_cast<float4*>(yf + (size_t)row * d)[c] = o;
        if (yb) reinterpret_cast<uint2*>(yb + (size_t)row * d)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

template <int VPT, bool DY_F32>
__global__ void __launch_bounds__(256, (VPT <= 4) ? 2 : 1)
layernorm_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dxsum, int M, int d, int rev) {
  __shared__ float red[8][VPT * 128];
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nvec = d >> 2;
  const float inv_d = 1.f / d;
  float4 gm[VPT], ag[VPT], ab[VPT], ax[VPT];  // ax: column sums of dx (bias gradient of the upstream linear)
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int c = lane + 32 * k;
    gm[k] = c < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[k] = ab[k] = ax[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int rowi = blockIdx.x * wpb + warp; rowi < M; rowi += gridDim.x * wpb) {
    const int row = rev ? M - 1 - rowi : rowi;
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    float4 xh[VPT], dyv[VPT], rv[VPT];
    float s1 = 0.f, s2 = 0.f;
    // every load of the row is issued before the first use: one HBM round trip per row, not two
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      rv[k] = (dres && c < nvec) ? __ldg(reinterpret_cast<const float4*>(dres + (size_t)row * d) + c)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        if constexpr (DY_F32) {
          dyv[k] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * d)[c];
        } else {
          const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + (size_t)row * d)[c];
          const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
          dyv[k] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        const float4 xv = xr[c];
        xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      } else {
        dyv[k] = xh[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const float gx = dyv[k].x * gm[k].x, gy = dyv[k].y * gm[k].y, gz = dyv[k].z * gm[k].z, gw = dyv[k].w * gm[k].w;
      s1 += (gx + gy) + (gz + gw);
      s2 += (gx * xh[k].x + gy * xh[k].y) + (gz * xh[k].z + gw * xh[k].w);
      ag[k].x += dyv[k].x * xh[k].x;
      ag[k].y += dyv[k].y * xh[k].y;
      ag[k].z += dyv[k].z * xh[k].z;
      ag[k].w += dyv[k].w * xh[k].w;
      ab[k].x += dyv[k].x;
      ab[k].y += dyv[k].y;
      ab[k].z += dyv[k].z;
      ab[k].w += dyv[k].w;
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        float4 o;
        o.x = rs * (dyv[k].x * gm[k].x - m1 - xh[k].x * m2);
        o.y = rs * (dyv[k].y * gm[k].y - m1 - xh[k].y * m2);
        o.z = rs * (dyv[k].z * gm[k].z - m1 - xh[k].z * m2);
        o.w = rs * (dyv[k].w * gm[k].w - m1 - xh[k].w * m2);
        o.x += rv[k].x;
        o.y += rv[k].y;
        o.z += rv[k].z;
        o.w += rv[k].w;
        reinterpret_cast<float4*>(dx + (size_t)row * d)[c] = o;
        if (dxb)
          reinterpret_cast<uint2*>(dxb + (size_t)row * d)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (dxsum) {
          ax[k].x += o.x;
          ax[k].y += o.y;
          ax[k].z += o.z;
          ax[k].w += o.w;
        }
      }
    }
  }
  // per-CTA reduction of the column partials over its 8 warps, then one atomic per column
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 2 && dxsum == nullptr) break;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const float4 val = pass == 0 ? ag[k] : (pass == 1 ? ab[k] : ax[k]);
      *reinterpret_cast<float4*>(&red[warp][(lane + 32 * k) * 4]) = val;
    }
    __syncthreads();
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) t += red[w2][c];
      atomicAdd(dst + c, t);
    }
  }
}

// ---- streaming variants (d = 128 * VPT exactly, 16-byte aligned rows): every warp owns a ring of row slots in shared
// memory that the bulk-copy engine fills (cp.async.bulk, one mbarrier per slot), so whole rows stay in flight while the
// warp computes on the previous one — the per-thread-load kernels above expose one HBM round trip per row and warp
// (their loads are only in flight between the issue and the first use).  Same arithmetic, same column ownership.
constexpr int kLnWarps = 8;

template <int VPT, int SLOTS>
struct LnFwdStream {
  static constexpr int kRowBytes = VPT * 128 * 4;
  static constexpr int kBarOff = kLnWarps * SLOTS * kRowBytes;
  static constexpr int kSmem = kBarOff + kLnWarps * SLOTS * 8;
};

template <int VPT, int SLOTS>
__global__ void __launch_bounds__(kLnWarps * 32, 3)
layernorm_fwd_stream_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                            __nv_bfloat16* __restrict__ yb, float* __restrict__ yf, float* __restrict__ mean,
                            float* __restrict__ rstd, int M, float eps, int rev) {
  using L = LnFwdStream<VPT, SLOTS>;
  constexpr int D = VPT * 128;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31;
  // warp index through a shuffle (the compiler then keeps row indices and ring addresses in uniform registers) and an
  // elected issuing lane: the bulk copies are issued without an R2UR ... BRA.U.ANY loop around each of them
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const bool leader = elect_one();
  uint8_t* ring = ln_smem + warp * SLOTS * L::kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + L::kBarOff) + warp * SLOTS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncwarp();
  pdl_wait();
  const int stride = gridDim.x * kLnWarps;
  const int row0 = blockIdx.x * kLnWarps + warp;
  if (leader) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const long long r = row0 + (long long)s * stride;
      if (r < M) {
        mbar_expect_tx(&bars[s], L::kRowBytes);
        bulk_load_1d(smem_u32(ring + s * L::kRowBytes), x + (size_t)(rev ? M - 1 - r : r) * D, L::kRowBytes, &bars[s]);
      }
    }
  }
  float4 g[VPT], bt[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    bt[k] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * k);
  }
  constexpr float inv_d = 1.f / D;
  int it = 0;
  for (long long rowi = row0; rowi < M; rowi += stride, ++it) {
    const long long row = rev ? M - 1 - rowi : rowi;
    const int slot = it % SLOTS;
    mbar_wait(&bars[slot], (it / SLOTS) & 1);
    const float4* xr = reinterpret_cast<const float4*>(ring + slot * L::kRowBytes);
    float4 v[VPT];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      v[k] = xr[lane + 32 * k];
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    __syncwarp();  // every lane has its part of the row in registers: the slot can be refilled
    if (leader) {
      const long long rn = rowi + (long long)SLOTS * stride;
      if (rn < M) {
        mbar_expect_tx(&bars[slot], L::kRowBytes);
        bulk_load_1d(smem_u32(ring + slot * L::kRowBytes), x + (size_t)(rev ? M - 1 - rn : rn) * D, L::kRowBytes,
                     &bars[slot]);
      }
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const float a = v[k].x - mu, b = v[k].y - mu, e = v[k].z - mu, f = v[k].w - mu;
      q += (a * a + b * b) + (e * e + f * f);
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      float4 o;
      o.x = (v[k].x - mu) * rs * g[k].x + bt[k].x;
      o.y = (v[k].y - mu) * rs * g[k].y + bt[k].y;
      o.z = (v[k].z - mu) * rs * g[k].z + bt[k].z;
      o.w = (v[k].w - mu) * rs * g[k].w + bt[k].w;
      if (yf) reinterpret_cast<float4*>(yf + (size_t)row * D)[c] = o;
      if (yb) reinterpret_cast<uint2*>(yb + (size_t)row * D)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

template <int VPT, int SLOTS>
struct LnBwdStream {
  static constexpr int D = VPT * 128;
  static constexpr int kXOff = 0, kResOff = D * 4, kDyOff = D * 8;
  static constexpr int kRowBytes = D * 10;  // x fp32 | dres fp32 | dy bf16
  static constexpr int kBarOff = kLnWarps * SLOTS * kRowBytes;
  static constexpr int kSmem = kBarOff + kLnWarps * SLOTS * 8;
  static_assert(kLnWarps * SLOTS * kRowBytes >= kLnWarps * D * 4, "the column reduction reuses the ring");
};

template <int VPT, int SLOTS>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
layernorm_bwd_stream_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                            const float* __restrict__ gamma, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const float* __restrict__ dres, float* __restrict__ dx,
                            __nv_bfloat16* __restrict__ dxb, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            float* __restrict__ dxsum, int M, int rev) {
  using L = LnBwdStream<VPT, SLOTS>;
  constexpr int D = L::D;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31;
  // warp index through a shuffle (the compiler then keeps row indices and ring addresses in uniform registers) and an
  // elected issuing lane: the bulk copies are issued without an R2UR ... BRA.U.ANY loop around each of them
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const bool leader = elect_one();
  uint8_t* ring = ln_smem + warp * SLOTS * L::kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + L::kBarOff) + warp * SLOTS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncwarp();
  pdl_wait();
  const int stride = gridDim.x * kLnWarps;
  const int row0 = blockIdx.x * kLnWarps + warp;
  const uint32_t row_tx = dres ? L::kRowBytes : D * 6;
  auto issue = [&](int slot, long long ri) {  // lane 0 only
    const long long r = rev ? M - 1 - ri : ri;
    uint8_t* sb = ring + slot * L::kRowBytes;
    mbar_expect_tx(&bars[slot], row_tx);
    bulk_load_1d(smem_u32(sb + L::kXOff), x + (size_t)r * D, D * 4, &bars[slot]);
    bulk_load_1d(smem_u32(sb + L::kDyOff), dy + (size_t)r * D, D * 2, &bars[slot]);
    if (dres) bulk_load_1d(smem_u32(sb + L::kResOff), dres + (size_t)r * D, D * 4, &bars[slot]);
  };
  if (leader) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const long long r = row0 + (long long)s * stride;
      if (r < M) issue(s, r);
    }
  }
  constexpr float inv_d = 1.f / D;
  float4 gm[VPT], ag[VPT], ab[VPT], ax[VPT];  // ax: column sums of dx (bias gradient of the upstream linear)
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    gm[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    ag[k] = ab[k] = ax[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mu_n = 0.f, rs_n = 0.f;
  if (row0 < M) {
    mu_n = mean[rev ? M - 1 - row0 : row0];
    rs_n = rstd[rev ? M - 1 - row0 : row0];
  }
  int it = 0;
  for (long long rowi = row0; rowi < M; rowi += stride, ++it) {
    const long long row = rev ? M - 1 - rowi : rowi;
    const int slot = it % SLOTS;
    const float mu = mu_n, rs = rs_n;
    if (rowi + stride < M) {  // the next row's statistics travel under this row's math
      const long long rn = rev ? M - 1 - (rowi + stride) : rowi + stride;
      mu_n = mean[rn];
      rs_n = rstd[rn];
    }
    mbar_wait(&bars[slot], (it / SLOTS) & 1);
    const uint8_t* sb = ring + slot * L::kRowBytes;
    const float4* xr = reinterpret_cast<const float4*>(sb + L::kXOff);
    const uint2* dyr = reinterpret_cast<const uint2*>(sb + L::kDyOff);
    const float4* rr = reinterpret_cast<const float4*>(sb + L::kResOff);
    float4 xh[VPT], dyv[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const uint2 u = dyr[lane + 32 * k];
      const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
      dyv[k] = make_float4(lo.x, lo.y, hi.x, hi.y);
      const float4 xv = xr[lane + 32 * k];
      xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const float gx = dyv[k].x * gm[k].x, gy = dyv[k].y * gm[k].y, gz = dyv[k].z * gm[k].z, gw = dyv[k].w * gm[k].w;
      s1 += (gx + gy) + (gz + gw);
      s2 += (gx * xh[k].x + gy * xh[k].y) + (gz * xh[k].z + gw * xh[k].w);
      ag[k].x += dyv[k].x * xh[k].x;
      ag[k].y += dyv[k].y * xh[k].y;
      ag[k].z += dyv[k].z * xh[k].z;
      ag[k].w += dyv[k].w * xh[k].w;
      ab[k].x += dyv[k].x;
      ab[k].y += dyv[k].y;
      ab[k].z += dyv[k].z;
      ab[k].w += dyv[k].w;
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = lane + 32 * k;
      float4 o;
      o.x = rs * (dyv[k].x * gm[k].x - m1 - xh[k].x * m2);
      o.y = rs * (dyv[k].y * gm[k].y - m1 - xh[k].y * m2);
      o.z = rs * (dyv[k].z * gm[k].z - m1 - xh[k].z * m2);
      o.w = rs * (dyv[k].w * gm[k].w - m1 - xh[k].w * m2);
      if (dres) {
        const float4 rv = rr[c];
        o.x += rv.x;
        o.y += rv.y;
        o.z += rv.z;
        o.w += rv.w;
      }
      reinterpret_cast<float4*>(dx + (size_t)row * D)[c] = o;
      if (dxb) reinterpret_cast<uint2*>(dxb + (size_t)row * D)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      if (dxsum) {
        ax[k].x += o.x;
        ax[k].y += o.y;
        ax[k].z += o.z;
        ax[k].w += o.w;
      }
    }
    __syncwarp();  // every lane is done with the slot: refill it with the row SLOTS iterations ahead
    if (leader) {
      const long long rn = rowi + (long long)SLOTS * stride;
      if (rn < M) issue(slot, rn);
    }
  }
  // per-CTA reduction of the column partials over its 8 warps (the ring is idle now: every copy that was issued has
  // been waited for), then one atomic per column
  float* red = reinterpret_cast<float*>(ln_smem);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 2 && dxsum == nullptr) break;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const float4 val = pass == 0 ? ag[k] : (pass == 1 ? ab[k] : ax[k]);
      *reinterpret_cast<float4*>(&red[warp * D + (lane + 32 * k) * 4]) = val;
    }
    __syncthreads();
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kLnWarps; ++w2) t += red[w2 * D + c];
      atomicAdd(dst + c, t);
    }
  }
}

// ============================================================ casts / column sums
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                                     long long ld_out, long long rows, long long cols) {
  const long long total = rows * ld_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out, c = i - r * ld_out;
    out[i] = __float2bfloat16_rn(c < cols ? in[r * ld_in + c] : 0.f);
  }
}

__global__ void cast_f32_bf16_vec_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

__global__ void split3_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                              long long rows, long long cols, long long cols_pad, int partner) {
  const long long total = rows * cols_pad;
  const long long ld_out = 3 * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad, c = i - r * cols_pad;
    const float v = c < cols ? in[r * ld_in + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * ld_out + c;
    o[0] = hi;
    o[cols_pad] = partner ? hi : lo;
    o[2 * cols_pad] = partner ? lo : hi;
  }
}

// dst[r,c] += S[r,c] + S[r,c+co] + S[r+ro,c] + S[r+ro,c+co]: the four hi/lo cross products of a stacked split-head
// weight-gradient GEMM ([g_hi|g_lo]^T [x_hi|x_lo]) folded into the fp32 gradient.
__global__ void fold_quadrants_add_kernel(const float* __restrict__ s, long long lds, float* __restrict__ dst,
                                          long long ldd, int rows, int cols, int ro, int co) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const float* a = s + r * lds + c;
    const float* b = s + (r + ro) * lds + c;
    dst[r * ldd + c] += (a[0] + a[co]) + (b[0] + b[co]);
  }
}

// ---------------------------------------------------------------------------------------- incremental decode
// One new token per sequence against a (K, V) cache: out[b,h,:] = softmax(q·K[lo..t]ᵀ·scale)·V[lo..t], where the
// row at position t is the new token's own (k, v) (taken from qkv_new and appended to the cache by the first head of
// each kv group).  Memory-bound: every cached K and V row is read once per query head.  CTA = one (b, h), 128 threads.
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv_new, __nv_bfloat16* __restrict__ k_cache,
                   __nv_bfloat16* __restrict__ v_cache, const int32_t* __restrict__ lo, __nv_bfloat16* __restrict__ out,
                   const int32_t* __restrict__ t_dev, int t_host, int Tmax, int H, int Hk, int hd, int window,
                   float scale_log2) {
  extern __shared__ float dsm[];
  const int t = t_dev ? min(*t_dev, Tmax - 1) : t_host;  // position of the new token (device-resident under graph replay)
  float* q_s = dsm;                  // [hd]
  float* red = dsm + hd;             // [128 / (hd/8)] x hd partial outputs, also the block reductions
  float* sc = red + 128 * 8;         // [t + 1 - jlo] scores / probabilities
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int rep = H / Hk, kvh = h / rep;
  const int W = (H + 2 * Hk) * hd, KW = Hk * hd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* qrow = qkv_new + (size_t)b * W + h * hd;
  const __nv_bfloat16* knew = qkv_new + (size_t)b * W + (H + kvh) * hd;
  const __nv_bfloat16* vnew = qkv_new + (size_t)b * W + (H + Hk + kvh) * hd;
  const __nv_bfloat16* kc = k_cache + (size_t)b * Tmax * KW + kvh * hd;
  const __nv_bfloat16* vc = v_cache + (size_t)b * Tmax * KW + kvh * hd;
  int jlo = lo ? lo[b] : 0;
  if (window > 0) jlo = max(jlo, t - window + 1);
  for (int c = tid; c < hd; c += 128) q_s[c] = __bfloat162float(qrow[c]) * scale_log2;
  __syncthreads();
  const int n = t + 1 - jlo, nv = hd / 8;
  float mx = -INFINITY;
  for (int r = tid; r < n; r += 128) {
    const int j = jlo + r;
    const uint4* kr = reinterpret_cast<const uint4*>(j == t ? knew : kc + (size_t)j * KW);
    float s = 0.f;
    for (int v8 = 0; v8 < nv; ++v8) {
      const uint4 u = kr[v8];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s = fmaf(q_s[8 * v8 + 2 * e], __uint_as_float(w[e] << 16), s);
        s = fmaf(q_s[8 * v8 + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u), s);
      }
    }
    sc[r] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int r = tid; r < n; r += 128) {
    const float p = exp2f(sc[r] - mx);
    sc[r] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  __syncthreads();
  // P·V: a thread owns 8 output columns (one 16-byte piece of a V row) of every `rows`-th cached position
  const int rows = 128 / nv, rg = tid / nv, dl = tid - rg * nv;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rg < rows) {
    for (int r = rg; r < n; r += rows) {
      const int j = jlo + r;
      const uint4 u = reinterpret_cast<const uint4*>(j == t ? vnew : vc + (size_t)j * KW)[dl];
      const float p = sc[r];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(p, __uint_as_float(w[e] << 16), acc[2 * e]);
        acc[2 * e + 1] = fmaf(p, __uint_as_float(w[e] & 0xffff0000u), acc[2 * e + 1]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[rg * hd + dl * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int c = tid; c < hd; c += 128) {
    float o = 0.f;
    for (int g = 0; g < rows; ++g) o += red[g * hd + c];
    out[(size_t)b * H * hd + h * hd + c] = __float2bfloat16_rn(o * inv);
  }
  // append the new (k, v) to the cache (once per kv group; no CTA reads position t from the cache)
  if (h == kvh * rep && tid < 2 * nv) {
    const bool is_v = tid >= nv;
    const int piece = is_v ? tid - nv : tid;
    uint4* dst = reinterpret_cast<uint4*>((is_v ? v_cache : k_cache) + ((size_t)b * Tmax + t) * KW + kvh * hd);
    dst[piece] = reinterpret_cast<const uint4*>(is_v ? vnew : knew)[piece];
  }
}

// ---------------------------------------------------------------------------------------- token feed
// (xb, yb) of one micro-batch gathered on the device from a resident packed token array:
// xb[r, c] = seq_r[c], yb[r, c] = seq_r[c + 1] for c < len_r - 1, PAD (0) elsewhere.
__global__ void pack_lm_batch_kernel(const int32_t* __restrict__ tokens, const long long* __restrict__ offsets,
                                     const long long* __restrict__ lengths, const long long* __restrict__ indices,
                                     int B, int T_out, long long* __restrict__ xb, long long* __restrict__ yb) {
  const long long total = (long long)B * T_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / T_out), c = (int)(i - (long long)r * T_out);
    const long long s = indices[r];
    const long long usable = lengths[s] - 1;
    long long x = 0, y = 0;
    if (c < usable) {
      const int32_t* seq = tokens + offsets[s];
      x = seq[c];
      y = seq[c + 1];
    }
    xb[i] = x;
    yb[i] = y;
  }
}

// ---------------------------------------------------------------------------------------- shape guidance
// out[m, c] = x[m, c] + b[c] + sum_k w[c, k] * s[m, k]   (nn.Linear(3, d) on the DNA-shape features, added to the
// embedding: model_tiny_gpt.py:226-229, 310-311).  One thread per 4 columns.
__global__ void shape_proj_fwd_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                      const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ out, long long M, int d) {
  const int d4 = d >> 2;
  const long long total = M * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / d4;
    const int c = (int)(i - m * d4) * 4;
    const float s0 = s[m * 3], s1 = s[m * 3 + 1], s2 = s[m * 3 + 2];
    float4 v = *reinterpret_cast<const float4*>(x + m * d + c);
    float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* wr = w + (size_t)(c + e) * 3;
      pv[e] += b[c + e] + wr[0] * s0 + wr[1] * s1 + wr[2] * s2;
    }
    *reinterpret_cast<float4*>(out + m * d + c) = v;
  }
}

// dw[c, k] += sum_m dx[m, c] s[m, k], db[c] += sum_m dx[m, c].  CTA = 64 columns x a row range, 4 row lanes.
__global__ void __launch_bounds__(256)
shape_proj_wgrad_kernel(const float* __restrict__ dx, const float* __restrict__ s, float* __restrict__ dw,
                        float* __restrict__ db, long long M, int d, int rows_per_cta) {
  __shared__ float red[4][64][4];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
  if (c < d) {
    for (long long m = r0 + rl; m < r1; m += 4) {
      const float g = dx[m * d + c];
      a0 = fmaf(g, s[m * 3], a0);
      a1 = fmaf(g, s[m * 3 + 1], a1);
      a2 = fmaf(g, s[m * 3 + 2], a2);
      ab += g;
    }
  }
  red[rl][cl][0] = a0;
  red[rl][cl][1] = a1;
  red[rl][cl][2] = a2;
  red[rl][cl][3] = ab;
  __syncthreads();
  if (rl == 0 && c < d) {
    float t[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = red[0][cl][e] + red[1][cl][e] + red[2][cl][e] + red[3][cl][e];
    atomicAdd(dw + (size_t)c * 3, t[0]);
    atomicAdd(dw + (size_t)c * 3 + 1, t[1]);
    atomicAdd(dw + (size_t)c * 3 + 2, t[2]);
    atomicAdd(db + c, t[3]);
  }
}

// ds[m, k] = sum_c dx[m, c] w[c, k]: one warp per row (only when the shape encoder is trained, loop.py:695)
__global__ void shape_proj_dgrad_kernel(const float* __restrict__ dx, const float* __restrict__ w,
                                        float* __restrict__ ds, long long M, int d) {
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float g = dx[m * d + c];
    a0 = fmaf(g, w[(size_t)c * 3], a0);
    a1 = fmaf(g, w[(size_t)c * 3 + 1], a1);
    a2 = fmaf(g, w[(size_t)c * 3 + 2], a2);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  if (lane == 0) {
    ds[m * 3] = a0;
    ds[m * 3 + 1] = a1;
    ds[m * 3 + 2] = a2;
  }
}

// out[n] += sum_m x[m,n]; CTA = 64 columns x a row range; 8 warps stride rows, lanes own column pairs.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, int M, int N,
                   int rows_per_cta) {
  __shared__ float2 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + 2 * lane;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float2 acc = make_float2(0.f, 0.f);
  if (col + 1 < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (size_t)r * ld + col));
      acc.x += v.x;
      acc.y += v.y;
    }
  } else if (col < N) {
    for (int r = r0 + warp; r < r1; r += 8) acc.x += __bfloat162float(x[(size_t)r * ld + col]);
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float2 t = make_float2(0.f, 0.f);
    for (int w2 = 0; w2 < 8; ++w2) {
      t.x += red[w2][lane].x;
      t.y += red[w2][lane].y;
    }
    if (col < N) atomicAdd(out + col, t.x);
    if (col + 1 < N) atomicAdd(out + col + 1, t.y);
  }
}

// Wide variant (N % 8 == 0, 16-byte aligned rows): a lane owns 8 consecutive columns (one 16-byte load), a warp 256
// columns = 512 contiguous bytes of a row, four rows in flight per warp.
__global__ void __launch_bounds__(256)
colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, int M, int N,
                       int rows_per_cta) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + 8 * lane;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < N) {
    const __nv_bfloat16* base = x + col;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(base + (size_t)(r + 8 * u) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16(w4[q]);
          acc[2 * q] += f.x;
          acc[2 * q + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)r * ld);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16(w4[q]);
        acc[2 * q] += f.x;
        acc[2 * q + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][8 * lane + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += red[w2][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ============================================================ RoPE on the q|k blocks of packed qkv (in place)
__global__ void rope_qk_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ cos_t,
                               const float* __restrict__ sin_t, int M, int T, int nheads, int hd, long long ld,
                               int inverse) {
  const int half = hd >> 1;
  const int pairs_per_row = nheads * (half >> 1);  // each thread rotates 2 adjacent (i, i+half) pairs
  const long long total = (long long)M * pairs_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / pairs_per_row;
    const int r = (int)(i - m * pairs_per_row);
    const int h = r / (half >> 1);
    const int j = (r - h * (half >> 1)) * 2;
    const int t = (int)(m % T);
    __nv_bfloat16* base = qkv + m * ld + (long long)h * hd;
    const float2 x1 = unpack_bf16(*reinterpret_cast<uint32_t*>(base + j));
    const float2 x2 = unpack_bf16(*reinterpret_cast<uint32_t*>(base + half + j));
    const float2 c = *reinterpret_cast<const float2*>(cos_t + (size_t)t * half + j);
    float2 s = *reinterpret_cast<const float2*>(sin_t + (size_t)t * half + j);
    if (inverse) {
      s.x = -s.x;
      s.y = -s.y;
    }
    const float o1x = x1.x * c.x - x2.x * s.x, o1y = x1.y * c.y - x2.y * s.y;
    const float o2x = x2.x * c.x + x1.x * s.x, o2y = x2.y * c.y + x1.y * s.y;
    *reinterpret_cast<uint32_t*>(base + j) = pack_bf16(o1x, o1y);
    *reinterpret_cast<uint32_t*>(base + half + j) = pack_bf16(o2x, o2y);
  }
}

// ============================================================ SwiGLU gate (gu = [g | u], each h wide, h % 2 == 0)
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu, __nv_bfloat16* __restrict__ act,
                                  long long ldact, int M, int h) {
  const int h2 = h >> 1;
  const long long total = (long long)M * h2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / h2;
    const int j = (int)(i - m * h2) * 2;
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(gu + m * ldgu + j));
    const float2 u = unpack_bf16(*reinterpret_cast<const uint32_t*>(gu + m * ldgu + h + j));
    *reinterpret_cast<uint32_t*>(act + m * ldact + j) =
        pack_bf16(g.x * sigmoidf_(g.x) * u.x, g.y * sigmoidf_(g.y) * u.y);
  }
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu,
                                  const __nv_bfloat16* __restrict__ dact, long long ldact,
                                  __nv_bfloat16* __restrict__ dgu, int M, int h) {
  const int h2 = h >> 1;
  const long long total = (long long)M * h2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / h2;
    const int j = (int)(i - m * h2) * 2;
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(gu + m * ldgu + j));
    const float2 u = unpack_bf16(*reinterpret_cast<const uint32_t*>(gu + m * ldgu + h + j));
    const float2 da = unpack_bf16(*reinterpret_cast<const uint32_t*>(dact + m * ldact + j));
    const float sx = sigmoidf_(g.x), sy = sigmoidf_(g.y);
    const float dgx = da.x * u.x * sx * (1.f + g.x * (1.f - sx));
    const float dgy = da.y * u.y * sy * (1.f + g.y * (1.f - sy));
    const float dux = da.x * g.x * sx, duy = da.y * g.y * sy;
    *reinterpret_cast<uint32_t*>(dgu + m * ldgu + j) = pack_bf16(dgx, dgy);
    *reinterpret_cast<uint32_t*>(dgu + m * ldgu + h + j) = pack_bf16(dux, duy);
  }
}

// ============================================================ dropout (Philox mask recomputed from seed/offset)
// out = (residual) + x * mask / (1-p); x fp32; out fp32 or bf16.  4 elements per Philox call.
template <bool OUT_BF16>
__global__ void dropout_kernel(const float4* __restrict__ x, const float4* __restrict__ residual, void* __restrict__ out,
                               long long n4, DropoutCfg d_in) {
  const DropoutCfg d = resolve_dropout(d_in);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(d.key, make_uint4((uint32_t)i + d.off_lo, (uint32_t)(i >> 32) + d.off_hi, 0u, 0x656c7700u));
    float4 v = x[i];
    v.x = r.x >= d.thresh ? v.x * d.inv_keep : 0.f;
    v.y = r.y >= d.thresh ? v.y * d.inv_keep : 0.f;
    v.z = r.z >= d.thresh ? v.z * d.inv_keep : 0.f;
    v.w = r.w >= d.thresh ? v.w * d.inv_keep : 0.f;
    if (residual) {
      const float4 q = residual[i];
      v.x += q.x;
      v.y += q.y;
      v.z += q.z;
      v.w += q.w;
    }
    if constexpr (OUT_BF16)
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    else
      reinterpret_cast<float4*>(out)[i] = v;
  }
}

// ============================================================ AdamW (torch.optim.AdamW semantics)
template <typename GradT>
__global__ void adamw_kernel(float* __restrict__ p, const GradT* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, long long n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                             const float* __restrict__ hyper) {
  if (hyper) {  // step-dependent scalars live on the device so that a captured CUDA graph can be replayed
    lr = hyper[0];
    bc1 = hyper[1];
    bc2_sqrt = hyper[2];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gr;
    if constexpr (sizeof(GradT) == 2) gr = __bfloat162float(g[i]) * gscale; else gr = g[i] * gscale;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
    if (shadow) shadow[i] = __float2bfloat16_rn(pv);
  }
}

// CGPT_LN_STREAM: 0 = per-thread-load LayerNorm kernels, 1 (default) = streaming backward, 2 = streaming forward too
// (measured slower than the per-thread-load forward: 41.0 vs 38.9 us at 65536 x 512); read once
inline int ln_stream_enabled() {
  static const int on = [] {
    const char* e = getenv("CGPT_LN_STREAM");
    return e ? atoi(e) : 1;
  }();
  return on;
}
// CGPT_LN_REVERSE=1: LAST rows first (the GEMM in front of a LayerNorm produces its rows in ascending order, so its
// newest lines could still be in L2, and the LayerNorm would end on the rows the next GEMM reads first).  Measured on
// the captured step: 29.33 ms against 29.37 / 29.40 ms in ascending order on the same box — within the noise, so the
// default stays ascending.
inline bool ln_reverse_enabled() {
  static const bool on = [] {
    const char* e = getenv("CGPT_LN_REVERSE");
    return e && e[0] == '1';
  }();
  return on;
}

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace cgpt

using namespace cgpt;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int cgpt_segment_ids(const int64_t* idx, int32_t* seg, int B, int T, int sep_id, cgpt_stream_t stream) {
  CGPT_REQUIRE(idx && seg && B > 0 && T > 0, "segment_ids: bad arguments");
  segment_ids_kernel<<<(B + 3) / 4, 128, 0, ST(stream)>>>(idx, seg, B, T, sep_id);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_segment_starts(const int64_t* idx, int32_t* start, int B, int T, int sep_id, cgpt_stream_t stream) {
  CGPT_REQUIRE(idx && start && B > 0 && T > 0, "segment_starts: bad arguments");
  segment_starts_kernel<<<(B + 3) / 4, 128, 0, ST(stream)>>>(idx, start, B, T, sep_id);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_next_in_set(const int64_t* yb, int32_t* next, int B, int T, const int64_t* ids_host, int n_ids,
                     cgpt_stream_t stream) {
  CGPT_REQUIRE(yb && next && B > 0 && T > 0, "next_in_set: bad arguments");
  CGPT_REQUIRE(n_ids >= 0 && n_ids <= 8, "next_in_set: at most 8 ids (got %d)", n_ids);
  IdSet ids;
  ids.n = n_ids;
  for (int i = 0; i < n_ids; ++i) ids.v[i] = ids_host[i];
  next_in_set_kernel<<<(B + 3) / 4, 128, 0, ST(stream)>>>(yb, next, B, T, ids);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_termination_labels(const int64_t* yb, const int32_t* next_stop, int64_t* labels, int B, int T,
                            const int64_t* edges_host, int n_edges, int64_t ignore_index, cgpt_stream_t stream) {
  CGPT_REQUIRE(yb && next_stop && labels && B > 0 && T > 0, "termination_labels: bad arguments");
  CGPT_REQUIRE(n_edges >= 0 && n_edges <= 8, "termination_labels: at most 8 bucket edges (got %d)", n_edges);
  EdgeSet e;
  e.n = n_edges;
  for (int i = 0; i < n_edges; ++i) e.v[i] = edges_host[i];
  const size_t n = (size_t)B * T;
  termination_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(yb, next_stop, labels, B, T, e,
                                                                                 ignore_index);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_embed_fwd(const int64_t* idx, const float* tok_w, const float* pos_w, float* x, int B, int T, int d,
                   int vocab, cgpt_stream_t stream) {
  CGPT_REQUIRE(idx && tok_w && x && B > 0 && T > 0, "embed_fwd: bad arguments");
  CGPT_REQUIRE(d % 4 == 0, "embed_fwd: d=%d must be a multiple of 4", d);
  const size_t n4 = (size_t)B * T * (d / 4);
  embed_fwd_kernel<<<grid_for((long long)n4, 256), 256, 0, ST(stream)>>>(
      idx, reinterpret_cast<const float4*>(tok_w), reinterpret_cast<const float4*>(pos_w),
      reinterpret_cast<float4*>(x), n4, T, d / 4, vocab);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_embed_bwd(const int64_t* idx, const float* dx, float* dtok_w, float* dpos_w, int B, int T, int d,
                   int vocab, cgpt_stream_t stream) {
  CGPT_REQUIRE(idx && dx && dtok_w && B > 0 && T > 0, "embed_bwd: bad arguments");
  CGPT_REQUIRE(d % 4 == 0, "embed_bwd: d=%d must be a multiple of 4", d);
  const int M = B * T;
  const int max_smem = 200 * 1024;
  if ((size_t)kEmbTokWarps * vocab * 32 * 4 <= 96 * 1024) {
    const size_t smem = (size_t)kEmbTokWarps * vocab * 32 * 4;
    const int col_tiles = (d + 31) / 32;
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    int tok_ctas = (per_sm * num_sms() + col_tiles - 1) / col_tiles;
    if (tok_ctas > (M + 63) / 64) tok_ctas = (M + 63) / 64;
    if (tok_ctas < 1) tok_ctas = 1;
    const int toks = (M + tok_ctas - 1) / tok_ctas;
    tok_ctas = (M + toks - 1) / toks;
    static size_t configured_small = 0;
    if (smem > 48 * 1024 && smem > configured_small) {
      CGPT_CHECK(cudaFuncSetAttribute(embed_bwd_tok_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured_small = smem;
    }
    embed_bwd_tok_small_kernel<<<dim3(tok_ctas, col_tiles), kEmbTokWarps * 32, smem, ST(stream)>>>(idx, dx, dtok_w, M, d,
                                                                                                vocab, toks);
    count_launch();
    CGPT_LAUNCH_CHECK();
    if (dpos_w) {
      const size_t n = (size_t)T * (d / 4);
      embed_bwd_pos_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ST(stream)>>>(
          reinterpret_cast<const float4*>(dx), reinterpret_cast<float4*>(dpos_w), B, T, d / 4);
      count_launch();
      CGPT_LAUNCH_CHECK();
    }
    return 0;
  }
  int cols = d;
  while ((size_t)vocab * cols * 4 > (size_t)max_smem) cols = (cols + 1) / 2;
  cols = (cols + 3) / 4 * 4;
  const int col_tiles = (d + cols - 1) / cols;
  int tok_ctas = (2 * num_sms()) / col_tiles;
  if (tok_ctas < 1) tok_ctas = 1;
  if (tok_ctas > M) tok_ctas = M;
  const int toks = (M + tok_ctas - 1) / tok_ctas;
  tok_ctas = (M + toks - 1) / toks;
  const size_t smem = (size_t)vocab * cols * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CGPT_CHECK(cudaFuncSetAttribute(embed_bwd_tok_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  embed_bwd_tok_kernel<<<dim3(tok_ctas, col_tiles), 256, smem, ST(stream)>>>(idx, dx, dtok_w, M, d, vocab, cols,
                                                                            toks);
  count_launch();
  CGPT_LAUNCH_CHECK();
  if (dpos_w) {
    const size_t n = (size_t)T * (d / 4);
    embed_bwd_pos_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ST(stream)>>>(
        reinterpret_cast<const float4*>(dx), reinterpret_cast<float4*>(dpos_w), B, T, d / 4);
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  return 0;
}

int cgpt_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                       float* mean, float* rstd, int M, int d, float eps, cgpt_stream_t stream) {
  CGPT_REQUIRE(x && gamma && beta && mean && rstd && (y_bf16 || y_f32) && M > 0, "layernorm_fwd: bad arguments");
  CGPT_REQUIRE(d % 4 == 0 && d <= 128 * kLnMaxVec, "layernorm: d=%d must be a multiple of 4 and <= %d", d,
               128 * kLnMaxVec);
  const int vpt = (d / 4 + 31) / 32;
  const int rev = ln_reverse_enabled() ? 1 : 0;
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  if (d == 512 && ln_stream_enabled() > 1 && M >= 1024 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
    using L = LnFwdStream<4, 4>;
    static bool attr_set = false;
    if (!attr_set) {
      CGPT_CHECK(cudaFuncSetAttribute(layernorm_fwd_stream_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmem));
      attr_set = true;
    }
    const int grid = grid_for((long long)M * 32, kLnWarps * 32, 3);
    CGPT_CHECK(launch_pdl(layernorm_fwd_stream_kernel<4, 4>, dim3(grid), dim3(kLnWarps * 32), L::kSmem, ST(stream), 1, M, x, gamma,
                          beta, yb, y_f32, mean, rstd, M, eps, rev));
    count_launch();
    CGPT_LAUNCH_CHECK();
    return 0;
  }
  const int grid = grid_for((long long)M * 32, 256, 4);
#define LN_FWD(V) CGPT_CHECK(launch_pdl(layernorm_fwd_kernel<V>, dim3(grid), dim3(256), 0, ST(stream), 1, M, x, gamma, beta, yb, y_f32, mean, rstd, M, d, eps, rev))
  if (vpt <= 1) LN_FWD(1); else if (vpt <= 2) LN_FWD(2); else if (vpt <= 4) LN_FWD(4); else LN_FWD(8);
#undef LN_FWD
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_layernorm_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, const float* mean,
                       const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                       float* dx_colsum, int M, int d, cgpt_stream_t stream) {
  CGPT_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && M > 0, "layernorm_bwd: bad arguments");
  CGPT_REQUIRE(d % 4 == 0 && d <= 128 * kLnMaxVec, "layernorm: d=%d must be a multiple of 4 and <= %d", d,
               128 * kLnMaxVec);
  const int vpt = (d / 4 + 31) / 32;
  const int rev = ln_reverse_enabled() ? 1 : 0;
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (d == 512 && !dy_is_f32 && ln_stream_enabled() > 0 && M >= 1024 &&
      (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0)) {
    using L = LnBwdStream<4, 2>;
    static bool attr_set = false;
    if (!attr_set) {
      CGPT_CHECK(cudaFuncSetAttribute(layernorm_bwd_stream_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmem));
      attr_set = true;
    }
    const int grid = grid_for((long long)M * 32, kLnWarps * 32, 2);
    CGPT_CHECK(launch_pdl(layernorm_bwd_stream_kernel<4, 2>, dim3(grid), dim3(kLnWarps * 32), L::kSmem, ST(stream), 1, M,
                          reinterpret_cast<const __nv_bfloat16*>(dy), x, gamma, mean, rstd, dres, dx, dxb, dgamma, dbeta,
                          dx_colsum, M, rev));
    count_launch();
    CGPT_LAUNCH_CHECK();
    return 0;
  }
  const int grid = grid_for((long long)M * 32, 256, vpt <= 4 ? 2 : 1);
#define LN_BWD(V, F) CGPT_CHECK(launch_pdl(layernorm_bwd_kernel<V, F>, dim3(grid), dim3(256), 0, ST(stream), 1, M, dy, x, gamma, mean, rstd, dres, dx, dxb, dgamma, dbeta, dx_colsum, M, d, rev))
  if (dy_is_f32) {
    if (vpt <= 1) LN_BWD(1, true); else if (vpt <= 2) LN_BWD(2, true); else if (vpt <= 4) LN_BWD(4, true); else LN_BWD(8, true);
  } else {
    if (vpt <= 1) LN_BWD(1, false); else if (vpt <= 2) LN_BWD(2, false); else if (vpt <= 4) LN_BWD(4, false); else LN_BWD(8, false);
  }
#undef LN_BWD
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_cast_f32_bf16(const float* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int64_t cols,
                       cgpt_stream_t stream) {
  CGPT_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= cols, "cast: bad arguments");
  const bool dense = (ld_in == cols) && (ld_out == cols) && ((rows * cols) % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  if (dense) {
    const long long n4 = rows * cols / 4;
    cast_f32_bf16_vec_kernel<<<grid_for(n4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(in),
                                                                        reinterpret_cast<uint2*>(out), n4);
  } else {
    cast_f32_bf16_kernel<<<grid_for(rows * ld_out, 256), 256, 0, ST(stream)>>>(
        in, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  }
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_split3_f32_bf16(const float* in, int64_t ld_in, void* out, int64_t rows, int64_t cols, int64_t cols_pad,
                         int partner, cgpt_stream_t stream) {
  CGPT_REQUIRE(in && out && rows > 0 && cols > 0 && cols_pad >= cols && ld_in >= cols, "split3: bad arguments");
  split3_kernel<<<grid_for(rows * cols_pad, 256), 256, 0, ST(stream)>>>(in, ld_in, reinterpret_cast<__nv_bfloat16*>(out),
                                                                        rows, cols, cols_pad, partner);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_fold_quadrants_add(const float* s, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int row_off,
                            int col_off, cgpt_stream_t stream) {
  CGPT_REQUIRE(s && dst && rows > 0 && cols > 0 && row_off >= rows && col_off >= cols && lds >= col_off + cols &&
                   ldd >= cols,
               "fold_quadrants_add: bad arguments");
  fold_quadrants_add_kernel<<<grid_for((long long)rows * cols, 256), 256, 0, ST(stream)>>>(s, lds, dst, ldd, rows, cols,
                                                                                           row_off, col_off);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_attn_decode(const void* qkv_new, void* k_cache, void* v_cache, const int32_t* lo, void* out, int B, int t,
                     const int32_t* t_dev, int Tmax, int H, int Hk, int hd, int window, float scale,
                     cgpt_stream_t stream) {
  CGPT_REQUIRE(qkv_new && k_cache && v_cache && out && B > 0 && (t_dev || (t >= 0 && t < Tmax)) && Tmax > 0 && H > 0 &&
                   Hk > 0 && H % Hk == 0 && hd % 16 == 0 && hd <= 128 && window >= 0,
               "attn_decode: bad arguments");
  const size_t smem = (size_t)(hd + 128 * 8 + (t_dev ? Tmax : t + 1)) * sizeof(float);
  CGPT_REQUIRE(smem <= 200 * 1024, "attn_decode: context too long for the score buffer");
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    CGPT_CHECK(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  attn_decode_kernel<<<B * H, 128, smem, ST(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_new), reinterpret_cast<__nv_bfloat16*>(k_cache),
      reinterpret_cast<__nv_bfloat16*>(v_cache), lo, reinterpret_cast<__nv_bfloat16*>(out), t_dev, t, Tmax, H, Hk, hd,
      window, scale * 1.4426950408889634f);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_pack_lm_batch(const int32_t* tokens, const int64_t* offsets, const int64_t* lengths, const int64_t* indices,
                       int B, int T_out, int64_t* xb, int64_t* yb, cgpt_stream_t stream) {
  CGPT_REQUIRE(tokens && offsets && lengths && indices && xb && yb && B > 0 && T_out > 0, "pack_lm_batch: bad arguments");
  pack_lm_batch_kernel<<<grid_for((long long)B * T_out, 256), 256, 0, ST(stream)>>>(
      tokens, reinterpret_cast<const long long*>(offsets), reinterpret_cast<const long long*>(lengths),
      reinterpret_cast<const long long*>(indices), B, T_out, reinterpret_cast<long long*>(xb),
      reinterpret_cast<long long*>(yb));
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_shape_proj_fwd(const float* x, const float* s, const float* w, const float* b, float* out, int64_t M, int d,
                        cgpt_stream_t stream) {
  CGPT_REQUIRE(x && s && w && b && out && M > 0 && d > 0 && d % 4 == 0, "shape_proj_fwd: bad arguments");
  shape_proj_fwd_kernel<<<grid_for(M * (d / 4), 256), 256, 0, ST(stream)>>>(x, s, w, b, out, M, d);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_shape_proj_bwd(const float* dx, const float* s, const float* w, float* dw, float* db, float* ds, int64_t M,
                        int d, cgpt_stream_t stream) {
  CGPT_REQUIRE(dx && s && w && dw && db && M > 0 && d > 0, "shape_proj_bwd: bad arguments");
  const int rows_per_cta = 2048;
  dim3 grid((d + 63) / 64, (unsigned)((M + rows_per_cta - 1) / rows_per_cta));
  shape_proj_wgrad_kernel<<<grid, 256, 0, ST(stream)>>>(dx, s, dw, db, M, d, rows_per_cta);
  count_launch();
  CGPT_LAUNCH_CHECK();
  if (ds) {
    shape_proj_dgrad_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, ST(stream)>>>(dx, w, ds, M, d);
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  return 0;
}

int cgpt_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, cgpt_stream_t stream) {
  CGPT_REQUIRE(x && out && M > 0 && N > 0 && ld >= N && ld % 2 == 0, "colsum: bad arguments");
  if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int ct = (N + 255) / 256;
    int rc = (4 * num_sms() + ct - 1) / ct;
    int rows = (M + rc - 1) / rc;
    if (rows < 64) rows = 64;
    rc = (M + rows - 1) / rows;
    colsum_bf16_vec_kernel<<<dim3(ct, rc), 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, out, M, N,
                                                                 rows);
    count_launch();
    CGPT_LAUNCH_CHECK();
    return 0;
  }
  const int col_tiles = (N + 63) / 64;
  int row_ctas = (4 * num_sms() + col_tiles - 1) / col_tiles;
  int rows = (M + row_ctas - 1) / row_ctas;
  if (rows < 64) rows = 64;
  row_ctas = (M + rows - 1) / rows;
  colsum_bf16_kernel<<<dim3(col_tiles, row_ctas), 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld,
                                                                       out, M, N, rows);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_rope_qk(void* qkv, const float* cos_t, const float* sin_t, int B, int T, int H, int Hk, int hd,
                 int inverse, cgpt_stream_t stream) {
  CGPT_REQUIRE(qkv && cos_t && sin_t && B > 0 && T > 0, "rope: bad arguments");
  CGPT_REQUIRE(hd % 4 == 0, "rope: head_dim=%d must be a multiple of 4", hd);
  const long long ld = (long long)(H + 2 * Hk) * hd;
  const long long total = (long long)B * T * (H + Hk) * (hd / 4);
  rope_qk_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(reinterpret_cast<__nv_bfloat16*>(qkv), cos_t, sin_t,
                                                               B * T, T, H + Hk, hd, ld, inverse);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_swiglu_fwd(const void* gu, int64_t ldgu, void* act, int64_t ldact, int M, int h, cgpt_stream_t stream) {
  CGPT_REQUIRE(gu && act && M > 0 && h > 0 && h % 2 == 0 && ldgu % 2 == 0 && ldact % 2 == 0, "swiglu_fwd: bad arguments");
  swiglu_fwd_kernel<<<grid_for((long long)M * h / 2, 256), 256, 0, ST(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(gu), ldgu, reinterpret_cast<__nv_bfloat16*>(act), ldact, M, h);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_swiglu_bwd(const void* gu, int64_t ldgu, const void* dact, int64_t ldact, void* dgu, int M, int h,
                    cgpt_stream_t stream) {
  CGPT_REQUIRE(gu && dact && dgu && M > 0 && h > 0 && h % 2 == 0 && ldgu % 2 == 0 && ldact % 2 == 0,
               "swiglu_bwd: bad arguments");
  swiglu_bwd_kernel<<<grid_for((long long)M * h / 2, 256), 256, 0, ST(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(gu), ldgu, reinterpret_cast<const __nv_bfloat16*>(dact), ldact,
      reinterpret_cast<__nv_bfloat16*>(dgu), M, h);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_dropout(const float* x, const float* residual, void* out, int out_bf16, int64_t n, float p, uint64_t seed,
                 uint64_t offset, cgpt_stream_t stream) {
  CGPT_REQUIRE(x && out && n > 0 && n % 4 == 0, "dropout: n=%lld must be a positive multiple of 4", (long long)n);
  CGPT_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f must be in [0,1)", p);
  const DropoutCfg d = make_dropout(p, seed, offset);
  const long long n4 = n / 4;
  if (out_bf16)
    dropout_kernel<true><<<grid_for(n4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(x),
                                                                     reinterpret_cast<const float4*>(residual), out, n4, d);
  else
    dropout_kernel<false><<<grid_for(n4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(x),
                                                                      reinterpret_cast<const float4*>(residual), out, n4, d);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, const float* dev_hyper,
               cgpt_stream_t stream) {
  CGPT_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<float><<<grid_for(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n,
                                                                lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, dev_hyper);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_adamw_bf16grad(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        const float* dev_hyper, cgpt_stream_t stream) {
  CGPT_REQUIRE(p && g_bf16 && m && v && n > 0 && step >= 1, "adamw_bf16grad: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, ST(stream)>>>(
      p, reinterpret_cast<const __nv_bfloat16*>(g_bf16), m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n, lr, beta1,
      beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, dev_hyper);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
