// Small-vocabulary heads: out = x·wᵀ (+b) with N <= 128 output columns (the ~68-codon LM head and
// the 5-class termination head), their backward, and the softmax cross-entropy (ignore_index,
// label smoothing, class weights, offset-validity mask).  Kept in fp32 FMA: 2·d·V FLOP/token is
// < 0.1 % of the step and the reference's argmax must be reproduced (SURVEY §7 hard part 1).
// Replaces model_tiny_gpt.py:327,330,336,343-349 and objectives.py:39-57,100-105.
#include "common.cuh"

namespace cgpt {
namespace {

constexpr int kKC = 32;       // k-chunk
constexpr int kPitch = 36;    // smem row pitch in floats (16-byte aligned, conflict-free float4 reads)

// ------------------------------------------------------------------ forward: 64 rows x (16*J) cols per CTA
template <int J>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int M, int N, int d) {
  __shared__ __align__(16) float xs[64 * kPitch];
  __shared__ __align__(16) float ws[16 * J * kPitch];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * 64;
  float acc[4][J];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += kKC) {
    __syncthreads();
    for (int i = tid; i < 64 * (kKC / 4); i += 256) {
      const int r = i / (kKC / 4), c4 = i % (kKC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < M && k0 + c4 * 4 < d) v = *reinterpret_cast<const float4*>(x + (size_t)(row0 + r) * d + k0 + c4 * 4);
      *reinterpret_cast<float4*>(&xs[r * kPitch + c4 * 4]) = v;
    }
    for (int i = tid; i < 16 * J * (kKC / 4); i += 256) {
      const int r = i / (kKC / 4), c4 = i % (kKC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < N && k0 + c4 * 4 < d) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)r * d + k0 + c4 * 4));
      *reinterpret_cast<float4*>(&ws[r * kPitch + c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKC; kk += 4) {
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(&xs[(ty * 4 + i) * kPitch + kk]);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[(tx + 16 * j) * kPitch + kk]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[i][j] += (xv[i].x * wv.x + xv[i].y * wv.y) + (xv[i].z * wv.z + xv[i].w * wv.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = tx + 16 * j;
      if (c < N) out[(size_t)r * N + c] = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
    }
  }
}

// ------------------------------------------------------------------ dx (+)= dout·w : 64 rows x 64 d-cols per CTA
__global__ void __launch_bounds__(256)
skinny_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w, float* __restrict__ dx, int M, int N,
                 int d, int accumulate) {
  extern __shared__ __align__(16) float sm[];
  const int np = N + 1;
  float* ds = sm;                 // [64][np]
  float* ws = sm + 64 * np + ((4 - (64 * np) % 4) % 4);  // [N][64], 16-byte aligned
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int i = tid; i < 64 * N; i += 256) {
    const int r = i / N, n = i - r * N;
    ds[r * np + n] = (row0 + r < M) ? dout[(size_t)(row0 + r) * N + n] : 0.f;
  }
  for (int i = tid; i < N * 16; i += 256) {
    const int n = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + c4 * 4 < d) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * d + c0 + c4 * 4));
    *reinterpret_cast<float4*>(&ws[n * 64 + c4 * 4]) = v;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float4 b = *reinterpret_cast<const float4*>(&ws[n * 64 + tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = ds[(ty * 4 + i) * np + n];
      acc[i][0] += a * b.x;
      acc[i][1] += a * b.y;
      acc[i][2] += a * b.z;
      acc[i][3] += a * b.w;
    }
  }
  const int c = c0 + tx * 4;
  if (c >= d) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= M) continue;
    float4* o = reinterpret_cast<float4*>(dx + (size_t)r * d + c);
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (accumulate) {
      const float4 p = *o;
      v.x += p.x;
      v.y += p.y;
      v.z += p.z;
      v.w += p.w;
    }
    *o = v;
  }
}

// ------------------------------------------------------------------ dw += doutᵀ·x ; dbias += colsum(dout)
template <int J>
__global__ void __launch_bounds__(256)
skinny_dw_kernel(const float* __restrict__ dout, const float* __restrict__ x, float* __restrict__ dw,
                 float* __restrict__ dbias, int M, int N, int d, int rows_per_cta) {
  __shared__ __align__(16) float ds[32 * 16 * J];
  __shared__ __align__(16) float xs[32 * 64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.x * 64;
  const int r_begin = blockIdx.y * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  float acc[J][4];
  float bacc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  }
  for (int r0 = r_begin; r0 < r_end; r0 += 32) {
    __syncthreads();
    for (int i = tid; i < 32 * 16 * J; i += 256) {
      const int r = i / (16 * J), n = i - r * (16 * J);
      ds[i] = (r0 + r < r_end && n < N) ? dout[(size_t)(r0 + r) * N + n] : 0.f;
    }
    for (int i = tid; i < 32 * 16; i += 256) {
      const int r = i >> 4, c4 = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < r_end && c0 + c4 * 4 < d) v = *reinterpret_cast<const float4*>(x + (size_t)(r0 + r) * d + c0 + c4 * 4);
      *reinterpret_cast<float4*>(&xs[r * 64 + c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float4 b = *reinterpret_cast<const float4*>(&xs[r * 64 + tx * 4]);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float a = ds[r * 16 * J + ty + 16 * j];
        acc[j][0] += a * b.x;
        acc[j][1] += a * b.y;
        acc[j][2] += a * b.z;
        acc[j][3] += a * b.w;
        bacc[j] += a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int n = ty + 16 * j;
    if (n >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + tx * 4 + i;
      if (c < d) atomicAdd(dw + (size_t)n * d + c, acc[j][i]);
    }
    if (dbias && blockIdx.x == 0 && tx == 0) atomicAdd(dbias + n, bacc[j]);
  }
}

// dw += doutᵀ·x for heads with N <= 8 outputs (the 5-class termination head): the generic kernel keeps 11 of its 16
// output lanes idle there and stages x through shared memory.  Here a thread owns 4 consecutive columns of x for ALL N
// outputs (N x 4 fp32 accumulators), the CTA's 256 threads are (d/4) column owners x several row groups, rows are walked
// with 4 loads in flight per thread and the row's N gradients come as broadcast loads: x is read once, coalesced, at HBM
// rate.  Row groups are combined through shared memory, then one atomic per accumulator and CTA.
__global__ void __launch_bounds__(256)
skinny_dw_small_kernel(const float* __restrict__ dout, const float* __restrict__ x, float* __restrict__ dw,
                       float* __restrict__ dbias, int M, int N, int d, int rows_per_cta) {
  extern __shared__ float red[];  // [groups-1][8][d]
  const int d4 = d >> 2;
  const int groups = 256 / d4;          // row groups inside the CTA (d = 512: 2)
  const int tid = threadIdx.x;
  const int grp = tid / d4, c4 = tid - grp * d4;
  const bool active = grp < groups;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  float acc[8][4];
  float bacc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    bacc[n] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  }
  if (active) {
    const float4* xp = reinterpret_cast<const float4*>(x) + c4;
    for (int r0 = r_begin + grp * 4; r0 < r_end; r0 += groups * 4) {
      float4 xv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        xv[k] = (r0 + k < r_end) ? __ldg(xp + (size_t)(r0 + k) * d4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (r0 + k >= r_end) break;
        const float* g = dout + (size_t)(r0 + k) * N;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          if (n < N) {
            const float a = __ldg(g + n);
            acc[n][0] += a * xv[k].x;
            acc[n][1] += a * xv[k].y;
            acc[n][2] += a * xv[k].z;
            acc[n][3] += a * xv[k].w;
            bacc[n] += a;
          }
        }
      }
    }
  }
  if (active && grp > 0) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) red[((grp - 1) * 8 + n) * d + c4 * 4 + i] = acc[n][i];
  }
  __syncthreads();
  if (active && grp == 0) {
    for (int gq = 1; gq < groups; ++gq)
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] += red[((gq - 1) * 8 + n) * d + c4 * 4 + i];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (n >= N) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dw + (size_t)n * d + c4 * 4 + i, acc[n][i]);
    }
  }
  // bias gradient: every row group of column owner 0 has summed its rows' gradients
  if (dbias && active && c4 == 0) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < N) atomicAdd(dbias + n, bacc[n]);
  }
}

// ------------------------------------------------------------------ cross entropy (one THREAD per row, V <= 128)
struct RowInfo {
  bool keep, bad;
  int64_t target;
};

// `bad`: a kept row whose label lies outside [0, V).  F.cross_entropy raises for such a label; a kernel cannot, and
// must neither read logits / class weights out of bounds nor return a plausible-looking number: the row's loss term is
// NaN (the total loss turns NaN, which the trainer's non-finite policy reports and aborts on) and its gradient row 0.
__device__ __forceinline__ RowInfo ce_row_info(const int64_t* targets, const int32_t* next_boundary, int row, int T,
                                               int shift, int64_t ignore_index, int V) {
  RowInfo ri;
  const int b = row / T, t = row - b * T;
  ri.keep = false;
  ri.bad = false;
  ri.target = 0;
  if (t + shift < T) {
    const int64_t tg = targets[(size_t)b * T + t + shift];
    ri.target = tg;
    ri.keep = (tg != ignore_index) && (next_boundary == nullptr || next_boundary[row] >= t + shift);
    if (ri.keep && (tg < 0 || tg >= V)) {
      ri.bad = true;
      ri.keep = false;
    }
  }
  return ri;
}

// A CTA of 128 threads owns 128 consecutive rows: their logits (one contiguous block of 128*V floats) are copied to
// shared memory with coalesced 16-byte loads, then every thread walks ITS row from shared memory (odd row pitch in
// 16-byte cells / words: conflict-free).  The per-row scalar work (row info, max, log-sum-exp, target pick) is done
// once per row instead of once per lane of a warp-per-row kernel, which is what bound the V = 68 heads: ~11
// warp-instructions per row instead of ~150.  The CTA writes ONE partial (loss, weight) pair; the last CTA to finish
// (ticket counter) adds the partials up in index order, so the sums do not depend on the order of arrival.
constexpr int kCeThreads = 128;
constexpr int kCeCounters = 64;
__device__ unsigned int g_ce_tickets[kCeCounters];  // zero at load; the last CTA of a launch re-arms its counter

// smem row pitch in floats: rows of 16-byte cells with an odd cell count (V % 4 == 0), else an odd word count
__host__ __device__ inline int ce_pitch(int V) {
  if ((V & 3) == 0) return ((V >> 2) & 1) ? V : V + 4;
  return (V & 1) ? V : V + 1;
}

// global [rows, V] block -> smem rows of pitch Vs (VEC: 16-byte pieces; a piece never straddles rows)
template <bool VEC>
__device__ __forceinline__ void ce_stage_rows(const float* __restrict__ src, float* z, int rows, int V, int Vs) {
  const int tid = threadIdx.x;
  if constexpr (VEC) {
    const int V4 = V >> 2, n4 = rows * V4;
    int r = tid / V4, c = tid - r * V4;
    const int dr = kCeThreads / V4, dc = kCeThreads - dr * V4;
#pragma unroll 4
    for (int i = tid; i < n4; i += kCeThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      *reinterpret_cast<float4*>(z + r * Vs + c * 4) = v;
      r += dr;
      c += dc;
      if (c >= V4) {
        c -= V4;
        ++r;
      }
    }
  } else {
    const int n = rows * V;
    int r = tid / V, c = tid - r * V;
    const int dr = kCeThreads / V, dc = kCeThreads - dr * V;
#pragma unroll 4
    for (int i = tid; i < n; i += kCeThreads) {
      z[r * Vs + c] = __ldg(src + i);
      r += dr;
      c += dc;
      if (c >= V) {
        c -= V;
        ++r;
      }
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kCeThreads)
ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
              const int32_t* __restrict__ next_boundary, const float* __restrict__ class_w, float* __restrict__ part,
              float* __restrict__ row_lse, float* __restrict__ sums, float* __restrict__ mean_out, int M, int T, int V,
              int Vs, int shift,
              float smoothing, int64_t ignore_index, int zero_if_empty, unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(16) float ce_smem[];
  __shared__ float red[2][kCeThreads / 32];
  __shared__ bool is_last;
  float* z = ce_smem;                    // [128][Vs]
  float* wcs = ce_smem + kCeThreads * Vs;  // [V] class weights (1 when there are none)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row_base = blockIdx.x * kCeThreads;
  const int rows = min(kCeThreads, M - row_base);
  const int n_part = gridDim.x;
  for (int c = tid; c < V; c += kCeThreads) wcs[c] = class_w ? __ldg(class_w + c) : 1.f;
  ce_stage_rows<VEC>(logits + (size_t)row_base * V, z, rows, V, Vs);
  RowInfo ri;
  ri.keep = ri.bad = false;
  ri.target = 0;
  if (tid < rows) ri = ce_row_info(targets, next_boundary, row_base + tid, T, shift, ignore_index, V);
  __syncthreads();
  float loss = 0.f, wt = 0.f;
  if (tid < rows) {
    const float* zr = z + tid * Vs;
    float mx = -INFINITY;
    if constexpr (VEC) {
      for (int c = 0; c < V; c += 4) {
        const float4 q = *reinterpret_cast<const float4*>(zr + c);
        mx = fmaxf(fmaxf(mx, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
      }
    } else {
      for (int c = 0; c < V; ++c) mx = fmaxf(mx, zr[c]);
    }
    float se = 0.f, swz = 0.f, sw = 0.f;
    const bool smooth = smoothing > 0.f && ri.keep;
    if constexpr (VEC) {
      for (int c = 0; c < V; c += 4) {
        const float4 q = *reinterpret_cast<const float4*>(zr + c);
        se += (expf(q.x - mx) + expf(q.y - mx)) + (expf(q.z - mx) + expf(q.w - mx));
        if (smooth) {
          const float4 w4 = *reinterpret_cast<const float4*>(wcs + c);
          swz += (w4.x * q.x + w4.y * q.y) + (w4.z * q.z + w4.w * q.w);
          sw += (w4.x + w4.y) + (w4.z + w4.w);
        }
      }
    } else {
      for (int c = 0; c < V; ++c) {
        const float q = zr[c];
        se += expf(q - mx);
        if (smooth) {
          swz += wcs[c] * q;
          sw += wcs[c];
        }
      }
    }
    const float lse = mx + logf(se);
    row_lse[row_base + tid] = lse;
    if (ri.keep) {
      const int tg = (int)ri.target;
      const float wy = wcs[tg];
      loss = (1.f - smoothing) * wy * (lse - zr[tg]);
      if (smoothing > 0.f) loss += (smoothing / V) * (sw * lse - swz);
      wt = wy;
    } else if (ri.bad) {
      loss = __int_as_float(0x7fc00000);  // label out of range: poison the loss, do not dereference
      wt = 1.f;
    }
  }
  // fixed order: lanes of a warp (shuffle tree), warps of the CTA, then CTAs (below)
  loss = warp_sum(loss);
  wt = warp_sum(wt);
  if (lane == 0) {
    red[0][warp] = loss;
    red[1][warp] = wt;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int w = 0; w < kCeThreads / 32; ++w) {
      a += red[0][w];
      b2 += red[1][w];
    }
    part[blockIdx.x] = a;
    part[n_part + blockIdx.x] = b2;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == (unsigned)(n_part - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float a = 0.f, b2 = 0.f;
  for (int i = tid; i < n_part; i += kCeThreads) {
    a += __ldcg(part + i);
    b2 += __ldcg(part + n_part + i);
  }
  a = warp_sum(a);
  b2 = warp_sum(b2);
  __syncthreads();  // red[] of the first reduction has been read
  if (lane == 0) {
    red[0][warp] = a;
    red[1][warp] = b2;
  }
  __syncthreads();
  if (tid == 0) {
    a = b2 = 0.f;
#pragma unroll
    for (int w = 0; w < kCeThreads / 32; ++w) {
      a += red[0][w];
      b2 += red[1][w];
    }
    sums[0] = a;
    sums[1] = b2;
    if (mean_out) *mean_out = (zero_if_empty && !(b2 > 0.f)) ? 0.f : a / b2;
    *ticket = 0u;
  }
}

// bf16 by-products of the logit gradient for the GEMMs that consume it (the head's input- and weight-gradient GEMMs
// read bf16 operands): mode 1 = [M, ld] bf16(g) with zero pad columns; mode 2 = [M, 3*ld] hi | lo | hi (the split
// operand of the fp32-accurate head), pads zero.  Written from the staged fp32 rows, 4 bytes per thread, coalesced.
template <bool VEC>
__global__ void __launch_bounds__(kCeThreads)
ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ row_lse, const int64_t* __restrict__ targets,
              const int32_t* __restrict__ next_boundary, const float* __restrict__ class_w,
              const float* __restrict__ sums, const float* __restrict__ gscale, float coef, float* __restrict__ dlogits,
              uint32_t* __restrict__ dl_bf16, int bf16_mode, int ld_bf16, int M, int T, int V, int Vs, int shift,
              float smoothing, int64_t ignore_index) {
  extern __shared__ __align__(16) float ce_smem[];
  float* z = ce_smem;
  float* wcs = ce_smem + kCeThreads * Vs;
  const int tid = threadIdx.x;
  const int row_base = blockIdx.x * kCeThreads;
  const int rows = min(kCeThreads, M - row_base);
  for (int c = tid; c < V; c += kCeThreads) wcs[c] = class_w ? __ldg(class_w + c) : 1.f;
  ce_stage_rows<VEC>(logits + (size_t)row_base * V, z, rows, V, Vs);
  RowInfo ri;
  ri.keep = ri.bad = false;
  ri.target = 0;
  float lse = 0.f;
  if (tid < rows) {
    ri = ce_row_info(targets, next_boundary, row_base + tid, T, shift, ignore_index, V);
    lse = row_lse[row_base + tid];
  }
  const float scale = coef * (gscale ? *gscale : 1.f) / sums[1];
  __syncthreads();
  if (tid < rows) {
    float* zr = z + tid * Vs;
    if (!ri.keep) {
      for (int c = 0; c < V; ++c) zr[c] = 0.f;
    } else {
      const int tg = (int)ri.target;
      const float wy = wcs[tg];
      const float a = (1.f - smoothing) * wy * scale;
      float sw = 0.f;
      if (smoothing > 0.f)
        for (int c = 0; c < V; ++c) sw += wcs[c];
      const float bsm = smoothing > 0.f ? (smoothing / V) * scale : 0.f;
      // g = scale * ((1-s) wy (p - onehot) + s/V (p sw - w_c))
      for (int c = 0; c < V; ++c) {
        const float p = expf(zr[c] - lse);
        float g = a * (p - (c == tg ? 1.f : 0.f));
        if (smoothing > 0.f) g += bsm * (p * sw - wcs[c]);
        zr[c] = g;
      }
    }
  }
  __syncthreads();
  // smem rows -> dlogits (same piece mapping as the staging), then the bf16 by-product
  {
    float* dst = dlogits + (size_t)row_base * V;
    if constexpr (VEC) {
      const int V4 = V >> 2, n4 = rows * V4;
      int r = tid / V4, c = tid - r * V4;
      const int dr = kCeThreads / V4, dc = kCeThreads - dr * V4;
      for (int i = tid; i < n4; i += kCeThreads) {
        reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<const float4*>(z + r * Vs + c * 4);
        r += dr;
        c += dc;
        if (c >= V4) {
          c -= V4;
          ++r;
        }
      }
    } else {
      const int n = rows * V;
      int r = tid / V, c = tid - r * V;
      const int dr = kCeThreads / V, dc = kCeThreads - dr * V;
      for (int i = tid; i < n; i += kCeThreads) {
        dst[i] = z[r * Vs + c];
        r += dr;
        c += dc;
        if (c >= V) {
          c -= V;
          ++r;
        }
      }
    }
  }
  if (bf16_mode != 0) {
    const int wpr1 = ld_bf16 >> 1;                       // 4-byte words per section of a row
    const int wpr = bf16_mode == 2 ? 3 * wpr1 : wpr1;    // words per output row
    uint32_t* dst = dl_bf16 + (size_t)row_base * wpr;
    const int n = rows * wpr;
    for (int i = tid; i < n; i += kCeThreads) {
      const int r = i / wpr, w = i - r * wpr;
      const int sec = w / wpr1, c = (w - sec * wpr1) * 2;
      const float v0 = c < V ? z[r * Vs + c] : 0.f, v1 = c + 1 < V ? z[r * Vs + c + 1] : 0.f;
      uint32_t o = pack_bf16(v0, v1);
      if (sec == 1) {  // lo = bf16(v - hi)
        const float2 hi = unpack_bf16(o);
        o = pack_bf16(v0 - hi.x, v1 - hi.y);
      }
      dst[i] = o;
    }
  }
}

}  // namespace
}  // namespace cgpt

using namespace cgpt;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int cgpt_skinny_linear_fwd(const float* x, const float* w, const float* bias, float* out, int M, int N, int d,
                           cgpt_stream_t stream) {
  CGPT_REQUIRE(x && w && out && M > 0 && N > 0 && d > 0, "skinny_linear_fwd: bad arguments");
  CGPT_REQUIRE(N <= 128, "skinny_linear: N=%d > 128 (use cgpt_gemm_bf16)", N);
  CGPT_REQUIRE(d % 4 == 0, "skinny_linear: d=%d must be a multiple of 4", d);
  const int grid = (M + 63) / 64;
  if (N <= 16)
    skinny_fwd_kernel<1><<<grid, 256, 0, ST(stream)>>>(x, w, bias, out, M, N, d);
  else if (N <= 80)
    skinny_fwd_kernel<5><<<grid, 256, 0, ST(stream)>>>(x, w, bias, out, M, N, d);
  else
    skinny_fwd_kernel<8><<<grid, 256, 0, ST(stream)>>>(x, w, bias, out, M, N, d);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_skinny_linear_bwd(const float* dout, const float* x, const float* w, float* dx, int dx_accumulate, float* dw,
                           float* dbias, int M, int N, int d, cgpt_stream_t stream) {
  CGPT_REQUIRE(dout && x && w && M > 0 && N > 0 && d > 0, "skinny_linear_bwd: bad arguments");
  CGPT_REQUIRE(N <= 128, "skinny_linear: N=%d > 128", N);
  CGPT_REQUIRE(d % 4 == 0, "skinny_linear: d=%d must be a multiple of 4", d);
  const int col_tiles = (d + 63) / 64;
  if (dx) {
    const size_t smem = (size_t)(64 * (N + 1) + 4 + N * 64) * 4;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      CGPT_CHECK(cudaFuncSetAttribute(skinny_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    skinny_dx_kernel<<<dim3((M + 63) / 64, col_tiles), 256, smem, ST(stream)>>>(dout, w, dx, M, N, d, dx_accumulate);
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  if (dw) {
    int row_ctas = (3 * num_sms() + col_tiles - 1) / col_tiles;
    int rows = (M + row_ctas - 1) / row_ctas;
    rows = (rows + 31) / 32 * 32;
    row_ctas = (M + rows - 1) / rows;
    dim3 grid(col_tiles, row_ctas);
    const int d4 = d / 4;
    if (N <= 8 && d4 <= 256 && 256 % d4 == 0 && (size_t)(256 / d4 - 1) * 8 * d * 4 <= 48 * 1024) {
      // small-head kernel: 2 CTAs per SM worth of row slices, each a multiple of the 4-row step of every row group
      const int groups = 256 / d4;
      int ctas = 2 * num_sms();
      int rpc = (M + ctas - 1) / ctas;
      rpc = (rpc + groups * 4 - 1) / (groups * 4) * (groups * 4);
      ctas = (M + rpc - 1) / rpc;
      skinny_dw_small_kernel<<<ctas, 256, (size_t)(groups - 1) * 8 * d * 4, ST(stream)>>>(dout, x, dw, dbias, M, N, d, rpc);
    } else if (N <= 16)
      skinny_dw_kernel<1><<<grid, 256, 0, ST(stream)>>>(dout, x, dw, dbias, M, N, d, rows);
    else if (N <= 80)
      skinny_dw_kernel<5><<<grid, 256, 0, ST(stream)>>>(dout, x, dw, dbias, M, N, d, rows);
    else
      skinny_dw_kernel<8><<<grid, 256, 0, ST(stream)>>>(dout, x, dw, dbias, M, N, d, rows);
    count_launch();
    CGPT_LAUNCH_CHECK();
  }
  return 0;
}

int cgpt_ce_fwd(const float* logits, const int64_t* targets, const int32_t* next_boundary, const float* class_w,
                float* sums, float* mean_out, float* row_lse, float* row_ws, int B, int T, int V, int shift,
                float smoothing, int64_t ignore_index, int zero_if_empty, cgpt_stream_t stream) {
  CGPT_REQUIRE(logits && targets && sums && row_lse && row_ws && B > 0 && T > 0, "ce_fwd: bad arguments");
  CGPT_REQUIRE(V >= 1 && V <= 128, "ce: V=%d must be in [1,128]", V);
  CGPT_REQUIRE(shift >= 0, "ce: shift must be >= 0");
  const int M = B * T;
  const int n_part = (M + kCeThreads - 1) / kCeThreads;  // per-CTA partial pairs live in row_ws (2*M floats >= 2*n_part)
  const int Vs = ce_pitch(V);
  const size_t smem = (size_t)(kCeThreads * Vs + V) * 4;
  const bool vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  static bool attr_set = false;
  if (!attr_set) {
    CGPT_CHECK(cudaFuncSetAttribute(ce_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CGPT_CHECK(cudaFuncSetAttribute(ce_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CGPT_CHECK(cudaFuncSetAttribute(ce_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CGPT_CHECK(cudaFuncSetAttribute(ce_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    attr_set = true;
  }
  // one ticket counter per launch in flight: launches on one stream are ordered anyway; launches on different streams
  // get different counters unless 64 other CE launches were issued in between
  static unsigned int next_ticket = 0;
  static unsigned int* ticket_base[64] = {nullptr};  // the symbol's address, looked up once per device
  int dev = 0;
  CGPT_CHECK(cudaGetDevice(&dev));
  CGPT_REQUIRE(dev >= 0 && dev < 64, "ce_fwd: device index %d out of range", dev);
  if (ticket_base[dev] == nullptr)
    CGPT_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&ticket_base[dev]), g_ce_tickets));
  unsigned int* tickets = ticket_base[dev];
  unsigned int* ticket = tickets + (next_ticket++ % kCeCounters);
  if (vec)
    ce_fwd_kernel<true><<<n_part, kCeThreads, smem, ST(stream)>>>(logits, targets, next_boundary, class_w, row_ws, row_lse,
                                                                  sums, mean_out, M, T, V, Vs, shift, smoothing, ignore_index,
                                                                  zero_if_empty, ticket);
  else
    ce_fwd_kernel<false><<<n_part, kCeThreads, smem, ST(stream)>>>(logits, targets, next_boundary, class_w, row_ws, row_lse,
                                                                   sums, mean_out, M, T, V, Vs, shift, smoothing, ignore_index,
                                                                   zero_if_empty, ticket);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

int cgpt_ce_bwd(const float* logits, const float* row_lse, const int64_t* targets, const int32_t* next_boundary,
                const float* class_w, const float* sums, const float* gscale, float coef, float* dlogits,
                void* dl_bf16, int bf16_mode, int64_t ld_bf16, int B, int T, int V, int shift, float smoothing,
                int64_t ignore_index, cgpt_stream_t stream) {
  CGPT_REQUIRE(logits && row_lse && targets && sums && dlogits && B > 0 && T > 0, "ce_bwd: bad arguments");
  CGPT_REQUIRE(V >= 1 && V <= 128, "ce: V=%d must be in [1,128]", V);
  CGPT_REQUIRE(bf16_mode == 0 || (dl_bf16 && (bf16_mode == 1 || bf16_mode == 2) && ld_bf16 >= V && ld_bf16 % 2 == 0 &&
                                  ld_bf16 <= 256 && (reinterpret_cast<uintptr_t>(dl_bf16) & 3) == 0),
               "ce_bwd: bf16 by-product needs mode 1|2, an even pitch >= V and a 4-byte aligned buffer");
  const int M = B * T;
  const int Vs = ce_pitch(V);
  const size_t smem = (size_t)(kCeThreads * Vs + V) * 4;
  const bool vec = (V % 4 == 0) && (((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0);
  static bool attr_set = false;
  if (!attr_set) {
    CGPT_CHECK(cudaFuncSetAttribute(ce_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CGPT_CHECK(cudaFuncSetAttribute(ce_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    attr_set = true;
  }
  const int grid = (M + kCeThreads - 1) / kCeThreads;
  uint32_t* ob = reinterpret_cast<uint32_t*>(dl_bf16);
  if (vec)
    ce_bwd_kernel<true><<<grid, kCeThreads, smem, ST(stream)>>>(logits, row_lse, targets, next_boundary, class_w, sums, gscale,
                                                                coef, dlogits, ob, bf16_mode, (int)ld_bf16, M, T, V, Vs, shift,
                                                                smoothing, ignore_index);
  else
    ce_bwd_kernel<false><<<grid, kCeThreads, smem, ST(stream)>>>(logits, row_lse, targets, next_boundary, class_w, sums, gscale,
                                                                 coef, dlogits, ob, bf16_mode, (int)ld_bf16, M, T, V, Vs, shift,
                                                                 smoothing, ignore_index);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
