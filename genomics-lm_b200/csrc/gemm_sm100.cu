// Dense bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double-buffered) -> fused epilogue.  Persistent, warp-specialised:
//   warp 0      TMA producer            (one elected lane)
//   warp 1      TMEM owner + MMA issuer (one elected lane)
//   warps 2..17 epilogue: tcgen05.ld -> swizzled smem transpose -> bias/GELU/aux/residual -> coalesced 16-B stores
// Tile 128 x BN x 64, BN in {64,128,256}.  Operands are K-major ([rows,K]) or MN-major ([K,rows]),
// which covers forward (x·Wᵀ), dgrad (dy·W) and wgrad (dyᵀ·x, split-K with fp32 atomics) without any
// transposed copies.  Replaces the nn.Linear calls of model_tiny_gpt.py:85-93,132,143-147,51-57,235-239.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cgpt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 16;                      // four per TMEM lane quadrant; they split the column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp0 TMA, warp1 MMA, warps 2..17 epilogue
constexpr int kStgBytesPerWarp = 32 * 64;          // 32 rows x 64 B (32 bf16 or 16 fp32), swizzled 16-byte cells

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, split_k, kb_total, kb_per_split;
  const float* bias;
  int epilogue;
  const __nv_bfloat16* aux;
  __nv_bfloat16* aux_out;
  long long ldaux;
  const float* residual;
  void* out;
  int out_f32;
  int accumulate;
  long long ldc;
  float* colsum;
  int tma_out;  // bit0: out leaves through tmC (store / reduce-add); bit1: tmX is valid (aux_out store, aux load
                // or residual load)
  int debug;  // bit0: skip global stores, bit1: skip TMEM loads, bit2: skip the whole epilogue body (probe only)
};

// gelu(x) = x*Phi(x) and gelu'(x) = Phi(x) + x*phi(x) with ONE special-function op (the exp of phi):
// Phi(-|x|) = mills(|x|) * phi(x), mills(t) = (1-Phi(t))/phi(t) as a degree-7 polynomial on [0, 5.5] fitted for
// relative error (6.9e-4 in fp32 Horner form: |dPhi| <= 3.3e-4, |dgelu| <= 1.1e-4, an order of magnitude below the
// bf16 rounding of the two outputs, and uniform in the tails); beyond 5.5 phi underflows the result anyway.
// The epilogue of the K=512 GEMMs is the long pole of those kernels (4 epilogue warps per scheduler, stalled on
// fixed-latency dependencies), so two columns are evaluated per instruction with packed fp32
// (FFMA2/FMUL2/FADD2), 1/sqrt(2 pi) rides in the exponent, and NP pairs advance in lock step.
// Returns the results packed to bf16x2 (low half = first column of the pair).
template <int NP>
__device__ __forceinline__ void gelu_and_grad2(const uint64_t (&x)[NP], uint32_t (&y_bf)[NP], uint32_t (&dy_bf)[NP]) {
  uint64_t t[NP], m[NP], phi[NP], cdf[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    float x0, x1;
    f2_unpack(x[k], x0, x1);
    t[k] = f2_pack(fminf(fabsf(x0), 5.5f), fminf(fabsf(x1), 5.5f));
  }
  // log2 of phi(x) = -x^2/2 * log2(e) + log2(1/sqrt(2 pi))
#pragma unroll
  for (int k = 0; k < NP; ++k) phi[k] = f2_mul(x[k], x[k]);
#pragma unroll
  for (int k = 0; k < NP; ++k) phi[k] = f2_fma(phi[k], f2_splat(-0.72134752f), f2_splat(-1.32574806f));
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    float a0, a1, e0, e1;
    f2_unpack(phi[k], a0, a1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    phi[k] = f2_pack(e0, e1);
  }
  // m = -mills(t): the sign is folded into the coefficients so that 1/2 - Phi(-|x|) is one FMA below
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(t[k], f2_splat(4.36487171e-05f), f2_splat(-0.00106563710f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(0.0110329464f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(-0.0638432875f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(0.231174618f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(-0.563592315f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(0.983467042f));
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], t[k], f2_splat(-1.25250721f));
  // Phi(x) = 1/2 + copysign(1/2 - Phi(-|x|), x); 1/2 - h >= 0, so the sign is a plain OR
#pragma unroll
  for (int k = 0; k < NP; ++k) m[k] = f2_fma(m[k], phi[k], f2_splat(0.5f));
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    float q0, q1, x0, x1;
    f2_unpack(m[k], q0, q1);
    f2_unpack(x[k], x0, x1);
    q0 = __uint_as_float(__float_as_uint(q0) | (__float_as_uint(x0) & 0x80000000u));
    q1 = __uint_as_float(__float_as_uint(q1) | (__float_as_uint(x1) & 0x80000000u));
    cdf[k] = f2_add(f2_pack(q0, q1), f2_splat(0.5f));
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    float y0, y1, d0, d1;
    f2_unpack(f2_mul(x[k], cdf[k]), y0, y1);
    f2_unpack(f2_fma(x[k], phi[k], cdf[k]), d0, d1);
    y_bf[k] = pack_bf16(y0, y1);
    dy_bf[k] = pack_bf16(d0, d1);
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// staging cell (16 bytes) of row `row` (64-byte rows), logical cell 0..3.  Swizzled so that both access
// patterns are conflict-free: "lane = row" 16-byte writes and "4 lanes per row, 8 rows" 16-byte reads.
__device__ __forceinline__ uint32_t stg_cell(uint32_t base, int row, int cell) {
  return base + row * 64 + ((cell ^ ((row >> 1) & 3)) << 4);
}

// bf16 outputs, 32 accumulator columns per step: bias / GELU in the accumulator's row layout (thread = row),
// pack to bf16, transpose 32 x 32 through 2 KB of smem, 16-byte coalesced stores (4 lanes per 64-byte segment).
template <bool GELU>
__device__ __forceinline__ void epi_chunk_bf16(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                               uint32_t taddr, uint32_t stg, int lane, bool leader, int row0, int n0,
                                               uint32_t bias_s /* smem address of this chunk's 32 bias floats */) {
  uint32_t r[32];
  if (!(p.debug & 2)) {
    tmem_ld32(taddr, r);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0x3f800000u + lane + i;
  }
  const bool vec_ok = (n0 + 32 <= p.N) && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                      (!GELU || p.aux_out == nullptr ||
                       (((p.ldaux & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0)));
  tmem_ld_wait();
  uint32_t act[16], dact[GELU ? 16 : 1];
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // 8 columns = 4 packed pairs per step
    const uint4 b0 = lds128(bias_s + j * 32), b1 = lds128(bias_s + j * 32 + 16);  // broadcast reads
    const uint32_t bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint64_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v[e] = f2_add(f2_pack(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1])),
                    f2_pack(__uint_as_float(bv[2 * e]), __uint_as_float(bv[2 * e + 1])));
    if constexpr (GELU) {
      uint32_t y4[4], d4[4];
      gelu_and_grad2<4>(v, y4, d4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        act[4 * j + e] = y4[e];
        dact[4 * j + e] = d4[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float o0, o1;
        f2_unpack(v[e], o0, o1);
        act[4 * j + e] = pack_bf16(o0, o1);
      }
    }
  }
  const int rl_in = lane >> 2, cell = lane & 3;
#pragma unroll
  for (int pass = 0; pass < (GELU ? 2 : 1); ++pass) {
    const uint32_t* src = (GELU && pass == 1) ? dact : act;
    __nv_bfloat16* dst = (GELU && pass == 1) ? p.aux_out : reinterpret_cast<__nv_bfloat16*>(p.out);
    const long long ld = (GELU && pass == 1) ? p.ldaux : p.ldc;
    if (dst == nullptr) break;
    if (p.tma_out & 1) {  // the staging buffer may still be feeding the previous bulk store
      if (leader) bulk_wait_read0();
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stg_cell(stg, lane, j), make_uint4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]));
    if (p.tma_out & 1) {
      // staging layout == CU_TENSOR_MAP_SWIZZLE_64B of a {32 col, 32 row} bf16 box: one asynchronous bulk store,
      // clipped by the hardware at the M / N edges; the warp moves on immediately
      fence_proxy_async_smem();
      __syncwarp();
      if (leader && !(p.debug & 1)) {
        tma_store_2d((GELU && pass == 1) ? tmX : tmC, stg, n0, row0);
        bulk_commit();
      }
      continue;
    }
    __syncwarp();
    if (vec_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int rl = g * 8 + rl_in;
        const long long row = row0 + rl;
        const uint4 v = lds128(stg_cell(stg, rl, cell));
        if (row < p.M && !(p.debug & 1)) *reinterpret_cast<uint4*>(dst + row * ld + n0 + cell * 8) = v;
      }
    } else {
      for (int g = 0; g < 4; ++g) {
        const int rl = g * 8 + rl_in;
        const long long row = row0 + rl;
        const uint4 v = lds128(stg_cell(stg, rl, cell));
        if (row >= p.M) continue;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = n0 + cell * 8 + e;
          const uint32_t w = e < 2 ? v.x : (e < 4 ? v.y : (e < 6 ? v.z : v.w));
          const float2 f = unpack_bf16(w);
          if (col < p.N) dst[row * ld + col] = __float2bfloat16_rn((e & 1) ? f.y : f.x);
        }
      }
    }
    __syncwarp();
  }
}

// fp32 outputs (residual add, plain store, split-K atomics) and the multiply-by-aux epilogue: 16 accumulator
// columns per step staged as fp32; afterwards lane l owns 4 consecutive columns of rows (l>>2) + 8g, and the
// per-element global loads (residual / aux) of all 4 row groups are issued before the accumulator is read.
__device__ __forceinline__ void epi_chunk_f32(const GemmParams& p, uint32_t taddr, uint32_t stg, int lane, bool leader, int row0,
                                              int n0, uint32_t bias_s /* smem address of this chunk's 16 bias floats */) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
  const int r_in = lane >> 2, cq = lane & 3;
  const int col = n0 + cq * 4;
  const bool out16 = !p.out_f32;
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                      (p.residual == nullptr || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) &&
                      (p.aux == nullptr || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 7) == 0)));
  const bool full = vec_ok && (col + 3 < p.N);
  uint2 aux4[4];
  float4 res4[4];
  if (full && p.epilogue == CGPT_EPI_MUL_AUX) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const long long row = row0 + g * 8 + r_in;
      aux4[g] = row < p.M ? *reinterpret_cast<const uint2*>(p.aux + row * p.ldaux + col) : make_uint2(0u, 0u);
    }
  }
  if (full && p.residual) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const long long row = row0 + g * 8 + r_in;
      res4[g] = row < p.M ? *reinterpret_cast<const float4*>(p.residual + row * p.ldc + col)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const uint4 bu = lds128(bias_s + cq * 16);
  const float4 b4 = make_float4(__uint_as_float(bu.x), __uint_as_float(bu.y), __uint_as_float(bu.z),
                                __uint_as_float(bu.w));
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) sts128(stg_cell(stg, lane, j), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
  __syncwarp();
  if (full) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int rl = g * 8 + r_in;
      const long long row = row0 + rl;
      const uint4 u = lds128(stg_cell(stg, rl, cq));
      float4 v = make_float4(__uint_as_float(u.x) + b4.x, __uint_as_float(u.y) + b4.y, __uint_as_float(u.z) + b4.z,
                             __uint_as_float(u.w) + b4.w);
      if (row >= p.M) continue;
      if (p.epilogue == CGPT_EPI_MUL_AUX) {
        const float2 a0 = unpack_bf16(aux4[g].x), a1 = unpack_bf16(aux4[g].y);
        v.x *= a0.x;
        v.y *= a0.y;
        v.z *= a1.x;
        v.w *= a1.y;
      }
      if (out16) {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + col) =
            make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
      } else {
        float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + col;
        if (p.residual) {
          v.x += res4[g].x;
          v.y += res4[g].y;
          v.z += res4[g].z;
          v.w += res4[g].w;
        }
        if (p.accumulate)
          red_add_v4(o, v);
        else
          *reinterpret_cast<float4*>(o) = v;
      }
    }
  } else {
    // scalar path: column tails and pitches that are not multiples of 4
    for (int g = 0; g < 4; ++g) {
      const int rl = g * 8 + r_in;
      const long long row = row0 + rl;
      const uint4 u = lds128(stg_cell(stg, rl, cq));
      if (row >= p.M) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col + e >= p.N) continue;
        float v = __uint_as_float(e == 0 ? u.x : (e == 1 ? u.y : (e == 2 ? u.z : u.w))) +
                  (e == 0 ? b4.x : (e == 1 ? b4.y : (e == 2 ? b4.z : b4.w)));
        if (p.epilogue == CGPT_EPI_MUL_AUX) v *= __bfloat162float(p.aux[row * p.ldaux + col + e]);
        if (out16) {
          reinterpret_cast<__nv_bfloat16*>(p.out)[row * p.ldc + col + e] = __float2bfloat16_rn(v);
        } else {
          float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + col + e;
          if (p.residual) v += p.residual[row * p.ldc + col + e];
          if (p.accumulate)
            atomicAdd(o, v);
          else
            *o = v;
        }
      }
    }
  }
  __syncwarp();
}

// MUL_AUX with prefetched aux (BN = 256, 3-stage variant): the aux piece already sits (or is about to land) in
// this chunk's own staging buffer; the product is written in place and leaves as one TMA bulk store.  `bar` is the
// warp's mbarrier for the first chunk of a tile (it covers the loads of both chunks), nullptr for the second.
__device__ __forceinline__ void epi_chunk_mulaux_pf(const GemmParams& p, const CUtensorMap* tmC, uint32_t taddr,
                                                    uint32_t stg, int lane, bool leader, int row0, int n0, uint32_t bias_s,
                                                    uint64_t* bar, uint32_t& phase) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  if (bar != nullptr) {
    mbar_wait(bar, phase);
    phase ^= 1;
  }
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t cell = stg_cell(stg, lane, j);
    const uint4 a = lds128(cell);
    const uint4 b0 = lds128(bias_s + j * 32), b1 = lds128(bias_s + j * 32 + 16);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w};
    const uint32_t bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint64_t acc2 = f2_add(f2_pack(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1])),
                                   f2_pack(__uint_as_float(bv[2 * e]), __uint_as_float(bv[2 * e + 1])));
      const uint64_t a2 = f2_pack(__uint_as_float(av[e] << 16), __uint_as_float(av[e] & 0xffff0000u));
      float o0, o1;
      f2_unpack(f2_mul(acc2, a2), o0, o1);
      o[e] = pack_bf16(o0, o1);
    }
    sts128(cell, make_uint4(o[0], o[1], o[2], o[3]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    tma_store_2d(tmC, stg, n0, row0);
    bulk_commit();
  }
  if (p.colsum) {
    // column sums of this 32 x 32 piece straight from the staged bf16 values (lane = column), one atomic per
    // column: the bias gradient of the layer in front of the GELU, without re-reading the output from HBM.
    // Rows past M were zero-filled by the aux load, so they add nothing.
    float acc = 0.f;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      uint16_t h;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(stg_cell(stg, rr, lane >> 3) + (lane & 7) * 2));
      acc += __uint_as_float(static_cast<uint32_t>(h) << 16);
    }
    if (n0 + lane < p.N) atomicAdd(p.colsum + n0 + lane, acc);
  }
}

// ---- fully TMA-fed epilogues: the per-element operand (aux / residual) arrives in the staging buffer through a
// bulk tensor load, is combined with the accumulator in the row layout (thread = row), written back to the same
// swizzled cells and leaves through a bulk tensor store (or reduce-add for split-K).  No per-thread global access.
__device__ __forceinline__ void epi_chunk_mulaux_tma(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                                     uint32_t taddr, uint32_t stg, int lane, bool leader, int row0, int n0,
                                                     uint32_t bias_s, uint64_t* bar, uint32_t& phase) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  if (leader) {
    bulk_wait_read0();  // the previous store has finished reading the staging buffer
    mbar_expect_tx(bar, 32 * 64);
    tma_load_2d_s(stg, tmX, bar, n0, row0);
  }
  tmem_ld_wait();
  mbar_wait(bar, phase);
  phase ^= 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t cell = stg_cell(stg, lane, j);
    const uint4 a = lds128(cell);
    const uint4 b0 = lds128(bias_s + j * 32), b1 = lds128(bias_s + j * 32 + 16);
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    uint4 o;
    o.x = pack_bf16((__uint_as_float(r[8 * j + 0]) + __uint_as_float(b0.x)) * a0.x,
                    (__uint_as_float(r[8 * j + 1]) + __uint_as_float(b0.y)) * a0.y);
    o.y = pack_bf16((__uint_as_float(r[8 * j + 2]) + __uint_as_float(b0.z)) * a1.x,
                    (__uint_as_float(r[8 * j + 3]) + __uint_as_float(b0.w)) * a1.y);
    o.z = pack_bf16((__uint_as_float(r[8 * j + 4]) + __uint_as_float(b1.x)) * a2.x,
                    (__uint_as_float(r[8 * j + 5]) + __uint_as_float(b1.y)) * a2.y);
    o.w = pack_bf16((__uint_as_float(r[8 * j + 6]) + __uint_as_float(b1.z)) * a3.x,
                    (__uint_as_float(r[8 * j + 7]) + __uint_as_float(b1.w)) * a3.y);
    sts128(cell, o);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    tma_store_2d(tmC, stg, n0, row0);
    bulk_commit();
  }
  if (p.colsum) {
    // column sums of this 32 x 32 piece straight from the staged bf16 values (lane = column), one atomic per
    // column: the bias gradient of the layer in front of the GELU, without re-reading the output from HBM.
    // Rows past M were zero-filled by the aux load, so they add nothing.
    float acc = 0.f;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      uint16_t h;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(stg_cell(stg, rr, lane >> 3) + (lane & 7) * 2));
      acc += __uint_as_float(static_cast<uint32_t>(h) << 16);
    }
    if (n0 + lane < p.N) atomicAdd(p.colsum + n0 + lane, acc);
    __syncwarp();  // everyone is done with the staging buffer before the next chunk reloads it
  }
}

__device__ __forceinline__ void epi_chunk_f32_tma(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                                  uint32_t taddr, uint32_t stg, int lane, bool leader, int row0, int n0,
                                                  uint32_t bias_s, uint64_t* bar, uint32_t& phase) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
  const bool has_res = p.residual != nullptr;
  if (leader) {
    bulk_wait_read0();
    if (has_res) {
      mbar_expect_tx(bar, 32 * 64);
      tma_load_2d_s(stg, tmX, bar, n0, row0);
    }
  }
  tmem_ld_wait();
  if (has_res) {
    mbar_wait(bar, phase);
    phase ^= 1;
  } else {
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t cell = stg_cell(stg, lane, j);
    const uint4 b = lds128(bias_s + j * 16);
    float4 v = make_float4(__uint_as_float(r[4 * j]) + __uint_as_float(b.x), __uint_as_float(r[4 * j + 1]) + __uint_as_float(b.y),
                           __uint_as_float(r[4 * j + 2]) + __uint_as_float(b.z), __uint_as_float(r[4 * j + 3]) + __uint_as_float(b.w));
    if (has_res) {
      const uint4 q = lds128(cell);
      v.x += __uint_as_float(q.x);
      v.y += __uint_as_float(q.y);
      v.z += __uint_as_float(q.z);
      v.w += __uint_as_float(q.w);
    }
    sts128(cell, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    if (p.accumulate)
      tma_reduce_add_2d(tmC, stg, n0, row0);
    else
      tma_store_2d(tmC, stg, n0, row0);
    bulk_commit();
  }
}

// bias + GELU with the two outputs (activation, GELU') staged in SEPARATE buffers (variants with two staging
// buffers per warp): every 8-column group is stored to smem as soon as it is computed, so only the accumulator row
// and one group's temporaries are live — the register budget that made ptxas serialise the chains — and the two
// bulk stores leave together without a read-completion wait between them.
__device__ __forceinline__ void epi_chunk_gelu_2buf(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                                    uint32_t taddr, uint32_t stg, int lane, bool leader, int row0, int n0,
                                                    uint32_t bias_s) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  if (leader) bulk_wait_read0();  // the previous chunk's stores have read both buffers
  __syncwarp();
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // 8 columns = 4 packed pairs per step
    const uint4 b0 = lds128(bias_s + j * 32), b1 = lds128(bias_s + j * 32 + 16);  // broadcast reads
    const uint32_t bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint64_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v[e] = f2_add(f2_pack(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1])),
                    f2_pack(__uint_as_float(bv[2 * e]), __uint_as_float(bv[2 * e + 1])));
    uint32_t y4[4], d4[4];
    gelu_and_grad2<4>(v, y4, d4);
    sts128(stg_cell(stg, lane, j), make_uint4(y4[0], y4[1], y4[2], y4[3]));
    sts128(stg_cell(stg + kStgBytesPerWarp, lane, j), make_uint4(d4[0], d4[1], d4[2], d4[3]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    tma_store_2d(tmC, stg, n0, row0);
    if (p.aux_out != nullptr) tma_store_2d(tmX, stg + kStgBytesPerWarp, n0, row0);
    bulk_commit();
  }
}

// fp32 output + fp32 residual with prefetched residual (pair mode, 3-stage variant): the residual piece already
// sits (or is about to land) in this chunk's own staging buffer; the sum is written in place and leaves as one TMA
// bulk store.  `bar` is the warp's mbarrier for the first chunk of a tile (it covers all four loads), else nullptr.
__device__ __forceinline__ void epi_chunk_f32_res_pf(const GemmParams& p, const CUtensorMap* tmC, uint32_t taddr,
                                                     uint32_t stg, int lane, bool leader, int row0, int n0, uint32_t bias_s,
                                                     uint64_t* bar, uint32_t& phase) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
  if (bar != nullptr) {
    mbar_wait(bar, phase);
    phase ^= 1;
  }
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t cell = stg_cell(stg, lane, j);
    const uint4 b = lds128(bias_s + j * 16);
    const uint4 q = lds128(cell);
    float4 v;
    v.x = __uint_as_float(r[4 * j]) + __uint_as_float(b.x) + __uint_as_float(q.x);
    v.y = __uint_as_float(r[4 * j + 1]) + __uint_as_float(b.y) + __uint_as_float(q.y);
    v.z = __uint_as_float(r[4 * j + 2]) + __uint_as_float(b.z) + __uint_as_float(q.z);
    v.w = __uint_as_float(r[4 * j + 3]) + __uint_as_float(b.w) + __uint_as_float(q.w);
    sts128(cell, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    tma_store_2d(tmC, stg, n0, row0);
    bulk_commit();
  }
}

// TWO: CTA-pair mode (cta_group::2).  The pair computes a 256 x BN tile: each CTA holds its own 128 rows of A and
// HALF of the B tile (the tensor cores of both SMs read both halves), and keeps its 128 x BN slice of the
// accumulator in its own TMEM.  Per CTA the operand traffic through shared memory drops from 48 KB to 32 KB per
// k-block, which is what the epilogue's staging traffic was competing with.
template <int BN, int STAGES, bool TWO = false>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStgOff = STAGES * kStageBytes;
  // BN = 256 with 3 stages is the MUL_AUX variant: the freed stage pays for a second staging buffer per warp, so
  // that the aux pieces of BOTH column chunks of a tile can be requested before the accumulator is ready
  static constexpr bool kAuxPrefetch = (BN == 256 && STAGES == (TWO ? 5 : 3));
  // pair mode with 3 stages is the fp32 + residual variant (proj / fc2 forward): four staging buffers per warp, so
  // that the residual pieces of ALL FOUR 16-column chunks of a tile are requested before the accumulator is ready
  static constexpr bool kResPrefetch = (BN == 256 && TWO && STAGES == 3);
  static constexpr int kStgPerWarp = (kResPrefetch ? 4 : (kAuxPrefetch ? 2 : 1)) * kStgBytesPerWarp;
  static constexpr int kBiasOff = kStgOff + kEpiWarps * kStgPerWarp;  // 2 x BN floats (per accumulator buffer)
  static constexpr int kBarOff = kBiasOff + 2 * BN * 4;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 4) * 8 + 16 + kEpiWarps * 8;  // + one mbarrier per epilogue warp
  // slack for aligning the base up to 1024 B (the kernel traps if it does not fit; in practice the dynamic
  // window starts 1 KB into the CTA's shared memory and is already aligned)
  static constexpr int kDynamic = (kTotal + 1024 <= 232448) ? kTotal + 1024 : 232448;
  static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

template <int BN, bool A_MN, bool B_MN, int STAGES, bool TWO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  using L = SmemLayout<BN, STAGES, TWO>;
  constexpr int BNL = TWO ? BN / 2 : BN;  // rows of B this CTA loads
  // persistent worker = one CTA, or one CTA pair; p.tiles_m counts 128-row (256-row for pairs) tiles
  const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
  const int unit = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_units = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) + L::kTotal > L::kDynamic) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* epi_bar = tempty_bar + 4;  // [kEpiWarps]: completion of the epilogue's own TMA loads

  // warp index through a shuffle: the compiler then knows it is warp-uniform, so everything derived from it (staging
  // addresses, tile coordinates) lives in uniform registers and the elected lane's TMA instructions are issued from
  // them directly — behind a divergent `lane == 0` test ptxas wraps every UTMASTG / UTMALDG in an
  // R2UR ... BRA.U.ANY loop
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total_work = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) {
      tma_prefetch_desc(&tmC);
      tma_prefetch_desc(&tmX);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], kEpiWarps * (TWO ? 2 : 1));  // pair mode: the leader's barrier hears both CTAs
    mbar_init(&tempty_bar[1], kEpiWarps * (TWO ? 2 : 1));
#pragma unroll
    for (int e = 0; e < kEpiWarps; ++e) mbar_init(&epi_bar[e], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (TWO) {
      tmem_alloc_2sm(tmem_slot, 2 * BN);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above ran while the previous kernel of the stream was still draining (programmatic dependent launch);
  // from here on its results are read
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = unit; w < total_work; w += n_units) {
        const int n_blk = w % p.tiles_n;
        const int m_blk = ((w / p.tiles_n) % p.tiles_m) * (TWO ? 2 : 1) + cta_rank;  // 128-row block
        const int ks = w / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int n_off = n_blk * BN + static_cast<int>(cta_rank) * BNL;  // pair mode: this CTA's half of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if constexpr (TWO) {
            // both CTAs' bytes are accounted on the LEADER's barrier, which its MMA issuer waits on
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d_2sm(sa, &tmA, fb, kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d_2sm(sa + j * (BK * 128), &tmA, fb, m_blk * BM + j * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d_2sm(sb, &tmB, fb, kb * BK, n_off);
            } else {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j)
                tma_load_2d_2sm(sb + j * (BK * 128), &tmB, fb, n_off + j * 64, kb * BK);
            }
          } else {
            mbar_expect_tx(&full_bar[s], L::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[s], m_blk * BM + j * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[s], n_blk * BN + j * 64, kb * BK);
            }
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
    if (elect_one() && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TWO ? 2 * BM : BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = unit; w < total_work; w += n_units, ++it) {
        const int ks = w / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024, kLayoutSW128)
                                     : umma_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128);
            const uint64_t bd = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024, kLayoutSW128)
                                     : umma_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128);
            if constexpr (TWO)
              umma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0) || (k > 0));
            else
              umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0) || (k > 0));
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (TWO) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator ready for the epilogue (of both CTAs of a pair)
        if constexpr (TWO) umma_commit_2sm(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    // TMEM lane = output row, so a thread first holds one row of the accumulator.  Every chunk goes through a
    // swizzled 4 KB smem transpose (explicit st/ld.shared) so that global traffic is coalesced 16-byte accesses.
    const int ew = warp - 2;
    const bool leader = elect_one();  // the one lane that issues (and waits for) this warp's bulk copies
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;  // 0..3: which quarter of the column chunks
    const uint32_t stg = smem_u32(smem + L::kStgOff + ew * L::kStgPerWarp);
    const bool bf16_rowmath = !p.out_f32 && p.epilogue != CGPT_EPI_MUL_AUX;
    const bool tma_io = (p.tma_out & 1) && (p.out_f32 || p.epilogue == CGPT_EPI_MUL_AUX);
    uint32_t epi_phase = 0;
    int it = 0;
    const uint32_t tempty_leader[2] = {TWO ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u,
                                       TWO ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0u};
    for (int w = unit; w < total_work; w += n_units, ++it) {
      const int n_blk = w % p.tiles_n;
      const int m_blk = ((w / p.tiles_n) % p.tiles_m) * (TWO ? 2 : 1) + cta_rank;  // 128-row block
      const int ks = w / (p.tiles_n * p.tiles_m);
      const int acc = it & 1;
      // this tile's bias slice -> smem (zeros when there is none), before the accumulator is even ready.
      // Double-buffered by accumulator parity; the barrier below also orders the reuse two tiles later.
      // Every warp fetches the entries of ITS column chunks into registers now and stores them once the accumulator
      // is ready (the four warps of a column group store identical values to the same words; a warp only ever reads
      // what it stored itself), so no CTA-wide barrier sits between tiles and the epilogue warps drift apart instead
      // of hitting the FMA / MUFU / shared-memory pipes in lock step.
      const uint32_t bias_tile = smem_u32(smem + L::kBiasOff) + acc * BN * 4;
      const bool legacy_sync = (p.debug & 8) != 0;
      int bcol[2];
      float bval[2] = {0.f, 0.f};
      if (legacy_sync) {
        if (ew * 32 < BN) {
          const int cb = n_blk * BN + ew * 32 + lane;
          float bv = 0.f;
          if (p.bias != nullptr && ks == 0 && cb < p.N) bv = __ldg(p.bias + cb);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_tile + (ew * 32 + lane) * 4), "f"(bv) : "memory");
        }
      } else {
        const bool chunk32 = bf16_rowmath || (tma_io && !p.out_f32);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = lane + 32 * h;
          bcol[h] = chunk32 ? (grp + 4 * h) * 32 + lane : (grp + 4 * (i >> 4)) * 16 + (i & 15);
          const int cb = n_blk * BN + bcol[h];
          if (bcol[h] < BN && p.bias != nullptr && ks == 0 && cb < p.N) bval[h] = __ldg(p.bias + cb);
        }
      }
      const int row0 = m_blk * BM + q * 32;
      if constexpr (L::kAuxPrefetch) {
        // MUL_AUX: the aux pieces of this warp's two column chunks start their trip from HBM now, while the
        // mainloop of the tile is still running (one mbarrier, armed with the bytes of both)
        if (tma_io && !p.out_f32) {
          if (leader) {
            bulk_wait_read0();  // the previous tile's stores have read both staging halves
            int nvalid = 0;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) nvalid += (n_blk * BN + (grp + 4 * ci) * 32 < p.N) ? 1 : 0;
            if (nvalid) {
              mbar_expect_tx(&epi_bar[ew], nvalid * kStgBytesPerWarp);
#pragma unroll
              for (int ci = 0; ci < 2; ++ci) {
                const int n0 = n_blk * BN + (grp + 4 * ci) * 32;
                if (n0 < p.N) tma_load_2d_s(stg + ci * kStgBytesPerWarp, &tmX, &epi_bar[ew], n0, row0);
              }
            }
          }
          __syncwarp();
        }
      }
      if constexpr (L::kResPrefetch) {
        // fp32 + residual: the residual pieces of this warp's four column chunks start their trip now
        if (tma_io && p.out_f32 && p.residual != nullptr) {
          if (leader) {
            bulk_wait_read0();  // the previous tile's stores have read all four staging buffers
            int nvalid = 0;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) nvalid += (n_blk * BN + (grp + 4 * ci) * 16 < p.N) ? 1 : 0;
            if (nvalid) {
              mbar_expect_tx(&epi_bar[ew], nvalid * kStgBytesPerWarp);
#pragma unroll
              for (int ci = 0; ci < 4; ++ci) {
                const int n0 = n_blk * BN + (grp + 4 * ci) * 16;
                if (n0 < p.N) tma_load_2d_s(stg + ci * kStgBytesPerWarp, &tmX, &epi_bar[ew], n0, row0);
              }
            }
          }
          __syncwarp();
        }
      }
      if (legacy_sync) named_bar_sync(1, kEpiWarps * 32);
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      if (!legacy_sync) {
        // safe only now: this accumulator parity's bias words were last read two tiles ago, and the MMA warp could
        // not start this tile before all sixteen warps had handed that accumulator back
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (bcol[h] < BN) asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_tile + bcol[h] * 4), "f"(bval[h]) : "memory");
        __syncwarp();
      }
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (p.debug & 4) {
      } else if (L::kResPrefetch && tma_io && p.out_f32 && p.residual != nullptr) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = grp + 4 * ci;
          const int n0 = n_blk * BN + c * 16;
          if (n0 < p.N)  // warp-uniform
            epi_chunk_f32_res_pf(p, &tmC, tbase + c * 16, stg + ci * kStgBytesPerWarp, lane, leader, row0, n0, bias_tile + c * 64,
                                 ci == 0 ? &epi_bar[ew] : nullptr, epi_phase);
        }
      } else if (L::kAuxPrefetch && tma_io && !p.out_f32) {
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = grp + 4 * ci;
          const int n0 = n_blk * BN + c * 32;
          if (n0 < p.N)  // warp-uniform
            epi_chunk_mulaux_pf(p, &tmC, tbase + c * 32, stg + ci * kStgBytesPerWarp, lane, leader, row0, n0, bias_tile + c * 128,
                                ci == 0 ? &epi_bar[ew] : nullptr, epi_phase);
        }
      } else if (tma_io && !p.out_f32) {
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 4) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N) break;  // warp-uniform
          epi_chunk_mulaux_tma(p, &tmC, &tmX, tbase + c * 32, stg, lane, leader, row0, n0, bias_tile + c * 128, &epi_bar[ew],
                               epi_phase);
        }
      } else if (tma_io) {
#pragma unroll 1
        for (int c = grp; c < BN / 16; c += 4) {
          const int n0 = n_blk * BN + c * 16;
          if (n0 >= p.N) break;  // warp-uniform
          epi_chunk_f32_tma(p, &tmC, &tmX, tbase + c * 16, stg, lane, leader, row0, n0, bias_tile + c * 64, &epi_bar[ew],
                            epi_phase);
        }
      } else if (bf16_rowmath) {
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 4) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N) break;  // warp-uniform
          if (L::kAuxPrefetch && p.epilogue == CGPT_EPI_GELU && (p.tma_out & 1) && (p.aux_out == nullptr || (p.tma_out & 2)))
            epi_chunk_gelu_2buf(p, &tmC, &tmX, tbase + c * 32, stg, lane, leader, row0, n0, bias_tile + c * 128);
          else if (p.epilogue == CGPT_EPI_GELU)
            epi_chunk_bf16<true>(p, &tmC, &tmX, tbase + c * 32, stg, lane, leader, row0, n0, bias_tile + c * 128);
          else
            epi_chunk_bf16<false>(p, &tmC, &tmX, tbase + c * 32, stg, lane, leader, row0, n0, bias_tile + c * 128);
        }
      } else {
#pragma unroll 1
        for (int c = grp; c < BN / 16; c += 4) {
          const int n0 = n_blk * BN + c * 16;
          if (n0 >= p.N) break;  // warp-uniform
          epi_chunk_f32(p, tbase + c * 16, stg, lane, leader, row0, n0, bias_tile + c * 64);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) mbar_arrive_cluster(tempty_leader[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
    }
    if ((p.tma_out & 1) && leader) bulk_wait0();  // smem must outlive the bulk stores that read it
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_2sm(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, bool A_MN, bool B_MN, int STAGES, bool TWO = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx,
           const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES, TWO>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, TWO>;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.split_k;
  if constexpr (TWO) {
    const int pairs = num_sms() / 2;
    CGPT_CHECK(launch_pdl(kern, dim3(2 * (total < pairs ? total : pairs)), dim3(kThreads), L::kDynamic, st, 2,
                          p.M > p.K ? p.M : p.K, ta, tb, tc, tx, p));
  } else {
    const int grid = total < num_sms() ? total : num_sms();
    CGPT_CHECK(launch_pdl(kern, dim3(grid), dim3(kThreads), L::kDynamic, st, 1, p.M > p.K ? p.M : p.K, ta, tb, tc, tx, p));
  }
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

template <int BN, int STAGES, bool TWO = false>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false, STAGES, TWO>(ta, tb, tc, tx, p, st);
  if (!a_mn && b_mn) return launch<BN, false, true, STAGES, TWO>(ta, tb, tc, tx, p, st);
  if (a_mn && b_mn) return launch<BN, true, true, STAGES, TWO>(ta, tb, tc, tx, p, st);
  return launch<BN, true, false, STAGES, TWO>(ta, tb, tc, tx, p, st);
}

}  // namespace
}  // namespace cgpt

extern "C" int cgpt_gemm_bf16(const cgpt_gemm_args* a, cgpt_stream_t stream) {
  using namespace cgpt;
  CGPT_REQUIRE(a && a->a && a->b && a->out, "gemm: null operand");
  CGPT_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  CGPT_REQUIRE(a->split_k >= 1, "gemm: split_k must be >= 1");
  CGPT_REQUIRE(a->split_k == 1 || (a->out_f32 && a->accumulate), "gemm: split_k>1 needs an fp32 accumulate output");
  CGPT_REQUIRE(!(a->accumulate && !a->out_f32), "gemm: accumulate needs an fp32 output");
  CGPT_REQUIRE(!(a->residual && !a->out_f32), "gemm: residual epilogue needs an fp32 output");
  CGPT_REQUIRE(a->epilogue == CGPT_EPI_NONE || a->split_k == 1, "gemm: activation epilogue with split_k");
  CGPT_REQUIRE(a->epilogue != CGPT_EPI_MUL_AUX || a->aux, "gemm: MUL_AUX needs aux");

  const int BN = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  // CTA-pair mode (cta_group::2) for the 256-wide tiles; CGPT_GEMM_2CTA=0 falls back to one CTA per tile (A/B probe)
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("CGPT_GEMM_2CTA");
    pair_env = e ? atoi(e) : 1;
  }
  const bool two = pair_env != 0 && BN == 256 && a->M > cgpt::BM && (num_sms() % 2 == 0);
  CUtensorMap ta, tb;
  int rc;
  {
    const uint64_t dimsK[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t dimsMN[2] = {(uint64_t)a->M, (uint64_t)a->K};
    const uint64_t str[1] = {(uint64_t)a->lda * 2};
    const uint32_t boxK[2] = {64, (uint32_t)cgpt::BM};
    const uint32_t boxMN[2] = {64, (uint32_t)cgpt::BK};
    rc = a->a_mn_major ? make_tmap_bf16(&ta, a->a, 2, dimsMN, str, boxMN, 128)
                       : make_tmap_bf16(&ta, a->a, 2, dimsK, str, boxK, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dimsK[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t dimsMN[2] = {(uint64_t)a->N, (uint64_t)a->K};
    const uint64_t str[1] = {(uint64_t)a->ldb * 2};
    const uint32_t boxK[2] = {64, (uint32_t)(two ? BN / 2 : BN)};  // pair mode: each CTA loads half of the B tile
    const uint32_t boxMN[2] = {64, (uint32_t)cgpt::BK};
    rc = a->b_mn_major ? make_tmap_bf16(&tb, a->b, 2, dimsMN, str, boxMN, 128)
                       : make_tmap_bf16(&tb, a->b, 2, dimsK, str, boxK, 128);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.tiles_m = two ? (a->M + 2 * cgpt::BM - 1) / (2 * cgpt::BM) : (a->M + cgpt::BM - 1) / cgpt::BM;
  p.tiles_n = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + cgpt::BK - 1) / cgpt::BK;
  int split = a->split_k < p.kb_total ? a->split_k : p.kb_total;
  p.kb_per_split = (p.kb_total + split - 1) / split;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.bias = a->bias;
  p.epilogue = a->epilogue;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(a->aux_out);
  p.ldaux = a->ldaux;
  p.residual = a->residual;
  p.out = a->out;
  p.out_f32 = a->out_f32;
  p.accumulate = a->accumulate;
  p.ldc = a->ldc;
  p.colsum = a->colsum;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("CGPT_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  // Outputs (and the per-element epilogue operands) that satisfy the TMA alignment rules move through
  // asynchronous bulk tensor copies: {32 bf16 | 16 fp32} x 32-row boxes in the 64-byte-swizzled staging layout.
  CUtensorMap tc, tx;
  memset(&tc, 0, sizeof(tc));
  memset(&tx, 0, sizeof(tx));
  p.tma_out = 0;
  {
    const int esz = a->out_f32 ? 4 : 2;
    const bool c_ok = ((reinterpret_cast<uintptr_t>(a->out) & 15) == 0) && ((a->ldc * esz) % 16 == 0);
    const void* xptr = a->epilogue == CGPT_EPI_MUL_AUX ? a->aux : (a->aux_out ? a->aux_out : (const void*)a->residual);
    const int xsz = (a->residual && a->epilogue != CGPT_EPI_MUL_AUX && !a->aux_out) ? 4 : 2;
    const long long xld = xsz == 4 ? a->ldc : a->ldaux;
    const bool x_ok = xptr == nullptr || (((reinterpret_cast<uintptr_t>(xptr) & 15) == 0) && ((xld * xsz) % 16 == 0));
    // combinations the TMA epilogues do not cover fall back to the per-thread path
    const bool combo_ok = !(a->epilogue == CGPT_EPI_MUL_AUX && (a->out_f32 || a->residual)) &&
                          !(a->out_f32 && a->epilogue != CGPT_EPI_NONE) && !(a->residual && a->accumulate);
    if (c_ok && x_ok && combo_ok) {
      const uint64_t dimsC[2] = {(uint64_t)a->N, (uint64_t)a->M};
      const uint32_t boxC[2] = {a->out_f32 ? 16u : 32u, 32u};
      const uint64_t strC[1] = {(uint64_t)a->ldc * esz};
      rc = a->out_f32 ? make_tmap_f32(&tc, a->out, 2, dimsC, strC, boxC, 64)
                      : make_tmap_bf16(&tc, a->out, 2, dimsC, strC, boxC, 64);
      if (rc) return rc;
      p.tma_out = 1;
      if (xptr) {
        const uint32_t boxX[2] = {xsz == 4 ? 16u : 32u, 32u};
        const uint64_t strX[1] = {(uint64_t)xld * xsz};
        rc = xsz == 4 ? make_tmap_f32(&tx, xptr, 2, dimsC, strX, boxX, 64)
                      : make_tmap_bf16(&tx, xptr, 2, dimsC, strX, boxX, 64);
        if (rc) return rc;
        p.tma_out |= 2;
      }
    }
  }
  CGPT_REQUIRE(a->colsum == nullptr || (a->epilogue == CGPT_EPI_MUL_AUX && (p.tma_out & 1)),
               "gemm: colsum needs the MUL_AUX epilogue on 16-byte aligned bf16 operands");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool amn = a->a_mn_major != 0, bmn = a->b_mn_major != 0;
  if (BN == 256) {
    // ×aux epilogue: one pipeline stage less + a second staging buffer per epilogue warp (aux prefetch), see SmemLayout
    const bool auxpf = a->epilogue == CGPT_EPI_MUL_AUX && (p.tma_out & 1) && !a->out_f32;
    if (two) {
      // fp32 + residual at short K (proj forward): epilogue-latency-bound, see SmemLayout::kResPrefetch
      static int respf_kmax = -1;
      if (respf_kmax < 0) {
        const char* e = getenv("CGPT_GEMM_RESPF_KMAX");
        respf_kmax = e ? atoi(e) : 1024;
      }
      if (a->out_f32 && a->residual && (p.tma_out & 1) && (p.tma_out & 2) && !a->accumulate && a->K <= respf_kmax)
        return dispatch_major<256, 3, true>(amn, bmn, ta, tb, tc, tx, p, st);
      // GELU(+GELU') shares the layout with two staging buffers per warp (one per output)
      if (auxpf || (a->epilogue == CGPT_EPI_GELU && (p.tma_out & 1) && !a->out_f32))
        return dispatch_major<256, 5, true>(amn, bmn, ta, tb, tc, tx, p, st);
      return dispatch_major<256, 6, true>(amn, bmn, ta, tb, tc, tx, p, st);
    }
    if (auxpf) return dispatch_major<256, 3>(amn, bmn, ta, tb, tc, tx, p, st);
    return dispatch_major<256, 4>(amn, bmn, ta, tb, tc, tx, p, st);
  }
  if (BN == 128) return dispatch_major<128, 6>(amn, bmn, ta, tb, tc, tx, p, st);
  return dispatch_major<64, 8>(amn, bmn, ta, tb, tc, tx, p, st);
}
