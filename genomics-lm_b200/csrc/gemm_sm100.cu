// Dense bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double-buffered) -> fused epilogue.  Persistent, warp-specialised:
//   warp 0      TMA producer            (one elected lane)
//   warp 1      TMEM owner + MMA issuer (one elected lane)
//   warps 2..9  epilogue: tcgen05.ld -> swizzled smem transpose -> bias/GELU/aux/residual -> coalesced 16-B stores
// Tile 128 x BN x 64, BN in {64,128,256}.  Operands are K-major ([rows,K]) or MN-major ([K,rows]),
// which covers forward (x·Wᵀ), dgrad (dy·W) and wgrad (dyᵀ·x, split-K with fp32 atomics) without any
// transposed copies.  Replaces the nn.Linear calls of model_tiny_gpt.py:85-93,132,143-147,51-57,235-239.
#include "common.cuh"

namespace cgpt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;                       // two per TMEM lane quadrant; they split the column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int kChunk = 32;                         // accumulator columns staged per step
constexpr int kStgBytesPerWarp = 32 * kChunk * 4;  // 32 rows x 128 B, XOR-swizzled 16-byte cells

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
};

// gelu(x) = x*Phi(x) and gelu'(x) = Phi(x) + x*phi(x) from one exp and one reciprocal
// (Abramowitz-Stegun 7.1.26 for erfc, |error| < 1.5e-7: far below the bf16 output rounding).
__device__ __forceinline__ void gelu_and_grad(float x, float& y, float& dy) {
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, fabsf(x), 1.0f)));  // 0.3275911/sqrt(2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752f * x * x));               // exp(-x^2/2)
  float p = fmaf(t, 0.5307027145f, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float h = p * t * e;  // Phi(-|x|)
  const float cdf = x >= 0.f ? 1.f - h : h;
  y = x * cdf;
  dy = fmaf(x, 0.3989422804f * e, cdf);
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStgOff = STAGES * kStageBytes;
  static constexpr int kBarOff = kStgOff + kEpiWarps * kStgBytesPerWarp;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 4) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024-B alignment
  static_assert(kDynamic <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], kEpiWarps);
    mbar_init(&tempty_bar[1], kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int n_blk = w % p.tiles_n;
        const int m_blk = (w / p.tiles_n) % p.tiles_m;
        const int ks = w / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
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
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
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
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0) || (k > 0));
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    // TMEM lane = output row, so a thread first holds one row x 32 columns.  The chunk goes through a
    // swizzled smem transpose so that global traffic is coalesced: afterwards lane l owns 4 consecutive
    // columns (l&7) of row 4*it + (l>>3); all bias / activation / aux / residual work happens there, with the
    // global loads of all 8 row groups issued before the first use.
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;  // which half of the column chunks
    uint8_t* stg = smem + L::kStgOff + ew * kStgBytesPerWarp;
    const int r_in = lane >> 3, cq = lane & 7;
    const bool out16 = !p.out_f32;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                        (p.residual == nullptr || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) &&
                        ((p.aux == nullptr && p.aux_out == nullptr) ||
                         (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 7) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.aux_out) & 7) == 0)));
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int n_blk = w % p.tiles_n;
      const int m_blk = (w / p.tiles_n) % p.tiles_m;
      const int ks = w / (p.tiles_n * p.tiles_m);
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      const int row0 = m_blk * BM + q * 32;
      const bool add_bias = (p.bias != nullptr) && (ks == 0);
#pragma unroll 1
      for (int c = grp; c < BN / kChunk; c += 2) {
        const int n0 = n_blk * BN + c * kChunk;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * kChunk, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int col = n0 + cq * 4;
        const bool full = vec_ok && (col + 3 < p.N);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add_bias) {
          if (col + 3 < p.N) {
            b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          } else {
            if (col < p.N) b4.x = __ldg(p.bias + col);
            if (col + 1 < p.N) b4.y = __ldg(p.bias + col + 1);
            if (col + 2 < p.N) b4.z = __ldg(p.bias + col + 2);
          }
        }
        if (full) {
          // ---------------- vector path
          uint2 aux8[8];
          float4 res8[8];
          if (p.epilogue == CGPT_EPI_MUL_AUX) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const long long row = row0 + g * 4 + r_in;
              aux8[g] = row < p.M ? *reinterpret_cast<const uint2*>(p.aux + row * p.ldaux + col) : make_uint2(0u, 0u);
            }
          }
          if (p.residual) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const long long row = row0 + g * 4 + r_in;
              res8[g] = row < p.M ? *reinterpret_cast<const float4*>(p.residual + row * p.ldc + col)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int rl = g * 4 + r_in;
            const long long row = row0 + rl;
            float4 v = *reinterpret_cast<const float4*>(stg + rl * 128 + ((cq ^ (rl & 7)) << 4));
            v.x += b4.x;
            v.y += b4.y;
            v.z += b4.z;
            v.w += b4.w;
            if (row >= p.M) continue;
            if (p.epilogue == CGPT_EPI_GELU) {
              float4 d;
              gelu_and_grad(v.x, v.x, d.x);
              gelu_and_grad(v.y, v.y, d.y);
              gelu_and_grad(v.z, v.z, d.z);
              gelu_and_grad(v.w, v.w, d.w);
              if (p.aux_out)
                *reinterpret_cast<uint2*>(p.aux_out + row * p.ldaux + col) =
                    make_uint2(pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
            } else if (p.epilogue == CGPT_EPI_MUL_AUX) {
              const float2 a0 = unpack_bf16(aux8[g].x), a1 = unpack_bf16(aux8[g].y);
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
                v.x += res8[g].x;
                v.y += res8[g].y;
                v.z += res8[g].z;
                v.w += res8[g].w;
              }
              if (p.accumulate)
                red_add_v4(o, v);
              else
                *reinterpret_cast<float4*>(o) = v;
            }
          }
        } else {
          // ---------------- scalar path: column tails and pitches that are not multiples of 4
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          for (int g = 0; g < 8; ++g) {
            const int rl = g * 4 + r_in;
            const long long row = row0 + rl;
            if (row >= p.M) continue;
            const float4 v4 = *reinterpret_cast<const float4*>(stg + rl * 128 + ((cq ^ (rl & 7)) << 4));
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
            for (int e = 0; e < 4; ++e) {
              if (col + e >= p.N) break;
              float v = vv[e] + bb[e];
              if (p.epilogue == CGPT_EPI_GELU) {
                float d;
                gelu_and_grad(v, v, d);
                if (p.aux_out) p.aux_out[row * p.ldaux + col + e] = __float2bfloat16_rn(d);
              } else if (p.epilogue == CGPT_EPI_MUL_AUX) {
                v *= __bfloat162float(p.aux[row * p.ldaux + col + e]);
              }
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES>;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, kThreads, L::kDynamic, st>>>(ta, tb, p);
  count_launch();
  CGPT_LAUNCH_CHECK();
  return 0;
}

template <int BN, int STAGES>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                   cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false, STAGES>(ta, tb, p, st);
  if (!a_mn && b_mn) return launch<BN, false, true, STAGES>(ta, tb, p, st);
  if (a_mn && b_mn) return launch<BN, true, true, STAGES>(ta, tb, p, st);
  return launch<BN, true, false, STAGES>(ta, tb, p, st);
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
    const uint32_t boxK[2] = {64, (uint32_t)BN};
    const uint32_t boxMN[2] = {64, (uint32_t)cgpt::BK};
    rc = a->b_mn_major ? make_tmap_bf16(&tb, a->b, 2, dimsMN, str, boxMN, 128)
                       : make_tmap_bf16(&tb, a->b, 2, dimsK, str, boxK, 128);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.tiles_m = (a->M + cgpt::BM - 1) / cgpt::BM;
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool amn = a->a_mn_major != 0, bmn = a->b_mn_major != 0;
  if (BN == 256) return dispatch_major<256, 4>(amn, bmn, ta, tb, p, st);
  if (BN == 128) return dispatch_major<128, 6>(amn, bmn, ta, tb, p, st);
  return dispatch_major<64, 8>(amn, bmn, ta, tb, p, st);
}
