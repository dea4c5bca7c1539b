// Dense bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double-buffered) -> fused epilogue.  Persistent, warp-specialised:
//   warp 0      TMA producer            (one elected lane)
//   warp 1      TMEM owner + MMA issuer (one elected lane)
//   warps 2..5  epilogue: tcgen05.ld -> smem transpose -> bias/GELU/GELU'/residual -> coalesced stores
// Tile 128 x BN x 64, BN in {64,128,256}.  Operands are K-major ([rows,K]) or MN-major ([K,rows]),
// which covers forward (x·Wᵀ), dgrad (dy·W) and wgrad (dyᵀ·x, split-K with fp32 atomics) without any
// transposed copies.  Replaces the nn.Linear calls of model_tiny_gpt.py:85-93,132,143-147,51-57,235-239.
#include "common.cuh"

namespace cgpt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
constexpr int kStgStride = 66;                                // floats per staged row (64 + 2 pad)
constexpr int kStgBytesPerWarp = 32 * kStgStride * 4;         // 8448

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

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStgOff = STAGES * kStageBytes;
  static constexpr int kBarOff = kStgOff + kEpiWarps * kStgBytesPerWarp;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 4) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024-B alignment
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
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    float* stg = reinterpret_cast<float*>(smem + L::kStgOff + (warp - 2) * kStgBytesPerWarp);
    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 7) == 0);
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
      for (int c = 0; c < BN / 64; ++c) {
        const int col = n_blk * BN + c * 64 + 2 * lane;
        if (n_blk * BN + c * 64 >= p.N) break;  // warp-uniform
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 64;
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          *reinterpret_cast<uint2*>(&stg[lane * kStgStride + j]) = make_uint2(r0[j], r0[j + 1]);
          *reinterpret_cast<uint2*>(&stg[lane * kStgStride + 32 + j]) = make_uint2(r1[j], r1[j + 1]);
        }
        __syncwarp();
        const bool c0ok = col < p.N, c1ok = col + 1 < p.N;
        float b0 = 0.f, b1 = 0.f;
        if (add_bias) {
          if (c0ok) b0 = __ldg(p.bias + col);
          if (c1ok) b1 = __ldg(p.bias + col + 1);
        }
        const int rmax = min(32, p.M - row0);
        for (int rr = 0; rr < rmax; ++rr) {
          const long long row = row0 + rr;
          float2 v = *reinterpret_cast<const float2*>(&stg[rr * kStgStride + 2 * lane]);
          v.x += b0;
          v.y += b1;
          if (!c0ok) continue;
          if (p.epilogue == CGPT_EPI_GELU) {
            if (p.aux_out) {
              __nv_bfloat16* ao = p.aux_out + row * p.ldaux + col;
              if (c1ok)
                *reinterpret_cast<uint32_t*>(ao) = pack_bf16(v.x, v.y);
              else
                ao[0] = __float2bfloat16_rn(v.x);
            }
            v.x = gelu_erf(v.x);
            v.y = gelu_erf(v.y);
          } else if (p.epilogue == CGPT_EPI_GELU_GRAD) {
            const __nv_bfloat16* ai = p.aux + row * p.ldaux + col;
            float a0, a1 = 0.f;
            if (c1ok) {
              float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(ai));
              a0 = a.x;
              a1 = a.y;
            } else {
              a0 = __bfloat162float(ai[0]);
            }
            v.x *= gelu_erf_grad(a0);
            v.y *= gelu_erf_grad(a1);
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + col;
            if (p.residual) {
              const float* rs = p.residual + row * p.ldc + col;
              if (c1ok && vec_ok) {
                float2 r = *reinterpret_cast<const float2*>(rs);
                v.x += r.x;
                v.y += r.y;
              } else {
                v.x += rs[0];
                if (c1ok) v.y += rs[1];
              }
            }
            if (p.accumulate) {
              atomicAdd(o, v.x);
              if (c1ok) atomicAdd(o + 1, v.y);
            } else if (c1ok && vec_ok) {
              *reinterpret_cast<float2*>(o) = v;
            } else {
              o[0] = v.x;
              if (c1ok) o[1] = v.y;
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + col;
            if (c1ok && vec_ok)
              *reinterpret_cast<uint32_t*>(o) = pack_bf16(v.x, v.y);
            else {
              o[0] = __float2bfloat16_rn(v.x);
              if (c1ok) o[1] = __float2bfloat16_rn(v.y);
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
  CGPT_REQUIRE(a->epilogue != CGPT_EPI_GELU_GRAD || a->aux, "gemm: GELU_GRAD needs aux");

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
