// Shared device/host helpers for the cgpt_b200 C-ABI library (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05 (TMEM + UMMA), error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/cgpt.h"

namespace cgpt {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define CGPT_CHECK(expr)                                              \
  do {                                                                \
    int _rc = ::cgpt::check_cuda((expr), #expr);                      \
    if (_rc) return _rc;                                              \
  } while (0)
#define CGPT_REQUIRE(cond, ...)                                       \
  do {                                                                \
    if (!(cond)) {                                                    \
      ::cgpt::set_error(__VA_ARGS__);                                 \
      return CGPT_ERR_INVALID;                                        \
    }                                                                 \
  } while (0)
#define CGPT_LAUNCH_CHECK() CGPT_CHECK(cudaGetLastError())

int num_sms();
void count_launch(int n = 1);
// Programmatic dependent launch for the kernels that carry pdl_wait(): CGPT_PDL=1 always, =0 never, unset: only for
// launches over at most 16384 token rows.  Measured under the captured step (same box, A/B): C4 training at B = 8
// (4096 rows) 2.52 -> 2.40 ms, C2 4.00 -> 3.92 ms; the headline shape (65536 rows) 27.57 -> 27.75 ms, hence the limit.
bool pdl_enabled(long long token_rows);

// TMA descriptor for a row-major bf16 tensor viewed as rank-`rank` (dims innermost first).
// box = tile extents (innermost first); swizzle_bytes in {0,32,64,128}.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                   int swizzle_bytes);
// same for an fp32 tensor (epilogue stores / reduce-adds)
int make_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// ---- programmatic dependent launch.  A kernel launched with launch_pdl() may have its CTAs placed (as SMs drain) and
// run its prologue — barrier init, TMEM allocation, descriptor prefetch, parameter staging — while the PREVIOUS kernel
// in the stream is still finishing; pdl_wait() then blocks until that kernel has completed and its writes are
// visible, so it must precede the first access to any global memory another kernel produces or still reads.
// Every thread of a kernel launched this way calls pdl_wait() exactly once, on every path: the completion of this
// grid is what the NEXT kernel's pdl_wait() observes, so the chain of dependencies stays transitive.  The trigger
// sits right behind the wait: at most one kernel runs ahead.  Both are no-ops under a plain launch.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              int cluster_x, long long token_rows, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled(token_rows)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100): one issue slot for two lanes of math
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_splat(float c) { return f2_pack(c, c); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (a warp that polls several barriers must not be parked on one of them)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- thread-block clusters / CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the bytes are accounted on the barrier at `bar_cluster_addr`
// (the leader CTA's), which is what the pair's single MMA issuer waits on
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[256 x N] (+)= A[256 x 16] * B[N x 16]^T over a CTA pair: rows 0..127 of A / D and the first N/2 rows of B live
// in the issuing (leader) CTA, the rest at the same smem / TMEM offsets of its peer
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// arrive on the mbarrier at this smem offset in BOTH CTAs of the pair once all prior MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ---- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA (cp.async.bulk.tensor), completion on an mbarrier
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d_s(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// smem -> global tile store / reduce-add through TMA (bulk async group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
// 1-D bulk copy global -> shared (no tensor map; 16-byte aligned addresses, size a multiple of 16), completion on
// an mbarrier.  The streaming kernels (LayerNorm) keep whole rows in flight with it instead of per-thread loads.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- UMMA (tcgen05.mma), single-CTA, bf16 x bf16 -> fp32 in TMEM
// Shared-memory operand descriptor (sm_100 "version 1"); see DESIGN.md §GEMM for the layouts.
//   K-major  SW128: rows of 128 B (64 bf16 of K), 8-row atoms 1024 B apart (SBO); LBO unused.
//   MN-major SW128: k-rows of 128 B (64 bf16 of M/N), 8-k atoms 1024 B apart (SBO),
//                   next 64-wide M/N block `lbo_bytes` away.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}
constexpr uint32_t kLayoutSW128 = 2, kLayoutSW64 = 4, kLayoutSW32 = 6, kLayoutNone = 0;

// Instruction descriptor, kind::f16: D=f32, A=B=bf16.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17)
         | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// Make all prior UMMAs of this thread arrive on an mbarrier when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers.  Warp w (w%4 = q) owns lanes 32q..32q+31; thread t gets lane 32q+t.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM (same lane / column mapping as the loads)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- Philox4x32-10 (counter-based RNG for dropout: the mask is recomputed in backward, never stored)
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint2 key, uint4 ctr) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) { return philox4x32<10>(key, ctr); }
struct DropoutCfg {
  uint32_t thresh;  // keep iff random u32 >= thresh  (thresh = p * 2^32); 0 = dropout off
  float inv_keep;   // 1 / (1 - p)
  uint2 key;        // seed
  uint32_t off_lo, off_hi;  // generator offset: separates successive calls under the same seed
  // attention-probability masks use 16 random bits per element (8 elements per Philox call): keep iff u16 >= thresh16,
  // scaled by inv_keep16 = 1 / (1 - thresh16 / 65536) so that the mask stays exactly unbiased
  uint32_t thresh16;
  float inv_keep16;
  // Device-resident generator state {seed, base offset} (cgpt_set_philox_state), or nullptr.  When set, the seed is
  // read from it and the base offset is ADDED to the by-value offset inside the kernel: a CUDA graph that captured
  // the launch draws fresh masks on every replay (the base is advanced on the device, cgpt_philox_advance).
  const unsigned long long* dev_state;
};
}  // namespace cgpt
// defined in core.cu; process-wide on purpose: backward kernels are launched from autograd's worker thread
extern const unsigned long long* volatile cgpt_philox_dev_state;
namespace cgpt {
__host__ inline DropoutCfg make_dropout(float p, uint64_t seed, uint64_t offset) {
  DropoutCfg d;
  const double t = (double)p * 4294967296.0;
  d.thresh = p <= 0.f ? 0u : (t >= 4294967295.0 ? 4294967295u : (uint32_t)t);
  d.inv_keep = p < 1.f ? 1.f / (1.f - p) : 0.f;
  d.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  d.off_lo = (uint32_t)offset;
  d.off_hi = (uint32_t)(offset >> 32);
  d.thresh16 = d.thresh >> 16;
  d.inv_keep16 = 65536.f / (65536.f - (float)d.thresh16);
  d.dev_state = d.thresh ? cgpt_philox_dev_state : nullptr;
  return d;
}
// the configuration a kernel works with: seed / offset resolved against the device-resident state, if any
__device__ __forceinline__ DropoutCfg resolve_dropout(DropoutCfg d) {
  if (d.dev_state != nullptr) {
    const unsigned long long seed = d.dev_state[0];
    const unsigned long long off = d.dev_state[1] + ((static_cast<unsigned long long>(d.off_hi) << 32) | d.off_lo);
    d.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    d.off_lo = static_cast<uint32_t>(off);
    d.off_hi = static_cast<uint32_t>(off >> 32);
  }
  return d;
}
// Attention-probability dropout: keep-mask of the 32 consecutive key positions jb .. jb+31 (jb % 32 == 0) of query
// row i, head bh; bit e set = keep position jb+e; 16 random bits per position.  ONE Philox4x32-7 call per (row, 32-key
// chunk) — 7 rounds pass BigCrush; cuRAND's 10 are a safety margin — gives 128 bits: the first 8 positions use them
// directly, the other 24 use twelve more words expanded from them by a bijective 32-bit finaliser (xor-shift /
// multiply, full avalanche) under three different whitening constants.  Four Philox calls per chunk cost more integer
// work than the whole softmax of the chunk (~10 against ~4 instructions per score); this is ~5.  Forward, backward and
// the dense probability kernel all go through these functions, so they see the same mask.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint4 attn_dropout_seed(const DropoutCfg& d, uint32_t bh, uint32_t i, uint32_t j32) {
  return philox4x32<7>(make_uint2(d.key.x ^ d.off_lo, d.key.y ^ d.off_hi), make_uint4(j32, i, bh, 0x61747400u));
}
// the 4 words (8 positions) of octet q = 0..3 of a chunk
__device__ __forceinline__ uint4 attn_dropout_octet(const uint4 s, uint32_t q) {
  if (q == 0) return s;
  const uint32_t w = 0x9E3779B9u * q;
  return make_uint4(mix32(s.x ^ w), mix32(s.y ^ w), mix32(s.z ^ w), mix32(s.w ^ w));
}
__device__ __forceinline__ uint32_t keep_bits8(const uint4 r, uint32_t t) {
  return (uint32_t)((r.x & 0xffffu) >= t) | ((uint32_t)((r.x >> 16) >= t) << 1) | ((uint32_t)((r.y & 0xffffu) >= t) << 2) |
         ((uint32_t)((r.y >> 16) >= t) << 3) | ((uint32_t)((r.z & 0xffffu) >= t) << 4) | ((uint32_t)((r.z >> 16) >= t) << 5) |
         ((uint32_t)((r.w & 0xffffu) >= t) << 6) | ((uint32_t)((r.w >> 16) >= t) << 7);
}
__device__ __forceinline__ uint32_t attn_keep_mask32(const DropoutCfg& d, uint32_t bh, uint32_t i, uint32_t jb) {
  const uint4 s = attn_dropout_seed(d, bh, i, jb >> 5);
  uint32_t m = 0;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) m |= keep_bits8(attn_dropout_octet(s, q), d.thresh16) << (8 * q);
  return m;
}
// the same mask as 32 multipliers (1 / (1 - p) or 0), octet by octet: a 16-bit field goes straight to a compare and a
// select — no keep-mask word is assembled and no bit of it is tested again
__device__ __forceinline__ void attn_keep_scale8(const uint4 r, uint32_t t, float inv, float (&mk)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mk[2 * k] = (w[k] & 0xffffu) >= t ? inv : 0.f;
    mk[2 * k + 1] = (w[k] >> 16) >= t ? inv : 0.f;
  }
}
__device__ __forceinline__ bool attn_keep(const DropoutCfg& d, uint32_t bh, uint32_t i, uint32_t j) {
  const uint4 s = attn_dropout_seed(d, bh, i, j >> 5);
  return (keep_bits8(attn_dropout_octet(s, (j >> 3) & 3u), d.thresh16) >> (j & 7u)) & 1u;
}

// ---- small math
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

#endif  // __CUDACC__
}  // namespace cgpt
