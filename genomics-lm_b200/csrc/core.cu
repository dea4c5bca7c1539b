// Library plumbing: error text, device check, TMA descriptor encoding, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

const unsigned long long* volatile cgpt_philox_dev_state = nullptr;

namespace cgpt {

__global__ void philox_advance_kernel(unsigned long long* state, unsigned long long inc) { state[1] += inc; }

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CGPT_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled(long long token_rows) {
  static const int mode = [] {
    const char* e = getenv("CGPT_PDL");
    return e ? (e[0] == '1' ? 1 : 0) : -1;
  }();
  return mode < 0 ? token_rows <= 16384 : mode == 1;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

int make_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

static int make_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CGPT_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA operand base %p is not 16-byte aligned", base);
    return CGPT_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0) {
      set_error("TMA operand pitch %llu bytes is not a multiple of 16", (unsigned long long)gstr[i]);
      return CGPT_ERR_INVALID;
    }
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                   bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu] box [%u,%u] pitch %llu", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return CGPT_ERR_CUDA;
  }
  return 0;
}

}  // namespace cgpt

extern "C" {

int cgpt_version(void) { return CGPT_VERSION; }

const char* cgpt_last_error(void) { return cgpt::g_err; }

int64_t cgpt_launch_count(void) { return cgpt::g_launches.load(); }

int cgpt_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cgpt::check_cuda(e, "cudaSetDevice");
  return 0;
}

int cgpt_set_philox_state(const uint64_t* dev_state) {
  cgpt_philox_dev_state = reinterpret_cast<const unsigned long long*>(dev_state);
  return 0;
}

int cgpt_philox_advance(uint64_t* dev_state, uint64_t increment, cgpt_stream_t stream) {
  if (dev_state == nullptr) {
    cgpt::set_error("cgpt_philox_advance: null state");
    return CGPT_ERR_INVALID;
  }
  cgpt::philox_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<unsigned long long*>(dev_state), static_cast<unsigned long long>(increment));
  cgpt::count_launch(1);
  return cgpt::check_cuda(cudaGetLastError(), "philox_advance_kernel");
}

int cgpt_device_ok(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cgpt::check_cuda(e, "cudaGetDeviceProperties");
  if (prop.major != 10 || prop.minor != 0) {
    cgpt::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    return CGPT_ERR_DEVICE;
  }
  if (!cgpt::get_encode()) {
    cgpt::set_error("driver does not export cuTensorMapEncodeTiled");
    return CGPT_ERR_CUDA;
  }
  return 0;
}

}  // extern "C"
