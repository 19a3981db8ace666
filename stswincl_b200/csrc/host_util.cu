#include "host_util.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace stswin {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("STSWIN_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, TmapDtype dt, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool swizzle64) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(kErrDriver, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (rank < 1 || rank > 5) return set_error(kErrInvalidArg, "tensor map rank %d out of range", rank);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(kErrInvalidArg, "tensor map base %p is not 16-byte aligned", base);
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0)
        return set_error(kErrInvalidArg, "tensor map stride %llu (dim %d) is not a multiple of 16 bytes",
                         (unsigned long long)gstr[i - 1], i);
    }
    if (box[i] == 0 || box[i] > 256) return set_error(kErrInvalidArg, "tensor map box[%d]=%u out of range", i, box[i]);
  }
  CUtensorMapDataType cdt = dt == TmapDtype::BF16  ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                            : dt == TmapDtype::F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                   : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(out, cdt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(kErrDriver,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box "
                     "[%u,%u,%u,%u,%u]",
                     (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                     (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                     (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0,
                     rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
  return kOk;
}

}  // namespace stswin
