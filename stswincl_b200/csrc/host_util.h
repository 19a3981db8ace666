// Host-side helpers shared by the launchers: error reporting for the C ABI,
// CUtensorMap construction through the driver entry point (no libcuda link-time
// dependency), device attribute cache.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stswin {

// Error codes of the C ABI (include/stswin_b200.h)
enum : int {
  kOk = 0,
  kErrInvalidArg = -1,
  kErrUnsupported = -2,
  kErrCuda = -3,
  kErrDriver = -4,
};

int set_error(int code, const char* fmt, ...);
const char* last_error();

#define STSWIN_CHECK_ARG(cond, ...)                                   \
  do {                                                                \
    if (!(cond)) return ::stswin::set_error(::stswin::kErrInvalidArg, __VA_ARGS__); \
  } while (0)

#define STSWIN_CUDA(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::stswin::set_error(::stswin::kErrCuda, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                                  \
  } while (0)

int num_sms();   // of the current device

enum class TmapDtype { BF16, F32, U8 };

// Tiled tensor map over `rank` dims (innermost first). strides_bytes has rank-1 entries
// (stride of dim 1.., the innermost is dense). swizzle128: 128-byte swizzle, else none (swizzle64: 64-byte swizzle instead).
// Returns kOk or a negative error (message in last_error()).
int make_tmap(CUtensorMap* out, TmapDtype dt, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool swizzle64 = false);

// Launch with programmatic dependent launch allowed (the kernel MUST call pdl_wait() before it touches anything a
// preceding kernel of the stream wrote).  STSWIN_PDL=0 in the environment falls back to plain serialised launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace stswin
