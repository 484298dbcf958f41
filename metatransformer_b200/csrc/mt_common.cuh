// mt_common.cuh — host-side helpers shared by the C-ABI translation units:
// error reporting, driver-entry-point TMA descriptor encoding, launch helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mt_b200.h"

namespace mt {

// thread-local last error message (returned by mt_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MT_CHECK_CUDA(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::mt::cuda_fail(_e, #expr); \
  } while (0)

#define MT_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::mt::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda dependency,
// so the library also loads on a box without a driver — required for the CPU-side symbol tests).
// dims/strides are innermost-first; strides (bytes) are given for dims 1..rank-1.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* gptr, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B);

int sm_count(int* out);

// next {start, end} slot of the device timeline (mt_timeline_begin), or nullptr when no timeline is being recorded
unsigned long long* timeline_next_slot();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch with an optional cluster dimension and, when MT_B200_PDL=1, programmatic dependent launch.  PDL is OFF by
// default: on the C2 encoder step it measured 3% slower than plain stream order (BENCH notes in DESIGN.md), the
// device-side griddepcontrol instructions are no-ops without the launch attribute.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster_x;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mt
