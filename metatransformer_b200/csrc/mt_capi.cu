// mt_capi.cu — library-level pieces of the C ABI: version, error string, device check, TMA descriptor
// encoding via the driver entry point, dtype casts.
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <string.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include <dlfcn.h>

namespace mt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", int(e), cudaGetErrorString(e), what);
  return MT_ERR_CUDA;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    else (void)cudaGetLastError();
  }
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* gptr, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                CUtensorMapL2promotion promo) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / no GPU?)");
    return MT_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, dtype, cuuint32_t(rank), const_cast<void*>(gptr), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu x %llu, box %u x %u, ptr %p)", int(r),
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0, gptr);
    return MT_ERR_CUDA;
  }
  return MT_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MT_B200_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int sm_count(int* out) {
  static int cached[64] = {0};
  int dev = 0;
  MT_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && cached[dev] > 0) {
    *out = cached[dev];
    return MT_OK;
  }
  int n = 0;
  MT_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev >= 0 && dev < 64) cached[dev] = n;
  *out = n;
  return MT_OK;
}

// ------------------------------------------------------------------------------------------------
// device timeline (mt_timeline_begin / mt_timeline_end): instrumented launches take consecutive {start, end} slots
// ------------------------------------------------------------------------------------------------
static unsigned long long* g_tl_buf = nullptr;
static int g_tl_cap = 0, g_tl_next = 0;

unsigned long long* timeline_next_slot() {
  if (g_tl_buf == nullptr || g_tl_next >= g_tl_cap) return nullptr;
  return g_tl_buf + 2 * size_t(g_tl_next++);
}

// split representation of fp32 data: hi = bf16(x), lo = bf16(x - hi)  (hi + lo carries 16 mantissa bits)
__global__ void split_f32_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(f[j]);
      l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<uint2*>(l);
  }
}

__global__ void merge_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                 float* __restrict__ dst, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += stride) {
    __nv_bfloat16 h[4], l[4];
    *reinterpret_cast<uint2*>(h) = *reinterpret_cast<const uint2*>(hi + i);
    *reinterpret_cast<uint2*>(l) = *reinterpret_cast<const uint2*>(lo + i);
    float4 v;
    v.x = __bfloat162float(h[0]) + __bfloat162float(l[0]);
    v.y = __bfloat162float(h[1]) + __bfloat162float(l[1]);
    v.z = __bfloat162float(h[2]) + __bfloat162float(l[2]);
    v.w = __bfloat162float(h[3]) + __bfloat162float(l[3]);
    *reinterpret_cast<float4*>(dst + i) = v;
  }
}

// in-place scale of an all-reduced gradient bucket (the 1 / world_size of DDP's averaging)
template <typename T>
__global__ void scale_kernel(T* __restrict__ p, size_t n, float s) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    if constexpr (sizeof(T) == 4) p[i] = p[i] * s;
    else p[i] = __float2bfloat16_rn(__bfloat162float(p[i]) * s);
  }
}

// per-tensor fp8 (e4m3) quantisation: q = sat(x * inv_scale), 8 elements per thread-iteration
__global__ void quantize_e4m3_kernel(const __nv_bfloat16* __restrict__ src, uint8_t* __restrict__ dst, size_t n, float inv_scale) {
  const size_t stride = size_t(gridDim.x) * blockDim.x * 8;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i + 7 < n; i += stride) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t out[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint16_t pr;
      // cvt packs {first operand -> upper byte, second -> lower byte}
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(pr) : "f"(bf16_hi(w[j]) * inv_scale), "f"(bf16_lo(w[j]) * inv_scale));
      out[j >> 1] |= uint32_t(pr) << (16 * (j & 1));
    }
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(out[0], out[1]);
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise cast (weight packing fp32 -> bf16, boundary conversions). Memory-bound, 16 B per thread-load.
// ------------------------------------------------------------------------------------------------
template <typename S>
MT_DEVINL float to_f32(S v);
template <> MT_DEVINL float to_f32<float>(float v) { return v; }
template <> MT_DEVINL float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> MT_DEVINL float to_f32<__half>(__half v) { return __half2float(v); }
template <typename D>
MT_DEVINL D from_f32(float v);
template <> MT_DEVINL float from_f32<float>(float v) { return v; }
template <> MT_DEVINL __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> MT_DEVINL __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      S s[4];
      if constexpr (sizeof(S) == 4) *reinterpret_cast<uint4*>(s) = *reinterpret_cast<const uint4*>(src + i);
      else *reinterpret_cast<uint2*>(s) = *reinterpret_cast<const uint2*>(src + i);
      D d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = from_f32<D>(to_f32<S>(s[j]));
      if constexpr (sizeof(D) == 4) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(d);
      else *reinterpret_cast<uint2*>(dst + i) = *reinterpret_cast<uint2*>(d);
    } else {
      for (size_t j = i; j < n; ++j) dst[j] = from_f32<D>(to_f32<S>(src[j]));
    }
  }
}

template <typename S, typename D>
static int launch_cast(const void* src, void* dst, size_t n, cudaStream_t s) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > size_t(sms) * 16) blocks = size_t(sms) * 16;
  if (blocks < 1) blocks = 1;
  cast_kernel<S, D><<<unsigned(blocks), 256, 0, s>>>(reinterpret_cast<const S*>(src), reinterpret_cast<D*>(dst), n);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

}  // namespace mt

extern "C" int mt_version(void) { return 200; }

extern "C" const char* mt_last_error(void) { return mt::g_err; }

extern "C" int mt_device_ok(int device) {
  using namespace mt;
  cudaDeviceProp prop;
  MT_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  MT_REQUIRE(prop.major == 10, MT_ERR_UNSUPPORTED,
             "device %d is compute capability %d.%d; libmt_b200 contains sm_100a code only", device, prop.major,
             prop.minor);
  return MT_OK;
}

extern "C" int mt_timeline_begin(void* slots, int capacity) {
  using namespace mt;
  MT_REQUIRE(slots != nullptr && capacity > 0, MT_ERR_BAD_SHAPE, "mt_timeline_begin: need a device buffer and a capacity");
  MT_REQUIRE((reinterpret_cast<uintptr_t>(slots) & 7) == 0, MT_ERR_BAD_ALIGN, "mt_timeline_begin: buffer must be 8-byte aligned");
  g_tl_buf = reinterpret_cast<unsigned long long*>(slots);
  g_tl_cap = capacity;
  g_tl_next = 0;
  return MT_OK;
}

extern "C" int mt_timeline_end(void) {
  const int used = mt::g_tl_next;
  mt::g_tl_buf = nullptr;
  mt::g_tl_cap = 0;
  mt::g_tl_next = 0;
  return used;
}

static int split_merge_grid(size_t n, unsigned* blocks) {
  int sms = 0;
  int rc = mt::sm_count(&sms);
  if (rc) return rc;
  size_t b = (n / 4 + 255) / 256;
  if (b > size_t(sms) * 16) b = size_t(sms) * 16;
  if (b < 1) b = 1;
  *blocks = unsigned(b);
  return MT_OK;
}

extern "C" int mt_split_f32(const float* src, void* hi, void* lo, size_t n, void* stream) {
  using namespace mt;
  MT_REQUIRE(src && hi && lo, MT_ERR_BAD_SHAPE, "mt_split_f32: null pointer");
  MT_REQUIRE(n % 4 == 0, MT_ERR_BAD_SHAPE, "mt_split_f32: n must be a multiple of 4");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_split_f32: pointers must be 16-byte aligned");
  if (n == 0) return MT_OK;
  unsigned blocks = 1;
  int rc = split_merge_grid(n, &blocks);
  if (rc) return rc;
  split_f32_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_merge_f32(const void* hi, const void* lo, float* dst, size_t n, void* stream) {
  using namespace mt;
  MT_REQUIRE(dst && hi && lo, MT_ERR_BAD_SHAPE, "mt_merge_f32: null pointer");
  MT_REQUIRE(n % 4 == 0, MT_ERR_BAD_SHAPE, "mt_merge_f32: n must be a multiple of 4");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_merge_f32: pointers must be 16-byte aligned");
  if (n == 0) return MT_OK;
  unsigned blocks = 1;
  int rc = split_merge_grid(n, &blocks);
  if (rc) return rc;
  merge_f32_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo), dst, n);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

// ncclAllReduce through dlopen: libmt_b200.so has no link-time NCCL dependency (it must load on a box without NCCL for the
// CPU-side symbol tests); the communicator belongs to the caller.
extern "C" int mt_allreduce(void* nccl_comm, void* buf, size_t count, int dtype, float scale, void* stream) {
  using namespace mt;
  typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static PFN_ncclAllReduce fn = nullptr;
  static bool tried = false;
  MT_REQUIRE(nccl_comm && buf, MT_ERR_BAD_SHAPE, "mt_allreduce: null pointer");
  MT_REQUIRE(dtype == MT_DT_F32 || dtype == MT_DT_BF16, MT_ERR_BAD_DTYPE, "mt_allreduce: dtype must be fp32 or bf16");
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = reinterpret_cast<PFN_ncclAllReduce>(dlsym(h, "ncclAllReduce"));
  }
  MT_REQUIRE(fn != nullptr, MT_ERR_UNSUPPORTED, "mt_allreduce: libnccl.so.2 / ncclAllReduce not found");
  if (count == 0) return MT_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int nccl_dt = (dtype == MT_DT_F32) ? 7 : 9;     // ncclFloat32, ncclBfloat16
  const int rc = fn(buf, buf, count, nccl_dt, 0 /* ncclSum */, nccl_comm, s);
  MT_REQUIRE(rc == 0, MT_ERR_CUDA, "mt_allreduce: ncclAllReduce failed with ncclResult %d", rc);
  if (scale != 1.0f) {
    int sms = 0;
    int rc2 = sm_count(&sms);
    if (rc2) return rc2;
    size_t blocks = (count + 255) / 256;
    if (blocks > size_t(sms) * 16) blocks = size_t(sms) * 16;
    if (dtype == MT_DT_F32) scale_kernel<float><<<unsigned(blocks), 256, 0, s>>>(reinterpret_cast<float*>(buf), count, scale);
    else scale_kernel<__nv_bfloat16><<<unsigned(blocks), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(buf), count, scale);
    MT_CHECK_CUDA(cudaGetLastError());
  }
  return MT_OK;
}

extern "C" int mt_quantize_fp8(const void* src_bf16, void* dst_e4m3, size_t n, float scale, void* stream) {
  using namespace mt;
  MT_REQUIRE(src_bf16 && dst_e4m3, MT_ERR_BAD_SHAPE, "mt_quantize_fp8: null pointer");
  MT_REQUIRE(n % 8 == 0 && scale > 0.f, MT_ERR_BAD_SHAPE, "mt_quantize_fp8: n %% 8 == 0 and scale > 0 required");
  MT_REQUIRE((reinterpret_cast<uintptr_t>(src_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_e4m3) & 7) == 0, MT_ERR_BAD_ALIGN,
             "mt_quantize_fp8: src 16-byte, dst 8-byte aligned");
  if (n == 0) return MT_OK;
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  quantize_e4m3_kernel<<<unsigned(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), reinterpret_cast<uint8_t*>(dst_e4m3), n, 1.0f / scale);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream) {
  using namespace mt;
  MT_REQUIRE(src && dst, MT_ERR_BAD_SHAPE, "mt_cast: null pointer");
  MT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_cast: pointers must be 16-byte aligned");
  if (n == 0) return MT_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (src_dtype == MT_DT_F32 && dst_dtype == MT_DT_BF16) return launch_cast<float, __nv_bfloat16>(src, dst, n, s);
  if (src_dtype == MT_DT_BF16 && dst_dtype == MT_DT_F32) return launch_cast<__nv_bfloat16, float>(src, dst, n, s);
  if (src_dtype == MT_DT_F16 && dst_dtype == MT_DT_BF16) return launch_cast<__half, __nv_bfloat16>(src, dst, n, s);
  if (src_dtype == MT_DT_BF16 && dst_dtype == MT_DT_F16) return launch_cast<__nv_bfloat16, __half>(src, dst, n, s);
  if (src_dtype == MT_DT_F16 && dst_dtype == MT_DT_F32) return launch_cast<__half, float>(src, dst, n, s);
  if (src_dtype == MT_DT_F32 && dst_dtype == MT_DT_F16) return launch_cast<float, __half>(src, dst, n, s);
  set_error("mt_cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  return MT_ERR_BAD_DTYPE;
}
