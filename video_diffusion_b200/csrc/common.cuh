// Shared helpers for libvdm.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <utility>

#include "../../include/vdm.h"

namespace vdm {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define VDM_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::vdm::set_error(__VA_ARGS__);  \
      return -1;                      \
    }                                 \
  } while (0)

// Call right after a kernel launch: counts it and converts launch errors.
#define VDM_AFTER_LAUNCH(name)                                                        \
  do {                                                                                \
    ::vdm::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::vdm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return (int)e__;                                                                \
    }                                                                                 \
  } while (0)

// Host-side caches (SM count, "function attributes already set") are kept per CUDA device: a process may hold
// models on several GPUs, and cudaFuncSetAttribute applies to the current device's context only.
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
template <typename T>
struct PerDevice {
  T v[64] = {};
  T& get() { return v[current_device() & 63]; }
};

inline int num_sms() {
  static PerDevice<int> cache;
  int& n = cache.get();
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute (launch_kernel below;
// VDM_PDL=1 switches it on) and follows one protocol: `pdl_launch_dependents()` first -- the NEXT kernel in the
// stream may then be scheduled as soon as this grid's CTAs are all resident and SM resources free up, so its launch
// latency and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap this kernel's tail -- and
// `pdl_wait()` before the first access to global memory: it returns once every prerequisite grid has completed and
// its writes are visible.  Nothing global is written before the wait, so there is no write-after-read hazard either.
// Captured into a CUDA graph the attribute becomes a programmatic edge between the two kernel nodes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef VDM_PDL_LATE
// Mixed protocol (build with -DVDM_PDL_LATE): the short-block kernels trigger at their top as before -- their
// dependent, usually a persistent GEMM, then brings up its CTAs (barriers, TMEM, descriptor fetch) as the last wave of
// blocks drains -- but a persistent GEMM, whose CTAs are all resident from the first cycle, triggers only when a CTA's
// TMA producer has issued its last load: triggered at the top, its dependents' blocks would sit on every SM for the
// whole kernel and keep the other micro-batch's kernels from running beside it.
__device__ __forceinline__ void pdl_launch_dependents_persistent() {}
__device__ __forceinline__ void pdl_trigger_late() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_launch_dependents_persistent() { pdl_launch_dependents(); }
__device__ __forceinline__ void pdl_trigger_late() {}
#endif

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VDM_PDL");
    on = (e != nullptr && atoi(e) != 0) ? 1 : 0;   // opt-in: measured slower than plain serialisation so far
  }
  return on != 0;
}

// cluster_x > 1: thread-block cluster of that many CTAs along x.  Launch errors are picked up by VDM_AFTER_LAUNCH.
template <typename... KArgs, typename... Args>
inline void launch_kernel_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             int cluster_x, bool programmatic, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (programmatic) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          int cluster_x, Args&&... args) {
  launch_kernel_ex(kernel, grid, block, smem, stream, cluster_x, pdl_enabled(), std::forward<Args>(args)...);
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// exact-ish SiLU used where the reference computes x*sigmoid(x) in fp32
__device__ __forceinline__ float silu_precise(float x) { return x * (1.0f / (1.0f + expf(-x))); }

// x*sigmoid(x) with sigmoid(x) = 0.5 + 0.5*tanh(x/2): ONE MUFU op per element (tanh.approx, rel. error ~2^-11, well
// below the bf16 rounding of the result) instead of ex2 + rcp.  Shared by gn_apply and the GEMM's transform warps, so
// the fused and the standalone GroupNorm-apply produce bit-identical bf16 operands.
__device__ __forceinline__ float silu_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

template <typename T>
__device__ __forceinline__ void store_elem(T* p, float v);
template <>
__device__ __forceinline__ void store_elem<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// fp32 pair -> packed IEEE half pair, round to nearest, SATURATING (a value beyond +-65504 clamps instead of
// becoming inf): the residual stream of the bf16 model is stored in fp16
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ uint16_t f32_to_f16_bits(float a) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float f16_bits_to_f32(uint16_t v) { return __half2float(__ushort_as_half(v)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace vdm
