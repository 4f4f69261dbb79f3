// Shared helpers for the snvc_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/snvc_b200.h"

namespace snvc {

// thread-local error text behind snvc_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);

#define SNVC_CHECK_ARG(cond, ...)                           \
  do {                                                      \
    if (!(cond)) return ::snvc::fail(SNVC_E_BADARG, __VA_ARGS__); \
  } while (0)

#define SNVC_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ::snvc::fail((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__));        \
  } while (0)

// kernels launched by this library in this process (monotonic; read by snvc_launch_count())
void count_launch();

inline int launch_status(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

int sm_count();  // cached per device

// Debug / A-B switches of the library (kernel-generation selection, grid clamps used by the ring wrap-around tests).
// Read from the environment ONCE, when the library is loaded; changed afterwards only through snvc_set_option().
// A launch never calls getenv.
enum OptId { OPT_CONV_MODE, OPT_CONV_STORE, OPT_CONV_OCC, OPT_CONV_MAXGRID, OPT_CV_SPLIT_OLD, OPT_CV_THREADS, OPT_ROI_MODE, OPT_LIFT_MODE, OPT_CV_WALK, OPT_COUNT };
const char* opt(OptId id);   // current value, or nullptr when unset

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (evict-first) 128-bit global store: outputs are written once and not re-read here
__device__ __forceinline__ void st_cs_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace snvc
