// Shared host/device helpers for the satb200 library.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace satb {

// ---- error plumbing (C ABI returns int codes; message via satb_last_error) ----
void set_last_error(const std::string& msg);
const char* get_last_error();

#define SATB_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::satb::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                             __FILE__ + ":" + std::to_string(__LINE__));                   \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

#define SATB_REQUIRE(cond, msg)                                                        \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::satb::set_last_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + \
                             std::to_string(__LINE__));                                \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define SATB_PROPAGATE(expr) \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 16-bit operand type: fp16 (default; what the reference's autocast path uses,
// models/transformer.py:498-504, inference/sampling.py:210) or bf16.
enum OperandType : int { OP_F16 = 0, OP_BF16 = 1 };

template <bool BF16>
struct Op16;
template <>
struct Op16<false> {
  using T = __half;
  using T2 = __half2;
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
  __device__ static __forceinline__ T from_float(float a) { return __float2half_rn(a); }
  __device__ static __forceinline__ float to_float(T a) { return __half2float(a); }
};
template <>
struct Op16<true> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  __device__ static __forceinline__ T from_float(float a) { return __float2bfloat16_rn(a); }
  __device__ static __forceinline__ float to_float(T a) { return __bfloat162float(a); }
};

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_launch_count;
static inline void count_launch(int n = 1) { g_launch_count += n; }

int device_sm_count();   // of the CURRENT device (cached per device)

// "done once per device" flag for cudaFuncSetAttribute: function attributes are per device, and a process may
// drive several devices (one handle each), so a process-wide static bool would skip the second device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {   // true exactly once per current device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Launch with programmatic stream serialization (PDL): the kernel must call pdl_wait() before its
// first global-memory access.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace satb
