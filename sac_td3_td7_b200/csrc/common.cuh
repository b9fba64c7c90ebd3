// common.cuh -- shared device/host helpers for the b200rl kernels (sm_100a).
//
// The same sources are compiled two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a      -> libb200rl.so   (the product)
//   * g++  -DB200_EMU -include tests/cuemu/cuemu.h      -> test-only CPU emulation of
//     the kernels' control flow (see tests/cuemu/cuemu.h; never loaded by the package)
#pragma once

#ifndef B200_EMU
#include <cuda_runtime.h>
#define B200_DYN_SMEM(type, name)                            \
  extern __shared__ __align__(16) unsigned char _dyn_smem[]; \
  type* name = reinterpret_cast<type*>(_dyn_smem)
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace b200rl {
// kernels launched by this thread since the counter was last reset (bench.py: gpu_launches)
inline int& launch_counter() {
  static thread_local int c = 0;
  return c;
}
}  // namespace b200rl

#ifdef B200_EMU
#undef B200_LAUNCH
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                        \
  do {                                                                             \
    ++b200rl::launch_counter();                                                    \
    cuemu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); });        \
  } while (0)
#else
#include <cstdlib>
#include <utility>
namespace b200rl {
// Programmatic dependent launch: every kernel begins with griddepcontrol.wait (block until the
// stream-order predecessor has completed and flushed) followed by griddepcontrol.launch_dependents, so
// the NEXT kernel's launch overlaps this kernel's execution instead of starting after it drains.
// Inside captured graphs the attribute becomes a programmatic edge.  Measured on B200 (bench.py, TD7
// HalfCheetah): 273 us/update with it, 267 us without -- the per-launch cost sits inside the kernels
// (pipeline fill, reduction, epilogue), not in the launch gap -- so it is OFF by default; B200RL_PDL=1 enables it.
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("B200RL_PDL"); return e ? atoi(e) != 0 : false; }();
  return on;
}
template <class... KArgs, class... Args>
inline void launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// launch with thread-block clusters of `cluster_x` consecutive CTAs along x (distributed shared memory)
template <class... KArgs, class... Args>
inline void launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace b200rl
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                              \
  do {                                                                                   \
    ++b200rl::launch_counter();                                                          \
    b200rl::launch_ex(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__);   \
  } while (0)
// Programmatic dependent launch: TRIGGER lets the next kernel of the stream be scheduled (its CTAs run their
// parameter-only setup and then block in WAIT); WAIT returns once every prerequisite grid has completed and its
// memory is visible.  No-ops unless the launch carried the attribute (B200RL_PDL=1).
#define B200_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define B200_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::)
#define B200_PDL_PROLOGUE() \
  do {                      \
    B200_PDL_WAIT();        \
    B200_PDL_TRIGGER();     \
  } while (0)
#endif
#ifdef B200_EMU
#define B200_PDL_PROLOGUE() do { } while (0)
#define B200_PDL_WAIT() do { } while (0)
#define B200_PDL_TRIGGER() do { } while (0)
#define B200_NOINLINE __attribute__((noinline))
#else
#define B200_NOINLINE __noinline__
#endif

namespace b200rl {

// ----------------------------------------------------------------------------- errors
enum Status : int {
  kOk = 0,
  kErrInvalidArg = -1,
  kErrCuda = -2,
  kErrNoDevice = -3,
  kErrState = -4,
  kErrUnsupported = -5,
};

void set_last_error(const std::string& msg);

#define B200_CHECK_CUDA(expr)                                                                       \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      b200rl::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + \
                             ":" + std::to_string(__LINE__));                                       \
      return b200rl::kErrCuda;                                                                      \
    }                                                                                               \
  } while (0)

#define B200_CHECK_LAUNCH() B200_CHECK_CUDA(cudaGetLastError())

#define B200_REQUIRE(cond, msg)                                        \
  do {                                                                 \
    if (!(cond)) {                                                     \
      b200rl::set_last_error(std::string("invalid argument: ") + msg); \
      return b200rl::kErrInvalidArg;                                   \
    }                                                                  \
  } while (0)

#define B200_TRY(expr)           \
  do {                           \
    int _s = (expr);             \
    if (_s != b200rl::kOk) return _s; \
  } while (0)

// ----------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive scan across the 32 lanes of a warp (double)
__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// ----------------------------------------------------------------------------- Philox4x32-10
// Counter-based generator (Salmon et al. 2011).  Used for on-device uniforms (replay
// sampling) and Gaussians (target smoothing / SAC reparameterisation) in the burst path;
// parity tests inject the draws instead.
struct Philox {
  uint32_t key0, key1;
  __host__ __device__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
  __host__ __device__ inline void operator()(uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
      uint32_t hi0 = mulhi(M0, c0), lo0 = M0 * c0;
      uint32_t hi1 = mulhi(M1, c2), lo1 = M1 * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
// 24-bit uniform in [0, 1 - 2^-24] (never 1.0, same grid as torch.rand float32)
__host__ __device__ inline float u01_from_bits(uint32_t b) { return (float)(b >> 8) * (1.0f / 16777216.0f); }
// Box-Muller from two 32-bit words; u1 in (0,1]
__device__ inline void gauss2_from_bits(uint32_t b0, uint32_t b1, float* z0, float* z1) {
  float u1 = ((float)(b0 >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u2 = (float)(b1 >> 8) * (1.0f / 16777216.0f);
  float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

// ----------------------------------------------------------------------------- small utils
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

}  // namespace b200rl
