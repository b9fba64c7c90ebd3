// cuemu.h -- a tiny host-side emulator for the CUDA C++ subset our kernels use.
//
// TEST / DEVELOPMENT INFRASTRUCTURE ONLY.  The build container has nvcc but no GPU,
// and GPU minutes are rationed; this header lets the *same kernel sources* that nvcc
// compiles for sm_100a be compiled by g++ (-DB200_EMU) and executed on the CPU so that
// indexing / synchronisation / orchestration logic can be checked against the oracle
// before a GPU run.  It is never built into, loaded by, or reachable from the product
// package: `sac_td3_td7_b200/_lib.py` only ever dlopens the nvcc-built library and
// raises if there is no CUDA device.  Numerics under emulation are NOT parity evidence
// (libm transcendentals, x86 FMA); the `-m gpu` tests are.
//
// Model: every CUDA thread of a block is a fiber (own stack, hand-written x86-64
// context switch); a block runs on one OS thread; blocks are distributed over a small
// pool of OS threads.  __syncthreads / warp shuffles / ballots are cooperative yields.
// Device memory is host memory; cudaMalloc places every allocation so that it ENDS at
// a PROT_NONE guard page, so reading or writing past the end faults immediately.
#pragma once
#ifndef B200_EMU
#error "cuemu.h is only for -DB200_EMU builds"
#endif

#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

// ----------------------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__ static
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) double2 { double x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ----------------------------------------------------------------------------- runtime types
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

extern "C" void cuemu_switch(void** from_sp, void* to_sp);

namespace cuemu {

enum FState { READY = 0, AT_BLOCK_BAR = 1, AT_WARP_BAR = 2, DONE = 3 };

struct Fiber {
  void* sp = nullptr;
  uint3 tid{0, 0, 0};
  int state = READY;
  int lin = 0;
};

struct Warp {
  int live = 0;
  int arrived = 0;
  uint64_t scratch[32];
};

constexpr size_t kStackBytes = 96 * 1024;

struct Block {
  uint3 bidx{0, 0, 0};
  dim3 bdim, gdim;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<char*> stacks;  // pooled
  int cur = 0;
  int live = 0;
  int bar_arrived = 0;
  void* sched_sp = nullptr;
  char* dyn_smem = nullptr;
  size_t dyn_cap = 0;
  const std::function<void()>* body = nullptr;
  ~Block() {
    for (char* s : stacks) munmap(s, kStackBytes);
    free(dyn_smem);
  }
};

inline Block*& tl_block() {
  static thread_local Block* b = nullptr;
  return b;
}

inline void yield_to_scheduler() {
  Block* b = tl_block();
  cuemu_switch(&b->fibers[b->cur].sp, b->sched_sp);
}

inline void release_block_bar(Block* b) {
  for (auto& f : b->fibers)
    if (f.state == AT_BLOCK_BAR) f.state = READY;
  b->bar_arrived = 0;
}
inline void release_warp_bar(Block* b, int w) {
  int lo = w * 32, hi = std::min<int>(lo + 32, (int)b->fibers.size());
  for (int i = lo; i < hi; ++i)
    if (b->fibers[i].state == AT_WARP_BAR) b->fibers[i].state = READY;
  b->warps[w].arrived = 0;
}

inline void block_barrier() {
  Block* b = tl_block();
  Fiber& f = b->fibers[b->cur];
  b->bar_arrived++;
  if (b->bar_arrived == b->live) {
    release_block_bar(b);
    return;
  }
  f.state = AT_BLOCK_BAR;
  yield_to_scheduler();
}

inline void warp_barrier() {
  Block* b = tl_block();
  Fiber& f = b->fibers[b->cur];
  Warp& w = b->warps[f.lin >> 5];
  w.arrived++;
  if (w.arrived == w.live) {
    release_warp_bar(b, f.lin >> 5);
    return;
  }
  f.state = AT_WARP_BAR;
  yield_to_scheduler();
}

extern "C" inline void cuemu_fiber_entry() {
  Block* b = tl_block();
  (*b->body)();
  // finished: retire and re-check barriers that were only waiting for this thread
  b = tl_block();
  Fiber& f = b->fibers[b->cur];
  f.state = DONE;
  b->live--;
  Warp& w = b->warps[f.lin >> 5];
  w.live--;
  if (b->live > 0 && b->bar_arrived == b->live) release_block_bar(b);
  if (w.live > 0 && w.arrived == w.live) release_warp_bar(b, f.lin >> 5);
  yield_to_scheduler();
  fprintf(stderr, "cuemu: resumed a finished fiber\n");
  abort();
}

inline void run_block(Block* b, const std::function<void()>& body, unsigned bx, unsigned by, unsigned bz) {
  const int nthreads = (int)(b->bdim.x * b->bdim.y * b->bdim.z);
  b->bidx = uint3{bx, by, bz};
  b->body = &body;
  b->fibers.resize(nthreads);
  b->warps.assign((nthreads + 31) / 32, Warp());
  while ((int)b->stacks.size() < nthreads) {
    char* s = (char*)mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s == MAP_FAILED) { perror("cuemu stack mmap"); abort(); }
    b->stacks.push_back(s);
  }
  for (int i = 0; i < nthreads; ++i) {
    Fiber& f = b->fibers[i];
    f.lin = i;
    f.state = READY;
    f.tid.x = i % b->bdim.x;
    f.tid.y = (i / b->bdim.x) % b->bdim.y;
    f.tid.z = i / (b->bdim.x * b->bdim.y);
    uintptr_t top = ((uintptr_t)(b->stacks[i] + kStackBytes) & ~(uintptr_t)15) - 64;
    // layout (low->high): r15 r14 r13 r12 rbx rbp | entry | fake return
    void** slot = (void**)top;  // 16-aligned address of the entry pointer
    slot[0] = (void*)&cuemu_fiber_entry;
    slot[1] = nullptr;
    void** regs = slot - 6;
    for (int r = 0; r < 6; ++r) regs[r] = nullptr;
    f.sp = (void*)regs;
    b->warps[i >> 5].live++;
  }
  b->live = nthreads;
  b->bar_arrived = 0;
  tl_block() = b;
  while (b->live > 0) {
    bool progressed = false;
    for (int i = 0; i < nthreads; ++i) {
      if (b->fibers[i].state == READY) {
        b->cur = i;
        cuemu_switch(&b->sched_sp, b->fibers[i].sp);
        progressed = true;
      }
    }
    if (!progressed) {
      fprintf(stderr, "cuemu: deadlock in block (%u,%u,%u): live=%d bar_arrived=%d\n", bx, by, bz, b->live, b->bar_arrived);
      abort();
    }
  }
}

inline int& num_workers() {
  static int n = [] {
    const char* e = getenv("CUEMU_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return v < 1 ? 1 : v;
  }();
  return n;
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const long total = (long)grid.x * grid.y * grid.z;
  if (total == 0) return;
  std::atomic<long> next{0};
  auto worker = [&]() {
    static thread_local Block blk_storage;
    Block* blk = &blk_storage;
    blk->bdim = block;
    blk->gdim = grid;
    if (blk->dyn_cap < smem + 16) {
      free(blk->dyn_smem);
      blk->dyn_smem = (char*)aligned_alloc(128, ((smem + 16 + 127) / 128) * 128);
      blk->dyn_cap = smem + 16;
    }
    for (;;) {
      long i = next.fetch_add(1);
      if (i >= total) break;
      unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y), bz = (unsigned)(i / ((long)grid.x * grid.y));
      run_block(blk, body, bx, by, bz);
    }
  };
  int nw = (int)std::min<long>(num_workers(), total);
  if (nw <= 1) {
    worker();
  } else {
    std::vector<std::thread> ts;
    for (int t = 0; t < nw; ++t) ts.emplace_back(worker);
    for (auto& t : ts) t.join();
  }
}

struct TidProxy { };
inline const uint3& cur_tid() { Block* b = tl_block(); return b->fibers[b->cur].tid; }
inline const uint3& cur_bid() { return tl_block()->bidx; }
inline const dim3& cur_bdim() { return tl_block()->bdim; }
inline const dim3& cur_gdim() { return tl_block()->gdim; }
inline int cur_lane() { Block* b = tl_block(); return b->fibers[b->cur].lin & 31; }
inline Warp& cur_warp() { Block* b = tl_block(); return b->warps[b->fibers[b->cur].lin >> 5]; }
inline char* dyn_smem() { return tl_block()->dyn_smem; }

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Warp& w = cur_warp();
  int lane = cur_lane();
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.scratch[lane] = bits;
  warp_barrier();
  uint64_t got = w.scratch[src_lane & 31];
  warp_barrier();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}

// ---- guarded device allocator ----
struct Alloc { void* base; size_t map_bytes; };
inline std::unordered_map<void*, Alloc>& allocs() { static std::unordered_map<void*, Alloc> m; return m; }
inline std::mutex& alloc_mu() { static std::mutex m; return m; }

}  // namespace cuemu

#define threadIdx (cuemu::cur_tid())
#define blockIdx (cuemu::cur_bid())
#define blockDim (cuemu::cur_bdim())
#define gridDim (cuemu::cur_gdim())
#define warpSize 32

static inline void __syncthreads() { cuemu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::warp_barrier(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// (width = segment size, a power of two: lanes exchange within their own segment, like the hardware)
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int l = cuemu::cur_lane(), seg = l & ~(width - 1);
  return cuemu::shfl_from(v, seg + (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cuemu::shfl_from(v, cuemu::cur_lane() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int l = cuemu::cur_lane(), seg = l & ~(width - 1);
  return cuemu::shfl_from(v, (l + (int)d) > seg + width - 1 ? l : l + (int)d);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int l = cuemu::cur_lane(), seg = l & ~(width - 1);
  return cuemu::shfl_from(v, (l - (int)d) < seg ? l : l - (int)d);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  cuemu::Warp& w = cuemu::cur_warp();
  int lane = cuemu::cur_lane();
  w.scratch[lane] = pred ? 1 : 0;
  cuemu::warp_barrier();
  unsigned r = 0;
  cuemu::Block* b = cuemu::tl_block();
  int wi = b->fibers[b->cur].lin >> 5;
  int lo = wi * 32, hi = std::min<int>(lo + 32, (int)b->fibers.size());
  for (int i = lo; i < hi; ++i)
    if (b->fibers[i].state != cuemu::DONE && w.scratch[i - lo]) r |= (1u << (i - lo));
  cuemu::warp_barrier();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }

// ----------------------------------------------------------------------------- atomics
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* a, float v) {
  uint32_t* p = (uint32_t*)a;
  uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  for (;;) {
    float f; memcpy(&f, &old, 4);
    float nf = f + v; uint32_t nb; memcpy(&nb, &nf, 4);
    if (__atomic_compare_exchange_n(p, &old, nb, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return f;
  }
}
static inline double atomicAdd(double* a, double v) {
  uint64_t* p = (uint64_t*)a;
  uint64_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  for (;;) {
    double f; memcpy(&f, &old, 8);
    double nf = f + v; uint64_t nb; memcpy(&nb, &nf, 8);
    if (__atomic_compare_exchange_n(p, &old, nb, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return f;
  }
}
template <class T> static inline T cuemu_atomic_minmax(T* a, T v, bool is_max) {
  T old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
  for (;;) {
    T nv = is_max ? (old > v ? old : v) : (old < v ? old : v);
    if (__atomic_compare_exchange_n(a, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return old;
  }
}
static inline int atomicMax(int* a, int v) { return cuemu_atomic_minmax(a, v, true); }
static inline unsigned atomicMax(unsigned* a, unsigned v) { return cuemu_atomic_minmax(a, v, true); }
static inline int atomicMin(int* a, int v) { return cuemu_atomic_minmax(a, v, false); }
static inline unsigned atomicMin(unsigned* a, unsigned v) { return cuemu_atomic_minmax(a, v, false); }
static inline unsigned atomicExch(unsigned* a, unsigned v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* a, int v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* a, unsigned cmp, unsigned v) {
  __atomic_compare_exchange_n(a, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long v) {
  __atomic_compare_exchange_n(a, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline unsigned atomicInc(unsigned* a, unsigned lim) {
  unsigned old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
  for (;;) {
    unsigned nv = (old >= lim) ? 0 : old + 1;
    if (__atomic_compare_exchange_n(a, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return old;
  }
}

// ----------------------------------------------------------------------------- intrinsics
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
// rounding-explicit arithmetic: volatile blocks contraction / reassociation
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
// glibc already declares __expf/__logf/__powf/__tanhf/__sincosf: map the CUDA fast intrinsics by macro
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __powf(x, y) powf((x), (y))
#define __tanhf(x) tanhf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }

// ----------------------------------------------------------------------------- runtime API
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "cuemu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t bytes) {
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  size_t b16 = (bytes + 15) & ~(size_t)15;
  if (b16 == 0) b16 = 16;
  size_t data_pages = (b16 + page - 1) / page;
  size_t map_bytes = (data_pages + 1) * page;
  char* base = (char*)mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
  if (mprotect(base + data_pages * page, page, PROT_NONE) != 0) return cudaErrorMemoryAllocation;
  char* user = base + data_pages * page - b16;  // allocation ends at the guard page
  memset(user, 0xCB, b16);                      // poison: uninitialised reads show up as garbage
  std::lock_guard<std::mutex> g(cuemu::alloc_mu());
  cuemu::allocs()[user] = cuemu::Alloc{base, map_bytes};
  *p = user;
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> g(cuemu::alloc_mu());
  auto it = cuemu::allocs().find(p);
  if (it == cuemu::allocs().end()) return cudaErrorInvalidValue;
  munmap(it->second.base, it->second.map_bytes);
  cuemu::allocs().erase(it);
  return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
#define cudaStreamLegacy ((cudaStream_t)0x1)
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }

#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cuemu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cuemu::dyn_smem())

#ifdef CUEMU_IMPLEMENTATION
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cuemu_switch,.-cuemu_switch
)");
#endif
