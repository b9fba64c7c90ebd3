// b200rl.cu -- single translation unit of libb200rl.so (all kernels + the C ABI).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC b200rl.cu -o libb200rl.so
//
// See include/b200rl.h for the exported surface and DESIGN.md for the kernel inventory.
#ifdef B200_EMU
#define CUEMU_IMPLEMENTATION
#include "cuemu.h"
#endif
#include "api.cuh"
