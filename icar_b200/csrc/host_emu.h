// TEST SUPPORT, never compiled into libicar_b200.so: what a plain C++ compiler needs in order to compile the
// kernel headers that are written to be emulated on the CPU (lw_kernels.cuh, upwind_p2.cuh) -- the CUDA built-ins
// those kernels use, as ordinary C++.  The harnesses under tests/cpp/ define blockIdx / threadIdx and call a kernel
// body once per thread of the grid the product launches.  Included only when __CUDACC__ is not defined.
#pragma once
#if defined(__CUDACC__)
#error "host_emu.h is for the host-only test harnesses"
#endif
#include <cmath>
#include <cstddef>
#include <cstdint>

struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return {x, y}; }
struct icar_emu_idx { unsigned x, y, z; };
extern thread_local icar_emu_idx blockIdx, threadIdx, blockDim;
inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned old = *p; *p = old | v; return old; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned __float_as_uint(float x) { unsigned u; __builtin_memcpy(&u, &x, 4); return u; }
inline float __uint_as_float(unsigned u) { float x; __builtin_memcpy(&x, &u, 4); return x; }
inline int __float_as_int(float x) { int u; __builtin_memcpy(&u, &x, 4); return u; }
inline float __int_as_float(int u) { float x; __builtin_memcpy(&x, &u, 4); return x; }
inline long long __double_as_longlong(double x) { long long u; __builtin_memcpy(&u, &x, 8); return u; }
inline double __longlong_as_double(long long u) { double x; __builtin_memcpy(&x, &u, 8); return x; }
inline float __frcp_rn(float x) { return 1.0f / x; }           // correctly rounded reciprocal == IEEE 1/x
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
