// TEST SUPPORT, never compiled into libicar_b200.so: what a plain C++ compiler needs in order to compile the kernel
// sources that the tests run on the CPU (lw_kernels.cuh, upwind_all.cuh, corrector_fused.inc, mp_simple_kernel.inc,
// mp_thompson_kernel.inc) -- the CUDA built-ins those kernels use, as ordinary C++, and host versions of the few
// device-only primitives the product defines under `#ifndef ICAR_HOST_EMULATION` (the f32x2 operations one IEEE half
// at a time, the unguarded division sequence as the IEEE quotient it equals).  Every harness under tests/cpp/ includes
// this file FIRST, defines blockIdx / threadIdx and calls a kernel body once per thread of the grid the product
// launches.
#pragma once
#if defined(__CUDACC__)
#error "host_emu.h is for the host-only test harnesses"
#endif
#define ICAR_HOST_EMULATION 1
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

#define __grid_constant__
#define __maxnreg__(n)

// ---- host versions of the device-only primitives of the product sources (each is defined there under
// `#ifndef ICAR_HOST_EMULATION`)
namespace icar {
// f32x2.cuh: the same IEEE roundings, one half at a time (the harnesses are built with -ffp-contract=off)
struct f2 { float x, y; };
inline f2 operator+(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
inline f2 operator-(f2 a, f2 b) { return f2{a.x - b.x, a.y - b.y}; }
inline f2 operator*(f2 a, f2 b) { return f2{a.x * b.x, a.y * b.y}; }
inline f2 fma2(f2 a, f2 b, f2 c) { return f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
// corrector_prims.inc: the unguarded div.rn sequence IS the IEEE quotient for the operands it is used with (checked
// against `/` on the device, tests/test_gpu_parity.py); the reciprocal it shares between quotients is not needed here
inline f2 rcp2n(f2 b) { return b; }
inline f2 quo2(f2 a, f2 b, f2) { return f2{a.x / b.x, a.y / b.y}; }
// mp_simple_kernel.inc
inline float div_rn(float a, float b) { return a / b; }
// mp_thompson_kernel.inc: the two MUFU-based helpers of fpow_third and the SFU start of the decade search
inline float th_fast_div(float a, float b) { return a / b; }
inline float th_fast_log(float x) { return logf(x); }
inline float th_fast_log10(float x) { return log10f(x); }
}  // namespace icar
