// Two fp32 values as one f32x2 register pair (sm_100: add / sub / mul / fma .rn.f32x2 -> SASS FADD2, FMUL2,
// FFMA2: one issue slot, each half rounded exactly like the scalar instruction).  ptxas keeps a pair in an
// aligned register pair; mov.b64 packs / unpacks cost nothing.  Everything else (compare, select, min / max,
// MUFU, shuffles) stays per component.
#pragma once
#include "common.cuh"

namespace icar {

#ifndef ICAR_HOST_EMULATION      // (the CPU harnesses of the tests bring the same operations one IEEE half at a time)
struct f2 { float x, y; };
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(f2 a) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ f2 up(u64 v) { f2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(r); }
// MUFU.RCP of both halves
__device__ __forceinline__ f2 rcp2(f2 a) {
    f2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(a.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(a.y));
    return r;
}
#endif
__device__ __forceinline__ f2 neg2(f2 a) { return f2{-a.x, -a.y}; }
__device__ __forceinline__ f2 splat(float v) { return f2{v, v}; }

}  // namespace icar
