// Pass 1 of MPDATA for all scalars on f32x2 pairs (the ICAR_UPWIND_PACKED=1 candidate of mpdata_staged.cu), in a
// header of its own so that tests/cpp/upwind_p2_host.cpp can run this very kernel body on the CPU (host_emu.h:
// one loop iteration per thread, the packed operations one IEEE half at a time) against the oracle's pass 1.
#pragma once
#include "common.cuh"
#include "f32x2.cuh"
#include "mpdata.cuh"

namespace icar {

// ---------------------------------------------------------------------------------------------
// Pass 1 for all scalars on f32x2 pairs (ICAR_UPWIND_PACKED=1; NOT the default: it has not run on hardware
// yet).  The four cells of a thread are two pairs.  The donor flux is taken in the Fortran's own form
// (adv_mpdata.f90:41), f = U+ * l + U- * r with U+ = (U + |U|) / 2 = max(U, 0) and U- = (U - |U|) / 2 =
// min(U, 0): one of the two products is an exact zero, so f = RN(U * l) or RN(U * r) like the select form --
// but it is one FMUL2 + one FFMA2 per PAIR of cells instead of FSETP + FSEL + FMUL per cell, and U+ / U- are
// shared by all N scalars.  Every other operation is a plain RN add / mul / fma whose packed form rounds each
// half identically, so the result equals k_upwind_all_v4's bit for bit (up to the sign of a zero flux).
// The bottom / top / boundary-ring cases are folded into the operands (W- below the ground = 0; above the top
// the "right" cell is the cell itself: outflow qin(kme) * w(kme), :98) instead of branches.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld22(const float* __restrict__ p, f2 (&o)[2]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    o[0] = f2{q.x, q.y}; o[1] = f2{q.z, q.w};
}
__device__ __forceinline__ f2 abs_of(f2 a) { return f2{fabsf(a.x), fabsf(a.y)}; }
// Markstein quotient of pairs: q0 = a*y, r = a - b*q0 (exact), q = q0 + r*y
__device__ __forceinline__ f2 div_shared2(f2 a, f2 b, f2 y) {
    const f2 q0 = a * y;
    const f2 r = fma2(neg2(b), q0, a);
    return fma2(r, y, q0);
}

template <int N>
__global__ void __launch_bounds__(128)
k_upwind_all_p2(Geom g, ScalarSet s, const float* __restrict__ U, const float* __restrict__ V,
                const float* __restrict__ W, const float* __restrict__ dz, const float* __restrict__ jaco) {
    const int i0 = 4 * (blockIdx.x * 128 + threadIdx.x), k = blockIdx.y, j = blockIdx.z;
    if (i0 >= g.nx) return;
    const int nx = g.nx, ld = g.ld;
    const size_t ps = g.ps();
    const size_t c = g.at(i0, k, j);
    const bool jring = (j == 0 || j == g.ny - 1);
    const bool bottom = (k == 0), top = (k == g.nz - 1), flat = (g.nz == 1);
    if (jring) {                                   // the whole row belongs to the boundary ring: copied through
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const float4 q = *reinterpret_cast<const float4*>(s.q[n] + c);
            *reinterpret_cast<float4*>(s.q2[n] + c) = q;
            const float qc[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) if (i0 + e < nx) s.out[n][c + e] = qc[e];
        }
        return;
    }
    const f2 half = splat(0.5f), zero = splat(0.f);
    f2 uc[2], vm[2], vp[2], wp[2], wm[2] = {zero, zero}, jc[2], dzc[2];
    ld22(U + c, uc);
    const float u_hi = (i0 + 4 < nx) ? U[c + 4] : 0.f;
    ld22(V + c, vm); ld22(V + c + ps, vp); ld22(W + c, wp);
    if (!bottom) ld22(W + c - ld, wm);
    ld22(jaco + c, jc); ld22(dz + c, dzc);
    // winds at the i+1 faces of the four cells: (u1, u2), (u3, u_hi)
    const f2 ur[2] = {f2{uc[0].y, uc[1].x}, f2{uc[1].y, u_hi}};
    f2 den[2], rj[2], rden[2];
    f2 uxm_p[2], uxm_m[2], uxp_p[2], uxp_m[2], vym_p[2], vym_m[2], vyp_p[2], vyp_m[2], wzp_p[2], wzp_m[2], wzm_p[2], wzm_m[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        den[h] = dzc[h] * jc[h];
        rj[h] = f2{__frcp_rn(jc[h].x), __frcp_rn(jc[h].y)};
        rden[h] = f2{__frcp_rn(den[h].x), __frcp_rn(den[h].y)};
        uxm_p[h] = (uc[h] + abs_of(uc[h])) * half; uxm_m[h] = (uc[h] - abs_of(uc[h])) * half;
        uxp_p[h] = (ur[h] + abs_of(ur[h])) * half; uxp_m[h] = (ur[h] - abs_of(ur[h])) * half;
        vym_p[h] = (vm[h] + abs_of(vm[h])) * half; vym_m[h] = (vm[h] - abs_of(vm[h])) * half;
        vyp_p[h] = (vp[h] + abs_of(vp[h])) * half; vyp_m[h] = (vp[h] - abs_of(vp[h])) * half;
        wzp_p[h] = (wp[h] + abs_of(wp[h])) * half; wzp_m[h] = (wp[h] - abs_of(wp[h])) * half;
        wzm_p[h] = (wm[h] + abs_of(wm[h])) * half; wzm_m[h] = (wm[h] - abs_of(wm[h])) * half;
    }
    // which of the four cells are boundary-ring cells in x (copied through)
    const bool ring0 = (i0 == 0) || (i0 >= nx - 1), ring1 = (i0 + 1 >= nx - 1), ring2 = (i0 + 2 >= nx - 1), ring3 = (i0 + 3 >= nx - 1);
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const float* __restrict__ qin = s.q[n];
        f2 qc[2], qym[2], qyp[2], qzm[2] = {zero, zero}, qzp[2];
        ld22(qin + c, qc);
        ld22(qin + c - ps, qym); ld22(qin + c + ps, qyp);
        if (!bottom) ld22(qin + c - ld, qzm);
        if (!top) ld22(qin + c + ld, qzp); else { qzp[0] = qc[0]; qzp[1] = qc[1]; }
        const float q_lo = (i0 > 0) ? qin[c - 1] : 0.f, q_hi = (i0 + 4 < nx) ? qin[c + 4] : 0.f;
        const f2 ql[2] = {f2{q_lo, qc[0].x}, f2{qc[0].y, qc[1].x}};          // cells i-1
        const f2 qr[2] = {f2{qc[0].y, qc[1].x}, f2{qc[1].y, q_hi}};          // cells i+1
        f2 val[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const f2 fxm = fma2(uxm_m[h], qc[h], uxm_p[h] * ql[h]);          // donor(u(i), q(i-1), q(i))
            const f2 fxp = fma2(uxp_m[h], qr[h], uxp_p[h] * qc[h]);          // donor(u(i+1), q(i), q(i+1))
            const f2 fym = fma2(vym_m[h], qc[h], vym_p[h] * qym[h]);
            const f2 fyp = fma2(vyp_m[h], qyp[h], vyp_p[h] * qc[h]);
            f2 v = qc[h] - div_shared2((fxp - fxm) + (fyp - fym), jc[h], rj[h]);
            if (!flat) {
                const f2 fzp = fma2(wzp_m[h], qzp[h], wzp_p[h] * qc[h]);     // top: qzp == qc, i.e. qc * w
                const f2 fzm = fma2(wzm_m[h], qc[h], wzm_p[h] * qzm[h]);     // bottom: w- == w+ == 0
                v = v - div_shared2(bottom ? fzp : (fzp - fzm), den[h], rden[h]);
            }
            val[h] = v;
        }
        const float4 o = make_float4(ring0 ? qc[0].x : val[0].x, ring1 ? qc[0].y : val[0].y,
                                     ring2 ? qc[1].x : val[1].x, ring3 ? qc[1].y : val[1].y);
        *reinterpret_cast<float4*>(s.q2[n] + c) = o;
        if (i0 == 0) s.out[n][c] = qc[0].x;                                   // the ring also goes to `out`
        if (i0 + 4 >= nx) {
            const float qq[4] = {qc[0].x, qc[0].y, qc[1].x, qc[1].y};
#pragma unroll
            for (int e = 0; e < 4; ++e) if (i0 + e == nx - 1) s.out[n][c + e] = qq[e];
        }
    }
}

}  // namespace icar
