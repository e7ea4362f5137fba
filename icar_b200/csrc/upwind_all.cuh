// Pass 1 of MPDATA for all scalars in one launch -- k_upwind_all_v4, four cells per thread -- in a header of its own so
// that tests/cpp/upwind_host.cpp can run this very kernel body on the CPU (host_emu.h) against the oracle's pass 1.
// Included by mpdata_staged.cu.
#pragma once
#include "common.cuh"
#include "mpdata.cuh"

namespace icar {

// ---------------------------------------------------------------------------------------------
// Pass 1 for all scalars (same arithmetic, bit for bit, as k_upwind of mpdata_staged.cu, built with -fmad=false).
// ---------------------------------------------------------------------------------------------
// Both divisors of a cell (J and dz*J) are shared by all N scalars, so each quotient is formed from ONE
// correctly rounded reciprocal per divisor, y = RN(1/b), as  q0 = RN(a*y), r = a - b*q0 (exact, FMA),
// q = RN(q0 + r*y)  -- Markstein's correction: q is the IEEE quotient RN(a/b), bit for bit (outside the
// subnormal range), for 3 instructions instead of the ~10 of a full division with its slow-path check.
// (tests/test_fp_identities.py checks the identity in exact rational arithmetic.)
__device__ __forceinline__ float div_shared(float a, float b, float y) {
    const float q0 = a * y;
    const float r = fmaf(-b, q0, a);
    return fmaf(r, y, q0);
}

// ---------------------------------------------------------------------------------------------
// Pass 1 for all scalars, four consecutive cells per thread: every x-row of the 7-point stencil is one
// 16-byte load (+ the i-1 / i+4 neighbours of the centre row), a quarter of the load and address
// instructions of a one-cell-per-thread kernel, which was issue-bound on exactly those.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const float* __restrict__ p, float (&o)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}

template <int N>
__global__ void __launch_bounds__(128)
k_upwind_all_v4(Geom g, ScalarSet s, const float* __restrict__ U, const float* __restrict__ V,
                const float* __restrict__ W, const float* __restrict__ dz, const float* __restrict__ jaco) {
    const int i0 = 4 * (blockIdx.x * 128 + threadIdx.x), k = blockIdx.y, j = blockIdx.z;
    if (i0 >= g.nx) return;
    const int nx = g.nx, ld = g.ld;
    const size_t ps = g.ps();
    const size_t c = g.at(i0, k, j);
    const bool jring = (j == 0 || j == g.ny - 1);
    const bool bottom = (k == 0), top = (k == g.nz - 1), flat = (g.nz == 1);
    float uc[4], vm[4], vp[4], wp[4], wm[4] = {0.f, 0.f, 0.f, 0.f}, jc[4], dzc[4];
    ld4(U + c, uc);
    const float u_hi = (i0 + 4 < nx) ? U[c + 4] : 0.f;
    ld4(V + c, vm); ld4(V + c + ps, vp); ld4(W + c, wp);
    if (!bottom) ld4(W + c - ld, wm);
    ld4(jaco + c, jc); ld4(dz + c, dzc);
    float rj[4], rden[4], den[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        den[e] = dzc[e] * jc[e];
        rj[e] = __frcp_rn(jc[e]);
        rden[e] = __frcp_rn(den[e]);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const float* __restrict__ qin = s.q[n];
        float qc[4], qym[4] = {0.f, 0.f, 0.f, 0.f}, qyp[4] = {0.f, 0.f, 0.f, 0.f}, qzm[4] = {0.f, 0.f, 0.f, 0.f},
              qzp[4] = {0.f, 0.f, 0.f, 0.f};
        ld4(qin + c, qc);
        if (!jring) { ld4(qin + c - ps, qym); ld4(qin + c + ps, qyp); }
        if (!bottom) ld4(qin + c - ld, qzm);
        if (!top) ld4(qin + c + ld, qzp);
        const float q_lo = (i0 > 0) ? qin[c - 1] : 0.f, q_hi = (i0 + 4 < nx) ? qin[c + 4] : 0.f;
        float val[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            if (jring || i == 0 || i >= nx - 1) { val[e] = qc[e]; continue; }      // boundary ring: copied through
            const float qxm = (e == 0) ? q_lo : qc[e > 0 ? e - 1 : 0], qxp = (e == 3) ? q_hi : qc[e < 3 ? e + 1 : 3];
            const float uxp = (e == 3) ? u_hi : uc[e < 3 ? e + 1 : 3];
            const float fxm = donor(uc[e], qxm, qc[e]), fxp = donor(uxp, qc[e], qxp);
            const float fym = donor(vm[e], qym[e], qc[e]), fyp = donor(vp[e], qc[e], qyp[e]);
            float v = qc[e] - div_shared((fxp - fxm) + (fyp - fym), jc[e], rj[e]);
            if (!flat) {
                float vert;
                if (bottom)   vert = donor(wp[e], qc[e], qzp[e]);
                else if (top) vert = qc[e] * wp[e] - donor(wm[e], qzm[e], qc[e]);
                else          vert = donor(wp[e], qc[e], qzp[e]) - donor(wm[e], qzm[e], qc[e]);
                v = v - div_shared(vert, den[e], rden[e]);
            }
            val[e] = v;
        }
        // rows are padded to a multiple of 32 floats: the whole float4 is storable; padding gets q's padding back
        *reinterpret_cast<float4*>(s.q2[n] + c) = make_float4(val[0], val[1], val[2], val[3]);
        if (jring || i0 == 0 || i0 + 4 >= nx - 1 + 1) {
            // the boundary ring also goes to `out` (the corrector only writes interior cells)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + e;
                if (i < nx && (jring || i == 0 || i == nx - 1)) s.out[n][c + e] = qc[e];
            }
        }
    }
}

}  // namespace icar
