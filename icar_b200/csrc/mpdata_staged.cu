// MPDATA advection, "staged" kernels: one kernel per reference sweep
// (wind normalisation, donor-cell pass, pseudo-velocities, limiter + corrector pass).
// General in mpdata_order / FCT / advect_density and templated on float|double;
// the fused fp32 order-2+FCT path lives in mpdata_fused.cu and is checked against this.
//
// Reference: src/physics/adv_mpdata.f90 (+ adv_mpdata_FCT_core.f90).  Statement order of
// every floating-point expression follows the Fortran so that, built with -fmad=false,
// results equal the CPU oracle bit for bit.
//
// Device layout of the staggered winds (differs from the Fortran module arrays, which are
// compact): U[i] is the x-face between cells i-1 and i (i>=1), V[j] the y-face between rows
// j-1 and j (j>=1), W[k] the z-face above cell k; all share the cell pitch `ld`.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "f32x2.cuh"
#include "mpdata.cuh"

namespace icar {

// ---------------------------------------------------------------------------
// adv_mpdata.f90:500-506 (scheme 2) / advect.f90:344-347 (scheme 1)
// ---------------------------------------------------------------------------
template <class R>
__global__ void k_wind_normalise(Geom g, const R* __restrict__ u, const R* __restrict__ v,
                                 const R* __restrict__ w, const R* __restrict__ ju,
                                 const R* __restrict__ jv, const R* __restrict__ jw,
                                 const R* __restrict__ rho, R dt, R dx, int scheme,
                                 R* __restrict__ U, R* __restrict__ V, R* __restrict__ W,
                                 const R* __restrict__ dz, R* __restrict__ Wn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    size_t c = g.at(i, k, j);
    R r0 = rho ? rho[c] : R(1);
    R Uv = R(0), Vv = R(0);
    if (i >= 1) {
        R rs = r0 + (rho ? rho[c - 1] : R(1));
        R uu = u[g.atu(i, k, j)], jj = ju[g.atu(i, k, j)];
        Uv = (scheme == 2) ? (uu * dt * rs * R(0.5) * jj / dx) : (uu * dt * jj * rs * R(0.5) / dx);
    }
    if (j >= 1) {
        R rs = r0 + (rho ? rho[c - g.ps()] : R(1));
        R vv = v[c], jj = jv[c];
        Vv = (scheme == 2) ? (vv * dt * rs * R(0.5) * jj / dx) : (vv * dt * jj * rs * R(0.5) / dx);
    }
    R ww = w[c], jj = jw[c];
    R Wv;
    if (k < g.nz - 1) Wv = ww * dt * jj * ((rho ? rho[c + g.ld] : R(1)) + r0) * R(0.5);
    else              Wv = ww * dt * jj * r0;
    U[c] = Uv;
    V[c] = Vv;
    W[c] = Wv;
    if (Wn) Wn[c] = Wv / dz[c];          // the (W_m/dz) temporary of adv_mpdata.f90:379
}

// The resident fp32 path (rho == 1, MPDATA operand order, W_m/dz wanted): four consecutive cells per
// thread with 16-byte loads and stores of the cell-pitched fields (rows start on 128-byte lines; u and
// jacobian_u rows are nx+1 long and stay scalar).  Same expressions as k_wind_normalise, bit for bit;
// a pure HBM stream (7 fields in, 4 out).
__global__ void __launch_bounds__(128)
k_wind_normalise_v4(Geom g, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
                    const float* __restrict__ ju, const float* __restrict__ jv, const float* __restrict__ jw,
                    float dt, float dx, float* __restrict__ U, float* __restrict__ V, float* __restrict__ W,
                    const float* __restrict__ dz, float* __restrict__ Wn) {
    const int i0 = 4 * (blockIdx.x * 128 + threadIdx.x), k = blockIdx.y, j = blockIdx.z;
    if (i0 >= g.ld) return;                                   // ld is a multiple of 32: whole float4s
    const size_t c = g.at(i0, k, j), cu = g.atu(i0, k, j);
    const float4 v4 = *reinterpret_cast<const float4*>(v + c), jv4 = *reinterpret_cast<const float4*>(jv + c);
    const float4 w4 = *reinterpret_cast<const float4*>(w + c), jw4 = *reinterpret_cast<const float4*>(jw + c);
    const float4 dz4 = *reinterpret_cast<const float4*>(dz + c);
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w}, jvv[4] = {jv4.x, jv4.y, jv4.z, jv4.w};
    const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, jww[4] = {jw4.x, jw4.y, jw4.z, jw4.w};
    const float dzz[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    float Uo[4], Vo[4], Wo[4], Wno[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + e;
        const bool in = i < g.nx;
        const float rs = 1.f + 1.f;
        Uo[e] = (in && i >= 1) ? (u[cu + e] * dt * rs * 0.5f * ju[cu + e] / dx) : 0.f;
        Vo[e] = (in && j >= 1) ? (vv[e] * dt * rs * 0.5f * jvv[e] / dx) : 0.f;
        Wo[e] = (k < g.nz - 1) ? (ww[e] * dt * jww[e] * (1.f + 1.f) * 0.5f) : (ww[e] * dt * jww[e] * 1.f);
        Wno[e] = Wo[e] / dzz[e];
        if (!in) { Wo[e] = 0.f; Wno[e] = 0.f; }               // row padding: keep it finite
    }
    *reinterpret_cast<float4*>(U + c) = make_float4(Uo[0], Uo[1], Uo[2], Uo[3]);
    *reinterpret_cast<float4*>(V + c) = make_float4(Vo[0], Vo[1], Vo[2], Vo[3]);
    *reinterpret_cast<float4*>(W + c) = make_float4(Wo[0], Wo[1], Wo[2], Wo[3]);
    *reinterpret_cast<float4*>(Wn + c) = make_float4(Wno[0], Wno[1], Wno[2], Wno[3]);
}

// (W_m/dz) and (jaco*rho), the array temporaries of adv_mpdata.f90:379
template <class R>
__global__ void k_wn_G(Geom g, const R* __restrict__ W, const R* __restrict__ dz,
                       const R* __restrict__ jaco, const R* __restrict__ rho,
                       R* __restrict__ Wn, R* __restrict__ G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    size_t c = g.at(i, k, j);
    Wn[c] = W[c] / dz[c];
    if (G) G[c] = jaco[c] * rho[c];     // rho == 1: jaco*rho is jaco itself, no copy needed
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:44-105: qout = qin - div(donor fluxes); boundary cells copied.
// ---------------------------------------------------------------------------
template <class R>
__global__ void k_upwind(Geom g, const R* __restrict__ qin, const R* __restrict__ U,
                         const R* __restrict__ V, const R* __restrict__ W,
                         const R* __restrict__ rho, const R* __restrict__ dz,
                         const R* __restrict__ jaco, R* __restrict__ qout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    size_t c = g.at(i, k, j);
    R qc = qin[c];
    if (i == 0 || i == g.nx - 1 || j == 0 || j == g.ny - 1) { qout[c] = qc; return; }
    const size_t ps = g.ps();
    const int ld = g.ld;
    R fxm = donor(U[c], qin[c - 1], qc);
    R fxp = donor(U[c + 1], qc, qin[c + 1]);
    R fym = donor(V[c], qin[c - ps], qc);
    R fyp = donor(V[c + ps], qc, qin[c + ps]);
    R rh = rho ? rho[c] : R(1);
    R jc = jaco[c];
    R val = qc - ((fxp - fxm) + (fyp - fym)) / (jc * rh);
    R den = dz[c] * jc * rh;
    if (g.nz > 1) {
        if (k == 0) {
            val = val - donor(W[c], qc, qin[c + ld]) / den;
        } else if (k == g.nz - 1) {
            val = val - (qc * W[c] - donor(W[c - ld], qin[c - ld], qc)) / den;
        } else {
            val = val - (donor(W[c], qc, qin[c + ld]) - donor(W[c - ld], qin[c - ld], qc)) / den;
        }
    }
    qout[c] = val;
}

// ---------------------------------------------------------------------------
// adv_mpdata.f90:107-255 + the scaling of :383-385.  q is the field after pass 1.
// u2[i], v2[j], w2[k] use the same face convention as U, V, W.
// ---------------------------------------------------------------------------
template <class R>
__global__ void k_pseudo_velocities(Geom g, const R* __restrict__ q, const R* __restrict__ U,
                                    const R* __restrict__ V, const R* __restrict__ Wn,
                                    const R* __restrict__ G, const R* __restrict__ dz,
                                    R* __restrict__ u2, R* __restrict__ v2, R* __restrict__ w2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    const size_t ps = g.ps();
    const int ld = g.ld;
    const R eps = R(1e-10f);
    size_t c = g.at(i, k, j);
    const bool jin = (j > 0 && j < g.ny - 1), kin = (k > 0 && k < g.nz - 1),
               iin = (i > 0 && i < g.nx - 1);
    R qc = q[c];
    // ---- U face (i-1,i)
    R tu = R(0);
    if (i >= 1) {
        R rx = qc, lx = q[c - 1];
        R uu = U[c];
        R GG = G[c] + G[c - 1];
        R t = ab(uu) * (R(1) - ab(uu) / (R(0.5) * GG));
        t = t * (rx - lx) / (rx + lx + eps);
        if (jin) {
            R a = q[c + ps], b = q[c - ps], cc = q[c - 1 + ps], d = q[c - 1 - ps];
            R eq = (a - b + cc - d) / (a + b + cc + d + eps);
            R ev = R(0.25) * (V[c] + V[c + ps] + V[c - 1] + V[c - 1 + ps]);
            t = t - R(0.5) * uu * ev * eq / GG;
        }
        if (kin) {
            R a = q[c + ld], b = q[c - ld], cc = q[c - 1 + ld], d = q[c - 1 - ld];
            R eq = (a - b + cc - d) / (a + b + cc + d + eps);
            R ev = R(0.25) * (Wn[c] + Wn[c - ld] + Wn[c - 1] + Wn[c - 1 - ld]);
            t = t - R(0.5) * uu * ev * eq / GG;
        }
        tu = t * R(0.5);
    }
    u2[c] = tu;
    // ---- V face (j-1,j)
    R tv = R(0);
    if (j >= 1) {
        R r = qc, l = q[c - ps];
        R vv = V[c];
        R GG = G[c] + G[c - ps];
        R t = ab(vv) * (R(1) - ab(vv) / (R(0.5) * GG));
        t = t * (r - l) / (r + l + eps);
        if (iin) {
            R a = q[c + 1 - ps], b = q[c - 1], cc = q[c + 1], d = q[c - 1 - ps];
            R eq = (a - b + cc - d) / (cc + a + b + d + eps);
            R ev = R(0.25) * (U[c + 1] + U[c + 1 - ps] + U[c] + U[c - ps]);
            t = t - R(0.5) * vv * ev * eq / GG;
        }
        if (kin) {
            R a = q[c + ld - ps], b = q[c - ld], cc = q[c + ld], d = q[c - ld - ps];
            R eq = (a - b + cc - d) / (a + b + cc + d + eps);
            R ev = R(0.25) * (Wn[c] + Wn[c - ld] + Wn[c - ps] + Wn[c - ld - ps]);
            t = t - R(0.5) * vv * ev * eq / GG;
        }
        tv = t * R(0.5);
    }
    v2[c] = tv;
    // ---- W face above k
    R tw = R(0);
    if (k < g.nz - 1) {
        R r = q[c + ld], l = qc;
        R ww = Wn[c];
        R GG = G[c + ld] + G[c];
        R t = ab(ww) * (R(1) - ab(ww) / (R(0.5) * GG));
        t = t * (r - l) / (r + l + eps);
        if (iin) {
            R a = q[c + 1 + ld], b = q[c - 1], cc = q[c + 1], d = q[c - 1 + ld];
            R eq = (a - b + cc - d) / (cc + a + b + d + eps);
            R ev = R(0.25) * (U[c + 1] + U[c + 1 + ld] + U[c] + U[c + ld]);
            t = t - R(0.5) * ww * ev * eq / GG;
        }
        if (jin) {
            R a = q[c + ld + ps], b = q[c - ps], cc = q[c + ps], d = q[c + ld - ps];
            R eq = (a - b + cc - d) / (cc + d + a + b + eps);
            R ev = R(0.25) * (V[c] + V[c + ld] + V[c + ps] + V[c + ld + ps]);
            t = t - R(0.5) * ww * ev * eq / GG;
        }
        tw = t * R(0.5) * dz[c];
    }
    w2[c] = tw;
}

// ---------------------------------------------------------------------------
// adv_mpdata_FCT_core.f90:47-116 for ONE face, in the independent-faces form
// (the Fortran's carried *_i2 -> *_i temporaries are value reuse only).
// Line of n cells with element stride s; q1/l point at cell 0, U2 at face 0
// (the face between cells 0 and 1).  Returns the limited U2 of face c.
// ---------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R fct_face(const R* __restrict__ q1, const R* __restrict__ l,
                                      const R* __restrict__ U2, size_t s, int c, int n,
                                      bool is_w) {
    const R u = U2[(size_t)c * s];
    if (!(u > R(0)) && !(u < R(0))) return u;
    auto Q = [&](int x) { return q1[(size_t)x * s]; };
    auto L = [&](int x) { return l[(size_t)x * s]; };
    auto F = [&](int x) { return donor(U2[(size_t)x * s], Q(x), Q(x + 1)); };
    auto hi = [&](int x) {
        if (x == 0) return mx(mx(Q(0), Q(1)), mx(L(0), L(1)));
        if (x == n - 1) return mx(mx(Q(n - 2), Q(n - 1)), L(n - 2));
        return mx(mx(mx(Q(x - 1), Q(x)), Q(x + 1)), mx(mx(L(x - 1), L(x)), L(x + 1)));
    };
    auto lo = [&](int x) {
        if (x == 0) return mn(mn(Q(0), Q(1)), mn(L(0), L(1)));
        if (x == n - 1) return mn(mn(Q(n - 2), Q(n - 1)), L(n - 2));
        return mn(mn(mn(Q(x - 1), Q(x)), Q(x + 1)), mn(mn(L(x - 1), L(x)), L(x + 1)));
    };
    auto fin = [&](int x) {
        if (x == 0) return is_w ? R(0) - mn(R(0), F(0)) : R(0);
        if (x == n - 1) { R f = F(n - 2); return is_w ? mx(R(0), f) - mn(R(0), f) : R(0); }
        return mx(R(0), F(x - 1)) - mn(R(0), F(x));
    };
    auto fout = [&](int x) {
        if (x == 0) return is_w ? mx(R(0), F(0)) : R(0);
        if (x == n - 1) { R f = F(n - 2); return is_w ? mx(R(0), f) - mn(R(0), f) : R(0); }
        return mx(R(0), F(x)) - mn(R(0), F(x - 1));
    };
    const R e15 = R(1e-15f);
    if (u > R(0)) {
        R beta_out = (Q(c) - lo(c)) / (fout(c) + e15);
        R beta_in2 = (hi(c + 1) - Q(c + 1)) / (fin(c + 1) + e15);
        return mn(mn(R(1), beta_in2), beta_out) * u;
    } else {
        R beta_in = (hi(c) - Q(c)) / (fin(c) + e15);
        R beta_out2 = (Q(c + 1) - lo(c + 1)) / (fout(c + 1) + e15);
        return mn(mn(R(1), beta_in), beta_out2) * u;
    }
}

// ---------------------------------------------------------------------------
// flux_limiter (adv_mpdata.f90:257-354) fused with the corrector donor-cell pass
// (:389): qout = q2 - div(donor(limited u2,v2,w2)).  l = field before pass 1.
// ---------------------------------------------------------------------------
template <class R, bool FCT>
__global__ void k_corrector(Geom g, const R* __restrict__ l, const R* __restrict__ q2,
                            const R* __restrict__ u2, const R* __restrict__ v2,
                            const R* __restrict__ w2, const R* __restrict__ rho,
                            const R* __restrict__ dz, const R* __restrict__ jaco,
                            R* __restrict__ qout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.nx || k >= g.nz || j >= g.ny) return;
    const size_t ps = g.ps();
    const int ld = g.ld;
    size_t c = g.at(i, k, j);
    R qc = q2[c];
    if (i == 0 || i == g.nx - 1 || j == 0 || j == g.ny - 1) { qout[c] = qc; return; }
    R uxm, uxp, vym, vyp, wzm = R(0), wzp = R(0);
    if (FCT) {
        size_t rowx = g.at(0, k, j);          // x line: faces live at u2[row+1+c]
        uxm = fct_face<R>(q2 + rowx, l + rowx, u2 + rowx + 1, 1, i - 1, g.nx, false);
        uxp = fct_face<R>(q2 + rowx, l + rowx, u2 + rowx + 1, 1, i, g.nx, false);
        size_t rowy = g.at(i, k, 0);          // y line: faces at v2[row + (c+1)*ps]
        vym = fct_face<R>(q2 + rowy, l + rowy, v2 + rowy + ps, ps, j - 1, g.ny, false);
        vyp = fct_face<R>(q2 + rowy, l + rowy, v2 + rowy + ps, ps, j, g.ny, false);
        if (g.nz > 1) {
            size_t rowz = g.at(i, 0, j);      // z line: face above k is w2[row + k*ld]
            if (k > 0) wzm = fct_face<R>(q2 + rowz, l + rowz, w2 + rowz, (size_t)ld, k - 1, g.nz, true);
            if (k < g.nz - 1) wzp = fct_face<R>(q2 + rowz, l + rowz, w2 + rowz, (size_t)ld, k, g.nz, true);
        }
    } else {
        uxm = u2[c]; uxp = u2[c + 1];
        vym = v2[c]; vyp = v2[c + ps];
        if (k > 0) wzm = w2[c - ld];
        wzp = w2[c];                          // top level: w2 == 0 (adv_mpdata.f90:214)
    }
    R fxm = donor(uxm, q2[c - 1], qc);
    R fxp = donor(uxp, qc, q2[c + 1]);
    R fym = donor(vym, q2[c - ps], qc);
    R fyp = donor(vyp, qc, q2[c + ps]);
    R rh = rho ? rho[c] : R(1);
    R jc = jaco[c];
    R val = qc - ((fxp - fxm) + (fyp - fym)) / (jc * rh);
    R den = dz[c] * jc * rh;
    if (g.nz > 1) {
        if (k == 0) {
            val = val - donor(wzp, qc, q2[c + ld]) / den;
        } else if (k == g.nz - 1) {
            val = val - (qc * wzp - donor(wzm, q2[c - ld], qc)) / den;   // wzp == 0 (:322, :214)
        } else {
            val = val - (donor(wzp, qc, q2[c + ld]) - donor(wzm, q2[c - ld], qc)) / den;
        }
    }
    qout[c] = val;
}

}  // namespace icar
#include "upwind_all.cuh"  // k_upwind_all_v4: pass 1 for all scalars, emulatable on the CPU
namespace icar {

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
// cell-wise kernels: one CTA = 128 consecutive x of one (k, j) row (measured faster than (32,4,4)
// blocks for the pass-1 and wind kernels: 212 vs 273 us and 116 vs 128 us at 512x512x40)
static inline dim3 cell_block() { return dim3(128, 1, 1); }
static inline dim3 cell_grid(const Geom& g) {
    return dim3((g.nx + 127) / 128, g.nz, g.ny);
}


int staged_upwind_all(const Geom& g, const ScalarSet& s, const float* U, const float* V, const float* W,
                      const float* dz, const float* jaco, cudaStream_t st) {
    int done = 0, launches = 0;
    while (done < s.n) {                 // batches of up to 5 scalars share the wind / metric loads
        ScalarSet b;
        b.n = (s.n - done >= 5) ? 5 : (s.n - done);
        for (int n = 0; n < b.n; ++n) { b.q[n] = s.q[done + n]; b.q2[n] = s.q2[done + n]; b.out[n] = s.out[done + n]; }
#define UPWIND_ALL(NN) k_upwind_all_v4<NN><<<dim3((g.ld / 4 + 127) / 128, g.nz, g.ny), 128, 0, st>>>(g, b, U, V, W, dz, jaco)
        switch (b.n) {
            case 5: UPWIND_ALL(5); break;
            case 4: UPWIND_ALL(4); break;
            case 3: UPWIND_ALL(3); break;
            case 2: UPWIND_ALL(2); break;
            default: UPWIND_ALL(1); break;
        }
#undef UPWIND_ALL
        done += b.n;
        ++launches;
    }
    return launches;
}

template <class R>
int staged_wind_normalise(const Geom& g, const R* u, const R* v, const R* w, const R* ju,
                          const R* jv, const R* jw, const R* rho, R dt, R dx, int scheme, R* U,
                          R* V, R* W, cudaStream_t st, const R* dz, R* Wn) {
    if constexpr (sizeof(R) == 4) {
        if (!rho && scheme == 2 && Wn && dz && g.ld % 4 == 0) {
            k_wind_normalise_v4<<<dim3((g.ld / 4 + 127) / 128, g.nz, g.ny), 128, 0, st>>>(
                g, (const float*)u, (const float*)v, (const float*)w, (const float*)ju, (const float*)jv,
                (const float*)jw, (float)dt, (float)dx, (float*)U, (float*)V, (float*)W, (const float*)dz, (float*)Wn);
            return 1;
        }
    }
    k_wind_normalise<R><<<cell_grid(g), cell_block(), 0, st>>>(g, u, v, w, ju, jv, jw, rho, dt, dx,
                                                            scheme, U, V, W, dz, Wn);
    return 1;
}

template <class R>
int staged_wn_G(const Geom& g, const R* W, const R* dz, const R* jaco, const R* rho, R* Wn, R* G,
                cudaStream_t st) {
    k_wn_G<R><<<cell_grid(g), cell_block(), 0, st>>>(g, W, dz, jaco, rho, Wn, G);
    return 1;
}

// advect3d (adv_mpdata.f90:356-418).  q: input field; result is left in *result (one of q, qa,
// qb).  qa, qb: two scratch fields; u2,v2,w2: three more.  Returns kernels launched.
template <class R>
int staged_advect3d(const Geom& g, R* q, const R* rho, const R* jaco, const R* dz, const R* U,
                    const R* V, const R* W, const R* Wn, const R* G, int order, bool fct, R* qa,
                    R* qb, R* u2, R* v2, R* w2, R** result, cudaStream_t st) {
    int launches = 0;
    dim3 grid = cell_grid(g);
    // iord = 1: q2 = upwind(q)
    k_upwind<R><<<grid, cell_block(), 0, st>>>(g, q, U, V, W, rho, dz, jaco, qa);
    ++launches;
    if (order <= 1) { *result = qa; return launches; }
    R* lfield = q;     // "q"  in the Fortran: field before the previous pass
    R* q2 = qa;        // "q2" in the Fortran: field after the previous pass
    R* out = qb;
    for (int iord = 2; iord <= order; ++iord) {
        k_pseudo_velocities<R><<<grid, cell_block(), 0, st>>>(g, q2, U, V, Wn, G, dz, u2, v2, w2);
        if (fct) k_corrector<R, true><<<grid, cell_block(), 0, st>>>(g, lfield, q2, u2, v2, w2, rho, dz, jaco, out);
        else     k_corrector<R, false><<<grid, cell_block(), 0, st>>>(g, lfield, q2, u2, v2, w2, rho, dz, jaco, out);
        launches += 2;
        if (iord != order) {
            // adv_mpdata.f90:393-401: q2 = q ; the next limiter then sees l == q1 == this result
            R* spare = (q2 == qa) ? qa : qb;   // the buffer holding the old q2 becomes the next output
            lfield = out;
            q2 = out;
            out = spare;
        } else {
            q2 = out;
        }
    }
    *result = q2;
    return launches;
}

#define INSTANTIATE(R)                                                                             \
    template int staged_wind_normalise<R>(const Geom&, const R*, const R*, const R*, const R*,     \
                                          const R*, const R*, const R*, R, R, int, R*, R*, R*,     \
                                          cudaStream_t, const R*, R*);                                           \
    template int staged_wn_G<R>(const Geom&, const R*, const R*, const R*, const R*, R*, R*,       \
                                cudaStream_t);                                                     \
    template int staged_advect3d<R>(const Geom&, R*, const R*, const R*, const R*, const R*,       \
                                    const R*, const R*, const R*, const R*, int, bool, R*, R*,     \
                                    R*, R*, R*, R**, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)

bool staged_preload() {
    return preload_kernel(k_wind_normalise<float>) && preload_kernel(k_wind_normalise_v4) && preload_kernel(k_wn_G<float>) &&
           preload_kernel(k_upwind<float>) && preload_kernel(k_pseudo_velocities<float>) &&
           preload_kernel(k_corrector<float, true>) && preload_kernel(k_corrector<float, false>) &&
           preload_kernel(k_upwind_all_v4<1>) && preload_kernel(k_upwind_all_v4<2>) && preload_kernel(k_upwind_all_v4<3>) &&
           preload_kernel(k_upwind_all_v4<4>) && preload_kernel(k_upwind_all_v4<5>);
}

}  // namespace icar
