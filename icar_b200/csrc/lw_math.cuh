// Per-wavenumber arithmetic of ICAR's linear-theory wind perturbation (Barstad & Gronas 2006, appendix A),
// src/physics/linear_winds.f90:181-237 (linear_perturbation_at_height), with the typing of linear_theory_type
// (src/main/data_structures.f90:184-195): k, l, kl, sig are REAL; denom, msq, m, ineta are complex double.
//
// The functions are __host__ __device__ so that tests/cpp/lw_math_host.cpp can evaluate exactly this code
// on the CPU against the NumPy restatement; the product only ever calls them from kernels.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define LW_HD __host__ __device__ __forceinline__
#else
#define LW_HD inline
#endif

namespace icar {
namespace lw {

struct cplx { double x, y; };

constexpr float kSmallValue = 1e-15f;                 // linear_winds.f90:131

LW_HD cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// complex division the way gfortran's default (-fcx-fortran-rules) evaluates it: Smith's range reduction
LW_HD cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) < fabs(b.y)) {
        const double r = b.x / b.y, d = b.x * r + b.y;
        return {(a.x * r + a.y) / d, (a.y * r - a.x) / d};
    }
    const double r = b.y / b.x, d = b.y * r + b.x;
    return {(a.y * r + a.x) / d, (a.y - a.x * r) / d};
}

// ii(i) of fftshift.f90:87-109 on 0-based indices: element i moves to (i + (n+1)/2) mod n under fftshift;
// ifftshift (:217-239) reads its element i from that position.
LW_HD int shifted(int i, int n) { return (i + (n + 1) / 2) % n; }

// the single-precision temporary both shifts go through (fftshift.f90:91, :221)
LW_HD cplx through_single(cplx a) { return {(double)(float)a.x, (double)(float)a.y}; }

// Everything of linear_perturbation_at_height that does not depend on the height z, for one wavenumber.
struct Mode {
    cplx m;         // vertical wavenumber: real (propagating, sign of sig) or +i sqrt(-msq) (evanescent)
    cplx iterrain;  // imaginary_number * fourier_terrain
    cplx den;       // kl / ((0 - m) * sig)
    float k, l;
};

LW_HD Mode make_mode(float U, float V, float Nsq, float k, float l, cplx fourier_terrain) {
    Mode q;
    q.k = k;
    q.l = l;
    float kl = k * k + l * l;                                   // :471
    if (kl == 0.0f) kl = kSmallValue;                           // :472
    const float uk = U * k, vl = V * l;
    float sig = uk + vl;                                        // :209 (REAL)
    if (sig == 0.0f) sig = kSmallValue;                         // :210
    const float sig2 = sig * sig;                               // sig**2 is REAL, stored in a complex double (:212)
    const double msq = ((double)Nsq / (double)sig2) * (double)kl;   // :214, imaginary part identically zero
    if (msq < 0.0) {
        q.m = {0.0, sqrt(-msq)};                                // :215-216, :220
    } else {
        const double r = sqrt(msq);                             // :218
        q.m = {sig < 0.0f ? -r : r, 0.0};                       // :219
    }
    q.iterrain = {-fourier_terrain.y, fourier_terrain.x};
    const cplx c = {(0.0 - q.m.x) * (double)sig, (0.0 - q.m.y) * (double)sig};
    q.den = cdiv({(double)kl, 0.0}, c);                         // :224
    return q;
}

// ineta at height z (:222-224): i * terrain_hat * exp(i m z) / (kl / (-m sig))
LW_HD cplx ineta_at(const Mode& q, float z) {
    const double zr = (double)z;
    const double amp = exp(-q.m.y * zr);
    const double ph = q.m.x * zr;
    const cplx e = {amp * cos(ph), amp * sin(ph)};
    return cdiv(cmul(q.iterrain, e), q.den);
}

// uhat, vhat of one height (:225-226) as ifftshift stores them (through single precision)
LW_HD void hats_at(const Mode& q, float z, cplx* uhat, cplx* vhat) {
    const cplx n = ineta_at(q, z);
    *uhat = through_single({(double)q.k * n.x, (double)q.k * n.y});
    *vhat = through_single({(double)q.l * n.x, (double)q.l * n.y});
}

// layer_fraction of linear_perturbation_varyingz (:326-328), all REAL
LW_HD float layer_fraction(float current_z, float zb, float zt, float step_size) {
    const float half = step_size / 2.0f;
    const float a = fminf(half, current_z - zb) + fminf(0.0f, zt - current_z);
    const float b = a + fminf(half, zt - current_z);
    const float c = b + fminf(0.0f, current_z - zb);
    return fmaxf(0.0f, c) / step_size;
}

}  // namespace lw
}  // namespace icar
