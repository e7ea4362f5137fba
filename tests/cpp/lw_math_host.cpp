// TEST INFRASTRUCTURE.  Evaluates the per-wavenumber arithmetic of the linear-winds kernels
// (icar_b200/csrc/lw_math.cuh -- the very functions k_lw_spectrum and k_lw_accumulate call) on the CPU, in
// the kernels' own order, so that tests/test_linear_winds_host.py can compare it with the NumPy restatement
// without a GPU.  Built by the test with g++ -ffp-contract=off; never part of libicar_b200.so.
#include "../../icar_b200/csrc/lw_math.cuh"

using namespace icar::lw;

extern "C" {

// k_lw_spectrum, one loop iteration per thread
void lw_host_spectrum(const double* tf, const float* kx, const float* ly, int fnx, int fny, float U, float V,
                      float Nsq, float z0, float step, int nsteps, double* uhat, double* vhat) {
    for (int j = 0; j < fny; ++j)
        for (int i = 0; i < fnx; ++i) {
            const int si = shifted(i, fnx), sj = shifted(j, fny);
            const double* t = tf + 2 * ((size_t)sj * fnx + si);
            const Mode q = make_mode(U, V, Nsq, kx[si], ly[sj], {t[0], t[1]});
            cplx us = {0.0, 0.0}, vs = {0.0, 0.0};
            float z = z0;
            for (int s = 0; s < nsteps; ++s) {
                cplx uh, vh;
                hats_at(q, z, &uh, &vh);
                us.x += uh.x; us.y += uh.y;
                vs.x += vh.x; vs.y += vh.y;
                z = z + step;
            }
            uhat[2 * ((size_t)j * fnx + i)] = us.x; uhat[2 * ((size_t)j * fnx + i) + 1] = us.y;
            vhat[2 * ((size_t)j * fnx + i)] = vs.x; vhat[2 * ((size_t)j * fnx + i) + 1] = vs.y;
        }
}

// k_lw_shift_norm
void lw_host_shift_norm(const double* in, double* out, int fnx, int fny) {
    const double cells = (double)(fnx * fny);
    for (int j = 0; j < fny; ++j)
        for (int i = 0; i < fnx; ++i) {
            const double* a = in + 2 * ((size_t)j * fnx + i);
            const cplx v = through_single({a[0] / cells, a[1] / cells});
            double* o = out + 2 * ((size_t)shifted(j, fny) * fnx + shifted(i, fnx));
            o[0] = v.x; o[1] = v.y;
        }
}

void lw_host_layer_fraction(const float* zb, const float* zt, int n, float current_z, float step_size, float* out) {
    for (int c = 0; c < n; ++c) out[c] = layer_fraction(current_z, zb[c], zt[c], step_size);
}

}
