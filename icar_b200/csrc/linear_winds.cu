// Linear-theory wind perturbation (Barstad & Gronas 2006) for sm_100a.
//
// Reference routines restated here (src/physics/linear_winds.f90 unless noted):
//   :351-418    add_buffer_topo                    host, once at set-up
//   :426-472    initialize_linear_theory_data      wavenumber grids k, l (REAL, -pi/dx .. +pi/dx inclusive)
//   :1180-1222  setup_linwinds                     buffered terrain -> FFT / N -> fftshift = terrain_frequency
//   :181-237    linear_perturbation_at_height      lw_math.cuh, one thread per wavenumber
//   :239-277    linear_perturbation_constz
//   :280-344    linear_perturbation_varyingz
//   :756-762    the stagger of the LUT entries;  :1094-1111  the update of u, v in spatial_winds
//   src/utilities/fftshift.f90:87-109, 217-239     the shifts (through a single-precision temporary)
//
// What is different from the reference, by design: there is no look-up table.  The reference tabulates the
// perturbation for 10 speeds x 36 directions x 10 stabilities and interpolates (3600 x nx x nz x ny x 2 REALs,
// 7 TB at 2048 x 2048 x 60); here the perturbation is evaluated for the background state it is needed for.
// The 2-D transforms are cuFFT's (the reference calls FFTW for them), bound with dlopen like NCCL; the
// wavenumber-space arithmetic, both shifts, the sub-layer mean, the layer-fraction weighting, the stagger and
// the wind update are kernels of this file.  ifftshift is folded into the spectrum kernel (it writes uhat and
// vhat at their FFT-order positions), and for a layer with constant bounds the mean over its sub-layer heights
// is taken BEFORE the inverse transform, so such a layer costs two FFTs instead of two per sub-layer.
#include "linear_winds.cuh"

#include <cufft.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "lw_kernels.cuh"

namespace icar {

namespace {

struct Cufft {
    void* h = nullptr;
    cufftResult (*Plan2d)(cufftHandle*, int, int, cufftType) = nullptr;
    cufftResult (*ExecZ2Z)(cufftHandle, cufftDoubleComplex*, cufftDoubleComplex*, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool load(std::string* err) {
        if (h) return true;
        for (const char* n : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"}) {
            h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (h) break;
        }
        if (!h) { *err = std::string("cannot dlopen libcufft: ") + dlerror(); return false; }
#define BIND(field, sym)                                                 \
    *(void**)(&field) = dlsym(h, sym);                                   \
    if (!field) { *err = std::string("libcufft lacks ") + sym; h = nullptr; return false; }
        BIND(Plan2d, "cufftPlan2d")
        BIND(ExecZ2Z, "cufftExecZ2Z")
        BIND(SetStream, "cufftSetStream")
        BIND(Destroy, "cufftDestroy")
#undef BIND
        return true;
    }
};
Cufft g_cufft;

inline dim3 grid2d(int nx, int ny) { return dim3((nx + lwk::kThreads - 1) / lwk::kThreads, ny); }
using namespace lwk;

}  // namespace

struct LinWinds {
    int nx = 0, ny = 0;            // the un-buffered domain (global_terrain)
    int fnx = 0, fny = 0;          // the FFT grid
    int buffer = 0;                // original buffer + 2
    cudaStream_t st = nullptr;
    cufftHandle plan = 0;
    bool have_plan = false;
    double2 *tf = nullptr, *uhat = nullptr, *vhat = nullptr, *up = nullptr, *vp = nullptr, *ua = nullptr, *va = nullptr;
    float *kx = nullptr, *ly = nullptr, *ibot = nullptr, *itop = nullptr, *count = nullptr, *zb = nullptr, *zt = nullptr;
    float *upert = nullptr, *vpert = nullptr;      // hi_u_perturbation / hi_v_perturbation of the tile (:1262-1268)
    size_t upert_elems = 0, vpert_elems = 0;
    bool have_perturbation = false;
    std::vector<float> terrain;    // domain%global_terrain, kept for the layer bounds of linear_perturb
    // spatial_winds
    float *z = nullptr, *nsq = nullptr, *u1d = nullptr, *v1d = nullptr, *planes_u = nullptr, *planes_v = nullptr;
    double* rowmeans = nullptr;
    unsigned* bitmap = nullptr;
    int* slot_of = nullptr;
    size_t cell_elems = 0;
    int plane_slots = 0;
};

#define LWCU(call)                                                               \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            *err = std::string(#call) + ": " + cudaGetErrorString(e__);          \
            return 2;                                                            \
        }                                                                        \
    } while (0)
#define LWFFT(call)                                                              \
    do {                                                                         \
        cufftResult r__ = (call);                                                \
        if (r__ != CUFFT_SUCCESS) {                                              \
            *err = std::string(#call) + ": cufft status " + std::to_string((int)r__); \
            return 2;                                                            \
        }                                                                        \
    } while (0)

void lw_add_buffer_topo(const float* terrain, int nx0, int ny0, int smooth_window, int buffer, double* out) {
    const int nx = nx0 + 2 * buffer, ny = ny0 + 2 * buffer;
    auto re = [&](int i, int j) -> double& { return out[2 * ((size_t)j * nx + i)]; };        // 0-based (i, j)
    auto im = [&](int i, int j) -> double& { return out[2 * ((size_t)j * nx + i) + 1]; };
    auto T = [&](int i, int j) { return terrain[(size_t)j * nx0 + i]; };
    float lowest = terrain[0];
    for (size_t c = 0; c < (size_t)nx0 * ny0; ++c) lowest = std::min(lowest, terrain[c]);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) { re(i, j) = (double)lowest; im(i, j) = 0.0; }
    for (int j = 0; j < ny0; ++j)
        for (int i = 0; i < nx0; ++i) re(buffer + i, buffer + j) = (double)T(i, j);
    // west / east rims: the first and last terrain columns blended into each other (:365-370, REAL expressions)
    for (int s = 1; s <= buffer; ++s) {
        const float weight = (float)s / ((float)buffer * 2.0f);
        const int pos = buffer - s;
        for (int j = 0; j < ny0; ++j) {
            const float first = T(0, j), last = T(nx0 - 1, j);
            re(pos, buffer + j) = (double)(first * (1.0f - weight) + last * weight);
            re(nx - pos - 1, buffer + j) = (double)(first * weight + last * (1.0f - weight));
        }
    }
    // south / north rims from the rows just inside them (:371-376): complex double times REAL
    for (int s = 1; s <= buffer; ++s) {
        const float weight = (float)s / ((float)buffer * 2.0f);
        const double w = (double)weight, omw = (double)(1.0f - weight);
        const int pos = buffer - s;
        for (int i = 0; i < nx; ++i) {
            const double lo_r = re(i, buffer), hi_r = re(i, ny - buffer - 1);
            const double lo_i = im(i, buffer), hi_i = im(i, ny - buffer - 1);
            re(i, pos) = lo_r * omw + hi_r * w;
            im(i, pos) = lo_i * omw + hi_i * w;
            re(i, ny - pos - 1) = lo_r * w + hi_r * omw;
            im(i, ny - pos - 1) = lo_i * w + hi_i * omw;
        }
    }
    if (smooth_window <= 0) return;
    // box means over a window that grows away from the terrain, in place and in this order (:380-414)
    auto box_mean = [&](int xs, int xe, int ys, int ye, int ti, int tj) {     // 1-based inclusive bounds and target
        double sr = 0.0, si = 0.0;
        for (int j = ys; j <= ye; ++j)                                         // Fortran sum(): column-major order
            for (int i = xs; i <= xe; ++i) { sr += re(i - 1, j - 1); si += im(i - 1, j - 1); }
        const double cells = (double)((xe - xs + 1) * (ye - ys + 1));
        re(ti - 1, tj - 1) = sr / cells;
        im(ti - 1, tj - 1) = si / cells;
    };
    for (int j = 1; j <= buffer; ++j) {
        const int window = std::min(j, smooth_window);
        for (int i = 1; i <= nx; ++i) {
            const int xs = std::max(1, i - window), xe = std::min(nx, i + window);
            int ys = std::max(1, buffer - j + 1 - window), ye = std::min(ny, buffer - j + 1 + window);
            box_mean(xs, xe, ys, ye, i, buffer - j + 1);
            ys = std::max(1, ny - (buffer - j) - window);
            ye = std::min(ny, ny - (buffer - j) + window);
            box_mean(xs, xe, ys, ye, i, ny - (buffer - j));
        }
        for (int i = 1; i <= ny; ++i) {
            int xs = std::max(1, buffer - j + 1 - window), xe = std::min(nx, buffer - j + 1 + window);
            const int ys = std::max(1, i - window), ye = std::min(ny, i + window);
            box_mean(xs, xe, ys, ye, buffer - j + 1, i);
            xs = std::max(1, nx - (buffer - j) - window);
            xe = std::min(nx, nx - (buffer - j) + window);
            box_mean(xs, xe, ys, ye, nx - (buffer - j), i);
        }
    }
}

void lw_destroy(LinWinds* lw) {
    if (!lw) return;
    if (lw->have_plan && g_cufft.Destroy) g_cufft.Destroy(lw->plan);
    void* bufs[] = {lw->tf, lw->uhat, lw->vhat, lw->up, lw->vp, lw->ua, lw->va, lw->kx, lw->ly, lw->ibot, lw->itop,
                    lw->count, lw->zb, lw->zt, lw->upert, lw->vpert, lw->z, lw->nsq, lw->u1d, lw->v1d, lw->planes_u,
                    lw->planes_v, lw->rowmeans, lw->bitmap, lw->slot_of};
    for (void* p : bufs) if (p) cudaFree(p);
    delete lw;
}

const float* lw_terrain(const LinWinds* lw, int* nx, int* ny) {
    if (nx) *nx = lw->nx;
    if (ny) *ny = lw->ny;
    return lw->terrain.data();
}

void lw_dims(const LinWinds* lw, int* fnx, int* fny, int* buffer) {
    if (fnx) *fnx = lw->fnx;
    if (fny) *fny = lw->fny;
    if (buffer) *buffer = lw->buffer;
}

int lw_create(LinWinds** out, const float* terrain, int nx, int ny, int buffer, float dx, cudaStream_t st,
              std::string* err) {
    *out = nullptr;
    if (!terrain || nx < 2 || ny < 2 || buffer < 1 || !(dx > 0.f)) {
        *err = "linear winds: need a terrain of at least 2 x 2 cells, buffer >= 1 (lt_parameters) and dx > 0";
        return 1;
    }
    if (!g_cufft.load(err)) return 2;
    // two buffering passes (:1199-1207)
    const int n1x = nx + 2 * buffer, n1y = ny + 2 * buffer;
    std::vector<double> first((size_t)2 * n1x * n1y);
    lw_add_buffer_topo(terrain, nx, ny, 5, buffer, first.data());
    std::vector<float> first_real((size_t)n1x * n1y);
    for (size_t c = 0; c < first_real.size(); ++c) first_real[c] = (float)first[2 * c];      // real(real(...))
    const int fnx = n1x + 4, fny = n1y + 4;
    std::vector<double> second((size_t)2 * fnx * fny);
    lw_add_buffer_topo(first_real.data(), n1x, n1y, 0, 2, second.data());

    LinWinds* lw = new LinWinds;
    lw->terrain.assign(terrain, terrain + (size_t)nx * ny);
    lw->nx = nx; lw->ny = ny; lw->fnx = fnx; lw->fny = fny; lw->buffer = buffer + 2; lw->st = st;
    auto bail = [&](int rc) { lw_destroy(lw); return rc; };
    const size_t n = (size_t)fnx * fny;
    for (double2** p : {&lw->tf, &lw->uhat, &lw->vhat, &lw->up, &lw->vp})
        if (cudaMalloc((void**)p, n * sizeof(double2)) != cudaSuccess) { *err = "linear winds: cudaMalloc failed"; return bail(2); }
    if (cudaMalloc((void**)&lw->kx, fnx * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void**)&lw->ly, fny * sizeof(float)) != cudaSuccess) { *err = "linear winds: cudaMalloc failed"; return bail(2); }
    // k, l (:447-463): REAL arithmetic, (i-1) * gain - offset
    const float pi = 3.1415927f;
    const float offset = pi / dx;
    std::vector<float> kx(fnx), ly(fny);
    float gain = 2.0f * offset / (float)(fnx - 1);
    for (int i = 0; i < fnx; ++i) { const float t = (float)i * gain; kx[i] = t - offset; }
    gain = 2.0f * offset / (float)(fny - 1);
    for (int j = 0; j < fny; ++j) { const float t = (float)j * gain; ly[j] = t - offset; }
    cudaError_t e = cudaMemcpyAsync(lw->kx, kx.data(), fnx * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lw->ly, ly.data(), fny * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lw->uhat, second.data(), n * sizeof(double2), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { *err = std::string("linear winds: upload: ") + cudaGetErrorString(e); return bail(2); }
    // Fortran (fnx, fny) with the first index fastest = a C array [fny][fnx]
    if (g_cufft.Plan2d(&lw->plan, fny, fnx, CUFFT_Z2Z) != CUFFT_SUCCESS) { *err = "linear winds: cufftPlan2d failed"; return bail(2); }
    lw->have_plan = true;
    if (g_cufft.SetStream(lw->plan, st) != CUFFT_SUCCESS) { *err = "linear winds: cufftSetStream failed"; return bail(2); }
    if (g_cufft.ExecZ2Z(lw->plan, lw->uhat, lw->up, CUFFT_FORWARD) != CUFFT_SUCCESS) { *err = "linear winds: forward FFT failed"; return bail(2); }
    k_lw_shift_norm<<<grid2d(fnx, fny), kThreads, 0, st>>>(lw->up, lw->tf, fnx, fny, (double)(fnx * fny));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // `second`, kx, ly go out of scope
    if (e != cudaSuccess) { *err = std::string("linear winds: set-up: ") + cudaGetErrorString(e); return bail(2); }
    *out = lw;
    return 0;
}

int lw_terrain_frequency(LinWinds* lw, double* host, std::string* err) {
    LWCU(cudaMemcpyAsync(host, lw->tf, (size_t)lw->fnx * lw->fny * sizeof(double2), cudaMemcpyDeviceToHost, lw->st));
    LWCU(cudaStreamSynchronize(lw->st));
    return 0;
}

int lw_download_perturbation(LinWinds* lw, double* u, double* v, std::string* err) {
    if (!lw->have_perturbation) { *err = "linear winds: no perturbation has been computed yet"; return 1; }
    const size_t bytes = (size_t)lw->fnx * lw->fny * sizeof(double2);
    if (u) LWCU(cudaMemcpyAsync(u, lw->up, bytes, cudaMemcpyDeviceToHost, lw->st));
    if (v) LWCU(cudaMemcpyAsync(v, lw->vp, bytes, cudaMemcpyDeviceToHost, lw->st));
    LWCU(cudaStreamSynchronize(lw->st));
    return 0;
}

static int zero_perturbation(LinWinds* lw, std::string* err) {          // U = V = 0  (:196-200)
    const size_t bytes = (size_t)lw->fnx * lw->fny * sizeof(double2);
    LWCU(cudaMemsetAsync(lw->up, 0, bytes, lw->st));
    LWCU(cudaMemsetAsync(lw->vp, 0, bytes, lw->st));
    lw->have_perturbation = true;
    return 0;
}

int lw_perturbation_constz(LinWinds* lw, float U, float V, float Nsq, float z_bottom, float z_top,
                           float minimum_step, long long* launches, std::string* err) {
    if (U == 0.f && V == 0.f) return zero_perturbation(lw, err);
    if (!(minimum_step > 0.f)) { *err = "linear winds: minimum_layer_size must be positive"; return 1; }
    const int n_steps = std::max(1, (int)std::ceil((z_top - z_bottom) / minimum_step));       // :257
    const float step_size = (z_top - z_bottom) / (float)n_steps;                              // :258
    const float half = step_size / 2.0f;
    const float z0 = z_bottom + half;                                                         // :263
    const size_t n = (size_t)lw->fnx * lw->fny;
    k_lw_spectrum<<<grid2d(lw->fnx, lw->fny), kThreads, 0, lw->st>>>(lw->tf, lw->kx, lw->ly, lw->fnx, lw->fny, U, V,
                                                                     Nsq, z0, step_size, n_steps, lw->uhat, lw->vhat);
    LWCU(cudaGetLastError());
    LWFFT(g_cufft.ExecZ2Z(lw->plan, lw->uhat, lw->up, CUFFT_INVERSE));     // FFTW_BACKWARD: unnormalised, like cuFFT
    LWFFT(g_cufft.ExecZ2Z(lw->plan, lw->vhat, lw->vp, CUFFT_INVERSE));
    int mine = 1;
    if (n_steps > 1) {
        k_lw_divide<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, lw->st>>>(lw->up, lw->vp, n, (double)n_steps);
        LWCU(cudaGetLastError());
        ++mine;
    }
    if (launches) *launches += mine;
    lw->have_perturbation = true;
    return 0;
}

int lw_perturbation_varyingz(LinWinds* lw, float U, float V, float Nsq, const float* z_bottom, const float* z_top,
                             float minimum_step, long long* launches, std::string* err) {
    if (U == 0.f && V == 0.f) return zero_perturbation(lw, err);
    if (!z_bottom || !z_top) { *err = "linear winds: z_bottom / z_top missing"; return 1; }
    const size_t cells = (size_t)lw->nx * lw->ny, n = (size_t)lw->fnx * lw->fny;
    float start_z = z_bottom[0], end_z = z_top[0], thinnest = z_top[0] - z_bottom[0];
    for (size_t c = 0; c < cells; ++c) {
        start_z = std::min(start_z, z_bottom[c]);                      // :301
        end_z = std::max(end_z, z_top[c]);                             // :302
        thinnest = std::min(thinnest, z_top[c] - z_bottom[c]);
    }
    const float step_size = std::min(minimum_step, thinnest);          // :315
    if (!(step_size > 0.f)) { *err = "linear winds: a layer of zero or negative thickness"; return 3; }
    if (!lw->ua) {
        for (double2** p : {&lw->ua, &lw->va}) LWCU(cudaMalloc((void**)p, n * sizeof(double2)));
        for (float** p : {&lw->ibot, &lw->itop, &lw->count}) LWCU(cudaMalloc((void**)p, n * sizeof(float)));
        for (float** p : {&lw->zb, &lw->zt}) LWCU(cudaMalloc((void**)p, cells * sizeof(float)));
    }
    LWCU(cudaMemcpyAsync(lw->zb, z_bottom, cells * sizeof(float), cudaMemcpyHostToDevice, lw->st));
    LWCU(cudaMemcpyAsync(lw->zt, z_top, cells * sizeof(float), cudaMemcpyHostToDevice, lw->st));
    LWCU(cudaMemsetAsync(lw->ua, 0, n * sizeof(double2), lw->st));
    LWCU(cudaMemsetAsync(lw->va, 0, n * sizeof(double2), lw->st));
    LWCU(cudaMemsetAsync(lw->count, 0, n * sizeof(float), lw->st));
    k_lw_layer_bounds<<<grid2d(lw->fnx, lw->fny), kThreads, 0, lw->st>>>(lw->zb, lw->zt, lw->nx, lw->ny, lw->buffer, start_z,
                                                                         end_z, lw->ibot, lw->itop, lw->fnx, lw->fny);
    LWCU(cudaGetLastError());
    int mine = 1;
    const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
    const float half = step_size / 2.0f;
    float current_z = start_z + half;                                  // :320
    while (current_z < end_z) {                                        // :325
        k_lw_spectrum<<<grid2d(lw->fnx, lw->fny), kThreads, 0, lw->st>>>(lw->tf, lw->kx, lw->ly, lw->fnx, lw->fny, U, V,
                                                                         Nsq, current_z, 0.f, 1, lw->uhat, lw->vhat);
        LWFFT(g_cufft.ExecZ2Z(lw->plan, lw->uhat, lw->up, CUFFT_INVERSE));
        LWFFT(g_cufft.ExecZ2Z(lw->plan, lw->vhat, lw->vp, CUFFT_INVERSE));
        k_lw_accumulate<<<blocks, kThreads, 0, lw->st>>>(lw->up, lw->vp, lw->ibot, lw->itop, current_z, step_size,
                                                         lw->ua, lw->va, lw->count, n);
        LWCU(cudaGetLastError());
        mine += 2;
        current_z = current_z + step_size;                             // :335
    }
    k_lw_divide_by_count<<<blocks, kThreads, 0, lw->st>>>(lw->ua, lw->va, lw->count, lw->up, lw->vp, n);
    LWCU(cudaGetLastError());
    // z_bottom / z_top are the caller's: the uploads must have left them before we return
    LWCU(cudaStreamSynchronize(lw->st));
    if (launches) *launches += mine + 1;
    lw->have_perturbation = true;
    return 0;
}

int ensure_perturbation_state(LinWinds* lw, const Geom& g, std::string* err);

int lw_apply_level(LinWinds* lw, const Geom& g, int i0, int j0, int k, float* u3d, float* v3d,
                   float linear_contribution, float update_fraction, long long* launches, std::string* err) {
    if (!lw->have_perturbation) { *err = "linear winds: no perturbation has been computed yet"; return 1; }
    if (i0 < 0 || j0 < 0 || i0 + g.nx > lw->nx || j0 + g.ny > lw->ny || k < 0 || k >= g.nz) {
        *err = "linear winds: the tile does not lie inside the terrain given to icar_linear_winds_init";
        return 1;
    }
    if (int rc = ensure_perturbation_state(lw, g, err)) return rc;
    k_lw_apply_u<<<grid2d(g.nx + 1, g.ny), kThreads, 0, lw->st>>>(lw->up, lw->fnx, lw->buffer, g.nx, g.nz, g.ldu, i0, j0, k, u3d, lw->upert,
                                                                  linear_contribution, update_fraction);
    k_lw_apply_v<<<grid2d(g.nx, g.ny + 1), kThreads, 0, lw->st>>>(lw->vp, lw->fnx, lw->buffer, g.nx, g.nz, g.ld, i0, j0, k, v3d, lw->vpert,
                                                                  linear_contribution, update_fraction);
    LWCU(cudaGetLastError());
    if (launches) *launches += 2;
    return 0;
}

// hi_u_perturbation / hi_v_perturbation of the tile, zero on first use (:1262-1268)
int ensure_perturbation_state(LinWinds* lw, const Geom& g, std::string* err) {
    const size_t ue = (size_t)g.ldu * g.nz * g.ny, ve = (size_t)g.ld * g.nz * (g.ny + 1);
    if (!lw->upert || lw->upert_elems != ue || lw->vpert_elems != ve) {
        if (lw->upert) cudaFree(lw->upert);
        if (lw->vpert) cudaFree(lw->vpert);
        lw->upert = lw->vpert = nullptr;
        LWCU(cudaMalloc((void**)&lw->upert, ue * sizeof(float)));
        LWCU(cudaMalloc((void**)&lw->vpert, ve * sizeof(float)));
        LWCU(cudaMemsetAsync(lw->upert, 0, ue * sizeof(float), lw->st));       // hi_u_perturbation = 0 (:1264)
        LWCU(cudaMemsetAsync(lw->vpert, 0, ve * sizeof(float), lw->st));
        lw->upert_elems = ue;
        lw->vpert_elems = ve;
    }
    return 0;
}

int lw_set_z(LinWinds* lw, const Geom& g, const float* z_host, std::string* err) {
    if (!z_host) { *err = "linear winds: z missing"; return 1; }
    const size_t elems = (size_t)g.ld * g.nz * g.ny;
    if (!lw->z || lw->cell_elems != elems) {
        if (lw->z) cudaFree(lw->z);
        lw->z = nullptr;
        LWCU(cudaMalloc((void**)&lw->z, elems * sizeof(float)));
        LWCU(cudaMemsetAsync(lw->z, 0, elems * sizeof(float), lw->st));
    }
    LWCU(cudaMemcpy2DAsync(lw->z, (size_t)g.ld * sizeof(float), z_host, (size_t)g.nx * sizeof(float),
                           (size_t)g.nx * sizeof(float), (size_t)g.nz * g.ny, cudaMemcpyHostToDevice, lw->st));
    LWCU(cudaStreamSynchronize(lw->st));
    lw->cell_elems = elems;
    return 0;
}

// linear_space (array_utilities.f90:233-235): (i - 1.0) / real(n - 1.0) * (max - min) + min
static void linear_space(float* out, float lo, float hi, int n) {
    for (int i = 0; i < n; ++i) {
        const float t = (float)i / (float)(n - 1);
        const float s = t * (hi - lo);
        out[i] = s + lo;
    }
}

int lw_spatial_winds(LinWinds* lw, const Geom& g, int i0, int j0, const LtFields& f, const LtOptions& o, float* u3d,
                     float* v3d, const float* global_z_interface, const float* global_dz_interface,
                     const float* layer_height, const float* dz_levels, long long* entries_computed,
                     long long* launches, std::string* err) {
    const bool varying = global_z_interface != nullptr;
    if (varying ? !global_dz_interface : !(layer_height && dz_levels)) { *err = "linear winds: layer bounds missing"; return 1; }
    if (i0 < 0 || j0 < 0 || i0 + g.nx > lw->nx || j0 + g.ny > lw->ny) { *err = "linear winds: the tile does not lie inside the terrain"; return 1; }
    if (o.n_dir_values < 2 || o.n_dir_values > 64 || o.n_nsq_values < 2 || o.n_nsq_values > 32 || o.n_spd_values < 2 || o.n_spd_values > 32) {
        *err = "linear winds: table sizes must be 2..64 directions, 2..32 speeds and stabilities";
        return 1;
    }
    if (o.smooth_nsq && (g.nx <= o.stability_window_size)) {
        *err = "linear winds: stability_window_size must be smaller than the tile (smooth_array reads rowmeans(2:windowsize))";
        return 1;
    }
    if (o.variable_N && (!lw->z || lw->cell_elems != (size_t)g.ld * g.nz * g.ny)) { *err = "linear winds: variable_N needs domain%z (icar_linear_winds_set_z)"; return 1; }
    if (int rc = ensure_perturbation_state(lw, g, err)) return rc;

    LutAxes ax{};
    ax.n_dir = o.n_dir_values; ax.n_nsq = o.n_nsq_values; ax.n_spd = o.n_spd_values;
    linear_space(ax.dir, o.dirmin, o.dirmax, ax.n_dir);                       // :657-661
    linear_space(ax.nsq, logf(o.min_stability), logf(o.max_stability), ax.n_nsq);   // nsqmin / nsqmax, options_obj.f90:1472-1473
    linear_space(ax.spd, o.spdmin, o.spdmax, ax.n_spd);
    SpatialOpts so{o.variable_N, o.smooth_nsq, o.vert_smooth, o.stability_window_size, o.min_stability, o.max_stability,
                   o.N_squared, o.linear_contribution, o.linear_update_fraction};
    const int n_entries = ax.n_dir * ax.n_nsq * ax.n_spd, words = (n_entries + 31) / 32;
    const size_t cells = (size_t)g.ld * g.nz * g.ny, nx1 = (size_t)g.nx + 1;
    if (!lw->nsq) {
        LWCU(cudaMalloc((void**)&lw->nsq, cells * sizeof(float)));
        LWCU(cudaMalloc((void**)&lw->rowmeans, cells * sizeof(double)));
        LWCU(cudaMalloc((void**)&lw->u1d, nx1 * (g.ny + 1) * sizeof(float)));
        LWCU(cudaMalloc((void**)&lw->v1d, nx1 * (g.ny + 1) * sizeof(float)));
        LWCU(cudaMalloc((void**)&lw->bitmap, 2048 * sizeof(unsigned)));       // 64 x 32 x 32 entries at most
        LWCU(cudaMalloc((void**)&lw->slot_of, 2048 * 32 * sizeof(int)));
        LWCU(cudaMemsetAsync(lw->nsq, 0, cells * sizeof(float), lw->st));
    }
    long long mine = 0, evaluated = 0;

    // log N^2, vertical running mean, horizontal box mean (:898-959)
    CellFields cf{f.th, f.pii, lw->z, f.qv, f.qc, f.qi, f.qr, f.qs};
    k_lw_nsq_column<<<grid2d(g.nx, g.ny), kThreads, 0, lw->st>>>(g.nx, g.nz, g.ld, cf, so, lw->nsq);
    ++mine;
    if (o.smooth_nsq) {
        const dim3 grid((g.nx + kThreads - 1) / kThreads, g.nz, g.ny);
        k_lw_smooth_y<<<grid, kThreads, 0, lw->st>>>(lw->nsq, lw->rowmeans, g.nx, g.nz, g.ny, g.ld, o.stability_window_size);
        k_lw_smooth_x<<<grid, kThreads, 0, lw->st>>>(lw->rowmeans, lw->nsq, g.nx, g.nz, g.ny, g.ld, o.stability_window_size);
        mine += 2;
    }
    LWCU(cudaGetLastError());

    const size_t stride_u = (size_t)(lw->nx + 1) * lw->ny, stride_v = (size_t)lw->nx * (lw->ny + 1);
    std::vector<float> zb, zt;
    if (varying) { zb.resize((size_t)lw->nx * lw->ny); zt.resize(zb.size()); }
    std::vector<unsigned> bits(words);
    std::vector<int> slot_of(n_entries);

    // rows 0 .. ny-1 carry u and v; row ny (the extra v row) is looked up with the column means of the UPDATED u of
    // row ny-1, because the reference walks the rows in order (:961-973 inside `do k = 1, nyv`)
    for (int phase = 0; phase < 2; ++phase) {
        const int row0 = phase == 0 ? 0 : g.ny, nrows = phase == 0 ? g.ny : 1;
        k_lw_colmean<<<grid2d(g.nx + 1, nrows), kThreads, 0, lw->st>>>(u3d, v3d, g.nx, g.nz, g.ny, g.ld, g.ldu, row0, lw->u1d, lw->v1d);
        ++mine;
        for (int j = 0; j < g.nz; ++j) {
            LWCU(cudaMemsetAsync(lw->bitmap, 0, words * sizeof(unsigned), lw->st));
            k_lw_mark_entries<<<grid2d(g.nx + 1, nrows), kThreads, 0, lw->st>>>(ax, lw->u1d, lw->v1d, lw->nsq, g.nx, g.nz, g.ny, g.ld, j,
                                                                              o.stability_window_size, row0, lw->bitmap);
            ++mine;
            LWCU(cudaGetLastError());
            LWCU(cudaMemcpyAsync(bits.data(), lw->bitmap, words * sizeof(unsigned), cudaMemcpyDeviceToHost, lw->st));
            LWCU(cudaStreamSynchronize(lw->st));
            int used = 0;
            for (int e = 0; e < n_entries; ++e) slot_of[e] = ((bits[e >> 5] >> (e & 31)) & 1u) ? used++ : 0;
            if (used > lw->plane_slots) {
                if (lw->planes_u) cudaFree(lw->planes_u);
                if (lw->planes_v) cudaFree(lw->planes_v);
                lw->planes_u = lw->planes_v = nullptr;
                lw->plane_slots = 0;
                LWCU(cudaMalloc((void**)&lw->planes_u, (size_t)used * stride_u * sizeof(float)));
                LWCU(cudaMalloc((void**)&lw->planes_v, (size_t)used * stride_v * sizeof(float)));
                lw->plane_slots = used;
            }
            if (varying)       // z_interface(:, z, :) - terrain and ... + dz_interface(:, z, :)   (:723-726)
                for (int y = 0; y < lw->ny; ++y)
                    for (int x = 0; x < lw->nx; ++x) {
                        const size_t c3 = (size_t)x + (size_t)lw->nx * ((size_t)j + (size_t)g.nz * (size_t)y);
                        const float b = global_z_interface[c3] - lw->terrain[(size_t)y * lw->nx + x];
                        zb[(size_t)y * lw->nx + x] = b;
                        zt[(size_t)y * lw->nx + x] = b + global_dz_interface[c3];
                    }
            for (int e = 0; e < n_entries; ++e) {
                if (!((bits[e >> 5] >> (e & 31)) & 1u)) continue;
                const int in = e % ax.n_nsq, is = (e / ax.n_nsq) % ax.n_spd, id = e / (ax.n_nsq * ax.n_spd);
                const float U = sinf(ax.dir[id]) * ax.spd[is];           // calc_u / calc_v (atm_utilities.f90:373-391)
                const float V = cosf(ax.dir[id]) * ax.spd[is];
                const float N2 = expf(ax.nsq[in]);                       // :712
                int rc;
                if (varying) {
                    rc = lw_perturbation_varyingz(lw, U, V, N2, zb.data(), zt.data(), o.minimum_layer_size, &mine, err);
                } else {
                    const float half = dz_levels[j] / 2.0f;              // :730-732
                    rc = lw_perturbation_constz(lw, U, V, N2, layer_height[j] - half, layer_height[j] + half,
                                                o.minimum_layer_size, &mine, err);
                }
                if (rc) return rc;
                k_lw_stagger<<<grid2d(lw->nx + 1, lw->ny + 1), kThreads, 0, lw->st>>>(
                    lw->up, lw->vp, lw->fnx, lw->buffer, lw->nx, lw->ny, lw->planes_u + (size_t)slot_of[e] * stride_u,
                    lw->planes_v + (size_t)slot_of[e] * stride_v);
                ++mine;
                ++evaluated;
            }
            LWCU(cudaMemcpyAsync(lw->slot_of, slot_of.data(), n_entries * sizeof(int), cudaMemcpyHostToDevice, lw->st));
            if (phase == 0)
                k_lw_interp<<<grid2d(g.nx + 1, g.ny), kThreads, 0, lw->st>>>(ax, so, lw->u1d, lw->v1d, lw->nsq, g.nx, g.nz, g.ny, g.ld, g.ldu,
                                                                           j, 0, 0, lw->slot_of, lw->planes_u, stride_u, lw->nx, i0, j0,
                                                                           u3d, lw->upert);
            k_lw_interp<<<grid2d(g.nx, nrows), kThreads, 0, lw->st>>>(ax, so, lw->u1d, lw->v1d, lw->nsq, g.nx, g.nz, g.ny, g.ld, g.ld, j,
                                                                     row0, 1, lw->slot_of, lw->planes_v, stride_v, lw->nx, i0, j0, v3d,
                                                                     lw->vpert);
            mine += phase == 0 ? 2 : 1;
            LWCU(cudaGetLastError());
            LWCU(cudaStreamSynchronize(lw->st));                         // slot_of (host) is rewritten for the next level
        }
    }
    if (launches) *launches += mine;
    if (entries_computed) *entries_computed = evaluated;
    return 0;
}

}  // namespace icar
