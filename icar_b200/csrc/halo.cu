// Halo pack / unpack for an x-slab decomposition.
// Reference: src/objects/exchangeable_obj.f90:249-356 -- put_east sends
// data_3d(n-2h+1:n-h,:,:) into the east neighbour's halo_west_in, put_west sends
// data_3d(start+h:start+2h-1,:,:) into the west neighbour's halo_east_in; retrieve_* copy
// halo_*_in into data_3d(start:start+h-1) / data_3d(n-h+1:n).  x is the unit-stride
// dimension, so a face is ny*nz runs of h floats: one kernel gathers the faces of all S
// exchanged fields into one contiguous buffer [s][j][k][h] (one message per neighbour).
#include "common.cuh"
#include "halo.cuh"

namespace icar {

namespace {

struct FieldList {
    float* p[ICAR_MAX_EXCHANGED];
};
struct Sides {
    int n;             // 1 or 2 strips
    int x0[2];         // first local column (0-based) of each h-wide strip
    float* buf[2];
};

template <bool PACK>
__global__ void k_halo_copy(Geom g, FieldList f, int nfields, int h, Sides sd) {
    const size_t n = (size_t)nfields * g.ny * g.nz * h;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n * sd.n;
         t += (size_t)gridDim.x * blockDim.x) {
        const int side = (int)(t / n);
        const size_t u = t - (size_t)side * n;
        int x = (int)(u % h);
        size_t r = u / h;
        int k = (int)(r % g.nz);
        r /= g.nz;
        int j = (int)(r % g.ny);
        int s = (int)(r / g.ny);
        size_t c = g.at(sd.x0[side] + x, k, j);
        if (PACK) sd.buf[side][u] = f.p[s][c];
        else      f.p[s][c] = sd.buf[side][u];
    }
}

template <bool PACK>
int launch(const Geom& g, float* const* fields, int nfields, int h, int nsides, const int* x0,
           float* const* buf, cudaStream_t st) {
    if (nsides <= 0) return 0;
    FieldList f;
    for (int s = 0; s < nfields; ++s) f.p[s] = fields[s];
    Sides sd;
    sd.n = nsides;
    for (int i = 0; i < nsides; ++i) { sd.x0[i] = x0[i]; sd.buf[i] = buf[i]; }
    size_t n = (size_t)nfields * g.ny * g.nz * h * nsides;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_halo_copy<PACK><<<blocks, 256, 0, st>>>(g, f, nfields, h, sd);
    return 1;
}

}  // namespace

int halo_pack_launch(const Geom& g, float* const* fields, int nfields, int h, int nsides, const int* x0,
                     float* const* buf, cudaStream_t st) {
    return launch<true>(g, fields, nfields, h, nsides, x0, buf, st);
}

int halo_unpack_launch(const Geom& g, float* const* fields, int nfields, int h, int nsides, const int* x0,
                       float* const* buf, cudaStream_t st) {
    return launch<false>(g, fields, nfields, h, nsides, x0, buf, st);
}

}  // namespace icar
