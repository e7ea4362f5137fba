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

// x0: first local column (0-based) of the h-wide strip
template <bool PACK>
__global__ void k_halo_copy(Geom g, FieldList f, int nfields, int h, int x0,
                            float* __restrict__ buf) {
    const size_t n = (size_t)nfields * g.ny * g.nz * h;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n;
         t += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(t % h);
        size_t r = t / h;
        int k = (int)(r % g.nz);
        r /= g.nz;
        int j = (int)(r % g.ny);
        int s = (int)(r / g.ny);
        size_t c = g.at(x0 + x, k, j);
        if (PACK) buf[t] = f.p[s][c];
        else      f.p[s][c] = buf[t];
    }
}

}  // namespace

int halo_pack_launch(const Geom& g, float* const* fields, int nfields, int h, int x0, float* buf,
                     cudaStream_t st) {
    FieldList f;
    for (int s = 0; s < nfields; ++s) f.p[s] = fields[s];
    size_t n = (size_t)nfields * g.ny * g.nz * h;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_halo_copy<true><<<blocks, 256, 0, st>>>(g, f, nfields, h, x0, buf);
    return 1;
}

int halo_unpack_launch(const Geom& g, float* const* fields, int nfields, int h, int x0,
                       const float* buf, cudaStream_t st) {
    FieldList f;
    for (int s = 0; s < nfields; ++s) f.p[s] = fields[s];
    size_t n = (size_t)nfields * g.ny * g.nz * h;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_halo_copy<false><<<blocks, 256, 0, st>>>(g, f, nfields, h, x0, const_cast<float*>(buf));
    return 1;
}

}  // namespace icar
