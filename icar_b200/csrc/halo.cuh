// Halo exchange between the images of a slab decomposition, written by the producer straight into the consumer's
// ghost cells over NVLink (peer-mapped device memory): kernels of halo.cu.  Return = kernels queued on `st`.
#pragma once
#include "common.cuh"

#define ICAR_MAX_EXCHANGED 12   /* domain%halo_send lists 12 exchangeable scalars (domain_obj.f90:104-122) */

namespace icar {

// The flag block every image keeps in its (peer-mapped) arena.  All values are exchange counters: they only grow.
enum {
    HF_STEP = 0,        // exchanges this image has started (written by its own k_halo_push)
    HF_TICKET,          // block ticket of k_halo_push
    HF_DATA_LO,         // written by the lower neighbour: "my boundary cells of exchange s are in your ghost cells"
    HF_DATA_HI,         // ... by the upper neighbour
    HF_READY_LO,        // written by the lower neighbour: "everything that read or wrote my ghost cells before exchange s
    HF_READY_HI,        //     has finished: you may overwrite them" ... by the upper neighbour
    HF_COUNT = 32
};

// one neighbour (side 0: lower image -- west / south; side 1: upper image -- east / north)
struct HaloSide {
    int present;
    int src0;                               // first local index (along the cut axis) of the h owned cells it needs
    int dst0;                               // first local index, in ITS tile, of the ghost cells they fill
    int dst_ld, dst_ny;                     // its row pitch / rows (x-slabs: pitch differs; y-slabs: rows differ)
    float* dst[ICAR_MAX_EXCHANGED];         // its fields (the buffers that hold them during this exchange)
    unsigned* flags;                        // its flag block
};
struct HaloLinks {
    int nfields, axis, h;                   // axis 0: x-slabs (h columns per row), 1: y-slabs (h contiguous planes)
    const float* src[ICAR_MAX_EXCHANGED];   // this image's fields
    unsigned* flags;                        // this image's flag block
    HaloSide side[2];
};

// icar_halo_send = begin (tell both neighbours that this image's ghost cells may be overwritten, wait until they have
// said the same) + push (copy the boundary cells of all fields into the neighbours' ghost cells, then raise their data
// flags); icar_halo_retrieve = wait (until both neighbours' data flags have reached this image's exchange count).
int halo_begin_launch(const HaloLinks& L, cudaStream_t st);
int halo_push_launch(const Geom& g, const HaloLinks& L, cudaStream_t st);
int halo_wait_launch(const HaloLinks& L, cudaStream_t st);
bool halo_preload();           // see preload_kernel (common.cuh)

}  // namespace icar
