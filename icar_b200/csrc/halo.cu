// Halo exchange of a slab decomposition without a message layer: every image's advected scalars live in one
// peer-mapped arena (CUDA IPC between processes, plain pointers inside one), and the image that owns the boundary cells
// stores them straight into its neighbours' ghost cells over NVLink, then raises a flag there.
//
// Reference: src/objects/exchangeable_obj.f90:249-356 -- put_east sends data_3d(n-2h+1:n-h,:,:) into the east
// neighbour's halo_west_in, put_west sends data_3d(start+h:start+2h-1,:,:) into the west neighbour's halo_east_in (coarray
// puts: one-sided writes into the neighbour's memory, which is what this is), `sync images` orders them, retrieve_* copy
// halo_*_in into the ghost cells.  Here the put lands in the ghost cells themselves, so retrieve is only the wait.
// All exchanged fields (domain%halo_send, domain_obj.f90:104-122) and both sides travel in ONE kernel.
//
// Ordering (all flags are exchange counters in the receiving image's memory, written with release / read with acquire at
// system scope; s = this image's exchange number):
//   k_halo_begin  READY(s) -> both neighbours; wait READY(s) from both.  Queued after everything of this image that
//                 reads or writes its own ghost cells (the previous advect), so "ready" means exactly that.
//   k_halo_push   copy; __threadfence_system; the last block raises DATA(s) at both neighbours and sets STEP = s.
//   k_halo_wait   wait DATA(STEP) from both neighbours; queued before the first consumer of the ghost cells.
// Both neighbours signal before they wait, so the rendezvous cannot deadlock; no kernel takes a step number as an
// argument, so the sequence can be replayed from a CUDA graph.  Only single-warp kernels spin (begin, wait): several
// images on ONE device (the single-GPU tests) cannot starve each other.
#include "common.cuh"
#include "halo.cuh"

namespace icar {

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spins until *p >= want; a neighbour that never arrives (a crashed image) ends in a trap after ~30 s instead of a hang
__device__ __forceinline__ void wait_at_least(const unsigned* p, unsigned want) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - want) < 0) {
        __nanosleep(64);
        if (clock64() - t0 > 60000000000LL) __trap();
    }
}

__global__ void k_halo_begin(HaloLinks L) {
    const int side = threadIdx.x;
    if (side >= 2 || !L.side[side].present) return;
    const unsigned s = *(volatile unsigned*)(L.flags + HF_STEP) + 1u;
    // I am the upper neighbour of my lower neighbour, and the other way round
    st_release_sys(L.side[side].flags + (side == 0 ? HF_READY_HI : HF_READY_LO), s);
    wait_at_least(L.flags + (side == 0 ? HF_READY_LO : HF_READY_HI), s);
}

__global__ void __launch_bounds__(256)
k_halo_push(Geom g, HaloLinks L) {
    const size_t ps = g.ps();
    for (int side = 0; side < 2; ++side) {
        const HaloSide& S = L.side[side];
        if (!S.present) continue;
        if (L.axis == 1) {
            // y-slabs: h contiguous planes per field; rows are padded to 32 floats, so everything moves as 16-byte words
            const size_t n4 = (size_t)L.h * ps / 4;
            for (int f = 0; f < L.nfields; ++f) {
                const float4* __restrict__ src = reinterpret_cast<const float4*>(L.src[f] + (size_t)S.src0 * ps);
                float4* __restrict__ dst = reinterpret_cast<float4*>(S.dst[f] + (size_t)S.dst0 * ps);
                for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n4; t += (size_t)gridDim.x * blockDim.x)
                    dst[t] = src[t];
            }
        } else {
            // x-slabs: ny * nz runs of h floats per field
            const size_t rows = (size_t)g.ny * g.nz, n = rows * L.h;
            for (int f = 0; f < L.nfields; ++f) {
                const float* __restrict__ src = L.src[f];
                float* __restrict__ dst = S.dst[f];
                for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
                    const int x = (int)(t % L.h);
                    const size_t r = t / L.h;                      // row (k, j) in both tiles: same nz, same ny
                    dst[r * S.dst_ld + S.dst0 + x] = src[r * g.ld + S.src0 + x];
                }
            }
        }
    }
    // every block's stores are visible system-wide before it takes its ticket; the last one to arrive raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(L.flags + HF_TICKET, 1u);
        if (t == gridDim.x - 1) {
            L.flags[HF_TICKET] = 0u;
            const unsigned s = *(volatile unsigned*)(L.flags + HF_STEP) + 1u;
            *(volatile unsigned*)(L.flags + HF_STEP) = s;
            __threadfence_system();
            if (L.side[0].present) st_release_sys(L.side[0].flags + HF_DATA_HI, s);
            if (L.side[1].present) st_release_sys(L.side[1].flags + HF_DATA_LO, s);
        }
    }
}

__global__ void k_halo_wait(HaloLinks L) {
    const int side = threadIdx.x;
    if (side >= 2 || !L.side[side].present) return;
    const unsigned s = *(volatile unsigned*)(L.flags + HF_STEP);
    wait_at_least(L.flags + (side == 0 ? HF_DATA_LO : HF_DATA_HI), s);
}

}  // namespace

int halo_begin_launch(const HaloLinks& L, cudaStream_t st) {
    k_halo_begin<<<1, 32, 0, st>>>(L);
    return 1;
}

int halo_push_launch(const Geom& g, const HaloLinks& L, cudaStream_t st) {
    const size_t cells = (L.axis == 1) ? (size_t)L.h * g.ps() / 4 : (size_t)g.ny * g.nz * L.h;
    int blocks = (int)((cells + 255) / 256);
    if (blocks > 64) blocks = 64;               // a few hundred KB per field: latency-bound, keep the ticket cheap
    if (blocks < 1) blocks = 1;
    k_halo_push<<<blocks, 256, 0, st>>>(g, L);
    return 1;
}

int halo_wait_launch(const HaloLinks& L, cudaStream_t st) {
    k_halo_wait<<<1, 32, 0, st>>>(L);
    return 1;
}

bool halo_preload() { return preload_kernel(k_halo_begin) && preload_kernel(k_halo_push) && preload_kernel(k_halo_wait); }

}  // namespace icar
