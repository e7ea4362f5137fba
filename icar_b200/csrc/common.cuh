// Shared device helpers for the icar_b200 kernels (sm_100a).
#pragma once
#ifndef ICAR_HOST_EMULATION   // (defined by tests/cpp/host_emu.h, which the CPU harnesses of the tests include first)
#include <cuda_runtime.h>
#endif
#include <cstddef>
#include <cstdint>

namespace icar {

// Geometry of one image's tile on the device.  Every 3-D array is stored in the
// reference's order (x fastest, then z, then y) with x-rows padded to `ld`
// elements (a multiple of 32 => every row starts on a 128-byte boundary, and
// TMA's 16-byte stride rule holds for any nx).
struct Geom {
    int nx, nz, ny;      // cells: (ime-ims+1, kme-kms+1, jme-jms+1)
    int ld;              // row pitch in elements for (nx,nz,ny[+1]) arrays
    int ldu;             // row pitch for the x-staggered (nx+1,nz,ny) arrays
    __host__ __device__ inline size_t ps() const { return (size_t)ld * (size_t)nz; }
    __host__ __device__ inline size_t at(int i, int k, int j) const {
        return (size_t)i + (size_t)ld * ((size_t)k + (size_t)nz * (size_t)j);
    }
    __host__ __device__ inline size_t atu(int i, int k, int j) const {
        return (size_t)i + (size_t)ldu * ((size_t)k + (size_t)nz * (size_t)j);
    }
};

__device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float ab(float a) { return fabsf(a); }
__device__ __forceinline__ double ab(double a) { return fabs(a); }

// Donor-cell flux, adv_mpdata.f90:41  f = ((U+|U|)*l + (U-|U|)*r)/2.
// (U+|U|) is exactly 2U or 0 and (U-|U|) exactly 0 or 2U, so the expression equals
// U*l for U>0 and U*r otherwise, bit for bit outside the subnormal range.
template <class R>
__device__ __forceinline__ R donor(R U, R l, R r) {
    return U * (U > R(0) ? l : r);
}

// Forces a kernel's code onto the device now.  CUDA loads kernels lazily, at their first launch, and that load can
// wait for the device to go idle -- which never happens while an image's halo kernels are waiting on the device for a
// neighbour whose own first launches are blocked behind that load.  icar_comm_init* therefore touches every kernel of
// the multi-image step before the first exchange (each translation unit lists its own in a *_preload function).
#ifndef ICAR_HOST_EMULATION
template <class K>
inline bool preload_kernel(K kernel) {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, kernel) == cudaSuccess;
}
#endif

}  // namespace icar
