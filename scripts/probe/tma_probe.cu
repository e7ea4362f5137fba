// Standalone TMA probe: one CTA loads a 32 x nz x 1 box of a (nx,nz,ny) fp32 field.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

struct alignas(64) Maps { CUtensorMap m[4]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ Maps maps, int which, int x0, int plane, int nz, float* out, int variant) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* dst = reinterpret_cast<float*>(sm);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + nz * 32 * 4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        if (variant & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nz * 32 * 4) : "memory");
        const CUtensorMap* mp = (variant & 2) ? &maps.m[which] : &maps.m[0];
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(mp)), "r"(smem_u32(bar)), "r"(x0), "r"(0), "r"(plane) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    for (int t = threadIdx.x; t < nz * 32; t += blockDim.x) out[t] = dst[t];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int x0 = argc > 2 ? atoi(argv[2]) : 0;
    int plane = argc > 3 ? atoi(argv[3]) : 1;
    const int nx = 64, nz = 16, ny = 8, ld = 64;
    std::vector<float> h((size_t)ld * nz * ny);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, nz * 32 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    Maps maps;
    cuuint64_t gdim[3] = {nx, nz, ny}; cuuint64_t gstr[2] = {ld * 4, (cuuint64_t)ld * nz * 4};
    cuuint32_t box[3] = {32, nz, 1}; cuuint32_t es[3] = {1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        CUresult r = fn(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    size_t smem = nz * 32 * 4 + 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(maps, 2, x0, plane, nz, o, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d x0 %d plane %d: %s\n", variant, x0, plane, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> r(nz * 32);
        cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int k = 0; k < nz; ++k) for (int x = 0; x < 32; ++x) {
            int gx = x0 + x; float want = (gx >= 0 && gx < nx && plane >= 0 && plane < ny) ? h[(size_t)gx + ld * (k + (size_t)nz * plane)] : 0.f;
            if (r[k * 32 + x] != want) ++bad;
        }
        printf("  mismatches: %d (r[0]=%g r[33]=%g)\n", bad, r[0], r[33]);
    }
    return 0;
}
