// Host side of the TMA-fed kernels: tensor maps of (nx, nz, ny) fp32 fields read in boxes of 32 x nz x 1 (cached per field
// pointer and geometry).  cuTensorMapEncodeTiled is looked up through the runtime, so libcuda is not a link dependency.
#pragma once
#include <cuda.h>

#include <unordered_map>

#include "common.cuh"

namespace icar {
namespace tma {
constexpr int BOX_W = 32;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* p; int nx, nz, ny, ld;
    bool operator==(const MapKey& o) const { return p == o.p && nx == o.nx && nz == o.nz && ny == o.ny && ld == o.ld; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const { return std::hash<const void*>()(k.p) ^ (size_t)k.nx * 1315423911u ^ (size_t)k.ny * 2654435761u; }
};

// a (nx, nz, ny) fp32 field with row pitch ld, read in boxes of 32 x nz x 1
inline bool field_map(const Geom& g, const float* ptr, CUtensorMap* out) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, g.nx, g.nz, g.ny, g.ld};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.nz, (cuuint64_t)g.ny};
    cuuint64_t gstr[2] = {(cuuint64_t)g.ld * sizeof(float), (cuuint64_t)g.ld * g.nz * sizeof(float)};
    cuuint32_t bx[3] = {(cuuint32_t)BOX_W, (cuuint32_t)g.nz, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}

}  // namespace tma
}  // namespace icar
