// wiggle_kernels.cuh -- plumbing kernels of the wiggle liftover (the mapping + scatter itself is the wiggle mode of
// liftoverKernel, liftover_kernel.cuh): the per-target-base key array that replaces WiggleTiles<double>
// (liftover/inc/halWiggleTiles.h), the --append preload (WiggleLoader::visitLine, liftover/impl/halWiggleLoader.cpp:37-48)
// and the read-out of the bases that hold a value (the exists() sweep of WiggleLiftover::write,
// liftover/impl/halWiggleLiftover.cpp:160-198).
#pragma once
#include "liftover_kernel.cuh"

namespace halgpu {

struct WigFillParams {
    unsigned long long *keys;
    int64_t n;
};
__global__ void wigFillKernel(const WigFillParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) p.keys[i] = WIG_UNSET;
}

struct WigPreloadParams { // WiggleTiles::set: plain store (positions are distinct: the host keeps the last line of each)
    unsigned long long *keys;
    const int64_t *pos;
    const double *val;
    int64_t n, genomeLen;
};
__global__ void wigPreloadKernel(const WigPreloadParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const int64_t x = p.pos[i];
        if (x < 0 || x >= p.genomeLen) continue; // validated on the host; never write outside the array
        unsigned long long k = wigKey(p.val[i]);
        if (k == WIG_UNSET) k = WIG_ZERO; // a preloaded -0.0 is kept as +0.0 (the key of -0.0 is the "unset" mark)
        p.keys[x] = k;
    }
}

struct WigFlagParams { // 1 where a base holds a value
    const unsigned long long *keys;
    uint8_t *flag;
    int64_t n;
};
__global__ void wigFlagKernel(const WigFlagParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) p.flag[i] = p.keys[i] != WIG_UNSET ? 1 : 0;
}

struct WigGatherParams {
    const unsigned long long *keys;
    const int64_t *pos; // ascending positions of the set bases
    double *val;
    int64_t n;
};
__global__ void wigGatherKernel(const WigGatherParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) p.val[i] = wigUnkey(p.keys[p.pos[i]]);
}

} // namespace halgpu
