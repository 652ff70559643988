// Staging-time kernels: re-pack the file's AoS segment records into the device layout of
// device_index.cuh, build the position -> segment bucket tables that replace the interpolation search
// of SegmentIterator::toSite (api/impl/halSegmentIterator.cpp:240-299), and assemble CSR output.
#pragma once
#include "device_index.cuh"

namespace halgpu {

struct PackTopParams {
    const uint8_t *raw; // (n) x 40 B: start, bottomParseIndex, nextParalogyIndex, parentIndex, reversed (mmapTopSegmentData.h:40-44)
    TopRec *out;
    int64_t n;
};
__global__ void packTopKernel(const PackTopParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const int64_t *r = reinterpret_cast<const int64_t *>(p.raw + 40 * i);
        const int64_t start = r[0], botParse = r[1], para = r[2], parent = r[3];
        const bool rev = p.raw[40 * i + 32] != 0;
        TopRec o;
        o.start = start;
        o.parentEnc = parent < 0 ? -1 : ((parent << 1) | (rev ? 1 : 0));
        o.botParse = botParse;
        o.nextPara = para;
        p.out[i] = o;
    }
}

struct PackBotParams {
    const uint8_t *raw; // n x stride: start, topParseIndex, childIndex[nc], childReversed[nc], pad (mmapBottomSegmentData.h:35-52)
    BotCore *core;      // n entries
    int64_t *child;     // nc columns of numBot entries
    int64_t n, numBot;
    int32_t nc, stride;
};
__global__ void packBotKernel(const PackBotParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const uint8_t *rec = p.raw + (int64_t)p.stride * i;
        const int64_t *r = reinterpret_cast<const int64_t *>(rec);
        BotCore o;
        o.start = r[0];
        o.topParse = r[1];
        p.core[i] = o;
        if (i < p.numBot) {
            for (int k = 0; k < p.nc; ++k) {
                const int64_t c = r[2 + k];
                const bool rev = rec[16 + 8 * p.nc + k] != 0;
                p.child[(int64_t)k * p.numBot + i] = c < 0 ? -1 : ((c << 1) | (rev ? 1 : 0));
            }
        }
    }
}

struct BucketParams {
    const void *arr;
    uint32_t *bucket;
    int64_t N, numBuckets;
    int32_t isTop, shift;
};
__global__ void bucketKernel(const BucketParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < p.numBuckets; b += step) {
        const int64_t pos = b << p.shift;
        int64_t lo = 0, hi = p.N; // largest i with start(i) <= pos
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            const int64_t s = p.isTop ? static_cast<const TopRec *>(p.arr)[mid].start : static_cast<const BotCore *>(p.arr)[mid].start;
            if (s <= pos) lo = mid; else hi = mid;
        }
        p.bucket[b] = (uint32_t)lo;
    }
}

struct IotaParams {
    uint32_t *out;
    uint64_t *keys;
    const int64_t *gs;
    int64_t n;
};
__global__ void iotaKeysKernel(const IotaParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        p.out[i] = (uint32_t)i;
        p.keys[i] = (uint64_t)p.gs[i];
    }
}

// compact the ids of intervals whose status == want into list (order irrelevant)
struct CollectParams {
    const uint32_t *status;
    uint32_t *list;
    unsigned long long *count;
    const uint32_t *subset; // optional: only look at these ids
    int64_t n;
    uint32_t want;
};
__global__ void collectKernel(const CollectParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const uint32_t id = p.subset ? p.subset[i] : (uint32_t)i;
        if (p.status[id] == p.want) {
            const unsigned long long at = atomicAdd(p.count, 1ull);
            p.list[at] = id;
        }
    }
}

// how many intervals ended in each state (one pass; the id lists are only built when a retry is needed)
struct StatusCountParams {
    const uint32_t *status;
    unsigned long long *counts; // 4 entries, indexed by ST_*
    int64_t n;
};
__global__ void statusCountKernel(const StatusCountParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    unsigned local[4] = {0u, 0u, 0u, 0u};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const uint32_t s = p.status[i];
        if (s < 4u) local[s]++;
    }
    for (int k = 1; k < 4; ++k)
        if (local[k]) atomicAdd(p.counts + k, (unsigned long long)local[k]);
}

struct AddBaseParams {
    uint64_t *v;
    int64_t n;
    uint64_t base;
};
__global__ void addBaseKernel(const AddBaseParams p) { // chunk-local CSR offsets -> offsets of the whole batch
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) p.v[i] += p.base;
}

// pool (allocation order) -> CSR (input order)
struct GatherParams {
    const uint32_t *outCount;
    const uint64_t *outOffset;
    const uint64_t *csr;
    const halgpu_lift_rec *pool;
    halgpu_lift_rec *recs;
    const uint32_t *pslPool; // optional
    uint32_t *psl;
    int64_t n;
};
__global__ void gatherKernel(const GatherParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const uint32_t c = p.outCount[i];
        const uint64_t from = p.outOffset[i], to = p.csr[i];
        for (uint32_t k = 0; k < c; ++k) p.recs[to + k] = p.pool[from + k];
        if (p.pslPool)
            for (uint32_t k = 0; k < 4 * c; ++k) p.psl[4 * to + k] = p.pslPool[4 * from + k];
    }
}

} // namespace halgpu
