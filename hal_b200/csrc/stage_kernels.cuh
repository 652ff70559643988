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
        o.parentEnc = parent < 0 ? -1 : makeLink(parent, rev, 0); // the run field is filled in by linkRunKernel
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
                p.child[(int64_t)k * p.numBot + i] = c < 0 ? -1 : makeLink(c, rev, 0);
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

// Collinear runs of the vertical links (device_index.cuh).  linkBreakKernel marks the segments after which the run ends,
// a suffix-min scan turns the marks into "index of the next break at or after i", linkRunKernel stores the distance in bases.
struct LinkRunParams {
    int64_t *links;        // link of segment i at links[i * linkStride]
    const int64_t *starts; // start of segment i at starts[i * startStride]; entry n is the sentinel (genome length)
    int64_t linkStride, startStride, n;
    const TopRec *landTop; // child links: top array of the child genome (paralogy rings end a run); NULL for parent links
    uint32_t *mark;        // n entries: i if the run ends with segment i, else 0xffffffff  ->  after the scan: next break >= i
};
__device__ __forceinline__ bool linkRing(const LinkRunParams &p, int64_t e) {
    return p.landTop != nullptr && p.landTop[linkIdx(e)].nextPara >= 0;
}
__global__ void linkBreakKernel(const LinkRunParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        bool cont = false;
        if (i + 1 < p.n) {
            const int64_t a = p.links[i * p.linkStride], b = p.links[(i + 1) * p.linkStride];
            if (a >= 0 && b >= 0 && linkRev(a) == linkRev(b) && linkIdx(b) == linkIdx(a) + (linkRev(a) ? -1 : 1)) {
                cont = !linkRing(p, a) && !linkRing(p, b);
            }
        }
        p.mark[i] = cont ? 0xffffffffu : (uint32_t)i;
    }
}
__global__ void linkRunKernel(const LinkRunParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const int64_t e = p.links[i * p.linkStride];
        if (e < 0) continue;
        int64_t run = 0;
        if (!linkRing(p, e)) {
            run = p.starts[((int64_t)p.mark[i] + 1) * p.startStride] - p.starts[i * p.startStride];
            if (run > LINK_RUN_CAP) run = LINK_RUN_CAP;
        }
        p.links[i * p.linkStride] = makeLink(linkIdx(e), linkRev(e), run);
    }
}

// xlate constant of every link (device_index.cuh): with s0 / L the start / length of segment i and t0 the start of the
// segment it links to, forward q -> q + (t0 - s0); reversed q -> (t0 + L - 1) - (q - s0) = (t0 + L - 1 + s0) - q
struct LinkXlateParams {
    const int64_t *links, *starts, *otherStarts;
    int64_t linkStride, startStride, otherStride, n;
    int64_t *xlate;
};
__global__ void linkXlateKernel(const LinkXlateParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const int64_t e = p.links[i * p.linkStride];
        int64_t x = 0;
        if (e >= 0) {
            const int64_t s0 = p.starts[i * p.startStride], L = p.starts[(i + 1) * p.startStride] - s0;
            const int64_t t0 = p.otherStarts[linkIdx(e) * p.otherStride];
            x = linkRev(e) ? t0 + L - 1 + s0 : t0 - s0;
        }
        p.xlate[i] = x;
    }
}

// FastRec table of one vertical transition (device_index.cuh)
struct FastIndexParams {
    const uint32_t *bucket;
    const int64_t *links, *starts, *xlate;
    int64_t linkStride, startStride, numBuckets;
    FastRec *out;
};
__global__ void fastIndexKernel(const FastIndexParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < p.numBuckets; b += step) {
        const int64_t i = (int64_t)p.bucket[b];
        FastRec r;
        r.start = p.starts[i * p.startStride]; r.link = p.links[i * p.linkStride]; r.xlate = p.xlate[i]; r.idx = i;
        p.out[b] = r;
    }
}

// sort input: key = source start (sorted on its upper bits only), value = interval id | min(length, 2^32 - 1) << 32
struct IotaParams {
    uint64_t *vals;             // NULL: no (key, value) sort input wanted
    uint64_t *keys;
    uint64_t *packed;           // optional, instead of vals/keys: one 64-bit word per interval, source start << 32 | interval id
                                // (source genomes shorter than 2^32: half the sort traffic; the kernels re-read the end)
    const int64_t *gs, *ge;
    unsigned long long *directLoc; // optional: outLoc[i] = "one record in pool slot i" (what fastLiftKernel leaves behind on success)
    int64_t n;
};
__global__ void iotaKeysKernel(const IotaParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        if (p.vals) {
            const int64_t a = p.gs[i], len = p.ge[i] - a + 1;
            const uint64_t l32 = (len <= 0 || len >= 0xffffffffll) ? 0xffffffffull : (uint64_t)len;
            p.vals[i] = (uint64_t)i | (l32 << 32);
            p.keys[i] = (uint64_t)a;
        }
        if (p.packed) {
            const int64_t a = p.gs[i];
            // an out-of-range start becomes 2^32 - 1, beyond any such genome: the interval fails the range test like the original
            p.packed[i] = ((a < 0 || a >= 0xffffffffll ? 0xffffffffull : (uint64_t)a) << 32) | (uint64_t)i;
        }
        if (p.directLoc) p.directLoc[i] = ((unsigned long long)i << HG_LOC_COUNT_BITS) | 1ull;
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

struct AddBaseParams {
    uint64_t *v;
    int64_t n;
    uint64_t base;
};
__global__ void addBaseKernel(const AddBaseParams p) { // chunk-local CSR offsets -> offsets of the whole batch
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) p.v[i] += p.base;
}

// CSR offsets = exclusive scan of the per-interval record counts (low bits of outLoc), as three small kernels so that the
// whole scan can be skipped on the device: when fastLiftKernel finished the batch alone (*skipIfZero == 0) every interval
// has exactly one record and offsets[i] = i was already written by csrIdentityKernel.
#define HG_SCAN_BLOCK 256
#define HG_SCAN_ITEMS 8 // per thread: a block covers 2048 intervals
struct CsrScanParams {
    const unsigned long long *outLoc; // n entries
    uint64_t *csr;                    // n + 1 entries
    uint64_t *blockSums;              // ceil(n / 2048) + 1 entries
    const unsigned long long *skipIfZero;
    int64_t n;
};
__global__ void csrIdentityKernel(const CsrScanParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= p.n; i += step) p.csr[i] = (uint64_t)i;
}
__device__ __forceinline__ uint64_t csrBlockScan(uint64_t v, uint64_t *warpSums, uint64_t &blockTotal) { // exclusive, 256 threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = v;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    uint64_t base = 0, total = 0;
    for (int w = 0; w < HG_SCAN_BLOCK / 32; ++w) {
        const uint64_t sW = warpSums[w];
        if (w < warp) base += sW;
        total += sW;
    }
    __syncthreads();
    blockTotal = total;
    return base + incl - v;
}
template <int PHASE> // 0: block sums, 2: final offsets
__global__ void __launch_bounds__(HG_SCAN_BLOCK) csrScanKernel(const CsrScanParams p) {
    if (p.skipIfZero != nullptr && *p.skipIfZero == 0ull) return;
    __shared__ uint64_t warpSums[HG_SCAN_BLOCK / 32];
    const uint64_t mask = (1ull << HG_LOC_COUNT_BITS) - 1ull;
    const int64_t perBlock = (int64_t)HG_SCAN_BLOCK * HG_SCAN_ITEMS;
    const int64_t nBlocks = (p.n + perBlock - 1) / perBlock;
    for (int64_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const int64_t first = b * perBlock + (int64_t)threadIdx.x * HG_SCAN_ITEMS;
        uint64_t c[HG_SCAN_ITEMS], mine = 0;
        for (int k = 0; k < HG_SCAN_ITEMS; ++k) {
            c[k] = first + k < p.n ? (uint64_t)p.outLoc[first + k] & mask : 0ull;
            mine += c[k];
        }
        uint64_t total;
        uint64_t at = csrBlockScan(mine, warpSums, total);
        if (PHASE == 0) {
            if (threadIdx.x == 0) p.blockSums[b] = total;
        } else {
            at += p.blockSums[b];
            for (int k = 0; k < HG_SCAN_ITEMS; ++k) {
                if (first + k < p.n) p.csr[first + k] = at;
                at += c[k];
            }
            if (b == nBlocks - 1 && threadIdx.x == HG_SCAN_BLOCK - 1) p.csr[p.n] = at;
        }
    }
}
// phase 1: exclusive scan of the block sums by one block
__global__ void __launch_bounds__(HG_SCAN_BLOCK) csrScanSumsKernel(const CsrScanParams p) {
    if (p.skipIfZero != nullptr && *p.skipIfZero == 0ull) return;
    __shared__ uint64_t warpSums[HG_SCAN_BLOCK / 32];
    const int64_t perBlock = (int64_t)HG_SCAN_BLOCK * HG_SCAN_ITEMS;
    const int64_t nBlocks = (p.n + perBlock - 1) / perBlock;
    uint64_t carry = 0;
    for (int64_t base = 0; base < nBlocks; base += HG_SCAN_BLOCK) {
        const int64_t i = base + threadIdx.x;
        const uint64_t v = i < nBlocks ? p.blockSums[i] : 0ull;
        uint64_t total;
        const uint64_t at = csrBlockScan(v, warpSums, total);
        if (i < nBlocks) p.blockSums[i] = carry + at;
        carry += total;
    }
    if (p.n == 0 && threadIdx.x == 0) p.csr[0] = 0;
}

// pool (allocation order) -> CSR (input order).  One warp per 32 intervals: when every interval of the tile has at most two
// records each lane copies its own, otherwise the warp copies interval after interval with 16-byte units across the lanes.
struct GatherParams {
    const unsigned long long *outLoc;
    const uint64_t *csr;
    const halgpu_lift_rec *pool;
    halgpu_lift_rec *recs;
    const uint32_t *pslPool; // optional
    uint32_t *psl;
    const unsigned long long *skipIfZero; // optional: nothing to do when this device counter is 0 (every record already sits in
                                          // its final slot: a batch fastLiftKernel finished alone)
    int64_t n;
};
__global__ void __launch_bounds__(256) gatherKernel(const GatherParams p) {
    if (p.skipIfZero != nullptr && *p.skipIfZero == 0ull) return;
    const int lane = threadIdx.x & 31;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t nRound = (p.n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRound; i += step) {
        uint32_t c = 0;
        uint64_t from = 0, to = 0;
        if (i < p.n) {
            const unsigned long long loc = p.outLoc[i];
            c = (uint32_t)(loc & ((1ull << HG_LOC_COUNT_BITS) - 1ull));
            from = loc >> HG_LOC_COUNT_BITS; to = p.csr[i];
        }
        if (!__any_sync(0xffffffffu, c > 2u)) {
            const longlong2 *src = reinterpret_cast<const longlong2 *>(p.pool + from);
            longlong2 *dst = reinterpret_cast<longlong2 *>(p.recs + to);
            for (uint32_t k = 0; k < 2 * c; ++k) dst[k] = src[k];
            if (p.pslPool)
                for (uint32_t k = 0; k < 4 * c; ++k) p.psl[4 * to + k] = p.pslPool[4 * from + k];
            continue;
        }
        unsigned todo = __ballot_sync(0xffffffffu, c > 0u);
        while (todo) {
            const int j = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const uint32_t cj = __shfl_sync(0xffffffffu, c, j);
            const uint64_t fj = __shfl_sync(0xffffffffu, from, j), tj = __shfl_sync(0xffffffffu, to, j);
            const longlong2 *src = reinterpret_cast<const longlong2 *>(p.pool + fj);
            longlong2 *dst = reinterpret_cast<longlong2 *>(p.recs + tj);
            for (uint32_t u = (uint32_t)lane; u < 2 * cj; u += 32) dst[u] = src[u];
            if (p.pslPool) {
                const longlong2 *ps = reinterpret_cast<const longlong2 *>(p.pslPool + 4 * fj);
                longlong2 *pd = reinterpret_cast<longlong2 *>(p.psl + 4 * tj);
                for (uint32_t u = (uint32_t)lane; u < cj; u += 32) pd[u] = ps[u];
            }
        }
    }
}

} // namespace halgpu
