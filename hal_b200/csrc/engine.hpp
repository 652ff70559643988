// Host-side engine behind the C ABI: owns the mapped HAL file, the staged device index, the per-(src,tgt)
// path plans and the batch pipeline (sort -> map kernel -> retry ladder -> scan -> gather).
#pragma once
#include "device_index.cuh"
#include "halmmap.hpp"
#include "rt.hpp"
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

namespace halgpu {

struct GenomeDev {
    TopRec *top = nullptr;     // numTop + 1
    BotCore *bot = nullptr;    // numBottom + 1
    int64_t *child = nullptr;  // nc columns x numBottom
    FastRec *topFast = nullptr;            // topBuckets: fastLiftKernel's records of the transition to the parent
    std::vector<FastRec *> childFast;      // per child slot: botBuckets records of the transition to that child
    int64_t *topX = nullptr;   // numTop: xlate constant of every parent link
    int64_t *childX = nullptr; // nc columns x numBottom: xlate constant of every child link
    uint8_t *dna = nullptr;    // (length + 1) / 2
    int64_t *seqStart = nullptr; // numSeq + 1
    int32_t *childGenome = nullptr; // nc entries
    uint32_t *topBucket = nullptr, *botBucket = nullptr;
    int topShift = 0, botShift = 0;
    int64_t topBuckets = 0, botBuckets = 0;
    bool staged = false;              // top / bot / child / seqStart / buckets are resident
    std::vector<char> childLinked;    // per child slot: run + xlate fields of that column are filled in
};

struct Plan {
    int src = -1, tgt = -1, mrca = -1;
    int coal = -1;         // coalescence limit genome when it lies above the MRCA, else -1
    std::vector<int> path; // src .. mrca [.. child of the limit .. mrca] .. tgt (genome of every path position)
    int upSteps = 0;
    PathStep *dSteps = nullptr;
    bool fastOk = false;   // every transition has its FastRec table: fastLiftKernel can run this path
    mutable double linesPerInterval = 0; // largest output lines / interval ratio seen on this path: sizes the next batch's record pool
};

struct LiftOutput { // device-side result of one batch; the buffers belong to the context's cache (Context::release)
    uint64_t *offsets = nullptr; // n + 1
    halgpu_lift_rec *recs = nullptr;
    uint32_t *psl = nullptr;     // 4 per record when HALGPU_PSL
    size_t n = 0, nRec = 0, nRetry = 0;
    size_t nComplex = 0;         // intervals the one-lane-per-interval kernel handed to the warp-per-interval walk
    size_t nRedo = 0;            // intervals the fused walk handed back to the piece-by-piece walk
    bool recsExternal = false;   // recs is the caller's ExternalPool buffer (not to be handed back to the cache)
    float kernelMs = 0;          // mapping kernels of the batch (fast + walk + retries)
    float fastMs = 0;            // fastLiftKernel alone (0 when the batch did not use it)
    int launches = 0;
};

// Device buffers recycled across batches: once warm, a batch performs no cudaMalloc / cudaFree at all.
class DeviceCache {
  public:
    ~DeviceCache();
    void *take(size_t bytes);
    void give(void *p);
    size_t bytesHeld() const { return _held; }

  private:
    std::mutex _m;
    std::multimap<size_t, void *> _idle;
    std::map<void *, size_t> _sizeOf;
    size_t _held = 0;
};

// A caller-owned buffer for the record pool of one batch (multi.cu: the communicator's send slot).  Used when it is large
// enough; a batch that the one-lane-per-interval kernel finishes alone then leaves its result right there.
struct ExternalPool {
    void *buf = nullptr;
    size_t bytes = 0;
};

struct WigScatter { // wiggle mode of the mapping kernel (device pointers)
    unsigned long long *keys; // one key per base of the target genome
    const int64_t *valOff;    // per interval
    const double *vals;
};

struct WigOutput { // bases of the target genome that hold a value, ascending (pinned host memory, rt::hostFree)
    int64_t *pos = nullptr;
    double *val = nullptr;
    size_t n = 0, nRetry = 0;
    float kernelMs = 0;
    int launches = 0;
};

class Context {
  public:
    Context(const std::string &path, int device);
    ~Context();
    const HalFile &file() const { return *_file; }
    rt::Stream stream() const { return _stream; }
    rt::Stream copyStream() const { return _copy; }     // host->device traffic of the pipelined host-buffer entry point
    rt::Stream copyBackStream() const { return _copyBack; } // device->host traffic (its own stream: PCIe is full duplex)
    size_t stagedBytes() const { return _staged; }
    int device() const { return _device; }
    // device pointers in, device result out (the caller hands offsets/recs/psl back with release())
    // offsetBase is added to every CSR offset (the pipelined host entry point lifts a batch chunk by chunk)
    void liftover(int src, int tgt, uint32_t flags, size_t n, const int64_t *dGs, const int64_t *dGe,
                  const uint8_t *dStrand, LiftOutput &out, uint64_t offsetBase = 0, const WigScatter *wig = nullptr, int coal = -1,
                  const ExternalPool *ext = nullptr);
    // bytes of record pool the next batch of n intervals on this path will ask for (what an ExternalPool has to hold)
    size_t poolBytesFor(int src, int tgt, size_t n, int coal = -1);

    // wiggle liftover of nRuns source ranges [first, last] (genome coordinates) carrying per-base values (valOff >= 0) or
    // one value (valOff < 0: ~index), onto the target genome preloaded with nPre (position, value) pairs; host in, host out
    void wiggle(int src, int tgt, uint32_t flags, size_t nRuns, const int64_t *first, const int64_t *last, const int64_t *valOff,
                const double *vals, size_t nVals, size_t nPre, const int64_t *prePos, const double *preVal, WigOutput &out);

    // per-base alignment depth of reference positions first, first+step, ... <= last (device output, n = count)
    void depth(int ref, int64_t first, int64_t last, int64_t step, const std::vector<int> &targets, uint32_t flags,
               int32_t *dOut, float *kernelMs);

    // text of MAF rows (maf_kernels.cuh); every piece must lie inside its genome (checked here)
    void mafText(size_t nRows, const halgpu_maf_row *rows, size_t nPieces, const halgpu_maf_piece *pieces, const char *prefix,
                 size_t prefixBytes, size_t outBytes, char *out, float *kernelMs);

    // column runs of reference positions first..last; host (pinned) output owned by the caller
    // windowFirst (COL_UNIQUE only): where the ColumnIterator sweep this range belongs to started (-1: at `first`)
    void columnRuns(int ref, int64_t first, int64_t last, const std::vector<int> &targets, uint32_t flags, halgpu_col_runs &out,
                    int64_t windowFirst = -1);

    void release(void *deviceBuffer) { _cache.give(deviceBuffer); } // result buffers of liftover()
    DeviceCache &cache() { return _cache; }

    // genomes are staged on first use; these make one genome (its DNA / all of them) resident now
    void ensureGenome(int g);
    void ensureDna(int g);
    void ensureAll(bool dna);
    const uint8_t *deviceDna(int g) { ensureDna(g); return _g[(size_t)g].dna; }

  private:
    struct Stager;
    void ensureUpLinks(int g);
    void ensureDownLinks(int g, int slot);
    void linkFields(int64_t *links, int64_t linkStride, const int64_t *starts, int64_t startStride, int64_t n, const TopRec *landTop,
                    const int64_t *otherStarts, int64_t otherStride, int64_t *xlateOut);
    void buildGenomeTab(int ref, const std::vector<int> &targets, std::vector<GenomeTab> &tab);
    void buildBucket(const void *arr, bool isTop, int64_t N, int64_t len, uint32_t *&table, int &shift, int64_t &nb);
    const Plan &plan(int src, int tgt, int coal = -1);
    void *alloc(size_t bytes);

    std::unique_ptr<HalFile> _file;
    int _device;
    rt::Stream _stream, _copy, _copyBack, _aux; // _aux: sorts slice c+1 of a batch while the lane kernel runs slice c
    std::vector<GenomeDev> _g;
    std::map<std::tuple<int, int, int>, Plan> _plans; // (src, tgt, coalescence limit or -1)
    std::vector<void *> _owned;
    size_t _staged = 0;
    int _sms = 0;
    DeviceCache _cache;
    std::unique_ptr<Stager> _stager;
    unsigned long long *_hostCtr = nullptr; // pinned: the counters of one batch, read back with its single synchronisation
    std::unique_ptr<rt::Event> _ev[4], _sliceEv[5];
};

} // namespace halgpu

// the opaque context of include/halgpu.h (shared by capi.cu and multi.cu)
struct halgpu_ctx {
    std::unique_ptr<halgpu::Context> impl;
    std::vector<std::vector<halgpu_seq>> seqTables;
};
