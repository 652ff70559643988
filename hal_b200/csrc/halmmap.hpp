// HAL-MMAP file reader (host side of the staging step).
//
// Parses the reference's raw-struct file format straight from the mapping -- header, alignment root,
// genome array, sequence tables, newick tree -- and hands out bounds-checked pointers to the top /
// bottom segment arrays and packed DNA so they can be copied to the GPU.  The name->index perfect
// hash (mmapPerfectHashTable.*) and the site->sequence BST (mmapGenomeSiteMap.cpp:99) are not needed:
// genomes are resolved by scanning the (small) genome array and sequences are stored in start order.
//
// Format references (reference repo): api/mmap_impl/mmapFile.h:23-31 (header), mmapFile.cpp:76-102
// (validation rules kept: format string, major version 1, dirty flag), mmapAlignment.h:14-31,
// mmapGenome.h:19-46, mmapSequenceData.h:21-30, mmapTopSegmentData.h:40-44,
// mmapBottomSegmentData.h:35-52, mmapArray.h:6-11.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace halgpu {

struct HalError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct SequenceInfo {
    std::string name;
    int64_t start = 0, length = 0;
    int64_t topFirst = 0, bottomFirst = 0, numTop = 0, numBottom = 0;
};

struct GenomeInfo {
    std::string name;
    int parent = -1;
    int slotInParent = -1;
    std::vector<int> children; // newick order == child slot order (mmapAlignment.h:145-153)
    int64_t length = 0, numTop = 0, numBottom = 0;
    size_t bottomStride = 16; // 8*(2+nc) + roundup8(nc), mmapBottomSegmentData.h:44-49
    const uint8_t *top = nullptr;    // (numTop+1) x 40 B
    const uint8_t *bottom = nullptr; // (numBottom+1) x bottomStride
    const uint8_t *dna = nullptr;    // (length+1)/2 bytes, even index = high nibble
    std::vector<SequenceInfo> sequences;
    std::map<std::string, int> sequenceByName;
    std::map<std::string, std::string> metadata; // Genome::getMetaData()->getMap() (api/mmap_impl/mmapMetaData.h)
};

class HalFile {
  public:
    explicit HalFile(const std::string &path);
    ~HalFile();
    HalFile(const HalFile &) = delete;
    HalFile &operator=(const HalFile &) = delete;

    const std::string &path() const { return _path; }
    const std::string &newick() const { return _newick; }
    const std::string &version() const { return _version; }
    int root() const { return _root; }
    const std::vector<GenomeInfo> &genomes() const { return _genomes; }
    int genomeId(const std::string &name) const;
    // file offset of a pointer handed out by this object (GenomeInfo::top / bottom / dna)
    uint64_t offsetOf(const void *p) const { return (uint64_t)(static_cast<const uint8_t *>(p) - static_cast<const uint8_t *>(_map)); }
    // lowest common ancestor (api/impl/halCommon.cpp:123-152)
    int mrca(int a, int b) const;

  private:
    const uint8_t *at(uint64_t off, uint64_t len, const char *what) const;
    uint64_t u64(uint64_t off) const;
    std::string _path, _newick, _version;
    void *_map = nullptr;
    size_t _size = 0;
    int _root = -1;
    std::vector<GenomeInfo> _genomes;
};

} // namespace halgpu
