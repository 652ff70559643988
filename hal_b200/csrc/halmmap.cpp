#include "halmmap.hpp"
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace halgpu {

namespace {
constexpr uint64_t HEADER_BYTES = 376, ALIGNMENT_BYTES = 312, GENOME_BYTES = 96, SEQUENCE_BYTES = 328,
                   ARRAY_HEADER_BYTES = 24, TOP_BYTES = 40;

struct Newick {
    std::string name;
    std::vector<Newick> kids;
};

Newick parseNewick(const std::string &s, size_t &i) {
    Newick n;
    if (i < s.size() && s[i] == '(') {
        ++i;
        while (true) {
            n.kids.push_back(parseNewick(s, i));
            if (i < s.size() && s[i] == ',') { ++i; continue; }
            if (i < s.size() && s[i] == ')') { ++i; break; }
            throw HalError("malformed newick tree in HAL file");
        }
    }
    size_t b = i;
    while (i < s.size() && s[i] != ':' && s[i] != ',' && s[i] != ')' && s[i] != ';') ++i;
    n.name = s.substr(b, i - b);
    if (i < s.size() && s[i] == ':') {
        while (i < s.size() && s[i] != ',' && s[i] != ')' && s[i] != ';') ++i;
    }
    return n;
}
} // namespace

// record count x record size of an array named by a (possibly corrupt) file: a product that wraps around must not slip
// through the bounds check of at()
static uint64_t arrayBytes(uint64_t count, uint64_t each, const std::string &path, const char *what) {
    if (each != 0 && count > UINT64_MAX / each) {
        throw HalError(path + ": " + what + " out of file bounds, probably file corruption");
    }
    return count * each;
}

const uint8_t *HalFile::at(uint64_t off, uint64_t len, const char *what) const {
    if (off > _size || len > _size - off) {
        throw HalError(_path + ": " + what + " out of file bounds, probably file corruption");
    }
    return static_cast<const uint8_t *>(_map) + off;
}

uint64_t HalFile::u64(uint64_t off) const {
    uint64_t v;
    std::memcpy(&v, at(off, 8, "field"), 8);
    return v;
}

HalFile::HalFile(const std::string &path) : _path(path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) {
        throw HalError(path + ": can't open HAL file: " + std::strerror(errno));
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        ::close(fd);
        throw HalError(path + ": fstat failed");
    }
    _size = static_cast<size_t>(st.st_size);
    if (_size < HEADER_BYTES) {
        ::close(fd);
        throw HalError(path + ": file size of " + std::to_string(_size) + " is less that header size of " +
                       std::to_string(HEADER_BYTES));
    }
    _map = mmap(nullptr, _size, PROT_READ, MAP_SHARED, fd, 0);
    ::close(fd);
    if (_map == MAP_FAILED) {
        _map = nullptr;
        throw HalError(path + ": mmap failed: " + std::strerror(errno));
    }
    try { // a constructor that throws never runs its destructor: give the mapping back here
    const char *hdr = reinterpret_cast<const char *>(_map);
    if (std::string(hdr, strnlen(hdr, 32)) != "HAL-MMAP") {
        throw HalError(path + ": invalid file header, expected format name of 'HAL-MMAP'");
    }
    _version.assign(hdr + 32, strnlen(hdr + 32, 32));
    if (std::atoi(_version.c_str()) != 1) {
        throw HalError(path + ": incompatible mmap major versions: file version " + _version + ", mmap API version 1.1");
    }
    const uint64_t rootOff = u64(104);
    if (hdr[112] != 0) {
        throw HalError(path + ": file is marked as dirty, most likely an inconsistent state.");
    }
    at(rootOff, ALIGNMENT_BYTES, "alignment root");
    const uint64_t numGenomes = u64(rootOff), nwOff = u64(rootOff + 8), nwLen = u64(rootOff + 16),
                   gaOff = u64(rootOff + 24);
    if (numGenomes == 0) {
        throw HalError("hal alignment is empty");
    }
    if (nwLen == 0) {
        throw HalError("hal alignment has no tree");
    }
    const char *nw = reinterpret_cast<const char *>(at(nwOff, nwLen, "newick string"));
    _newick.assign(nw, strnlen(nw, nwLen));
    at(gaOff, arrayBytes(numGenomes, GENOME_BYTES, _path, "genome array"), "genome array");
    _genomes.resize(numGenomes);
    for (uint64_t g = 0; g < numGenomes; ++g) {
        const uint64_t b = gaOff + g * GENOME_BYTES;
        GenomeInfo &gi = _genomes[g];
        gi.length = static_cast<int64_t>(u64(b));
        const uint64_t nseq = u64(b + 8);
        gi.numTop = static_cast<int64_t>(u64(b + 16));
        gi.numBottom = static_cast<int64_t>(u64(b + 24));
        const uint64_t nameOff = u64(b + 32), seqOff = u64(b + 56);
        const uint64_t nameLen = u64(nameOff + 16); // MMapArrayData::_length, includes the NUL
        const char *nm = reinterpret_cast<const char *>(at(nameOff + ARRAY_HEADER_BYTES, nameLen, "genome name"));
        gi.name.assign(nm, strnlen(nm, nameLen));
        at(seqOff, arrayBytes(nseq, SEQUENCE_BYTES, _path, "sequence array"), "sequence array");
        gi.sequences.resize(nseq);
        int64_t expectStart = 0;
        for (uint64_t s = 0; s < nseq; ++s) {
            const uint64_t sb = seqOff + s * SEQUENCE_BYTES;
            SequenceInfo &si = gi.sequences[s];
            si.start = static_cast<int64_t>(u64(sb));
            si.length = static_cast<int64_t>(u64(sb + 16));
            si.topFirst = static_cast<int64_t>(u64(sb + 24));
            si.bottomFirst = static_cast<int64_t>(u64(sb + 32));
            si.numTop = static_cast<int64_t>(u64(sb + 40));
            si.numBottom = static_cast<int64_t>(u64(sb + 48));
            const uint64_t nl = u64(sb + 56), no = u64(sb + 64);
            const char *sn = reinterpret_cast<const char *>(at(no, nl, "sequence name"));
            si.name.assign(sn, strnlen(sn, nl));
            if (si.start != expectStart) {
                throw HalError(path + ": sequences of genome " + gi.name + " are not stored in coordinate order");
            }
            expectStart += si.length;
            gi.sequenceByName[si.name] = static_cast<int>(s);
        }
        if (expectStart != gi.length) {
            throw HalError("Problem: genome has length " + std::to_string(gi.length) + ", however sequences total " +
                           std::to_string(expectStart));
        }
        // metadata: MMapMetaDataData {keysOffset, valuesOffset} -> two MMapArray<size_t> of offsets to MMapStrings
        // (api/mmap_impl/mmapMetaData.h:10-16,66-76)
        const uint64_t metaOff = u64(b + 64);
        if (metaOff != 0) {
            const uint64_t keysOff = u64(metaOff), valsOff = u64(metaOff + 8);
            const uint64_t nk = u64(keysOff + 16), nv = u64(valsOff + 16);
            if (nk != nv) throw HalError("# keys != # values in metadata");
            auto str = [&](uint64_t off) {
                const uint64_t len = u64(off + 16);
                const char *p = reinterpret_cast<const char *>(at(off + ARRAY_HEADER_BYTES, len, "metadata string"));
                return std::string(p, strnlen(p, len));
            };
            for (uint64_t i = 0; i < nk; ++i) {
                gi.metadata.insert(std::make_pair(str(u64(keysOff + ARRAY_HEADER_BYTES + 8 * i)), str(u64(valsOff + ARRAY_HEADER_BYTES + 8 * i))));
            }
        }
    }
    // tree
    size_t pos = 0;
    Newick tree = parseNewick(_newick, pos);
    std::function<void(const Newick &, int, int)> link = [&](const Newick &n, int parent, int slot) {
        int id = genomeId(n.name);
        if (id < 0) {
            throw HalError("genome " + n.name + " not found in alignment.");
        }
        _genomes[id].parent = parent;
        _genomes[id].slotInParent = slot;
        if (parent < 0) {
            _root = id;
        }
        for (size_t k = 0; k < n.kids.size(); ++k) {
            int cid = genomeId(n.kids[k].name);
            _genomes[id].children.push_back(cid);
            link(n.kids[k], id, static_cast<int>(k));
        }
    };
    link(tree, -1, -1);
    // segment arrays (their stride needs the child count)
    for (uint64_t g = 0; g < numGenomes; ++g) {
        const uint64_t b = gaOff + g * GENOME_BYTES;
        GenomeInfo &gi = _genomes[g];
        const uint64_t nc = gi.children.size();
        gi.bottomStride = 8 * (2 + nc) + ((nc + 7) / 8) * 8;
        gi.dna = at(u64(b + 72), static_cast<uint64_t>((gi.length + 1) / 2), "dna array");
        gi.top = at(u64(b + 80), arrayBytes(static_cast<uint64_t>(gi.numTop) + 1, TOP_BYTES, _path, "top segment array"), "top segment array");
        gi.bottom = at(u64(b + 88), arrayBytes(static_cast<uint64_t>(gi.numBottom) + 1, gi.bottomStride, _path, "bottom segment array"), "bottom segment array");
    }
    } catch (...) {
        munmap(_map, _size);
        _map = nullptr;
        throw;
    }
}

HalFile::~HalFile() {
    if (_map != nullptr) {
        munmap(_map, _size);
    }
}

int HalFile::genomeId(const std::string &name) const {
    for (size_t i = 0; i < _genomes.size(); ++i) {
        if (_genomes[i].name == name) {
            return static_cast<int>(i);
        }
    }
    return -1;
}

int HalFile::mrca(int a, int b) const {
    std::vector<int> pa, pb;
    for (int g = a; g >= 0; g = _genomes[g].parent) pa.push_back(g);
    for (int g = b; g >= 0; g = _genomes[g].parent) pb.push_back(g);
    int r = -1;
    while (!pa.empty() && !pb.empty() && pa.back() == pb.back()) {
        r = pa.back();
        pa.pop_back();
        pb.pop_back();
    }
    return r;
}

} // namespace halgpu
