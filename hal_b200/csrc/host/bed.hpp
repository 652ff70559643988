// BED text layer of the liftover CLI: parsing and printing with the reference's exact conventions
// (liftover/impl/halBedLine.cpp:27-151): TAB-only column split, bedType = min(#cols, 12) unless forced,
// fields of a previous, wider line persist into a narrower one (the reference reuses one BedLine object,
// liftover/inc/halBedScanner.h), extra columns carried through verbatim.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace halgpu {

struct BedBlock {
    int64_t start = 0, length = 0;
};

struct PslInfo { // liftover/inc/halBedLine.h:31-48
    uint64_t matches = 0, misMatches = 0, repMatches = 0, nCount = 0, qNumInsert = 0, qBaseInsert = 0, tNumInsert = 0, tBaseInsert = 0;
    std::string qSeqName;
    uint64_t qSeqSize = 0, qEnd = 0, qChromOffset = 0, tSeqSize = 0;
    char qStrand = '+';
    std::vector<int64_t> qBlockStarts; // absolute (genome) coordinates
};

struct BedLine {
    std::string chrName, name;
    int64_t start = -1, end = -1, score = 0, thickStart = 0, thickEnd = 0, itemR = 0, itemG = 0, itemB = 0;
    char strand = '+';
    int bedType = -1;
    std::vector<BedBlock> blocks;
    std::vector<std::string> extra;
    int64_t srcStart = -1; // genome-global source start of a lifted line (ordering key, not printed)
    char srcStrand = '+';

    // Parses one line into *this (fields beyond the line's width keep their previous values).
    // Throws std::runtime_error with the reference's messages on malformed input.
    void parse(const std::string &line, int forcedBedType);
    void append(std::string &out) const; // BedLine::write
    std::vector<PslInfo> psl;            // 0 or 1 element, as in the reference
    void expandToBed12();                // BedLine::expandToBed12 (halBedLine.cpp:153-182)
    bool validatePSL() const;            // BedLine::validatePSL (:251-334)
    void appendPSL(std::string &out, bool prefixWithName) const; // BedLine::writePSL (:206-249)
};

std::vector<std::string> chopString(const std::string &s, char sep); // hal::chopString (api/impl/halCommon.cpp:28-43)
int64_t strToInt(const std::string &s);                              // hal::strToInt (api/impl/halCommon.cpp:45-53)

} // namespace halgpu
