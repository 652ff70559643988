#include "bed.hpp"
#include <algorithm>
#include <charconv>
#include <stdexcept>

namespace halgpu {

std::vector<std::string> chopString(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t start = 0, end;
    while ((end = s.find(sep, start)) != std::string::npos) {
        out.emplace_back(s, start, end - start);
        start = end + 1;
    }
    if (start < s.size()) {
        out.emplace_back(s, start);
    }
    return out;
}

int64_t strToInt(const std::string &s) {
    // stringstream >> int64 semantics: optional leading blanks and sign, then digits; trailing text ignored
    size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == '\f' || s[i] == '\v')) ++i;
    size_t b = i;
    if (i < s.size() && s[i] == '+') { ++i; b = i; }
    int64_t v = 0;
    auto r = std::from_chars(s.data() + b, s.data() + s.size(), v);
    if (r.ec != std::errc() || r.ptr == s.data() + b) {
        throw std::runtime_error("Error converting string to int: " + s);
    }
    return v;
}

void BedLine::parse(const std::string &line, int forcedBedType) {
    bedType = forcedBedType;
    const std::vector<std::string> row = chopString(line, '\t');
    if (row.size() < 3) {
        throw std::runtime_error("Expected at least three columns in BED record: " + line);
    }
    if (bedType == 0) {
        bedType = std::min<int>((int)row.size(), 12);
    }
    chrName = row[0];
    start = strToInt(row[1]);
    end = strToInt(row[2]);
    if (start >= end) {
        throw std::runtime_error("Error zero or negative length BED range: " + line);
    }
    if (bedType > 3) name = row.at(3);
    if (bedType > 4) score = strToInt(row.at(4));
    if (bedType > 5) {
        strand = row.at(5).empty() ? '\0' : row[5][0];
        if (strand != '.' && strand != '+' && strand != '-') {
            throw std::runtime_error("Strand character must be + or - or ." + line);
        }
    }
    if (bedType > 6) thickStart = strToInt(row.at(6));
    if (bedType > 7) thickEnd = strToInt(row.at(7));
    if (bedType > 8) {
        const std::vector<std::string> rgb = chopString(row.at(8), ',');
        if (rgb.size() > 3 || rgb.empty()) {
            throw std::runtime_error("Error parsing BED itemRGB: " + line);
        }
        itemR = strToInt(rgb[0]);
        itemG = itemB = itemR;
        if (rgb.size() > 1) itemG = strToInt(rgb[1]);
        if (rgb.size() == 3) itemB = strToInt(rgb[2]);
    }
    if (bedType > 9) {
        if (bedType < 12) {
            throw std::runtime_error("Error parsing BED, insufficient columns for blocks: " + line);
        }
        const size_t numBlocks = (size_t)strToInt(row.at(9));
        const std::vector<std::string> sizes = chopString(row.at(10), ',');
        if (sizes.size() != numBlocks) {
            throw std::runtime_error("Error parsing BED blockSizes: " + line);
        }
        const std::vector<std::string> starts = chopString(row.at(11), ',');
        if (starts.size() != numBlocks) {
            throw std::runtime_error("Error parsing BED blockStarts: " + line);
        }
        blocks.resize(numBlocks);
        for (size_t i = 0; i < numBlocks; ++i) {
            blocks[i].length = strToInt(sizes[i]);
            blocks[i].start = strToInt(starts[i]);
            if (start + blocks[i].start + blocks[i].length > end) {
                throw std::runtime_error("Error BED block out of range: " + line);
            }
        }
    }
    extra.clear();
    for (size_t i = (size_t)bedType; i < row.size(); ++i) {
        extra.push_back(row[i]);
    }
}

namespace {
inline void putInt(std::string &out, int64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    out.append(buf, r.ptr);
}
} // namespace

void BedLine::append(std::string &out) const {
    out += chrName; out += '\t'; putInt(out, start); out += '\t'; putInt(out, end);
    if (bedType > 3) { out += '\t'; out += name; }
    if (bedType > 4) { out += '\t'; putInt(out, score); }
    if (bedType > 5) { out += '\t'; out += strand; }
    if (bedType > 6) { out += '\t'; putInt(out, thickStart); }
    if (bedType > 7) { out += '\t'; putInt(out, thickEnd); }
    if (bedType > 8) { out += '\t'; putInt(out, itemR); out += ','; putInt(out, itemG); out += ','; putInt(out, itemB); }
    if (bedType > 9) {
        out += '\t'; putInt(out, (int64_t)blocks.size());
        for (size_t i = 0; i < blocks.size(); ++i) { out += i == 0 ? '\t' : ','; putInt(out, blocks[i].length); }
        for (size_t i = 0; i < blocks.size(); ++i) { out += i == 0 ? '\t' : ','; putInt(out, blocks[i].start); }
    }
    for (const std::string &e : extra) { out += '\t'; out += e; }
    out += '\n';
}

void BedLine::expandToBed12() {
    if (bedType <= 3) name = "";
    if (bedType <= 4) score = 0;
    if (bedType <= 5) strand = '+';
    if (bedType <= 6) thickStart = start;
    if (bedType <= 7) thickEnd = end;
    if (bedType <= 8) itemR = itemG = itemB = 0;
    if (bedType <= 9) {
        blocks.resize(1);
        blocks[0].start = 0;
        blocks[0].length = end - start;
    }
    bedType = 12;
}

bool BedLine::validatePSL() const {
    if (psl.size() != 1 || blocks.empty()) return false;
    const PslInfo &p = psl[0];
    if (blocks.size() != p.qBlockStarts.size()) return false;
    uint64_t tot = 0;
    for (const BedBlock &b : blocks) tot += (uint64_t)b.length;
    if (tot != p.matches + p.misMatches + p.repMatches + p.nCount) return false;
    if (tot + p.qBaseInsert != p.qEnd - (uint64_t)srcStart) return false;
    if (tot + p.tBaseInsert != (uint64_t)(end - start)) return false;
    if (strand != '-') {
        if (blocks[0].start != 0 || blocks.back().start + blocks.back().length + start != end) return false;
    } else {
        if (blocks.back().start != 0 || blocks[0].start + blocks[0].length + start != end) return false;
    }
    if (p.qStrand != '-') {
        if (p.qBlockStarts[0] != srcStart || (uint64_t)(p.qBlockStarts.back() + blocks.back().length) != p.qEnd) return false;
    } else {
        if (p.qBlockStarts.back() != srcStart || (uint64_t)(p.qBlockStarts[0] + blocks[0].length) != p.qEnd) return false;
    }
    return true;
}

void BedLine::appendPSL(std::string &out, bool prefixWithName) const {
    if (!validatePSL()) throw std::runtime_error("Internal error: PSL does not validate");
    const PslInfo &p = psl[0];
    auto num = [&](uint64_t v) { char b[24]; auto r = std::to_chars(b, b + sizeof b, v); out.append(b, r.ptr); };
    auto snum = [&](int64_t v) { putInt(out, v); };
    if (prefixWithName) { out += name; out += '\t'; }
    num(p.matches); out += '\t'; num(p.misMatches); out += '\t'; num(p.repMatches); out += '\t'; num(p.nCount); out += '\t';
    num(p.qNumInsert); out += '\t'; num(p.qBaseInsert); out += '\t'; num(p.tNumInsert); out += '\t'; num(p.tBaseInsert); out += '\t';
    out += p.qStrand; out += strand; out += '\t';
    out += p.qSeqName; out += '\t'; num(p.qSeqSize); out += '\t'; snum(srcStart - (int64_t)p.qChromOffset); out += '\t';
    num(p.qEnd - p.qChromOffset); out += '\t'; out += chrName; out += '\t'; num(p.tSeqSize); out += '\t'; snum(start); out += '\t';
    snum(end); out += '\t'; num(blocks.size()); out += '\t';
    for (const BedBlock &b : blocks) { snum(b.length); out += ','; }
    out += '\t';
    for (size_t i = 0; i < p.qBlockStarts.size(); ++i) {
        int64_t st = p.qBlockStarts[i] - (int64_t)p.qChromOffset;
        if (p.qStrand == '-') st = (int64_t)p.qSeqSize - st - blocks[i].length;
        snum(st); out += ',';
    }
    out += '\t';
    for (const BedBlock &b : blocks) {
        int64_t st = b.start + start;
        if (strand == '-') st = (int64_t)p.tSeqSize - st - b.length;
        snum(st); out += ',';
    }
    out += '\n';
}

} // namespace halgpu
