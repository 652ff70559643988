/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 * mmap-only replacement for api/impl/halAlignmentInstance.cpp (which needs libhdf5): same
 * public entry points (api/inc/halAlignmentInstance.h:96-111), HDF5 branch reports an error. */
#include "halAlignmentInstance.h"
#include "halCLParser.h"
#include "halCommon.h"
#include "hdf5Alignment.h"
#include "mmapAlignment.h"
#include <fstream>

using namespace hal;

const std::string hal::STORAGE_FORMAT_HDF5 = "hdf5";
const std::string hal::STORAGE_FORMAT_MMAP = "mmap";

Alignment *hal::mmapAlignmentInstance(const std::string &path, unsigned mode, size_t fileSize) {
    return new MMapAlignment(path, mode, fileSize);
}

const std::string &hal::detectHalAlignmentFormat(const std::string &path, const CLParser *) {
    static const std::string none;
    std::ifstream fh(path.c_str());
    if (!fh) {
        throw hal_errno_exception(path, "can't open HAL file", errno);
    }
    char buf[64];
    fh.read(buf, sizeof buf);
    std::string head(buf, 0, fh.gcount());
    return MMapFile::isMmapFile(head) ? STORAGE_FORMAT_MMAP : none;
}

AlignmentPtr hal::openHalAlignment(const std::string &path, const CLParser *options, unsigned mode,
                                   const std::string &overrideFormat) {
    std::string fmt = overrideFormat;
    if (fmt.empty()) {
        if ((mode & CREATE_ACCESS) == 0) {
            fmt = detectHalAlignmentFormat(path, options);
            if (fmt.empty()) {
                throw hal_exception("unable to determine HAL storage format of " + path);
            }
        } else if (options != NULL) {
            fmt = options->getOption<const std::string &>("format");
        } else {
            fmt = STORAGE_FORMAT_MMAP;
        }
    }
    if (fmt != STORAGE_FORMAT_MMAP) {
        throw hal_exception("oracle build is mmap-only; format '" + fmt + "' unavailable (use --format mmap)");
    }
    if (options == NULL) {
        return AlignmentPtr(new MMapAlignment(path, mode));
    }
    return AlignmentPtr(new MMapAlignment(path, mode, options));
}
