/* TEST INFRASTRUCTURE ONLY (oracle build shim).  libhdf5 is not installed in this image, so the
 * oracle is an mmap-only build of the reference: this stub satisfies the two non-hdf5 translation
 * units that name Hdf5Alignment (api/impl/halCLParser.cpp:8,23 and the instance factory). */
#ifndef ORACLE_SHIM_HDF5ALIGNMENT_H
#define ORACLE_SHIM_HDF5ALIGNMENT_H
#include <string>
namespace hal {
    class CLParser;
    struct Hdf5Alignment {
        static bool isHdf5File(const std::string &) { return false; }
        static void defineOptions(CLParser *, unsigned) {}
    };
}
#endif
