/* TEST INFRASTRUCTURE ONLY -- CPU oracle: halWiggleLiftover restated (see wiggle.cpp). */
#ifndef ORACLE_WIGGLE_H
#define ORACLE_WIGGLE_H
#include "liftover.h"
#include <string>

namespace oracle {

/* Output text of `halWiggleLiftover hal src in.wig tgt out.wig [--noDupes] [--append]`; preloadText = the existing
 * out.wig for --append (else NULL).  correctPath = true maps along the src -> MRCA -> tgt path even where the reference takes its
 * wrong turn at the MRCA (see wiggle.cpp) -- what the GPU build does there.  Throws std::runtime_error with the reference's message on the inputs it rejects. */
std::string wiggleLiftover(const HalView &v, int src, int tgt, bool dupes, const std::string &inText, const std::string *preloadText,
                           bool correctPath = false);

} // namespace oracle
#endif
