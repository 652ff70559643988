// maf_kernels.cuh -- the text of MAF rows, written on the device.
//
// Replaces the per-character loop of MafBlock (maf/impl/halMafBlock.cpp:370-395: appendColumn pushes one base per row and
// column through DnaIterator::getBase, api/inc/halDnaIterator.h:131-138 = nibble unpack + case bit, complemented when the
// row runs on the reverse strand, api/inc/halCommon.h:45-67,187-196) and the row printer (:452-456, :499-519).  The block
// state machine above this (which columns share a block, which rows it has) is sequential and stays on the host
// (csrc/host/maf_export.cpp); what it hands down is, per row, where its bytes go in the output, its already formatted
// prefix ("a\n" for a block's first row, "s\t<name>\t<start>\t<length>\t<strand>\t<srcLength>\t") and its pieces:
// gap runs and runs of consecutive bases.  One warp per row; a lane per output character.
#pragma once
#include "device_index.cuh"

namespace halgpu {

struct MafTextParams {
    const halgpu_maf_row *rows;
    const halgpu_maf_piece *pieces;
    const char *prefix;          // all prefixes, back to back
    const uint8_t *const *dna;   // per genome: packed nibbles (NULL for genomes without rows)
    char *out;
    int64_t nRows;
};

// dnaUnpack + reverseComplement as two 16-entry tables in one 64-bit constant each would not fit ('?' entries): plain switch-free
// lookups from constant memory
__device__ __constant__ char HG_NIB[16] = {'a', 'c', 'g', 't', 'n', '?', '?', '?', 'A', 'C', 'G', 'T', 'N', '?', '?', '?'};
__device__ __constant__ char HG_NIBRC[16] = {'t', 'g', 'c', 'a', 'n', '?', '?', '?', 'T', 'G', 'C', 'A', 'N', '?', '?', '?'};

__global__ void __launch_bounds__(256) mafTextKernel(const MafTextParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < P.nRows; r += nWarps) {
        const halgpu_maf_row row = P.rows[r];
        char *o = P.out + row.out_offset;
        for (uint32_t i = (uint32_t)lane; i < row.prefix_len; i += 32) o[i] = P.prefix[row.prefix_offset + i];
        o += row.prefix_len;
        const uint8_t *d = P.dna[row.genome];
        for (uint32_t k = 0; k < row.num_pieces; ++k) {
            const halgpu_maf_piece pc = P.pieces[row.first_piece + k];
            const int64_t count = pc.count_kind >> 2;
            const int kind = (int)(pc.count_kind & 3);
            if (kind == 0) {
                for (int64_t i = lane; i < count; i += 32) o[i] = '-';
            } else if (kind == 1) {
                for (int64_t i = lane; i < count; i += 32) {
                    const int64_t p = pc.pos + i;
                    const uint8_t b = d[p >> 1];
                    o[i] = HG_NIB[(p & 1) ? (b & 0xF) : (b >> 4)];
                }
            } else {
                for (int64_t i = lane; i < count; i += 32) {
                    const int64_t p = pc.pos - i;
                    const uint8_t b = d[p >> 1];
                    o[i] = HG_NIBRC[(p & 1) ? (b & 0xF) : (b >> 4)];
                }
            }
            o += count;
        }
        if (lane < (int)row.tail_newlines) o[lane] = '\n';
    }
}

} // namespace halgpu
