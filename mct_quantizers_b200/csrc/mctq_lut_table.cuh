// mctq_lut_table.cuh -- layout of the LUT search table blob shared by mctq_lut.cu (generic kernels, host builder) and
// mctq_lutp.cu (prepared per-channel decision tables).
#pragma once
#include <stdint.h>

namespace mctq {

// LUT search table as it lives in global memory (built on the host by mctq_lut_build_table)
struct LutTableHeader {
    uint32_t magic;        // 'MQLT'
    int32_t K;             // original number of centroids
    int32_t Ks;            // sorted unique centroids
    int32_t P;             // power of two >= Ks: the search walks P - 1 padded thresholds
    int32_t levels;        // log2(P)
    int32_t pos_of_idx0;   // sorted position of original index 0 (NaN inputs select index 0)
    int32_t bw, is_signed;
    float mult;            // 2^(bw - signed)
    float reserved[7];
};
// followed by: float tau[P - 1] (thresholds in the q = x / d domain, +inf padded),
//              float cq[P]      (lut_sorted / mult), uint8_t orig[P] (original index), padded to 16 bytes
constexpr uint32_t kLutMagic = 0x4d514c54u;
constexpr int kLutMaxK = 256;


inline int lut_geometry_from_K(int K, int* P, int* levels) {
    if (K < 1 || K > kLutMaxK) return MCTQ_E_LUT;
    int p = 1, l = 0;
    while (p < K) { p <<= 1; ++l; }
    *P = p;
    *levels = l;
    return 0;
}

}  // namespace mctq
