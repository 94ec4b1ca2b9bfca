// Device-side arithmetic of the reference's DPU program, restated for sm_100a registers.
//
//   dequantise  src/decoder_dpu.c:158-177   c = (short)(c * QT)
//   IDCT        src/decoder_dpu.c:210-321   8 row passes, results stored as short, 8 column passes, stored as short
//   colour      src/decoder_dpu.c:361-390   nearest-neighbour chroma, 22-bit fixed point, clamp
//
// Representation trick (exact, see DESIGN.md "K2 arithmetic"): a 16-bit value v that the reference keeps in a
// `short` is carried as X = v << 16 in a 32-bit register.  Then
//   * the (short) wrap is free: (c * (q << 16)) mod 2^32 == ((c*q) mod 2^16) << 16, sign included;
//   * (v * K) >> s  ==  __mulhi(X, K << (16 - s))   (floor division of the exact 64-bit product by 2^32),
// so dequantise+wrap is one IMAD and every first-stage multiply of a 1-D pass is one IMAD.HI.
#pragma once
#include <stdint.h>

namespace bj {

// Standard zig-zag (index -> natural position).  The reference's table differs only at index 48 (38 instead of
// 58, src/headers/common.h:16): natural 58 is never written, natural 38 gets index 48 and then index 52.
// Closed form used below:  nat[38] = zz[52] != 0 ? zz[52] : zz[48];  nat[58] = 0.
__host__ __device__ constexpr int zz2nat(int k) {
    constexpr int T[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return T[k];
}

// One 1-D pass over X[base + k*STRIDE], k = 0..7 (src/decoder_dpu.c:219-267 == :271-319).
// Inputs are in the X = v<<16 form.  MID: outputs are re-wrapped to 16 bits and left in X form for the next
// pass (the reference stores shorts between the passes, :260-267).  !MID: outputs are the raw (sum >> 4) ints;
// the caller keeps their low 16 bits (the final short store, :312-319).
template <int STRIDE, bool MID>
__device__ __forceinline__ void idct_pass(int (&X)[64], const int base) {
    const int g0 = __mulhi(X[base + 0 * STRIDE], 181 << 11);   // (v0*181)>>5
    const int g1 = __mulhi(X[base + 4 * STRIDE], 181 << 11);   // (v4*181)>>5
    const int g2 = __mulhi(X[base + 2 * STRIDE], 59 << 13);    // (v2*59)>>3
    const int g3 = __mulhi(X[base + 6 * STRIDE], 49 << 12);    // (v6*49)>>4
    const int g4 = __mulhi(X[base + 5 * STRIDE], 71 << 12);    // (v5*71)>>4
    const int g5 = __mulhi(X[base + 1 * STRIDE], 251 << 11);   // (v1*251)>>5
    const int g6 = __mulhi(X[base + 7 * STRIDE], 25 << 12);    // (v7*25)>>4
    const int g7 = __mulhi(X[base + 3 * STRIDE], 213 << 11);   // (v3*213)>>5

    const int f4 = g4 - g7, f5 = g5 + g6, f6 = g5 - g6, f7 = g4 + g7;
    const int e2 = g2 - g3, e3 = g2 + g3, e5 = f5 - f7, e7 = f5 + f7, e8 = f4 + f6;
    const int d2 = (e2 * 181) >> 7, d4 = (f4 * 277) >> 8, d5 = (e5 * 181) >> 7, d6 = (f6 * 669) >> 8, d8 = (e8 * 49) >> 6;
    const int c0 = g0 + g1, c1 = g0 - g1, c2 = d2 - e3, c4 = d4 + d8, c5 = d5 + e7, c6 = d6 - d8, c8 = c5 - c6;
    const int b0 = c0 + e3, b1 = c1 + c2, b2 = c1 - c2, b3 = c0 - e3, b4 = c4 - c8, b6 = c6 - e7;

    int o[8];
    o[0] = (b0 + e7) >> 4; o[1] = (b1 + b6) >> 4; o[2] = (b2 + c8) >> 4; o[3] = (b3 + b4) >> 4;
    o[4] = (b3 - b4) >> 4; o[5] = (b2 - c8) >> 4; o[6] = (b1 - b6) >> 4; o[7] = (b0 - e7) >> 4;
#pragma unroll
    for (int k = 0; k < 8; k++) X[base + k * STRIDE] = MID ? (int)((unsigned)o[k] << 16) : o[k];
}

// Full 8x8: X in natural order, X form in; raw column-pass outputs out (low 16 bits = the reference's shorts).
__device__ __forceinline__ void idct8x8(int (&X)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++) idct_pass<1, true>(X, r * 8);
#pragma unroll
    for (int c = 0; c < 8; c++) idct_pass<8, false>(X, c);
}

// Chroma contributions of one (cb, cr) sample, src/decoder_dpu.c:376-378.  cb/cr are the sign-extended shorts;
// the 32-bit products wrap exactly like the reference's `int` multiplies.
struct ChromaTerms { int r, g, b; };
__device__ __forceinline__ ChromaTerms chroma_terms(int cb, int cr) {
    ChromaTerms t;
    t.r = ((int)(5880414u * (unsigned)cr) >> 22) + 128;
    t.g = 128 - ((int)(1442840u * (unsigned)cb) >> 22) - ((int)(2994733u * (unsigned)cr) >> 22);
    t.b = ((int)(7432306u * (unsigned)cb) >> 22) + 128;
    return t;
}

// d = [sat8(b0), sat8(b1), sat8(b2), sat8(b3)] little-endian, unsigned saturation = the reference's clamp to
// [0,255] (src/decoder_dpu.c:380-382).  Two I2IP instructions.
__device__ __forceinline__ unsigned pack_sat4(int b0, int b1, int b2, int b3) {
    unsigned hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(b3), "r"(b2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b1), "r"(b0), "r"(hi));
    return d;
}

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ int sext_lo(unsigned w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int sext_hi(unsigned w) { return (int)w >> 16; }

}  // namespace bj
