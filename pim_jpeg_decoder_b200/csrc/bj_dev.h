// Structures shared by the host side of the library and its kernels (plain data, lives in HBM).
#pragma once
#include <stdint.h>

namespace bj {

constexpr int kTileThreads = 192;     // K2/K3 CTA size = data units per tile (192 = lcm-friendly for bpm 1,2,3,4,6)
constexpr int kQPitch = 68;           // words between per-component dequant tables in smem (bank skew, 16 B aligned)

// Per image: geometry + where its data lives.  Data units ("du") are 64 coefficients = 128 B, zig-zag order,
// stored in decode order (MCU after MCU; inside an MCU: Y units row-major, then Cb, then Cr) - the order of
// the reference's loop nest src/jpeg_scanner.cpp:721-732.
struct ImgDev {
    uint32_t width, height;
    uint32_t nmx, nmy;            // MCUs per row / per column
    uint32_t du_base;             // index of this image's first data unit in the coefficient buffer
    uint32_t out_pitch;           // bytes between consecutive output rows
    uint64_t out_row0;            // byte offset (in the batch output buffer) of image row y = 0
    int32_t  row_dir;             // +1: row y at out_row0 + y*pitch (RGB8); -1: out_row0 - y*pitch (BMP, bottom-up)
    uint32_t row_bytes;           // payload bytes per row (width*3)
    uint32_t row_pad;             // zero bytes after the payload (BMP: width % 4)
    uint8_t  hs, vs, ncomp, bpm;  // luma sampling, components, data units per MCU
    uint8_t  bgr;                 // 1: B,G,R byte order (BMP)
    uint8_t  valid;
    uint8_t  dc_sep;              // 1: slot 0 of every unit is unused, the DC value comes from the DC plane (informative: the kernel is told by its dc_plane argument)
    uint8_t  pad_[1];
    uint32_t qslot;               // this image's quantiser set in the batch's pool of QTab (deduplicated across the batch)
    uint32_t tile0;               // index of this image's first K2/K3 tile in the batch's tile list
    uint32_t ref_blocks;          // BJ_OUT_REF_MCUS: 2x2-position blocks in this image's output (whole chunks of MAX_MCU_PER_DPU / 4)
};

// One set of quantisation tables as the K2 kernel stages it: per component (quantiser << 16) in ZIG-ZAG (file) order.
struct QTab {
    uint32_t q16[3][kQPitch];
};

// One CTA of the fused dequant/IDCT/colour kernel: `nm` consecutive MCUs of MCU-row `my`, starting at `mx0`.
// 16 bytes, one load; du0 / ndu let the CTA start fetching its coefficient units before it has seen the image record.
struct TileDev {
    uint32_t img;
    uint16_t my, mx0, nm;
    uint16_t ndu;                 // data units of the tile (nm * units per MCU), <= kTileThreads
    uint32_t du0;                 // index of the tile's first data unit in the coefficient buffer / DC plane
};
static_assert(sizeof(TileDev) == 16, "TileDev is loaded as one uint4");

}  // namespace bj
