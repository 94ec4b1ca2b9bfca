// b200jpeg_ref.h - glue for a host that keeps the reference's OWN scanner and BMP writer and replaces only what lies
// between them: decode_Huffman_data (src/jpeg_scanner.cpp:707-756) + the DPU round trip (src/decoder_host.cpp:268-312).
// Compiled against the reference's headers (-I <reference>/src); nothing of the reference is copied - its `Header`
// type and its `zigzag_map` are used through its own header files.
//
//   Header *h = read_JPEG(path);                                   // unchanged (src/jpeg_scanner.cpp:345)
//   bj_image_desc d;  bj_desc_from_header(*h, &d);                 // field for field
//   bj_decode_batch_desc(ctx, &d, &scan, &len, &kind, 1, BJ_OUT_REF_MCUS, &mcus, &status);
//   write_BMP(metadata, chunks, 0, name);                          // unchanged (src/bmp_writer.cpp:19)
#ifndef B200JPEG_REF_H
#define B200JPEG_REF_H
#include <cstdint>
#include <cstring>
#include <vector>

#include "b200jpeg.h"
#include "headers/jpeg.h"       // (brings headers/common.h with zigzag_map; that header has no include guard)

// `Header` (src/headers/jpeg.h:146-179) -> bj_image_desc.  The reference keeps its quantisation tables de-zigzagged
// through its own zigzag_map (src/jpeg_scanner.cpp:306,311), the descriptor wants them in file order: entry i of the
// file is table[zigzag_map[i]].  (The map sends both 48 and 52 to 38 - src/headers/common.h:16 - so the file's entry 48
// cannot be recovered; it is never used: the coefficient at index 48 is multiplied by the entry read at index 52, like
// in the reference.)
inline void bj_desc_from_header(const Header &h, bj_image_desc *d) {
    std::memset(d, 0, sizeof(*d));
    d->width = h.width; d->height = h.height;
    d->mcu_w = h.mcu_width; d->mcu_h = h.mcu_height;
    d->mcu_w_real = h.mcu_width_real; d->mcu_h_real = h.mcu_height_real;
    d->restart_interval = h.restart_interval;
    d->ncomp = h.num_components;
    d->hs = h.h_sampling_factor; d->vs = h.v_sampling_factor;
    d->frame_type = h.frame_type;
    d->scan_ncomp = h.components_in_scan;
    for (int j = 0; j < 3; j++) {
        const ColorComponent &c = h.color_components[j];
        d->comp_h[j] = c.h_sampling_factor; d->comp_v[j] = c.v_sampling_factor;
        d->qt_id[j] = c.QT_ID; d->dc_id[j] = c.DHT_ID; d->ac_id[j] = c.AHT_ID;
    }
    for (int t = 0; t < 4; t++) {
        d->qt_set[t] = h.quantization_tables[t].set;
        for (int i = 0; i < 64; i++) d->qt_zz[t][i] = (uint16_t)h.quantization_tables[t].table[zigzag_map[i]];
        d->dc_set[t] = h.huffman_DC_tables[t].set;
        d->ac_set[t] = h.huffman_AC_tables[t].set;
        std::memcpy(d->dc_offsets[t], h.huffman_DC_tables[t].offsets, 17);
        std::memcpy(d->dc_symbols[t], h.huffman_DC_tables[t].symbols, 162);
        std::memcpy(d->ac_offsets[t], h.huffman_AC_tables[t].offsets, 17);
        std::memcpy(d->ac_symbols[t], h.huffman_AC_tables[t].symbols, 162);
    }
}

// The 276-word metadata record of one image, as mcu_prepare fills it (src/decoder_host.cpp:156-178): what write_BMP
// and the DPU program read.
inline std::vector<uint32_t> bj_metadata_from_header(const Header &h, int max_mcu_per_dpu) {
    std::vector<uint32_t> md(20 + 4 * 64, 0);
    md[0] = h.mcu_height; md[1] = h.mcu_width; md[2] = h.mcu_height_real; md[3] = h.mcu_width_real;
    md[4] = h.num_components; md[5] = h.v_sampling_factor; md[6] = h.h_sampling_factor;
    for (unsigned j = 0; j < h.num_components; j++) {
        md[7 + j] = h.color_components[j].QT_ID;
        md[7 + h.num_components + j] = h.color_components[j].h_sampling_factor;
        md[7 + 2 * h.num_components + j] = h.color_components[j].v_sampling_factor;
    }
    md[17] = h.height; md[18] = h.width; md[19] = (uint32_t)max_mcu_per_dpu;
    for (int t = 0; t < 4 && h.quantization_tables[t].set; t++)
        for (int k = 0; k < 64; k++) md[20 + t * 64 + k] = h.quantization_tables[t].table[k];
    return md;
}
#endif
