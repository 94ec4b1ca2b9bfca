// TEST INFRASTRUCTURE (oracle build only): stage-level access to the *real* reference code.
// Links the reference's own jpeg_scanner.cpp / bmp_writer.cpp / decoder_dpu.c (compiled from /root/reference,
// never copied) and exposes, as plain C, the intermediate buffers of one image so tests can pin every stage:
//   post-Huffman  MCU_buffer  (decoder_host.cpp:181 -> jpeg_scanner.cpp:707)
//   post-exec     mcus        (decoder_host.cpp:292,308 -> decoder_dpu.c:82)
//   BMP bytes                 (decoder_host.cpp:330 -> bmp_writer.cpp:19)
// The metadata record is filled exactly as decoder_host.cpp:156-178 does.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "headers/jpeg.h"
#include "headers/bmp.h"

extern "C" {
int  oracle_dpu_mcus_len(void);
void oracle_dpu_run(const uint32_t *metadata276, short *mcus_inout);
}

#ifndef MAX_MCU_PER_DPU
#error "MAX_MCU_PER_DPU must be defined"
#endif

extern "C" {

struct ref_info {
    uint32_t width, height, ncomp, h_samp, v_samp, mcu_w, mcu_h, mcu_w_real, mcu_h_real, restart_interval;
    uint32_t nchunks, valid, huffman_ok, scan_bytes;
};

// Number of DPU chunks an image needs (decoder_host.cpp:125-128).  Returns 0 for an invalid file.
int ref_probe(const char *path, ref_info *info) {
    Header *h = read_JPEG(path);
    std::memset(info, 0, sizeof(*info));
    if (h == nullptr) return 0;
    info->valid = h->valid;
    if (!h->valid) { delete h; return 0; }
    info->width = h->width; info->height = h->height; info->ncomp = h->num_components;
    info->h_samp = h->h_sampling_factor; info->v_samp = h->v_sampling_factor;
    info->mcu_w = h->mcu_width; info->mcu_h = h->mcu_height;
    info->mcu_w_real = h->mcu_width_real; info->mcu_h_real = h->mcu_height_real;
    info->restart_interval = h->restart_interval;
    info->scan_bytes = (uint32_t)h->huffman_data.size();
    int pw = (h->mcu_width_real + 1) / 2 * 2, ph = (h->mcu_height_real + 1) / 2 * 2;
    info->nchunks = (pw * ph + MAX_MCU_PER_DPU - 1) / MAX_MCU_PER_DPU;
    int n = info->nchunks;
    delete h;
    return n;
}

// Full reference decode of one file with stage dumps.  Buffers are caller-allocated:
//   metadata[276], mcus_pre/mcus_post [nchunks * 64*M*3] (either may be NULL).  bmp_path may be NULL.
// Returns nchunks, or <0 on error.
int ref_decode_file(const char *path, uint32_t *metadata, short *mcus_pre, short *mcus_post, const char *bmp_path, ref_info *info) {
    ref_info local;
    if (info == nullptr) info = &local;
    Header *h = read_JPEG(path);
    std::memset(info, 0, sizeof(*info));
    if (h == nullptr || !h->valid) { if (h) delete h; return -1; }
    int pw = (h->mcu_width_real + 1) / 2 * 2, ph = (h->mcu_height_real + 1) / 2 * 2;
    int nchunks = (pw * ph + MAX_MCU_PER_DPU - 1) / MAX_MCU_PER_DPU;
    const int len = oracle_dpu_mcus_len();

    std::vector<uint32_t> md(20 + 4 * 64, 0);
    md[0] = h->mcu_height; md[1] = h->mcu_width; md[2] = h->mcu_height_real; md[3] = h->mcu_width_real;
    md[4] = h->num_components; md[5] = h->v_sampling_factor; md[6] = h->h_sampling_factor;
    for (uint j = 0; j < h->num_components; j++) md[j + 7] = h->color_components[j].QT_ID;
    for (uint j = 0; j < h->num_components; j++) md[j + h->num_components + 7] = h->color_components[j].h_sampling_factor;
    for (uint j = 0; j < h->num_components; j++) md[j + h->num_components * 2 + 7] = h->color_components[j].v_sampling_factor;
    md[17] = h->height; md[18] = h->width; md[19] = MAX_MCU_PER_DPU;
    for (uint j = 0; j < 4; j++) {
        if (!h->quantization_tables[j].set) break;
        for (uint k = 0; k < 64; k++) md[20 + j * 64 + k] = h->quantization_tables[j].table[k];
    }
    if (metadata) std::memcpy(metadata, md.data(), md.size() * sizeof(uint32_t));

    std::vector<std::vector<short>> buf(nchunks, std::vector<short>(len, 0));
    bool ok = decode_Huffman_data(h, buf, 0);
    if (mcus_pre) for (int i = 0; i < nchunks; i++) std::memcpy(mcus_pre + (size_t)i * len, buf[i].data(), len * sizeof(short));
    for (int i = 0; i < nchunks; i++) oracle_dpu_run(md.data(), buf[i].data());
    if (mcus_post) for (int i = 0; i < nchunks; i++) std::memcpy(mcus_post + (size_t)i * len, buf[i].data(), len * sizeof(short));
    if (bmp_path) write_BMP(md, buf, 0, bmp_path);

    info->valid = 1; info->huffman_ok = ok;
    info->width = h->width; info->height = h->height; info->ncomp = h->num_components;
    info->h_samp = h->h_sampling_factor; info->v_samp = h->v_sampling_factor;
    info->mcu_w = h->mcu_width; info->mcu_h = h->mcu_height;
    info->mcu_w_real = h->mcu_width_real; info->mcu_h_real = h->mcu_height_real;
    info->restart_interval = h->restart_interval; info->nchunks = nchunks;
    info->scan_bytes = (uint32_t)h->huffman_data.size();
    delete h;
    return nchunks;
}

// The DPU program alone on caller-supplied chunks (the literal pim.exec() of decoder_host.cpp:292).
void ref_exec_mcus(const uint32_t *metadata /*[nchunk][276]*/, short *mcus /*[nchunk][64*M*3]*/, int nchunk) {
    const int len = oracle_dpu_mcus_len();
    for (int i = 0; i < nchunk; i++) oracle_dpu_run(metadata + (size_t)i * 276, mcus + (size_t)i * len);
}

// The reference's own write_BMP (bmp_writer.cpp:19-67) on caller-supplied post-exec chunks: what a host that keeps its
// BMP writer does with the buffers a back end hands back (decoder_host.cpp:330).
int ref_write_bmp(const uint32_t *metadata276, const short *mcus /*[nchunk][64*M*3]*/, int nchunk, const char *path) {
    const int len = oracle_dpu_mcus_len();
    std::vector<uint32_t> md(metadata276, metadata276 + 276);
    std::vector<std::vector<short>> buf(nchunk, std::vector<short>(len));
    for (int i = 0; i < nchunk; i++) std::memcpy(buf[i].data(), mcus + (size_t)i * len, len * sizeof(short));
    write_BMP(md, buf, 0, path);
    return 0;
}

int ref_max_mcu_per_dpu(void) { return MAX_MCU_PER_DPU; }

}  // extern "C"
