// libb200jpeg.so - C ABI (include/b200jpeg.h) and host orchestration of the sm_100a kernels.
// No CPU decode path exists in this library: every compute entry needs a CUDA device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <time.h>
#include <string>
#include <vector>

#include "../../include/b200jpeg.h"
#include "bj_dev.h"
#include "bj_host.h"
#include "kernels_idct.cuh"
#include "parse.h"
#include "batch.h"

using namespace bj;

// ------------------------------------------------------------------------------------------------ context

extern "C" const char *bj_status_string(int s) {
    switch (s) {
        case BJ_OK: return "ok";
        case BJ_ERR_ARG: return "bad argument";
        case BJ_ERR_CUDA: return "CUDA error";
        case BJ_ERR_NOMEM: return "out of memory";
        case BJ_ERR_INVALID_JPEG: return "invalid JPEG (the reference's read_JPEG rejects it)";
        case BJ_ERR_UNSUPPORTED: return "unsupported JPEG (not a single-scan baseline file)";
        case BJ_ERR_CORRUPT_SCAN: return "corrupt entropy-coded data";
    }
    return "unknown status";
}

extern "C" const char *bj_build_info(void) { return "b200jpeg sm_100a (CUDA " BJ_STR(CUDART_VERSION) "), built " __DATE__; }

extern "C" int bj_create(bj_ctx **out, int device) {
    if (!out) return BJ_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "b200jpeg: no CUDA device (%s) - this library has no CPU fallback\n", cudaGetErrorString(e));
        return BJ_ERR_CUDA;
    }
    if (device < 0 || device >= count) return BJ_ERR_ARG;
    bj_ctx *c = new (std::nothrow) bj_ctx();
    if (!c) return BJ_ERR_NOMEM;
    c->device = device;
    if (c->check(cudaSetDevice(device)) != BJ_OK) { delete c; return BJ_ERR_CUDA; }
    cudaDeviceProp prop;
    if (c->check(cudaGetDeviceProperties(&prop, device)) != BJ_OK) { delete c; return BJ_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    if (c->check(cudaFuncSetAttribute(k_idct_color, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemIdctColor)) != BJ_OK) { delete c; return BJ_ERR_CUDA; }
    if (batch_kernels_init(c) != BJ_OK) { delete c; return BJ_ERR_CUDA; }
    for (int i = 0; i < kSlots; i++)
        if (c->check(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking)) != BJ_OK) { delete c; return BJ_ERR_CUDA; }
    {
        const unsigned hw = std::thread::hardware_concurrency();
        c->host_pool.resize(std::max(1, std::min(4, (int)hw / 2)));   // measured on the 16-core B200 box: 4 is the knee, more only contends with the copy engines
    }
    *out = c;
    return BJ_OK;
}

extern "C" void bj_destroy(bj_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &p : c->pool) p.release();
    for (auto &b : c->slots) if (b) { b->release(); delete b; b = nullptr; }
    for (int i = 0; i < kSlots; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    delete c;
}

extern "C" const char *bj_last_error(const bj_ctx *c) { return c ? c->last_error.c_str() : "null context"; }
extern "C" int bj_device_sm_count(const bj_ctx *c) { return c ? c->sm_count : 0; }

extern "C" void *bj_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void bj_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int bj_set_option(bj_ctx *c, const char *name, long value) {
    if (!c || !name) return BJ_ERR_ARG;
    if (!strcmp(name, "subseq_bits")) { if (value != 0 && (value < 128 || value % 32 || value > (1 << 18))) return BJ_ERR_ARG; c->subseq_bits = (int)value; return BJ_OK; }
    if (!strcmp(name, "sub_batch_ramp")) { c->sub_batch_ramp = value != 0; return BJ_OK; }
    if (!strcmp(name, "sync_phased")) { c->sync_phased = value != 0; return BJ_OK; }
    if (!strcmp(name, "slices")) { if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return BJ_ERR_ARG; c->slices = (int)value; return BJ_OK; }
    if (!strcmp(name, "sub_batch_bytes")) { if (value < (1 << 16)) return BJ_ERR_ARG; c->sub_batch_bytes = (size_t)value; return BJ_OK; }
    if (!strcmp(name, "packed_outputs")) { c->packed_outputs = value != 0; return BJ_OK; }
    if (!strcmp(name, "packed_inputs")) { c->packed_inputs = value != 0; return BJ_OK; }
    if (!strcmp(name, "host_threads")) { if (value < 1 || value > 256) return BJ_ERR_ARG; c->host_pool.resize((int)value); return BJ_OK; }
    if (!strcmp(name, "sync_rounds")) { if (value < 0 || value > kMaxRounds) return BJ_ERR_ARG; c->sync_rounds = (int)value; return BJ_OK; }
    return BJ_ERR_ARG;
}

extern "C" int bj_get_stat(const bj_ctx *c, const char *name, double *value) {
    if (!c || !name || !value) return BJ_ERR_ARG;
    if (!strcmp(name, "exec_ms")) { *value = c->last_exec_ms; return BJ_OK; }                 // kernel time of the last bj_exec_mcus
    if (!strcmp(name, "decode_batch_sub_batches")) { *value = c->stats[0]; return BJ_OK; }    // of the last bj_decode_batch
    if (!strcmp(name, "decode_batch_launches")) { *value = c->stats[1]; return BJ_OK; }
    if (!strcmp(name, "decode_batch_h2d_bytes")) { *value = c->stats[2]; return BJ_OK; }
    if (!strcmp(name, "decode_batch_d2h_bytes")) { *value = c->stats[3]; return BJ_OK; }
    if (!strcmp(name, "decode_batch_host_ms")) { *value = c->stats[4]; return BJ_OK; }        // parse + layout + pack, summed
    if (!strcmp(name, "decode_batch_wait_ms")) { *value = c->stats[5]; return BJ_OK; }        // host blocked on the GPU, summed
    if (!strcmp(name, "decode_batch_d2h_copies")) { *value = c->stats[6]; return BJ_OK; }
    if (!strcmp(name, "host_threads")) { *value = c->host_pool.threads(); return BJ_OK; }
    return BJ_ERR_ARG;
}

// ------------------------------------------------------------------------------------------------ compat entry

extern "C" int bj_exec_mcus_device(bj_ctx *c, const uint32_t *d_md, int16_t *d_mcus, int nchunk, int M, void *stream) {
    if (!c || !d_md || !d_mcus || nchunk < 0 || M < 4) return BJ_ERR_ARG;
    if (nchunk == 0) return BJ_OK;
    const int blk_per_chunk = M / 4;
    const long long nblk = (long long)nchunk * blk_per_chunk;
    const int grid = (int)((nblk + 15) / 16);
    k_exec_mcus<<<grid, kTileThreads, 0, (cudaStream_t)stream>>>(d_md, d_mcus, nchunk, blk_per_chunk, 64 * M * 3);
    return c->check(cudaGetLastError());
}

extern "C" int bj_exec_mcus(bj_ctx *c, const uint32_t *metadata, int16_t *mcus, int nchunk) {
    if (!c || !metadata || !mcus || nchunk < 0) return BJ_ERR_ARG;
    if (nchunk == 0) return BJ_OK;
    int M = 0;                                   // MAX_MCU_PER_DPU as the host compiled it (metadata[19], decoder_host.cpp:172)
    for (int i = 0; i < nchunk && M == 0; i++) M = (int)metadata[(size_t)i * 276 + 19];
    if (M == 0) return BJ_OK;                    // only idle DPUs: the program touches nothing
    if (M < 4) return BJ_ERR_ARG;
    if (c->check(cudaSetDevice(c->device)) != BJ_OK) return BJ_ERR_CUDA;
    const size_t md_bytes = (size_t)nchunk * 276 * 4, mc_bytes = (size_t)nchunk * 64 * M * 3 * 2;
    DevBuf &dmd = c->pool[POOL_COMPAT_MD], &dmc = c->pool[POOL_COMPAT_MCUS];
    if (dmd.reserve(md_bytes) != BJ_OK || dmc.reserve(mc_bytes) != BJ_OK) return BJ_ERR_NOMEM;
    cudaStream_t s = c->streams[0];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // the three pim.copy calls + pim.exec of src/decoder_host.cpp:276-308
    int rc = c->check(cudaMemcpyAsync(dmd.p, metadata, md_bytes, cudaMemcpyHostToDevice, s));
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(dmc.p, mcus, mc_bytes, cudaMemcpyHostToDevice, s));
    cudaEventRecord(e0, s);
    if (rc == BJ_OK) rc = bj_exec_mcus_device(c, (const uint32_t *)dmd.p, (int16_t *)dmc.p, nchunk, M, s);
    cudaEventRecord(e1, s);
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(mcus, dmc.p, mc_bytes, cudaMemcpyDeviceToHost, s));
    if (rc == BJ_OK) rc = c->check(cudaStreamSynchronize(s));
    if (rc == BJ_OK) cudaEventElapsedTime(&c->last_exec_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

// ------------------------------------------------------------------------------------------------ descriptors

extern "C" int bj_parse_header(const uint8_t *file, size_t len, bj_image_desc *desc) {
    if (!file || !desc) return BJ_ERR_ARG;
    return parse_header(file, len, desc);
}

extern "C" size_t bj_output_size(const bj_image_desc *d, int format) {
    if (!d) return 0;
    const size_t w = d->width, h = d->height;
    if (format == BJ_OUT_RGB8) return w * h * 3;
    if (format == BJ_OUT_BMP) return 26 + h * (w * 3 + w % 4);
    return 0;
}

// ------------------------------------------------------------------------------------------------ stage entry

extern "C" int bj_stage_idct_color(bj_ctx *c, const bj_image_desc *desc, const int16_t *coef_zz, int format, uint8_t *out) {
    if (!c || !desc || !coef_zz || !out) return BJ_ERR_ARG;
    if (format != BJ_OUT_RGB8 && format != BJ_OUT_BMP) return BJ_ERR_ARG;
    if (c->check(cudaSetDevice(c->device)) != BJ_OK) return BJ_ERR_CUDA;
    ImgDev im;
    std::vector<TileDev> tiles;
    Geometry g = geometry_of(*desc);
    fill_imgdev(*desc, g, format, /*du_base=*/0, /*out_base=*/0, &im);
    append_tiles(g, 0, 0, &tiles);
    const size_t coef_bytes = (size_t)g.ndu * 128, out_bytes = bj_output_size(desc, format);
    DevBuf &dco = c->pool[POOL_COEF], &dout = c->pool[POOL_OUT], &dim = c->pool[POOL_IMGS], &dti = c->pool[POOL_TILES];
    if (dco.reserve(coef_bytes) || dout.reserve(out_bytes + 64) || dim.reserve(sizeof(im)) || dti.reserve(tiles.size() * sizeof(TileDev))) return BJ_ERR_NOMEM;
    cudaStream_t s = c->streams[0];
    int rc = c->check(cudaMemcpyAsync(dco.p, coef_zz, coef_bytes, cudaMemcpyHostToDevice, s));
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(dim.p, &im, sizeof(im), cudaMemcpyHostToDevice, s));
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(dti.p, tiles.data(), tiles.size() * sizeof(TileDev), cudaMemcpyHostToDevice, s));
    if (rc == BJ_OK) {
        k_idct_color<<<(unsigned)tiles.size(), kTileThreads, kSmemIdctColor, s>>>((const int16_t *)dco.p, nullptr, (const ImgDev *)dim.p, (const TileDev *)dti.p, (uint8_t *)dout.p);
        rc = c->check(cudaGetLastError());
    }
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(out, dout.p, out_bytes, cudaMemcpyDeviceToHost, s));
    if (rc == BJ_OK) rc = c->check(cudaStreamSynchronize(s));
    return rc;
}

// ------------------------------------------------------------------------------------------------ full path, staged

extern "C" int bj_batch_create(bj_ctx *c, const uint8_t *const *files, const size_t *lens, int n, int format, bj_batch **out) {
    if (!c || !out || n < 0 || (n > 0 && (!files || !lens))) return BJ_ERR_ARG;
    *out = nullptr;
    bj_batch *b = new (std::nothrow) bj_batch();
    if (!b) return BJ_ERR_NOMEM;
    const int rc = batch_assign(b, c, files, lens, n, format);
    if (rc != BJ_OK) { b->release(); delete b; return rc; }
    *out = b;
    return BJ_OK;
}

extern "C" void bj_batch_destroy(bj_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->last_stream) cudaStreamSynchronize(b->last_stream);
    b->release();
    delete b;
}

static cudaStream_t stream_or_default(bj_batch *b, void *stream) { return stream ? (cudaStream_t)stream : b->ctx->streams[0]; }

extern "C" int bj_batch_upload(bj_batch *b, void *stream) { return b ? batch_upload(b, stream_or_default(b, stream)) : BJ_ERR_ARG; }
extern "C" int bj_batch_decode(bj_batch *b, void *stream) { return b ? batch_decode(b, stream_or_default(b, stream)) : BJ_ERR_ARG; }
extern "C" int bj_batch_sync(bj_batch *b) { return b ? batch_sync(b) : BJ_ERR_ARG; }
extern "C" int bj_batch_download(bj_batch *b, uint8_t *const *outs, void *stream) {
    if (!b || (b->n > 0 && !outs)) return BJ_ERR_ARG;
    return batch_download(b, outs, stream_or_default(b, stream));
}

extern "C" int bj_batch_status(const bj_batch *b, int *status) {
    if (!b || !status) return BJ_ERR_ARG;
    for (int i = 0; i < b->n; i++) status[i] = batch_image_status(b, i);
    return BJ_OK;
}

extern "C" int bj_batch_get_info(const bj_batch *b, bj_batch_info *info) {
    if (!b || !info) return BJ_ERR_ARG;
    memset(info, 0, sizeof(*info));
    info->pixels = b->pixels; info->scan_bytes = b->scan_bytes; info->data_units = b->coef_units; info->out_bytes = b->out_bytes;
    info->h2d_bytes = b->files_bytes + b->meta_bytes; info->d2h_bytes = b->d2h_bytes;
    uint32_t nsub = 0;
    if (b->synced) for (int i = 0; i < b->n; i++) nsub += const_cast<bj_batch *>(b)->h_state()[i].nsub;
    info->subsequences = nsub; info->sync_rounds = b->sync_rounds; info->launches = b->launches;
    info->ms_entropy = b->ms_entropy; info->ms_idct = b->ms_idct;
    info->ms_unstuff = b->ms_unstuff; info->ms_sync = b->ms_sync; info->ms_write = b->ms_write;
    info->clean_bytes = 0;
    if (b->synced) for (int i = 0; i < b->n; i++) info->clean_bytes += const_cast<bj_batch *>(b)->h_state()[i].clean_len;
    return BJ_OK;
}

extern "C" int bj_batch_output_offset(const bj_batch *b, int i, size_t *offset, size_t *bytes) {
    if (!b || i < 0 || i >= b->n) return BJ_ERR_ARG;
    if (offset) *offset = b->out_off[i];
    if (bytes) *bytes = b->out_size[i];
    return b->parse_status[i];
}

extern "C" int bj_batch_device_output(const bj_batch *b, int i, void **dptr, size_t *bytes) {
    if (!b || i < 0 || i >= b->n || b->parse_status[i] != BJ_OK) return BJ_ERR_ARG;
    if (dptr) *dptr = (uint8_t *)b->d_out.p + b->out_off[i];
    if (bytes) *bytes = b->out_size[i];
    return BJ_OK;
}

extern "C" int bj_batch_device_coefficients(const bj_batch *b, int i, void **dptr, size_t *bytes, void **dc_dptr) {
    if (!b || i < 0 || i >= b->n || b->parse_status[i] != BJ_OK) return BJ_ERR_ARG;
    if (dptr) *dptr = (int16_t *)b->d_coef.p + (size_t)b->du_base[i] * 64;
    if (bytes) *bytes = (size_t)b->ndu[i] * 128;
    if (dc_dptr) *dc_dptr = (int16_t *)b->d_dc.p + b->du_base[i];
    return BJ_OK;
}

// Stage-level entry for known-answer tests: K0 + K1 only; coefficients (zig-zag, DC un-differenced) back to the host.
extern "C" int bj_stage_entropy(bj_ctx *c, const uint8_t *file, size_t len, int16_t *coef_zz, size_t capacity_bytes, int *status) {
    if (!c || !file || !coef_zz) return BJ_ERR_ARG;
    bj_batch *b = nullptr;
    const uint8_t *files[1] = {file};
    const size_t lens[1] = {len};
    int rc = bj_batch_create(c, files, lens, 1, BJ_OUT_RGB8, &b);
    if (rc != BJ_OK) return rc;
    if (b->parse_status[0] != BJ_OK) { rc = b->parse_status[0]; bj_batch_destroy(b); return rc; }
    if ((size_t)b->ndu[0] * 128 > capacity_bytes) { bj_batch_destroy(b); return BJ_ERR_ARG; }
    cudaStream_t s = c->streams[0];
    rc = batch_upload(b, s);
    if (rc == BJ_OK) rc = batch_decode(b, s);
    if (rc == BJ_OK) rc = batch_sync(b);
    std::vector<int16_t> dcs(b->ndu[0]);
    if (rc == BJ_OK) rc = c->check(cudaMemcpyAsync(coef_zz, b->d_coef.p, (size_t)b->ndu[0] * 128, cudaMemcpyDeviceToHost, s));
    if (rc == BJ_OK && b->ndu[0]) rc = c->check(cudaMemcpyAsync(dcs.data(), b->d_dc.p, (size_t)b->ndu[0] * 2, cudaMemcpyDeviceToHost, s));
    if (rc == BJ_OK) rc = c->check(cudaStreamSynchronize(s));
    if (rc == BJ_OK) for (uint32_t u = 0; u < b->ndu[0]; u++) coef_zz[(size_t)u * 64] = dcs[u];   // the device keeps DC in its own plane
    if (rc == BJ_OK && status) *status = batch_image_status(b, 0);
    bj_batch_destroy(b);
    return rc;
}

// ------------------------------------------------------------------------------------------------ full path, one call
// Sub-batches rotate over kSlots batch objects, each with its own stream: while sub-batch k copies out (PCIe D2H is
// the bound of this call: 3 bytes per pixel), sub-batch k+1 decodes and the host parses and packs sub-batch k+2
// on the worker pool.
static double wall_ms() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

extern "C" int bj_decode_batch(bj_ctx *c, const uint8_t *const *files, const size_t *lens, int n, int format,
                               uint8_t *const *outs, int *status) {
    if (!c || n < 0 || (n > 0 && (!files || !lens || !outs))) return BJ_ERR_ARG;
    if (format != BJ_OUT_RGB8 && format != BJ_OUT_BMP) return BJ_ERR_ARG;
    if (c->check(cudaSetDevice(c->device)) != BJ_OK) return BJ_ERR_CUDA;
    const size_t budget = c->sub_batch_bytes ? c->sub_batch_bytes : ((size_t)24 << 20);    // compressed bytes per sub-batch
    for (auto &b : c->slots) if (!b) { b = new (std::nothrow) bj_batch(); if (!b) return BJ_ERR_NOMEM; }
    int first[kSlots] = {}, count[kSlots] = {};
    bool busy[kSlots] = {};
    double nsub = 0, launches = 0, h2d = 0, d2h = 0, host_ms = 0, wait_ms = 0, d2h_copies = 0;
    int rc = BJ_OK;
    // B200JPEG_TRACE=1: one line per sub-batch on stderr - when its kernels ran and its copy-out ended (ms since the
    // first sub-batch was enqueued) - to see whether the copy-out engine is kept busy.  (No event is recorded in front
    // of the upload: on a stream whose last operation was a copy-out it waits for the copy-out engine.)
    static const bool trace = getenv("B200JPEG_TRACE") != nullptr;
    cudaEvent_t ev_base = nullptr;
    double host_t[kSlots][2] = {};
    const double wall0 = wall_ms();
    static const bool blocking_wait = getenv("B200JPEG_SPIN_WAIT") == nullptr;
    auto finish = [&](int slot) -> int {
        bj_batch *b = c->slots[slot];
        const double t0 = wall_ms();
        // sleep until the sub-batch's copy-out is done instead of polling for it (cudaStreamSynchronize spins by
        // default, and a polling host thread takes bandwidth from the copy-out it is waiting for)
        if (blocking_wait && b->ev_done) cudaEventSynchronize(b->ev_done);
        int r = batch_sync(b);
        if (trace && r == BJ_OK && ev_base) {
            float k0 = 0, k1 = 0, out = 0;
            cudaEventSynchronize(b->ev[5]);
            cudaEventElapsedTime(&k0, ev_base, b->ev[0]);
            cudaEventElapsedTime(&k1, ev_base, b->ev[4]); cudaEventElapsedTime(&out, ev_base, b->ev[5]);
            fprintf(stderr, "b200jpeg trace: sub-batch of %4d images (%6.1f MB out)  host prepare %6.2f..%6.2f  kernels %6.2f..%6.2f  copied out %6.2f\n",
                    b->n, b->d2h_bytes / 1e6, host_t[slot][0], host_t[slot][1], k0, k1, out);
        }
        if (r == BJ_OK && b->n_blk && b->h_flags()[b->rounds - 1] != 0)   // extra rounds ran: the early copy-out is stale
            r = batch_download(b, outs + first[slot], c->streams[slot]);
        wait_ms += wall_ms() - t0;
        if (r == BJ_OK && status) for (int i = 0; i < count[slot]; i++) status[first[slot] + i] = batch_image_status(b, i);
        launches += b->launches; h2d += (double)(b->files_bytes + b->meta_bytes); d2h += (double)b->d2h_bytes; d2h_copies += b->d2h_copies;
        busy[slot] = false;
        return r;
    };
    int i0 = 0, k = 0;
    while (rc == BJ_OK && i0 < n) {
        int i1 = i0;
        size_t bytes = 0;
        // the first sub-batches are small, so that the copy-out (the bound of this call) starts early
        const size_t cap = !c->sub_batch_ramp ? budget : (k == 0 ? budget / 8 : (k == 1 ? budget / 3 : budget));
        while (i1 < n && (i1 == i0 || bytes + lens[i1] <= cap)) bytes += lens[i1++];
        const int slot = k % kSlots;
        if (busy[slot]) rc = finish(slot);
        if (rc != BJ_OK) break;
        bj_batch *b = c->slots[slot];
        cudaStream_t s = c->streams[slot];
        const double t0 = wall_ms();
        rc = batch_assign(b, c, files + i0, lens + i0, i1 - i0, format, c->packed_inputs != 0);
        host_ms += wall_ms() - t0;
        host_t[slot][0] = t0 - wall0; host_t[slot][1] = wall_ms() - wall0;
        if (rc == BJ_OK && trace) {
            if (!ev_base) { cudaEventCreate(&ev_base); cudaEventRecord(ev_base, s); }
        }
        if (rc == BJ_OK) rc = batch_upload(b, s);
        if (rc == BJ_OK) rc = batch_decode(b, s);
        if (rc == BJ_OK) {                                            // enqueue the copy-out behind the kernels, no host wait
            b->synced = true;                                         // (checked for real in finish())
            rc = batch_download_async(b, outs + i0, s);
            b->synced = false;
            if (trace) cudaEventRecord(b->ev[5], s);
            if (!b->ev_done) cudaEventCreateWithFlags(&b->ev_done, cudaEventBlockingSync | cudaEventDisableTiming);
            if (b->ev_done) cudaEventRecord(b->ev_done, s);
        }
        first[slot] = i0; count[slot] = i1 - i0; busy[slot] = true;
        nsub += 1;
        i0 = i1; k++;
    }
    // drain in submission order
    for (int j = 0; j < kSlots; j++) { const int slot = (k + j) % kSlots; if (busy[slot]) { const int r = finish(slot); if (rc == BJ_OK) rc = r; } }
    if (ev_base) cudaEventDestroy(ev_base);
    c->stats[0] = nsub; c->stats[1] = launches; c->stats[2] = h2d; c->stats[3] = d2h; c->stats[4] = host_ms; c->stats[5] = wait_ms; c->stats[6] = d2h_copies;
    return rc;
}
