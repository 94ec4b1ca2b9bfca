// decoder_b200 - the reference CLI (`./bin/decoder <jpeg_1> ...`, src/decoder_host.cpp:352-394) on the B200 back end,
// full path: file bytes -> bj_submit (Huffman + dequant + IDCT + colour + BMP bytes on the GPU) -> one write() per image.
// Same user-visible behaviour as the reference:
//   * inputs are sorted ascending by file size                      (src/decoder_host.cpp:46-61, :360)
//   * `<name>.bmp` is written next to each input                    (:326-331)
//   * an unreadable / invalid file prints "<file>: Error - Invalid JPEG" and is skipped   (:120-123)
//   * a "Profiles:" block is printed at the end, with the per-stage lines   (:379-394)
// and the same shape of pipeline: the reference overlaps its producer (parse + Huffman) with its consumer (offload +
// BMP write) on two threads and a queue (:25-38, :364-365); here three stages run side by side on groups of images -
// a reader thread (file bytes straight into page-locked memory, header peek), the GPU decode (bj_submit / bj_wait,
// zero-copy upload) and a writer thread (one fwrite per BMP).
// Devices: like the reference's single process takes every DPU (DPU_ALLOCATE_ALL, :32), one process takes every visible
// GPU (bj_create_multi) unless told otherwise: B200JPEG_DEVICES=<n> limits the count, B200JPEG_DEVICE=<i> picks one;
// under a one-process-per-GPU launcher (RANK / WORLD_SIZE / LOCAL_RANK) each process takes its LOCAL_RANK GPU and its
// round-robin share of the sorted list - images are independent, there is no collective.
// There is no CPU fallback: without a CUDA device bj_create fails.
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200jpeg.h"

static double now_s() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + t.tv_nsec * 1e-9;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

struct Input {
    std::string path;
    size_t size = 0;
};

static std::string bmp_name(const std::string &p) {
    const size_t pos = p.find_last_of('.');
    return (pos == std::string::npos ? p : p.substr(0, pos)) + ".bmp";
}

// One group of images on its way through the three stages.  Its buffers are page-locked and reused.
struct Group {
    size_t i0 = 0, n = 0;                       // images [i0, i0 + n) of the sorted list
    uint8_t *in = nullptr, *out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    std::vector<const uint8_t *> files;
    std::vector<size_t> lens, out_bytes;
    std::vector<uint8_t *> outs;
    std::vector<int> status;
    bj_job *job = nullptr;
    bool last = false;
};

template <class T> class Channel {              // the queue between two stages (src/decoder_host.cpp:35-38: batched_queue, mtx, cv)
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
public:
    void push(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
    T pop() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
};

// fn(k) for k in [0, n) on `threads` threads (file reads and BMP writes of a group are independent per image)
template <class F> static void parallel_each(size_t n, int threads, F fn) {
    if (n == 0) return;
    std::atomic<size_t> next{0};
    auto work = [&] { for (size_t k; (k = next.fetch_add(1)) < n;) fn(k); };
    std::vector<std::thread> th;
    for (int t = 1; t < threads && (size_t)t < n; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

static bool grow(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return true;
    if (*p) bj_host_free(*p);
    *cap = need + need / 4 + 4096;
    *p = static_cast<uint8_t *>(bj_host_alloc(*cap));
    if (!*p) { fprintf(stderr, "decoder_b200: pinned allocation of %zu bytes failed\n", *cap); *cap = 0; return false; }
    return true;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        printf("Error - Invalid arguments\n");
        return 1;
    }
    const int rank = env_int("RANK", 0), world = std::max(1, env_int("WORLD_SIZE", 1));
    const size_t out_budget = (size_t)env_int("B200JPEG_OUT_MB", 192) << 20;      // decoded bytes per group
    const size_t in_budget = (size_t)env_int("B200JPEG_IN_MB", 32) << 20;         // compressed bytes per group
    const int io_threads = std::max(1, env_int("B200JPEG_IO_THREADS", 4));                  // file reads / BMP writes side by side
    const size_t max_images = (size_t)std::max(1, env_int("B200JPEG_GROUP_IMAGES", 1 << 20));   // images per group

    const double t_start = now_s();
    std::vector<Input> all(argc - 1);
    for (int i = 1; i < argc; i++) {
        all[i - 1].path = argv[i];
        struct stat sb;
        all[i - 1].size = stat(argv[i], &sb) == 0 ? (size_t)sb.st_size : 0;
    }
    std::stable_sort(all.begin(), all.end(), [](const Input &a, const Input &b) { return a.size < b.size; });
    std::vector<Input> in;
    for (size_t i = rank; i < all.size(); i += world) in.push_back(std::move(all[i]));

    bj_ctx *ctx = nullptr;
    int rc;
    if (world > 1 || getenv("B200JPEG_DEVICE")) {
        const int device = env_int("B200JPEG_DEVICE", env_int("LOCAL_RANK", 0));
        rc = bj_create(&ctx, device);
    } else {
        rc = bj_create_multi(&ctx, nullptr, env_int("B200JPEG_DEVICES", 0));       // 0 = every visible GPU
    }
    if (rc != BJ_OK) {
        fprintf(stderr, "decoder_b200: bj_create: %s\n", bj_status_string(rc));
        return 2;
    }
    bj_set_option(ctx, "packed_outputs", 1);
    const double t_ready = now_s();
    printf("B200 devices: %d (%d SMs each)\n", bj_device_count(ctx), bj_device_sm_count(ctx));

    constexpr int kGroups = 4;                   // one being read, one decoding, one queued behind it, one being written
    Group groups[kGroups];
    Channel<Group *> to_decode, to_write, to_read;
    for (auto &g : groups) to_read.push(&g);
    double t_read = 0, t_decode = 0, t_write = 0;
    int calls = 0, failures = 0;
    std::atomic<bool> fatal{false};

    // ---- stage 1: read the files of a group into page-locked memory, peek at the headers to size the outputs
    std::thread reader([&] {
        size_t i0 = 0;
        while (true) {
            Group *g = to_read.pop();
            const double t0 = now_s();
            g->i0 = i0; g->n = 0; g->last = false;
            size_t in_bytes = 0, out_total = 0, i1 = i0;
            while (i1 < in.size() && i1 - i0 < max_images && (i1 == i0 || in_bytes + in[i1].size + 16 <= in_budget)) { in_bytes += in[i1].size + 16; i1++; }
            if (!grow(&g->in, &g->in_cap, in_bytes + 64)) { fatal = true; g->last = true; to_decode.push(g); return; }
            g->files.clear(); g->lens.clear(); g->out_bytes.clear();
            // read the candidates side by side (each file at its place in the pinned buffer), then keep as many as the
            // output budget allows
            std::vector<size_t> in_off(i1 - i0 + 1, 0), got(i1 - i0, 0), obytes(i1 - i0, 0);
            for (size_t k = i0; k < i1; k++) in_off[k - i0 + 1] = in_off[k - i0] + in[k].size;
            parallel_each(i1 - i0, io_threads, [&](size_t j) {
                FILE *fp = fopen(in[i0 + j].path.c_str(), "rb");
                if (fp) { got[j] = in[i0 + j].size ? fread(g->in + in_off[j], 1, in[i0 + j].size, fp) : 0; fclose(fp); }
                bj_image_desc d;
                obytes[j] = (got[j] && bj_peek_header(g->in + in_off[j], got[j], &d) == BJ_OK) ? bj_output_size(&d, BJ_OUT_BMP) : 0;
            });
            size_t k = i0;
            for (; k < i1; k++) {
                const size_t padded = (obytes[k - i0] + 15) / 16 * 16;
                if (k > i0 && out_total + padded > out_budget) break;          // (the rest is read again with the next group)
                g->files.push_back(g->in + in_off[k - i0]); g->lens.push_back(got[k - i0]); g->out_bytes.push_back(obytes[k - i0]);
                out_total += padded;
            }
            g->n = k - i0;
            i0 = k;
            g->last = i0 >= in.size();
            if (!grow(&g->out, &g->out_cap, out_total + 64)) { fatal = true; g->last = true; g->n = 0; to_decode.push(g); return; }
            g->outs.assign(g->n, nullptr); g->status.assign(g->n, 0);
            size_t oo = 0;
            for (size_t j = 0; j < g->n; j++) { g->outs[j] = g->out_bytes[j] ? g->out + oo : nullptr; oo += (g->out_bytes[j] + 15) / 16 * 16; }
            t_read += now_s() - t0;
            const bool last = g->last;
            to_decode.push(g);
            if (last) return;
        }
    });

    // ---- stage 3: write the BMPs of a finished group.  Like the reference, a scan that fails to decode still produces
    // its (partial) image (the result of decode_Huffman_data is ignored, src/decoder_host.cpp:181).
    std::thread writer([&] {
        while (true) {
            Group *g = to_write.pop();
            const double t0 = now_s();
            std::atomic<int> bad{0};
            std::mutex print_m;
            parallel_each(g->n, io_threads, [&](size_t k) {
                const Input &f = in[g->i0 + k];
                if (g->status[k] != BJ_OK && g->status[k] != BJ_ERR_CORRUPT_SCAN) {
                    std::lock_guard<std::mutex> l(print_m);
                    printf("%s: Error - Invalid JPEG\n", f.path.c_str());
                    bad++;
                } else {
                    FILE *fp = fopen(bmp_name(f.path).c_str(), "wb");
                    if (!fp || fwrite(g->outs[k], 1, g->out_bytes[k], fp) != g->out_bytes[k]) { std::lock_guard<std::mutex> l(print_m); printf("%s: Error - cannot write BMP\n", f.path.c_str()); bad++; }
                    if (fp) fclose(fp);
                }
            });
            failures += bad;
            t_write += now_s() - t0;
            const bool last = g->last;
            to_read.push(g);
            if (last) return;
        }
    });

    // ---- stage 2 (this thread): hand each group to the GPU as soon as it is read; keep one job decoding while the next
    // is submitted, pass finished groups on to the writer
    Group *running = nullptr;
    double t_run0 = 0;
    auto finish_running = [&]() {
        if (!running) return;
        const int r = bj_wait(running->job);
        running->job = nullptr;
        if (r != BJ_OK) {
            fprintf(stderr, "decoder_b200: decode: %s (%s)\n", bj_status_string(r), bj_last_error(ctx));
            for (auto &s : running->status) s = r;
            fatal = true;
        }
        t_decode += now_s() - t_run0;
        calls++;
        to_write.push(running);
        running = nullptr;
    };
    while (true) {
        Group *g = to_decode.pop();
        bj_job *job = nullptr;
        int r = g->n ? bj_submit(ctx, g->files.data(), g->lens.data(), (int)g->n, BJ_OUT_BMP, g->outs.data(), g->status.data(), &job) : BJ_OK;
        if (r != BJ_OK) { fprintf(stderr, "decoder_b200: bj_submit: %s\n", bj_status_string(r)); fatal = true; for (auto &s : g->status) s = r; }
        finish_running();                          // (the new job is already queued behind it)
        if (job) { g->job = job; running = g; t_run0 = now_s(); }
        else to_write.push(g);
        if (g->last) break;
    }
    finish_running();
    reader.join();
    writer.join();

    double ms[4] = {0, 0, 0, 0}, direct = 0, subs = 0;
    bj_get_stat(ctx, "total_ms_unstuff", &ms[0]); bj_get_stat(ctx, "total_ms_sync", &ms[1]);
    bj_get_stat(ctx, "total_ms_write", &ms[2]); bj_get_stat(ctx, "total_ms_idct", &ms[3]);
    bj_get_stat(ctx, "total_direct_uploads", &direct); bj_get_stat(ctx, "total_sub_batches", &subs);
    for (auto &g : groups) { if (g.in) bj_host_free(g.in); if (g.out) bj_host_free(g.out); }
    bj_destroy(ctx);

    // (the reference prints: total, queue waiting, CPU-DPU transfer, DPU execution, DPU-CPU transfer, BMP write, and the
    // DPU program's own split into initialization / dequantization / IDCT / colour conversion)
    printf("\nProfiles:\n");
    printf("End-to-end execution time: %gs\n", now_s() - t_start);
    printf(" - Start-up (CUDA context, kernel image): %gs\n", t_ready - t_start);
    printf(" - File read + header peek time (reader thread): %gs\n", t_read);
    printf(" - B200 decode time (H2D + kernels + D2H, overlapped with reading and writing): %gs\n", t_decode);
    printf("    - scan filter / segmenter kernels: %gs\n", ms[0] * 1e-3);
    printf("    - Huffman synchronisation kernels: %gs\n", ms[1] * 1e-3);
    printf("    - Huffman write + DC prediction kernels: %gs\n", ms[2] * 1e-3);
    printf("    - dequantisation + IDCT + colour conversion kernel: %gs\n", ms[3] * 1e-3);
    printf(" - BMP write time (writer thread): %gs\n", t_write);
    printf(" - Total %d groups, %d sub-batches (%d uploaded without a staging copy), %zu images, %d failed\n", calls, (int)subs, (int)direct, in.size(), failures);
    return fatal ? 2 : 0;
}
