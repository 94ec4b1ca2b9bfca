// decoder_b200 - the reference CLI (`./bin/decoder <jpeg_1> ...`, src/decoder_host.cpp:352-394) on the B200 back end,
// full path: file bytes -> bj_decode_batch (Huffman + dequant + IDCT + colour + BMP bytes on the GPU) -> one
// write() per image.  Same user-visible behaviour as the reference:
//   * inputs are sorted ascending by file size                      (src/decoder_host.cpp:46-61, :360)
//   * `<name>.bmp` is written next to each input                    (:326-331)
//   * an unreadable / invalid file prints "<file>: Error - Invalid JPEG" and is skipped   (:120-123)
//   * a "Profiles:" block is printed at the end                     (:379-394)
// One process drives one GPU.  For a multi-GPU job start one process per GPU (torchrun-style environment:
// RANK / WORLD_SIZE / LOCAL_RANK, or B200JPEG_DEVICE); the sorted list is dealt round-robin to the ranks - images
// are independent, there is no collective.  There is no CPU fallback: without a CUDA device bj_create fails.
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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
    std::vector<uint8_t> bytes;
    bj_image_desc desc;
    int parse = BJ_ERR_INVALID_JPEG;
    size_t out_bytes = 0;
};

static std::string bmp_name(const std::string &p) {
    const size_t pos = p.find_last_of('.');
    return (pos == std::string::npos ? p : p.substr(0, pos)) + ".bmp";
}

int main(int argc, char **argv) {
    if (argc < 2) {
        printf("Error - Invalid arguments\n");
        return 1;
    }
    const int rank = env_int("RANK", 0), world = std::max(1, env_int("WORLD_SIZE", 1));
    const int device = env_int("B200JPEG_DEVICE", env_int("LOCAL_RANK", 0));
    const size_t out_budget = (size_t)env_int("B200JPEG_OUT_MB", 4096) << 20;     // pinned output bytes per call

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
    int rc = bj_create(&ctx, device);
    if (rc != BJ_OK) {
        fprintf(stderr, "decoder_b200: bj_create(device %d): %s\n", device, bj_status_string(rc));
        return 2;
    }
    bj_set_option(ctx, "packed_outputs", 1);
    printf("B200 device %d: %d SMs\n", device, bj_device_sm_count(ctx));

    double t_read = 0, t_decode = 0, t_write = 0;
    int calls = 0, failures = 0;
    uint8_t *pin = nullptr;
    size_t pin_cap = 0;
    size_t i0 = 0;
    while (i0 < in.size()) {
        // ---- read + parse a group whose outputs fit the pinned buffer
        double t0 = now_s();
        size_t i1 = i0, out_total = 0;
        while (i1 < in.size()) {
            Input &f = in[i1];
            FILE *fp = fopen(f.path.c_str(), "rb");
            if (fp) {
                f.bytes.resize(f.size);
                const size_t got = f.size ? fread(f.bytes.data(), 1, f.size, fp) : 0;
                fclose(fp);
                f.bytes.resize(got);
                f.parse = got ? bj_parse_header(f.bytes.data(), got, &f.desc) : BJ_ERR_INVALID_JPEG;
            }
            f.out_bytes = f.parse == BJ_OK ? bj_output_size(&f.desc, BJ_OUT_BMP) : 0;
            const size_t padded = (f.out_bytes + 15) / 16 * 16;
            if (i1 > i0 && out_total + padded > out_budget) { f.bytes.clear(); f.bytes.shrink_to_fit(); break; }
            out_total += padded;
            i1++;
        }
        t_read += now_s() - t0;

        // ---- decode (one call; the library double-buffers sub-batches over its own streams)
        t0 = now_s();
        if (out_total + 64 > pin_cap) {
            if (pin) bj_host_free(pin);
            pin_cap = out_total + out_total / 8 + 64;
            pin = static_cast<uint8_t *>(bj_host_alloc(pin_cap));
            if (!pin) { fprintf(stderr, "decoder_b200: pinned allocation of %zu bytes failed\n", pin_cap); return 2; }
        }
        const int n = (int)(i1 - i0);
        std::vector<const uint8_t *> files(n);
        std::vector<size_t> lens(n);
        std::vector<uint8_t *> outs(n);
        std::vector<int> status(n, 0);
        size_t o = 0;
        for (int k = 0; k < n; k++) {
            Input &f = in[i0 + k];
            files[k] = f.bytes.data(); lens[k] = f.bytes.size();
            outs[k] = f.parse == BJ_OK ? pin + o : nullptr;
            o += (f.out_bytes + 15) / 16 * 16;
        }
        rc = bj_decode_batch(ctx, files.data(), lens.data(), n, BJ_OUT_BMP, outs.data(), status.data());
        if (rc != BJ_OK) {
            fprintf(stderr, "decoder_b200: bj_decode_batch: %s (%s)\n", bj_status_string(rc), bj_last_error(ctx));
            return 2;
        }
        calls++;
        t_decode += now_s() - t0;

        // ---- write.  Like the reference, a scan that fails to decode still produces its (partial) image
        // (the result of decode_Huffman_data is ignored, src/decoder_host.cpp:181).
        t0 = now_s();
        for (int k = 0; k < n; k++) {
            Input &f = in[i0 + k];
            if (status[k] != BJ_OK && status[k] != BJ_ERR_CORRUPT_SCAN) {
                printf("%s: Error - Invalid JPEG\n", f.path.c_str());
                failures++;
            } else {
                FILE *fp = fopen(bmp_name(f.path).c_str(), "wb");
                if (!fp || fwrite(outs[k], 1, f.out_bytes, fp) != f.out_bytes) { printf("%s: Error - cannot write BMP\n", f.path.c_str()); failures++; }
                if (fp) fclose(fp);
            }
            f.bytes.clear(); f.bytes.shrink_to_fit();
        }
        t_write += now_s() - t0;
        i0 = i1;
    }
    if (pin) bj_host_free(pin);
    bj_destroy(ctx);

    printf("\nProfiles:\n");
    printf("End-to-end execution time: %gs\n", now_s() - t_start);
    printf(" - File read + header parse time: %gs\n", t_read);
    printf(" - B200 decode time (H2D + kernels + D2H): %gs\n", t_decode);
    printf(" - BMP write time: %gs\n", t_write);
    printf(" - Total %d calls, %zu images, %d failed\n", calls, in.size(), failures);
    return 0;
}
