// TEST INFRASTRUCTURE - exercises the product's host worker pool (pim_jpeg_decoder_b200/csrc/bj_host.h, bj::HostPool)
// without a GPU: only the header is compiled (no CUDA call is made).
#include <atomic>
#include <vector>
#include "../../pim_jpeg_decoder_b200/csrc/bj_host.h"

// Runs `rounds` parallel_for calls over n items with `threads` threads; every item must be visited exactly once per
// round.  Returns the number of violations.
extern "C" int emu_hostpool(int threads, int n, int chunk, int rounds) {
    bj::HostPool pool;
    pool.resize(threads);
    if (pool.threads() != (threads < 1 ? 1 : threads)) return -1;
    int bad = 0;
    for (int r = 0; r < rounds; r++) {
        std::vector<std::atomic<int>> hits(n > 0 ? n : 1);
        for (auto &h : hits) h.store(0);
        pool.parallel_for(n, chunk, [&](int b, int e) { for (int i = b; i < e; i++) hits[i].fetch_add(1); });
        for (int i = 0; i < n; i++) if (hits[i].load() != 1) bad++;
    }
    return bad;
}
