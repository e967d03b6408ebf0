// batch_test.cpp -- the multi-GPU batch driver of libsimdjson_b200.so driven from C++ alone (no Python, no torch):
// what a Mojo host would do through FFI.  One process, G GPUs, an NDJSON batch in host memory:
//   sjb200_batch_create(G) -> sjb200_batch_run(host batch) -> every segment compared with the CPU oracle on the same byte
//   range (indexes, trailer, n, verdict), twice (the two row-buffer sets alternate), plus a batch with one broken segment.
// Then the device-resident path on the same contexts: plan_resident / run_resident_async x3 / finish.
// The oracle (oracle/liboracle_stage1.so) is the checker only.
//
// usage: batch_test [gpus] [MiB] [segments per gpu]      (defaults: min(2, devices), 64, 2)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/simdjson_b200.h"

extern "C" {
int sjb200_gen_ndjson(uint8_t *out, uint64_t size, uint64_t seed);
int32_t oracle_stage1_fast(const uint8_t *buf, uint64_t len, uint32_t *idx, uint64_t cap, uint32_t *n_out, uint64_t *n_written, uint32_t flags);
}

#define REQUIRE(cond, ...)                          \
    do {                                            \
        if (!(cond)) {                              \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);           \
            fprintf(stderr, "\n");                  \
            return 1;                               \
        }                                           \
    } while (0)

static int check_segments(const uint8_t *buf, const uint32_t *idx, const uint64_t *seg_off, const uint64_t *seg_idx, const uint32_t *seg_n,
                          const int32_t *seg_err, uint32_t nseg, int32_t worst) {
    int32_t want_worst = 0;
    std::vector<uint32_t> ref;
    for (uint32_t s = 0; s < nseg; s++) {
        const uint64_t a = seg_off[s], e = seg_off[s + 1];
        REQUIRE(e > a && buf[e - 1] == '\n', "segment %u does not end on a line end", s);
        ref.assign(e - a + 8, 0);
        uint32_t n = 0xFFFFFFFFu;
        uint64_t nw = 0;
        const int32_t err = oracle_stage1_fast(buf + a, e - a, ref.data(), ref.size(), &n, &nw, 0);
        want_worst = err > want_worst ? err : want_worst;
        REQUIRE(seg_err[s] == err, "segment %u: verdict %d, oracle %d", s, seg_err[s], err);
        if (err == 0) {
            REQUIRE(seg_n[s] == n, "segment %u: n %u, oracle %u", s, seg_n[s], n);
            REQUIRE(seg_idx[s] == a + 3ull * s, "segment %u: index offset", s);
            REQUIRE(memcmp(idx + seg_idx[s], ref.data(), (size_t)(n + 3) * 4) == 0, "segment %u: index stream differs from the oracle", s);
        }
    }
    REQUIRE(worst == want_worst, "global verdict %d, expected %d", worst, want_worst);
    return 0;
}

int main(int argc, char **argv) {
    int devices = sjb200_device_count();
    REQUIRE(devices >= 1, "no CUDA device: the stage-1 path has no CPU fallback");
    const int G = argc > 1 ? atoi(argv[1]) : (devices >= 2 ? 2 : 1);
    const uint64_t size = (uint64_t)(argc > 2 ? atoi(argv[2]) : 64) << 20;
    const uint32_t per_gpu = argc > 3 ? (uint32_t)atoi(argv[3]) : 2u;
    REQUIRE(G >= 1 && G <= devices, "asked for %d GPUs, %d present", G, devices);
    const uint32_t max_seg = 16, max_total = max_seg * (uint32_t)G;
    uint8_t *buf = nullptr;
    uint32_t *idx = nullptr;
    REQUIRE(sjb200_pinned_alloc(size, (void **)&buf) == 0 && sjb200_pinned_alloc((size + 3ull * max_total) * 4, (void **)&idx) == 0, "pinned allocation");
    REQUIRE(sjb200_gen_ndjson(buf, size, 0x5EED0003ull) == 0, "generator");
    const uint64_t shard = (size + G - 1) / G, seg_bytes = shard / per_gpu;
    sjb200_batch *b = nullptr;
    int32_t rc = sjb200_batch_create(G, nullptr, shard + (4u << 20), seg_bytes, max_seg, 0, &b);
    REQUIRE(rc == 0 && b, "sjb200_batch_create(%d) -> %d", G, rc);
    REQUIRE(sjb200_batch_local_gpus(b) == G, "local gpus");
    std::vector<uint64_t> seg_off(max_total + 1), seg_idx(max_total + 1);
    std::vector<uint32_t> seg_n(max_total);
    std::vector<int32_t> seg_err(max_total);
    uint32_t nseg = 0;
    int32_t worst = -1;
    for (int pass = 0; pass < 3; pass++) {
        memset(idx, 0xFF, (size + 3ull * max_total) * 4);
        rc = sjb200_batch_run(b, buf, size, idx, size + 3ull * max_total, seg_off.data(), seg_idx.data(), seg_n.data(), seg_err.data(), max_total, &nseg, &worst, 0);
        REQUIRE(rc == 0, "sjb200_batch_run -> %d", rc);
        REQUIRE(nseg >= (uint32_t)G * per_gpu && seg_off[0] == 0 && seg_off[nseg] == size, "segments: %u", nseg);
        if (check_segments(buf, idx, seg_off.data(), seg_idx.data(), seg_n.data(), seg_err.data(), nseg, worst)) return 1;
    }
    printf("host batch: %d GPU(s), %u segments of ~%llu KiB, 3 passes bit-exact with the oracle, global verdict %d\n", G, nseg,
           (unsigned long long)(seg_bytes >> 10), worst);
    // one broken segment (the first quote of the last one blanked: it now ends inside a string): its verdict, the global
    // verdict, and the others untouched
    uint64_t hit = seg_off[nseg - 1];
    while (buf[hit] != '"') hit++;
    const uint8_t saved = buf[hit];
    buf[hit] = ' ';
    rc = sjb200_batch_run(b, buf, size, idx, size + 3ull * max_total, seg_off.data(), seg_idx.data(), seg_n.data(), seg_err.data(), max_total, &nseg, &worst, 0);
    REQUIRE(rc == 0, "sjb200_batch_run (broken) -> %d", rc);
    if (check_segments(buf, idx, seg_off.data(), seg_idx.data(), seg_n.data(), seg_err.data(), nseg, worst)) return 1;
    REQUIRE(seg_err[nseg - 1] != 0 && worst == seg_err[nseg - 1], "the broken segment must fail (got %d, global %d)", seg_err[nseg - 1], worst);
    for (uint32_t s = 0; s + 1 < nseg; s++) REQUIRE(seg_err[s] == 0, "segment %u must be unaffected", s);
    printf("broken last segment: verdict %d on it, global verdict %d, the other %u segments unchanged\n", seg_err[nseg - 1], worst, nseg - 1);
    buf[hit] = saved;

    // device-resident shards on the same driver: every GPU gets its own copy of a shard, plans it, three passes, one finish
    std::vector<uint8_t *> d_shard(G);
    std::vector<uint32_t *> d_idx(G);
    std::vector<uint64_t> caps(G);
    std::vector<std::vector<uint64_t>> offs(G), ioffs(G);
    std::vector<uint32_t> nsegs(G);
    const uint64_t lo = 0, hi = seg_off[per_gpu];   // the first GPU's shard of the host run (whole lines)
    for (int g = 0; g < G; g++) {
        REQUIRE(cudaSetDevice(g) == cudaSuccess, "cudaSetDevice");
        REQUIRE(cudaMalloc(&d_shard[g], hi - lo) == cudaSuccess, "cudaMalloc");
        REQUIRE(cudaMemcpy(d_shard[g], buf + lo, hi - lo, cudaMemcpyHostToDevice) == cudaSuccess, "copy");
        offs[g].resize(max_seg + 1);
        ioffs[g].resize(max_seg + 1);
        rc = sjb200_batch_plan_resident(b, g, d_shard[g], hi - lo, offs[g].data(), ioffs[g].data(), &nsegs[g]);
        REQUIRE(rc == 0 && nsegs[g] >= 1, "plan_resident -> %d", rc);
        caps[g] = ioffs[g][nsegs[g]];
        REQUIRE(cudaMalloc(&d_idx[g], caps[g] * 4) == cudaSuccess, "cudaMalloc");
    }
    for (int pass = 0; pass < 3; pass++) {
        rc = sjb200_batch_run_resident_async(b, d_idx.data(), caps.data(), 0);
        REQUIRE(rc == 0, "run_resident_async -> %d", rc);
    }
    std::vector<int32_t> rows((size_t)G * max_seg * 2);
    rc = sjb200_batch_finish(b, rows.data(), &worst);
    REQUIRE(rc == 0 && worst == 0, "finish -> %d, worst %d", rc, worst);
    std::vector<uint32_t> got, ref;
    for (int g = 0; g < G; g++) {
        REQUIRE(cudaSetDevice(g) == cudaSuccess, "cudaSetDevice");
        got.resize(caps[g]);
        REQUIRE(cudaMemcpy(got.data(), d_idx[g], caps[g] * 4, cudaMemcpyDeviceToHost) == cudaSuccess, "copy back");
        for (uint32_t s = 0; s < max_seg; s++) {
            const int32_t err = rows[((size_t)g * max_seg + s) * 2], n = rows[((size_t)g * max_seg + s) * 2 + 1];
            if (s >= nsegs[g]) {
                REQUIRE(err == -1 && n == -1, "GPU %d row %u should read 'no such segment'", g, s);
                continue;
            }
            const uint64_t a = offs[g][s], e = offs[g][s + 1];
            ref.assign(e - a + 8, 0);
            uint32_t wn = 0;
            uint64_t nw = 0;
            const int32_t werr = oracle_stage1_fast(buf + lo + a, e - a, ref.data(), ref.size(), &wn, &nw, 0);
            REQUIRE(err == werr && (uint32_t)n == wn, "GPU %d segment %u: {%d, %d}, oracle {%d, %u}", g, s, err, n, werr, wn);
            REQUIRE(memcmp(got.data() + ioffs[g][s], ref.data(), (size_t)(wn + 3) * 4) == 0, "GPU %d segment %u: index stream differs", g, s);
        }
        cudaFree(d_shard[g]);
        cudaFree(d_idx[g]);
    }
    printf("resident shards: %d GPU(s) x %u segment(s), 3 passes, gathered rows and index streams bit-exact with the oracle\n", G, nsegs[0]);
    REQUIRE(sjb200_batch_destroy(b) == 0, "destroy");
    sjb200_pinned_free(buf);
    sjb200_pinned_free(idx);
    printf("PASS\n");
    return 0;
}
