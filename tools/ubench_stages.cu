// tools/ubench_stages.cu -- standalone throughput of the two stages of the stage-1 kernel (no look-back, no tickets):
//   classify : 2 KiB per warp from shared memory -> dual structural masks (warp_compute)
//   flatten  : masks -> indexes staged in shared memory -> 16-byte global stores (flatten_to + copy_out)
// Tells how far the full kernel is from the sum of its parts.  Build with the same flags as the library.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../mojo_simdjson_b200/csrc/stage1_kernel.cuh"
using namespace sjb200;

template <int WARPS, bool UTF8>
__global__ void __launch_bounds__(WARPS * 32) k_classify(const uint8_t *in, uint64_t nchunks_per_warp, uint64_t total_bytes, uint4 *masks, Stage1Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *mine = smem + 16 + warp * 2048;
    const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp;
    uint64_t acc0 = 0, acc1 = 0;
    for (uint64_t it = 0; it < nchunks_per_warp; it++) {
        const uint64_t chunk = (gw * nchunks_per_warp + it) % (total_bytes / 2048);
        // plain coalesced copy of the warp's 2 KiB into its shared-memory slice (stands in for the bulk copy)
        const uint4 *src = reinterpret_cast<const uint4 *>(in + chunk * 2048);
        uint4 *dst = reinterpret_cast<uint4 *>(mine);
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q * 32 + lane] = src[q * 32 + lane];
        __syncwarp();
        LaneInput li;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(mine + lane * 64);
#pragma unroll
        for (int q = 0; q < 4; q++) { const uint4 v = s4[q]; li.w[4*q] = v.x; li.w[4*q+1] = v.y; li.w[4*q+2] = v.z; li.w[4*q+3] = v.w; }
        li.prev = *reinterpret_cast<const uint32_t *>(mine + lane * 64 - 4);
        li.wst.e = 0; li.wst.p = 0; li.wst.unresolved = 0; li.ends = 0;
        li.g0 = 64 + (int64_t)chunk * 2048 + lane * 64;
        __syncwarp();
        LanePhase1 ph;
        warp_compute<UTF8>(ph, li, lane, P);
        acc0 ^= ph.m0; acc1 ^= ph.m1 + ph.wc0;
    }
    masks[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = make_uint4((uint32_t)acc0, (uint32_t)(acc0 >> 32), (uint32_t)acc1, (uint32_t)(acc1 >> 32));
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_flatten(const uint4 *masks, uint64_t nmask_chunks, uint64_t nchunks_per_warp, uint32_t *out, uint64_t cap) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem) + warp * 516;
    const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp;
    uint64_t first = gw * nchunks_per_warp * 340 % (cap - 4096);
    for (uint64_t it = 0; it < nchunks_per_warp; it++) {
        const uint64_t chunk = (gw * nchunks_per_warp + it) % nmask_chunks;
        const uint4 mm = masks[chunk * 32 + lane];
        const uint64_t structural = join64(mm.x, mm.y);
        const uint32_t cnt = (uint32_t)__popcll(structural);
        const uint32_t incl = warp_inclusive_sum(cnt);
        const uint32_t wtotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t v0 = (uint32_t)(chunk * 2048 + lane * 64);
        if (wtotal <= 512) {
            const uint32_t a = ((uint32_t)first + out_phase(out)) & 3u;
            flatten_to(stage + a + (incl - cnt), structural, v0);
            __syncwarp();
            copy_out(stage, a, wtotal, out, first, cap, (uint32_t)lane, 32u);
            __syncwarp();
        }
        first += wtotal;
        if (first + 1024 > cap) first = 0;
    }
}

// real masks: run the oracle-free classifier once over real text to get realistic structural masks
template <bool UTF8>
__global__ void k_make_masks(const uint8_t *in, uint64_t nchunks, uint4 *masks, Stage1Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *mine = smem + 16 + warp * 2048;
    const uint64_t chunk = (uint64_t)blockIdx.x * (blockDim.x / 32) + warp;
    if (chunk >= nchunks) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(in + chunk * 2048);
    uint4 *dst = reinterpret_cast<uint4 *>(mine);
    for (int q = 0; q < 4; q++) dst[q * 32 + lane] = src[q * 32 + lane];
    __syncwarp();
    LaneInput li;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(mine + lane * 64);
    for (int q = 0; q < 4; q++) { const uint4 v = s4[q]; li.w[4*q] = v.x; li.w[4*q+1] = v.y; li.w[4*q+2] = v.z; li.w[4*q+3] = v.w; }
    li.prev = 0x20202020u; li.wst.e = 0; li.wst.p = 0; li.wst.unresolved = 0; li.ends = 0; li.g0 = 64 + (int64_t)chunk * 2048 + lane * 64;
    LanePhase1 ph;
    warp_compute<UTF8>(ph, li, lane, P);
    masks[chunk * 32 + lane] = make_uint4((uint32_t)ph.m0, (uint32_t)(ph.m0 >> 32), (uint32_t)ph.m1, (uint32_t)(ph.m1 >> 32));
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "gpurun_out/doc64m.bin";
    FILE *f = fopen(path, "rb");
    if (!f) { printf("cannot open %s\n", path); return 1; }
    std::vector<uint8_t> h(64 << 20);
    size_t n = fread(h.data(), 1, h.size(), f); fclose(f);
    n &= ~(size_t)2047;
    uint8_t *d_in; cudaMalloc(&d_in, n + 4096); cudaMemcpy(d_in + 2048, h.data(), n, cudaMemcpyHostToDevice);
    uint4 *d_masks; cudaMalloc(&d_masks, (n / 2048) * 32 * 16);
    uint32_t *d_out; const uint64_t cap = 64 << 20; cudaMalloc(&d_out, cap * 4);
    uint4 *d_sink; cudaMalloc(&d_sink, 16 * 2048 * 1024);
    Stage1Params P; memset(&P, 0, sizeof P); P.alen = 1ull << 40; P.mis = 0;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_make_masks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 + 8 * 2048 + 64);
    k_make_masks<true><<<(unsigned)((n / 2048 + 7) / 8), 256, 16 + 8 * 2048 + 64>>>(d_in + 2048, n / 2048, d_masks, P);
    cudaDeviceSynchronize();
    printf("input %zu bytes, %zu chunks; err=%s\n", n, n / 2048, cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *name, int ctas_per_sm, int warps, float ms, uint64_t chunks) {
        double clk_per_chunk_per_sm = ms * 1e-3 * 1.965e9 * sms / chunks;
        printf("%-22s %2d CTAs/SM x %2d warps: %.3f ms, %.0f clk per 2KiB chunk per SM, %.0f GB/s equivalent\n", name, ctas_per_sm, warps, ms,
               clk_per_chunk_per_sm, chunks * 2048.0 / (ms * 1e-3) / 1e9);
    };
    const uint64_t per_warp = 64;
#define RUN_CLASSIFY(W, U, CPS)                                                                                     \
    {                                                                                                                   \
        const int smem = 16 + W * 2048 + 64;                                                                           \
        cudaFuncSetAttribute(k_classify<W, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                     \
        k_classify<W, U><<<sms * CPS, W * 32, smem>>>(d_in + 2048, per_warp, n, d_sink, P);                            \
        cudaEventRecord(e0);                                                                                            \
        k_classify<W, U><<<sms * CPS, W * 32, smem>>>(d_in + 2048, per_warp, n, d_sink, P);                            \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                                                                  \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                                    \
        report(U ? "classify (utf8)" : "classify (no utf8)", CPS, W, ms, (uint64_t)sms * CPS * W * per_warp);          \
    }
    RUN_CLASSIFY(8, true, 2) RUN_CLASSIFY(8, true, 4) RUN_CLASSIFY(8, true, 6) RUN_CLASSIFY(16, true, 3) RUN_CLASSIFY(8, false, 4) RUN_CLASSIFY(8, false, 6)
#define RUN_FLATTEN(W, CPS)                                                                                         \
    {                                                                                                                   \
        const int smem = W * 516 * 4;                                                                                   \
        k_flatten<W><<<sms * CPS, W * 32, smem>>>(d_masks, n / 2048, per_warp, d_out, cap);                            \
        cudaEventRecord(e0);                                                                                            \
        k_flatten<W><<<sms * CPS, W * 32, smem>>>(d_masks, n / 2048, per_warp, d_out, cap);                            \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                                                                  \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                                    \
        report("flatten", CPS, W, ms, (uint64_t)sms * CPS * W * per_warp);                                             \
    }
    RUN_FLATTEN(8, 2) RUN_FLATTEN(8, 4) RUN_FLATTEN(8, 8) RUN_FLATTEN(16, 4)
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
