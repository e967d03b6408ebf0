// stage1_split.cuh -- stage 1 as two launches for large documents: classify, then flatten.
//
// The fused kernels (stage1_persistent.cuh, stage1_dataflow.cuh) are bound by latency, not by a pipe: a warp that
// classifies 2 KiB and then extracts its indexes spends long stretches waiting (bulk copy, look-back, the serial
// clear-lowest-bit chain of the extraction loop), and registers cap the SM at ~34 such warps.  Measured standalone
// (tools/ubench_stages.cu) both halves only reach their issue-bound rate with >= 32 warps per SM *each*.  So for
// documents large enough to amortise a second launch the two halves run as separate grids, each at the occupancy its own
// register budget allows, and neither ever waits for a carry:
//
//   stage1_classify_kernel : persistent CTAs of NW compute warps + 1 scan warp, tiles drawn from the ticket counter
//                            exactly as in the fused kernel.  A compute warp turns 2 KiB into the two structural masks
//                            (string state entering the chunk unknown -> one mask per parity), stores both (16 bytes per
//                            64 input bytes) and moves on.  The scan warp runs the decoupled look-back for every tile
//                            and leaves, per 2 KiB chunk, one 8-byte carry word: bit 63 = chunk starts inside a string,
//                            bits 0..39 = rank of its first structural in the output.  It also writes the verdict.
//   stage1_flatten_kernel  : one warp per chunk, no shared state between warps: carry word -> pick the mask plane ->
//                            popcount, warp scan, extract into the warp's staging area, 16-byte stores.
//
// Extra HBM traffic: the mask planes, 16 B written + 8 B read per 64 input bytes (+23 % over the algorithmic bytes at the
// bench document's density of 0.161 structurals per byte).
// Same arithmetic, same descriptors, same results as the fused kernels (reference json_structural_indexer.mojo:83-186).
#pragma once
#include "stage1_persistent.cuh"

#ifndef SJ_SPLITREG
#define SJ_SPLITREG 56  // registers per thread of the classify kernel
#endif

namespace sjb200 {

#if defined(__CUDACC__)

constexpr uint64_t CARRY_INSIDE = 1ull << 63;
constexpr uint64_t CARRY_RANK_MASK = (1ull << 40) - 1;

template <int NW>
struct SplitCfg {
    static constexpr int THREADS = (NW + 1) * 32;
    static constexpr int TILE = NW * 2048;
    static constexpr int NIN = 2;                                         // input buffers (bulk copy of tile i+2 overlaps tile i+1)
    static constexpr int IN_STRIDE = ((16 + TILE) + 127) & ~127;          // halo + tile
    static constexpr int SMEM_BYTES = NIN * IN_STRIDE;
    static constexpr int MAXREG = SJ_SPLITREG;
};

struct SplitSlot {                // hand-off from the compute warps to the scan warp
    uint32_t wc0[32], wc1[32], wflags[32];
    uint32_t R[32], off0[32], off1[32];
    uint64_t agg;
    uint32_t tail;
    uint32_t arrived;
    int32_t tile;
};

template <int NW, bool UTF8>
__global__ void __launch_bounds__((NW + 1) * 32) __maxnreg__(SplitCfg<NW>::MAXREG) stage1_classify_kernel(const Stage1Params P) {
    using Cfg = SplitCfg<NW>;
    constexpr int TILE = Cfg::TILE;
    constexpr int NIN = Cfg::NIN;
    constexpr int NS = 4;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[NIN + 2 * NS];   // in_full[NIN], sum_full[NS], slot_free[NS]
    __shared__ SplitSlot s_slot[NS];
    __shared__ int32_t s_tile_of[NIN];
    __shared__ uint32_t s_loaded[NIN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_in = smem_u32(&s_bar[0]), bar_sum = smem_u32(&s_bar[NIN]), bar_free = smem_u32(&s_bar[NIN + NS]);

    if (tid == 0) {
        for (int k = 0; k < NIN; k++) {
            mbar_init(bar_in + 8 * k, 1);
            s_loaded[k] = 0;
        }
        for (int k = 0; k < NS; k++) {
            mbar_init(bar_sum + 8 * k, 1);
            mbar_init(bar_free + 8 * k, 1);
            s_slot[k].arrived = 0;
        }
        fence_mbar_init();
        if (blockIdx.x == 0) P.ticket[(P.ticket_sel + 1) & 1u] = 0;   // the counter the NEXT launch will use
    }
    __syncthreads();
    uint32_t *ticket = P.ticket + (P.ticket_sel & 1u);

    // draw the ticket of iteration `it` and start its bulk copy into input buffer it % NIN (free by construction)
    auto produce = [&](int it) {
        const int b = it % NIN;
        const uint32_t k = P.tile_begin + atomicAdd(ticket, 1u);
        if (k < P.tile_end) {
            const int t = (int)k;
            s_tile_of[b] = t;
            const int64_t tb = (int64_t)t * TILE;
            int64_t nbytes = (int64_t)P.alen - tb;
            nbytes = nbytes > TILE ? TILE : nbytes;
            nbytes = (nbytes + 15) & ~15ll;
            const uint32_t halo = t > 0 ? 16u : 0u;
            mbar_expect_tx(bar_in + 8 * b, (uint32_t)nbytes + halo);
            bulk_load(smem_u32(smem_raw) + b * Cfg::IN_STRIDE + 16u - halo, P.abase + tb - halo, (uint32_t)nbytes + halo, bar_in + 8 * b);
        } else {
            s_tile_of[b] = -1;
            mbar_arrive(bar_in + 8 * b);
        }
    };
    if (tid == 0) {
        for (int k = 0; k < NIN; k++) produce(k);
    }

    if (warp == NW) {
        // =============================== scan warp ===============================
        for (int i = 0;; i++) {
            const int slot = i & (NS - 1);
            mbar_wait(bar_sum + 8 * slot, (uint32_t)(i / NS) & 1u);
            SplitSlot &S = s_slot[slot];
            const int cur = *reinterpret_cast<volatile int32_t *>(&S.tile);
            if (cur < 0) break;
            const TileAgg agg = desc_unpack_agg(S.agg);
            const bool have = lane < NW;
            const uint32_t R = have ? S.R[lane] : 0u, off0 = have ? S.off0[lane] : 0u, off1 = have ? S.off1[lane] : 0u;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8 * slot);   // everything we need is in registers: the slot may be reused
            LookbackResult lb = {0, 0, 0};
            if (cur > 0) lb = lookback(P.desc, P.gen, cur, lane);
            const uint32_t s_in = lb.s_in & 1u;
            const uint32_t total = s_in ? agg.c[1] : agg.c[0];
            TilePrefix pre;
            pre.s_out = s_in ^ agg.par;
            pre.e_out = agg.e_out;
            pre.p_out = agg.p_out;
            pre.err = lb.err | ((s_in ? agg.un[1] : agg.un[0]) ? EF_UNESCAPED : 0u) | (agg.u8 ? EF_UTF8 : 0u);
            pre.count = lb.base + total;
            if (lane == 0) {
                st_desc(P.desc + cur, desc_pack_prefix(P.gen, pre));
                if (cur == (int)P.tile_end - 1 && P.progress) *P.progress = pre.count;
                if (cur == (int)P.ntiles - 1) write_verdict(P, pre);
            }
            if (have) {
                const uint64_t first = (uint64_t)lb.base + (s_in ? off1 : off0);
                P.carry[(size_t)cur * NW + lane] = first | (((s_in ^ R) & 1u) ? CARRY_INSIDE : 0ull);
            }
        }
    } else {
        // =============================== compute warps ===============================
        int i = 0;
        while (true) {
            const int b = i % NIN;
            mbar_wait(bar_in + 8 * b, (uint32_t)(i / NIN) & 1u);
            const int tile = *reinterpret_cast<volatile int32_t *>(&s_tile_of[b]);
            if (tile < 0) break;
            const int slot = i & (NS - 1);
            const int64_t tb = (int64_t)tile * TILE;
            LanePhase1 ph;
            {
                LaneInput in;
                warp_load<UTF8>(in, smem_raw + b * Cfg::IN_STRIDE + 16, warp, lane, tile, tb, TILE, P);
                __syncwarp();
                if (lane == 0) {                            // this warp no longer needs the input buffer
                    __threadfence_block();
                    if (atomicAdd(&s_loaded[b], 1u) == NW - 1) {
                        s_loaded[b] = 0;                    // everyone has it in registers: refill the buffer
                        __threadfence_block();
                        produce(i + NIN);
                    }
                }
                warp_compute<UTF8>(ph, in, lane, P);
            }
            {   // both mask planes of the chunk: [chunk][parity][lane], 256 contiguous bytes per plane
                uint64_t *mp = P.masks + ((size_t)tile * NW + warp) * 64 + lane;
                __stcs(reinterpret_cast<unsigned long long *>(mp), (unsigned long long)ph.m0);
                __stcs(reinterpret_cast<unsigned long long *>(mp + 32), (unsigned long long)ph.m1);
            }
            if (i >= NS) mbar_wait(bar_free + 8 * slot, (uint32_t)(i / NS - 1) & 1u);   // the scan warp is done with iteration i - NS
            SplitSlot &S = s_slot[slot];
            uint32_t order = 0;
            if (lane == 0) {
                S.wc0[warp] = ph.wc0;
                S.wc1[warp] = ph.wc1;
                S.wflags[warp] = ph.wflags;
                if (warp == NW - 1) S.tail = ph.tail;
                __threadfence_block();
                order = atomicAdd(&S.arrived, 1u);
            }
            order = __shfl_sync(0xFFFFFFFFu, order, 0);
            if (order == NW - 1) {
                __threadfence_block();
                const bool have = lane < NW;
                uint32_t R, off0, off1;
                const TileAgg agg = tile_aggregate(have ? S.wflags[lane] : 0u, have ? S.wc0[lane] : 0u, have ? S.wc1[lane] : 0u,
                                                   S.tail, NW, lane, R, off0, off1);
                if (have) {
                    S.R[lane] = R;
                    S.off0[lane] = off0;
                    S.off1[lane] = off1;
                }
                const uint64_t packed = desc_pack_agg(P.gen, agg);
                if (lane == 0) {
                    if (tile > 0) st_desc(P.desc + tile, packed);
                    S.agg = packed;
                    S.tile = tile;
                    S.arrived = 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sum + 8 * slot);
            }
            i++;
        }
        if (warp == 0 && lane == 0) {   // tell the scan warp that iteration i does not exist
            const int slot = i & (NS - 1);
            if (i >= NS) mbar_wait(bar_free + 8 * slot, (uint32_t)(i / NS - 1) & 1u);
            s_slot[slot].tile = -1;
            mbar_arrive(bar_sum + 8 * slot);
        }
    }
}

#ifndef SJ_K3_MINCTAS
#define SJ_K3_MINCTAS 8   // resident CTAs per SM the flatten kernel's register budget allows (8 -> 32 registers)
#endif

template <int FW>
struct FlattenCfg {
    static constexpr int THREADS = FW * 32;
    static constexpr int WCAP = 512;
    static constexpr int SMEM_BYTES = FW * (WCAP + 4) * 4;
};

// one warp flattens chunk c: carry word -> plane -> popcount, warp scan, extraction into `stage` (the warp's staging area
// of WCAP + 4 entries in shared memory), 16-byte stores.  BitIndexer.write, reference json_structural_indexer.mojo:46-58.
template <int WCAP>
__device__ __forceinline__ void flatten_chunk(const Stage1Params &P, uint32_t c, uint64_t carry, uint64_t structural, uint32_t *stage, int lane) {
    const uint64_t first = carry & CARRY_RANK_MASK;
    const uint32_t lo = (uint32_t)structural, hi = (uint32_t)(structural >> 32);
    const uint32_t cnt_lo = (uint32_t)__popc(lo), cnt = cnt_lo + (uint32_t)__popc(hi);
    const uint32_t incl = warp_inclusive_sum(cnt);
    const uint32_t wtotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t v0 = c * 2048u + (uint32_t)lane * 64u - P.mis;
    if (wtotal <= (uint32_t)WCAP) {
        const uint32_t a = ((uint32_t)first + out_phase(P.out)) & 3u;
        const uint32_t sp = smem_u32(stage) + 4u * (a + (incl - cnt));
        flatten_word_pair(sp, lo, v0 + 31u);
        flatten_word_pair(sp + 4u * cnt_lo, hi, v0 + 63u);   // (the popcount of the low word is needed for the scan anyway)
        __syncwarp();
        copy_out_warp(stage, a, wtotal, P.out, first, P.cap, (uint32_t)lane);
    } else {
        flatten_direct(P.out, P.cap, first + (incl - cnt), structural, v0);
    }
}

#ifndef SJ_K3_CPW
#define SJ_K3_CPW 2   // consecutive chunks per warp of the flatten kernel: carries and mask words of all of them are requested up front
#endif

// chunks [chunk_begin, chunk_end): SJ_K3_CPW consecutive chunks per warp
template <int FW>
__global__ void __launch_bounds__(FW * 32, SJ_K3_MINCTAS * 8 / FW) stage1_flatten_kernel(const Stage1Params P, uint32_t chunk_begin, uint32_t chunk_end) {
    using Cfg = FlattenCfg<FW>;
    constexpr uint32_t CPW = SJ_K3_CPW;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw) + warp * (Cfg::WCAP + 4);
    const uint32_t c0 = chunk_begin + (blockIdx.x * FW + warp) * CPW;
    grid_dependency_wait();
    if (c0 >= chunk_end) return;
    const uint32_t gave_up = P.spec_flag ? __ldcg(P.spec_flag) : 0u;   // stream pipeline only; loaded together with the carries
    uint64_t carry[CPW], structural[CPW];
#pragma unroll
    for (uint32_t j = 0; j < CPW; j++)
        carry[j] = c0 + j < chunk_end ? __ldcg(reinterpret_cast<const unsigned long long *>(P.carry + c0 + j)) : 0ull;
    if (P.spec_flag && gave_up == P.gen) return;
#pragma unroll
    for (uint32_t j = 0; j < CPW; j++) {
        const uint32_t s_w = (uint32_t)(carry[j] >> 63);
        structural[j] = c0 + j < chunk_end ? __ldcs(reinterpret_cast<const unsigned long long *>(P.masks + (size_t)(c0 + j) * 64 + s_w * 32 + lane)) : 0ull;
    }
#pragma unroll
    for (uint32_t j = 0; j < CPW; j++) {
        if (c0 + j < chunk_end) {
            flatten_chunk<Cfg::WCAP>(P, c0 + j, carry[j], structural[j], stage, lane);
            if (j + 1 < CPW) __syncwarp();   // the staging area is reused by the next chunk
        }
    }
}

#endif  // __CUDACC__

}  // namespace sjb200
