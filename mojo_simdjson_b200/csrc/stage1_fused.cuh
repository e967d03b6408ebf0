// stage1_fused.cuh -- stage 1 of a device-resident document as ONE persistent launch in which every warp alternates
// between the two halves of the work and never waits for another warp in steady state:
//
//   classify : a run of RUN consecutive 2 KiB chunks drawn from an atomic counter, each fetched with its own bulk copy
//              (cp.async.bulk + mbarrier, DEPTH in flight per warp) -> both structural mask planes and a 16-byte chunk
//              summary to L2 (the string state entering the chunk is not known yet: one plane per parity).  Lanes with
//              bytes >= 0x80 park their bit planes in the warp's shared-memory slots; the warp validates them, 32 lanes
//              at a time, at the end of the run (no global traffic, no extra kernel).
//   scan     : chunks are grouped in blocks of BLOCK_CHUNKS.  The warp that completes the last run of a block scans the
//              block's summaries (ordered span_concat), publishes the block aggregate, runs the decoupled look-back over
//              the earlier blocks' descriptors (stage1_kernel.cuh: lookback), publishes the inclusive prefix and leaves
//              one carry word per chunk (bit 63 = starts inside a string, bits 0..39 = rank of its first index).
//   flatten  : after every classify run the warp claims a group of RUN chunks from a second counter and, as soon as the
//              block they belong to is scanned, turns their mask words into indexes (flatten_chunk, stage1_split.cuh).
//              Flatten work trails classify work by a block or two, so the masks are read back from L2.
//
// The ALU-bound classify half and the shared-memory / XU-bound flatten half thus share every SM at all times instead of
// running as two kernels one after the other.  Waiting happens only at the very end (classify work exhausted, the last
// blocks not scanned yet).  Everything is handed out by tickets, so no CTA depends on a CTA that is not resident.
//
// Escape / scalar carries entering a chunk come from a 32-byte look-behind (stage1_stream.cuh: chunk_load); if a backslash
// run covers all of it the warp walks back through global memory (up to WALK_MAX bytes, backslash_run_global).  Only a run
// longer than that raises `spec_flag`, and the persistent kernel enqueued behind this one redoes the document.
// Reference: json_structural_indexer.mojo:83-186 (step / next / finish), restated in oracle/stage1_oracle.c.
#pragma once
#include "stage1_stream.cuh"

#ifndef SJ_FUSED_NW
#define SJ_FUSED_NW 8
#endif
#ifndef SJ_FUSED_DEPTH
#define SJ_FUSED_DEPTH 2
#endif
#ifndef SJ_FUSED_REG
#define SJ_FUSED_REG 56
#endif
#ifndef SJ_FUSED_ABL
#define SJ_FUSED_ABL 0     // timing ablations (results are wrong): 1 = claim but do not flatten, 2 = no fence per run, 4 = classify stores nothing
#endif
#ifndef SJ_FUSED_LAG
#define SJ_FUSED_LAG 1     // flatten groups a warp may take per classify run once it has fallen behind
#endif

namespace sjb200 {

#if defined(__CUDACC__)

constexpr uint32_t BLOCK_CHUNKS = 64;        // chunks per look-back block (128 KiB of input; counts fit the 18-bit descriptor fields)

template <int NW>
struct FusedCfg {
    static constexpr int THREADS = NW * 32;
    static constexpr int DEPTH = SJ_FUSED_DEPTH;
    static constexpr int HALO = 32;
    static constexpr int BUF = 2048 + HALO;
    static constexpr int WCAP = 512;
    static constexpr int SCRATCH = 32 * 80;                     // 32 parked lanes of 80 B >= (WCAP + 4) * 4 bytes of index staging
    static constexpr int WARP_BYTES = DEPTH * BUF + SCRATCH;
    static constexpr int SMEM_BYTES = NW * WARP_BYTES;
    static constexpr int MAXREG = SJ_FUSED_REG;
    static_assert(SCRATCH >= (WCAP + 4) * 4, "staging area must fit the scratch");
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_cg_u64(const void *p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int NW, bool UTF8>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(FusedCfg<NW>::MAXREG) stage1_fused_kernel(const Stage1Params P, uint32_t nchunks) {
    using Cfg = FusedCfg<NW>;
    constexpr int DEPTH = Cfg::DEPTH;
    constexpr uint32_t RUN = TICKET_CHUNKS;
    static_assert(RUN * SJ_U8_DEFER_MAX <= 32, "a run must not park more lanes than the scratch holds");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[NW * DEPTH];
    __shared__ uint32_t s_chunk[NW * DEPTH];                  // chunk held by each buffer, NO_CHUNK = nothing more to do
    __shared__ uint32_t s_first_run;                           // first classify run of the CTA's warps (one atomic per CTA)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = smem_raw + warp * Cfg::WARP_BYTES;
    const uint32_t buf0 = smem_u32(wbase);
    const uint8_t *bufs = wbase;
    uint8_t *scratch = wbase + DEPTH * Cfg::BUF;               // parked UTF-8 lanes while classifying, index staging while flattening
    const uint32_t bar0 = smem_u32(&s_bar[warp * DEPTH]);
    volatile uint32_t *my_chunk = s_chunk + warp * DEPTH;
    uint32_t *ticket = P.ticket + 3, *fticket = P.ticket + 4, *exits = P.ticket + 5;
    const uint32_t last = nchunks - 1u;
    const uint32_t last_bytes = (uint32_t)(P.alen - (uint64_t)last * 2048u);   // 1 .. 2048
    const uint32_t last_tx = (last_bytes + 15u) & ~15u;
    const bool last_partial = last_bytes < 2048u;
    const uint32_t nblocks = (nchunks + BLOCK_CHUNKS - 1u) / BLOCK_CHUNKS;
    const uint32_t ngroups = (nchunks + RUN - 1u) / RUN;

    if (threadIdx.x == 0) s_first_run = atomicAdd(ticket, NW * RUN);
    __syncthreads();

    // ---- classify side: lane 0 draws runs and keeps DEPTH bulk copies in flight --------------------------------
    uint32_t t_cur = 0, t_left = 0, t_next = s_first_run + warp * RUN;
    auto next_chunk = [&]() -> uint32_t {
        if (t_left == 0) {
            t_cur = t_next;
            t_left = RUN;
            t_next = atomicAdd(ticket, RUN);
        }
        const uint32_t c = t_cur + (RUN - t_left);
        t_left--;
        return c;
    };
    auto fetch = [&](int b) {
        const uint32_t c = next_chunk();
        if (c < nchunks) {
            my_chunk[b] = c;
            const uint32_t halo = c > 0u ? (uint32_t)Cfg::HALO : 0u;
            const uint32_t tx = (c == last ? last_tx : 2048u) + halo;
            mbar_expect_tx(bar0 + 8 * b, tx);
            bulk_load(buf0 + b * Cfg::BUF + Cfg::HALO - halo, P.abase + (size_t)c * 2048u - halo, tx, bar0 + 8 * b);
        } else {
            my_chunk[b] = NO_CHUNK;
            mbar_arrive(bar0 + 8 * b);
        }
    };
    if (lane == 0) {
        for (int b = 0; b < DEPTH; b++) mbar_init(bar0 + 8 * b, 1);
        fence_mbar_init();
        for (int b = 0; b < DEPTH; b++) fetch(b);
    }
    __syncwarp();

    // ---- scan side: the warp that completed block `blk` scans it -------------------------------------------------
    auto finish_block = [&](uint32_t blk) {
        const uint32_t c0 = blk * BLOCK_CHUNKS + 2u * (uint32_t)lane;
        uint4 sum[2];
        SpanAcc mine = span_empty();
#pragma unroll
        for (int k = 0; k < 2; k++) {
            sum[k] = c0 + k < nchunks ? ld_cg_u4(reinterpret_cast<const uint4 *>(P.chunk_sum) + (c0 + k)) : make_uint4(0u, 0u, 0u, 0u);
            mine = span_concat(mine, span_from_summary(sum[k]));
        }
        const SpanAcc incl = warp_span_inclusive(mine, lane);
        // the block aggregate (lane 31's inclusive span), as a look-back descriptor
        TileAgg agg;
        const uint32_t fl = __shfl_sync(0xFFFFFFFFu, span_flags(incl), 31);
        agg.par = fl & 1u;
        agg.un[0] = (fl >> 1) & 1u;
        agg.un[1] = (fl >> 2) & 1u;
        agg.u8 = (fl >> 3) & 1u;
        agg.c[0] = __shfl_sync(0xFFFFFFFFu, incl.c[0], 31);
        agg.c[1] = __shfl_sync(0xFFFFFFFFu, incl.c[1], 31);
        agg.e_out = agg.p_out = 0;
        LookbackResult lb = {0, 0, 0};
        if (blk > 0) {
            if (lane == 0) st_desc(P.desc + blk, desc_pack_agg(P.gen, agg));
            lb = lookback(P.desc, P.gen, (int)blk, lane);
        }
        const uint32_t s_in = lb.s_in & 1u;
        TilePrefix pre;
        pre.s_out = s_in ^ agg.par;
        pre.e_out = pre.p_out = 0;
        pre.err = lb.err | ((s_in ? agg.un[1] : agg.un[0]) ? EF_UNESCAPED : 0u) | (agg.u8 ? EF_UTF8 : 0u);
        pre.count = lb.base + (s_in ? agg.c[1] : agg.c[0]);
        if (lane == 0) {
            st_desc(P.desc + blk, desc_pack_prefix(P.gen, pre));
            if (blk == nblocks - 1u) {
                // every run of the document has signalled its block by now, i.e. has validated its parked lanes
                __threadfence();
                if (UTF8 && ld_acquire_u32(P.spec_flag + 1) == P.gen) pre.err |= EF_UTF8;
                write_verdict(P, pre);
            }
        }
        // state entering each of this lane's chunks: the lane's inclusive span "minus" its chunks, walked backwards
        // (as a function of the parity entering the BLOCK, which is s_in)
        uint32_t par = incl.par & 1u, cnt0 = incl.c[0], cnt1 = incl.c[1];
#pragma unroll
        for (int k = 1; k >= 0; k--) {
            const uint32_t rel = (par ^ sum[k].z) & 1u;            // parity of the block's chunks before chunk k
            cnt0 -= rel ? sum[k].y : sum[k].x;                      // rank offset of chunk k if the block starts outside a string
            cnt1 -= rel ? sum[k].x : sum[k].y;                      //   ... inside a string
            par = rel;
            const uint32_t s_w = s_in ^ rel;
            const uint64_t first = (uint64_t)lb.base + (s_in ? cnt1 : cnt0);
            if (c0 + k < nchunks) P.carry[c0 + k] = first | (s_w ? CARRY_INSIDE : 0ull);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(P.blk_ready + blk, P.gen);
    };

    // ---- flatten side ------------------------------------------------------------------------------------------------
    uint32_t fg = NO_CHUNK;        // claimed flatten group (RUN chunks), NO_CHUNK = none
    bool f_ready = false, f_done = false;
    auto claim = [&]() {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(fticket, 1u);
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        if (g >= ngroups) {
            f_done = true;
            fg = NO_CHUNK;
        } else {
            fg = g;
            f_ready = false;
        }
    };
    auto flatten_group = [&]() {
        uint32_t *stage = reinterpret_cast<uint32_t *>(scratch);
        const uint32_t cb = fg * RUN, ce = cb + RUN < nchunks ? cb + RUN : nchunks;
        // all carry words, then all mask words, are requested before anything is consumed: two L2 round trips per group
        uint64_t carry[RUN], structural[RUN];
#pragma unroll
        for (uint32_t k = 0; k < RUN; k++) carry[k] = cb + k < ce ? ld_cg_u64(P.carry + cb + k) : 0ull;
#pragma unroll
        for (uint32_t k = 0; k < RUN; k++) {
            const uint32_t s_w = (uint32_t)(carry[k] >> 63);
            structural[k] = cb + k < ce ? ld_cg_u64(P.masks + (size_t)(cb + k) * 64 + s_w * 32 + lane) : 0ull;
        }
#pragma unroll
        for (uint32_t k = 0; k < RUN; k++) {
            if (cb + k < ce) {
                if (!(SJ_FUSED_ABL & 1) || structural[k] == 0x123456789ull) flatten_chunk<Cfg::WCAP>(P, cb + k, carry[k], structural[k], stage, lane);
                __syncwarp();   // the staging area is reused by the next chunk
            }
        }
        fg = NO_CHUNK;
    };
    // non-blocking unless `wait`: flatten the claimed group if its block has been scanned
    auto try_flatten = [&](bool wait) -> bool {
        if (fg == NO_CHUNK && !f_done) claim();
        if (fg == NO_CHUNK) return false;
        const uint32_t blk = (fg * RUN) / BLOCK_CHUNKS;
        while (!f_ready) {
            uint32_t r = 0;
            if (lane == 0) r = ld_acquire_u32(P.blk_ready + blk);
            r = __shfl_sync(0xFFFFFFFFu, r, 0);
            f_ready = (r == P.gen);
            if (f_ready || !wait) break;
            __nanosleep(200);
        }
        if (!f_ready) return false;
        flatten_group();
        return true;
    };

    // ---- main loop ----------------------------------------------------------------------------------------------------
    int b = 0;
    uint32_t phase = 0;
    uint32_t parked = 0;           // lanes parked in `scratch` during the current run
    bool u8_bad = false;
    while (true) {
        mbar_wait(bar0 + 8 * b, phase);
        const uint32_t c = my_chunk[b];
        if (c == NO_CHUNK) break;
        LanePhase1 ph;
        {
            LaneInput in;
            uint32_t unresolved;
            const bool edge = (c == 0u) || (c == last && last_partial);
            chunk_load<UTF8>(in, bufs + b * Cfg::BUF + Cfg::HALO, lane, c, edge, c == last, last_bytes, P, unresolved);
            __syncwarp();  // every lane has its bytes (and the chunk number) in registers: the buffer can be refilled
            if (lane == 0) fetch(b);
            if (unresolved && !resolve_long_runs(P, c, unresolved, lane, in.wst) && lane == 0) *P.spec_flag = P.gen;   // rare, warp-uniform
            warp_compute<UTF8, 2>(ph, in, lane, P, reinterpret_cast<uint4 *>(scratch) + 5 * parked);
            parked += (uint32_t)__popc(ph.u8_lanes);
        }
        uint64_t *mp = P.masks + (size_t)c * 64 + lane;
        if (!(SJ_FUSED_ABL & 4) || ph.m0 == 0x123456789ull) {
            __stcs(reinterpret_cast<unsigned long long *>(mp), (unsigned long long)ph.m0);
            __stcs(reinterpret_cast<unsigned long long *>(mp + 32), (unsigned long long)ph.m1);
        }
        if (lane == 0) reinterpret_cast<uint4 *>(P.chunk_sum)[c] = make_uint4(ph.wc0, ph.wc1, ph.wflags, 0u);
        if (++b == DEPTH) {
            b = 0;
            phase ^= 1u;
        }
        if ((c & (RUN - 1u)) == RUN - 1u || c == last) {
            // ---- end of a run: validate the parked lanes, signal the block, then take a turn at flattening ----
            if (UTF8 && parked) {
                __syncwarp();
                if (validate_parked_lanes(scratch, parked, lane)) u8_bad = true;
                parked = 0;
                __syncwarp();
            }
            const uint32_t blk = c / BLOCK_CHUNKS;
            const uint32_t run_len = (c & (RUN - 1u)) + 1u;
            const uint32_t blk_len = blk == nblocks - 1u ? nchunks - blk * BLOCK_CHUNKS : BLOCK_CHUNKS;
            uint32_t old = 0;
            __syncwarp();   // orders every lane's mask stores before lane 0's fence
            if (lane == 0) {
                if (UTF8 && u8_bad) P.spec_flag[1] = P.gen;
                if (!(SJ_FUSED_ABL & 2)) __threadfence();
                old = atomicAdd(P.blk_done + blk, run_len);
            }
            old = __shfl_sync(0xFFFFFFFFu, old, 0);
            if (old + run_len == blk_len) {
                if (lane == 0) P.blk_done[blk] = 0;   // every run of the block has arrived: the counter is free for the next document
                __threadfence();
                finish_block(blk);
            }
            for (int k = 0; k < SJ_FUSED_LAG; k++)
                if (!try_flatten(false)) break;
        }
    }
    // classify work is exhausted: flatten whatever is left, waiting for the last blocks to be scanned
    while (try_flatten(true)) {
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(exits, 1u) == gridDim.x - 1u) {
        *ticket = 0;
        *fticket = 0;
        *exits = 0;
    }
}

#endif  // __CUDACC__

}  // namespace sjb200
