// stage1_dataflow.cuh -- the stage-1 kernel as a dataflow pipeline inside every persistent CTA.
//
// Measured on the warp-specialised kernel of stage1_persistent.cuh (profiles/, tools/trace_run.py): phase 1 (bytes ->
// structural masks) saturates the integer ALU pipe while it runs, the flatten phase (masks -> indexes) is bound by
// dependent-chain latency, shared-memory stores and the XU pipe and leaves the ALU half idle -- and because the same
// warps alternate between the two, the SM sees one or the other, never both.  Here they are different warps:
//
//   bulk copy (cp.async.bulk) -> CLASSIFIER warps -> mask ring in shared memory -> FLATTENER warps -> global memory
//                                        |                                               ^
//                                        +--> tile aggregate --> SCAN warp (look-back) --+  (parity, output cursor)
//
//   classifier warp c : phase 1 of its 2 KiB of every tile; writes the lane's two candidate masks (16 B) to the ring,
//                       never waits for a look-back
//   flattener warp c  : for every tile, once the scan warp has delivered the parity and the cursor, reads the masks
//                       of classifier c, picks the right one, flattens into its private staging area and stores
//   scan warp         : look-backs, nothing else
// The ring is RD tiles deep, so the look-back latency (and its jitter) is absorbed without stalling classification.
#pragma once
#include "stage1_persistent.cuh"

namespace sjb200 {

#if defined(__CUDACC__)

template <int NC>
struct FlowCfg {
    static constexpr int THREADS = (2 * NC + 1) * 32;
    static constexpr int TILE = NC * 2048;
    static constexpr int RD = 4;                                   // ring depth in tiles
    static constexpr int WCAP = 512;                               // staged indexes per flattener warp
    static constexpr int NIN = 2;                                  // input buffers (the bulk copy runs two tiles ahead)
    static constexpr int IN_ONE = ((16 + TILE) + 127) & ~127;      // halo + tile
    static constexpr int IN_BYTES = NIN * IN_ONE;
    static constexpr int RING_BYTES = RD * NC * 32 * 16;           // {m0, m1} per lane
    static constexpr int STAGE_BYTES = NC * (WCAP + 4) * 4;
    static constexpr int SMEM_BYTES = IN_BYTES + RING_BYTES + STAGE_BYTES;
    static constexpr int MAXREG = NC == 8 ? SJ_FLOWREG8 : 64;
};

struct FlowSlot {
    uint32_t wc0[16], wc1[16], wflags[16];
    uint32_t R[16], off0[16], off1[16];
    uint64_t agg;
    uint32_t tail;
    uint32_t arrived;
    uint32_t s_in, base;
    int32_t tile;
};

template <int NC, bool UTF8>
__global__ void __launch_bounds__((2 * NC + 1) * 32) __maxnreg__(FlowCfg<NC>::MAXREG) stage1_dataflow_kernel(const Stage1Params P) {
    using Cfg = FlowCfg<NC>;
    constexpr int TILE = Cfg::TILE;
    constexpr int RD = Cfg::RD;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem_in = smem_raw;
    uint4 *ring = reinterpret_cast<uint4 *>(smem_raw + Cfg::IN_BYTES);                       // [RD][NC][32]
    uint32_t *smem_stage = reinterpret_cast<uint32_t *>(smem_raw + Cfg::IN_BYTES + Cfg::RING_BYTES);
    __shared__ __align__(8) uint64_t s_bar[2 + 3 * RD];  // 0-1 in_full[2], 2.. sum_full[RD], carry_full[RD], slot_free[RD]
    __shared__ FlowSlot s_slot[RD];
    __shared__ int32_t s_tile_of[2];
    __shared__ uint32_t s_loaded[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_in_full = smem_u32(&s_bar[0]);
    const uint32_t bar_sum = smem_u32(&s_bar[2]), bar_carry = smem_u32(&s_bar[2 + RD]), bar_free = smem_u32(&s_bar[2 + 2 * RD]);

    if (tid == 0) {
        mbar_init(bar_in_full, 1);
        mbar_init(bar_in_full + 8, 1);
        s_loaded[0] = s_loaded[1] = 0;
        for (int k = 0; k < RD; k++) {
            mbar_init(bar_sum + 8 * k, 1);
            mbar_init(bar_carry + 8 * k, 1);
            mbar_init(bar_free + 8 * k, NC);
            s_slot[k].arrived = 0;
        }
        fence_mbar_init();
        if (blockIdx.x == 0) P.ticket[(P.ticket_sel + 1) & 1u] = 0;
    }
    __syncthreads();
    uint32_t *ticket = P.ticket + (P.ticket_sel & 1u);

    // one thread draws the next ticket and starts its bulk copy into input buffer b (which must be free)
    auto produce = [&](int b) {
        const uint32_t k = P.tile_begin + atomicAdd(ticket, 1u);
        const uint32_t bar = bar_in_full + 8 * b;
        if (k < P.tile_end) {
            const int t = (int)k;
            s_tile_of[b] = t;
            TRACE(P, t, 0, gtime());  // ticket drawn / copy issued
            const int64_t tb = (int64_t)t * TILE;
            int64_t nbytes = (int64_t)P.alen - tb;
            nbytes = nbytes > TILE ? TILE : nbytes;
            nbytes = (nbytes + 15) & ~15ll;
            const uint32_t halo = t > 0 ? 16u : 0u;
            mbar_expect_tx(bar, (uint32_t)nbytes + halo);
            bulk_load(smem_u32(smem_in + b * Cfg::IN_ONE) + 16u - halo, P.abase + tb - halo, (uint32_t)nbytes + halo, bar);
        } else {
            s_tile_of[b] = -1;
            mbar_arrive(bar);
        }
    };
    if (tid == 0) {
        produce(0);
        produce(1);
    }

    if (warp < NC) {
        // =============================== classifier warps ===============================
        for (int i = 0;; i++) {
            const int slot = i & (RD - 1);
            const int b = i & 1;                            // input buffer of this iteration
#if SJ_TRACE
            const uint64_t tw0 = gtime();
#endif
            mbar_wait(bar_in_full + 8 * b, (uint32_t)(i >> 1) & 1u);
            const int tile = *reinterpret_cast<volatile int32_t *>(&s_tile_of[b]);
#if SJ_TRACE
            const uint64_t tw1 = gtime();
#endif
            // the ring slot must have been drained by the flatteners (tile i - RD)
            if (i >= RD) mbar_wait(bar_free + 8 * slot, (uint32_t)(i / RD - 1) & 1u);
#if SJ_TRACE
            if (tile >= 0 && lane == 0 && (warp == 0 || warp == NC - 1)) {
                TRACE(P, tile, warp == 0 ? 5 : 6, tw1 - tw0);              // input wait
                TRACE(P, tile, warp == 0 ? 14 : 15, gtime() - tw1);        // ring-full wait
                TRACE(P, tile, warp == 0 ? 1 : 7, gtime());                // phase-1 start
            }
#endif
            FlowSlot &S = s_slot[slot];
            if (tile < 0) {
                // no more work: tell the scan warp (which tells the flatteners) through this iteration's slot
                if (warp == 0 && lane == 0) {
                    S.tile = -1;
                    mbar_arrive(bar_sum + 8 * slot);
                }
                break;
            }
            const int64_t tb = (int64_t)tile * TILE;
            LanePhase1 ph;
            {
                LaneInput in;
                warp_load<UTF8>(in, smem_in + b * Cfg::IN_ONE + 16, warp, lane, tile, tb, TILE, P);
                __syncwarp();
                if (lane == 0) {                            // this warp no longer needs the input buffer
                    __threadfence_block();
                    if (atomicAdd(&s_loaded[b], 1u) == NC - 1) {
                        s_loaded[b] = 0;                    // everyone has it in registers: refill this buffer (tile i+2)
                        __threadfence_block();
                        produce(b);
                    }
                }
                warp_compute<UTF8>(ph, in, lane, P);
            }
            // hand the candidate masks to the flattener of this warp
            ring[(slot * NC + warp) * 32 + lane] =
                make_uint4((uint32_t)ph.m0, (uint32_t)(ph.m0 >> 32), (uint32_t)ph.m1, (uint32_t)(ph.m1 >> 32));
            uint32_t order = 0;
            if (lane == 0) {
                S.wc0[warp] = ph.wc0;
                S.wc1[warp] = ph.wc1;
                S.wflags[warp] = ph.wflags;
                if (warp == NC - 1) S.tail = ph.tail;
            }
            __syncwarp();                                   // the warp's ring entries are written
            if (lane == 0) {
                __threadfence_block();
                order = atomicAdd(&S.arrived, 1u);
            }
            order = __shfl_sync(0xFFFFFFFFu, order, 0);
            if (order == NC - 1) {
                // last classifier of the tile: build and publish the aggregate right away
                __threadfence_block();
                const bool have = lane < NC;
                uint32_t R, off0, off1;
                const TileAgg agg = tile_aggregate(have ? S.wflags[lane] : 0u, have ? S.wc0[lane] : 0u, have ? S.wc1[lane] : 0u,
                                                   S.tail, NC, lane, R, off0, off1);
                if (have) {
                    S.R[lane] = R;
                    S.off0[lane] = off0;
                    S.off1[lane] = off1;
                }
                const uint64_t packed = desc_pack_agg(P.gen, agg);
                if (lane == 0) {
                    if (tile > 0) st_desc(P.desc + tile, packed);  // tile 0 goes straight to its prefix
                    TRACE(P, tile, 2, gtime());  // aggregate published
                    S.agg = packed;
                    S.tile = tile;
                    S.arrived = 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sum + 8 * slot);     // release: slot + ring contents visible downstream
            }
        }
    } else if (warp < 2 * NC) {
        // =============================== flattener warps ===============================
        const int c = warp - NC;                            // the classifier this warp drains
        uint32_t *stage = smem_stage + c * (Cfg::WCAP + 4);
        for (int i = 0;; i++) {
            const int slot = i & (RD - 1);
#if SJ_TRACE
            const uint64_t tc0 = gtime();
#endif
            mbar_wait(bar_carry + 8 * slot, (uint32_t)(i / RD) & 1u);
            const FlowSlot &S = s_slot[slot];
            const int tile = *reinterpret_cast<const volatile int32_t *>(&S.tile);
            if (tile < 0) break;
#if SJ_TRACE
            if (lane == 0 && (c == 0 || c == NC - 1)) { TRACE(P, tile, c == 0 ? 8 : 9, gtime() - tc0); TRACE(P, tile, c == 0 ? 10 : 11, gtime()); }
#endif
            const uint4 mm = ring[(slot * NC + c) * 32 + lane];
            const uint32_t s_in = S.s_in & 1u;
            const uint32_t s_w = (s_in ^ S.R[c]) & 1u;
            const uint64_t first = (uint64_t)S.base + (s_in ? S.off1[c] : S.off0[c]);  // the warp's first index
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8 * slot);  // everything this warp needs from the slot is in registers
            const uint64_t structural = s_w ? join64(mm.z, mm.w) : join64(mm.x, mm.y);
            const uint32_t cnt = (uint32_t)__popcll(structural);
            const uint32_t incl = warp_inclusive_sum(cnt);
            const uint32_t wtotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t v0 = (uint32_t)((int64_t)tile * TILE + c * 2048 + lane * 64 - (int64_t)P.mis);
            if (wtotal <= (uint32_t)Cfg::WCAP) {
                const uint32_t a = ((uint32_t)first + out_phase(P.out)) & 3u;
                flatten_to(stage + a + (incl - cnt), structural, v0);
                __syncwarp();
                copy_out_warp(stage, a, wtotal, P.out, first, P.cap, (uint32_t)lane);
                __syncwarp();  // the staging area is reused by the next tile
            } else {
                flatten_direct(P.out, P.cap, first + (incl - cnt), structural, v0);
            }
#if SJ_TRACE
            if (lane == 0 && (c == 0 || c == NC - 1)) TRACE(P, tile, c == 0 ? 12 : 13, gtime());
#endif
        }
    } else {
        // =============================== scan warp ===============================
        for (int i = 0;; i++) {
            const int slot = i & (RD - 1);
            mbar_wait(bar_sum + 8 * slot, (uint32_t)(i / RD) & 1u);
            FlowSlot &S = s_slot[slot];
            const int cur = *reinterpret_cast<volatile int32_t *>(&S.tile);
            if (cur < 0) {
                if (lane == 0) mbar_arrive(bar_carry + 8 * slot);  // pass "no more work" on to the flatteners
                break;
            }
            TRACE(P, cur, 3, gtime());  // look-back starts
            const TileAgg agg = desc_unpack_agg(S.agg);
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
                S.s_in = s_in;
                S.base = lb.base;
                if (cur == (int)P.tile_end - 1 && P.progress) *P.progress = pre.count;
                if (cur == (int)P.ntiles - 1) write_verdict(P, pre);
                TRACE(P, cur, 4, gtime());  // look-back done
                mbar_arrive(bar_carry + 8 * slot);  // release: S.s_in / S.base visible to the flatteners
            }
            __syncwarp();
        }
    }
}

#endif  // __CUDACC__

}  // namespace sjb200
