// stage1_stream2.cuh -- the classify launch of the stream pipeline with lanes of 128 consecutive bytes.
//
// stage1_stream_classify_kernel (stage1_stream.cuh) gives a lane 64 bytes and a warp one 2 KiB chunk per iteration.  It is
// bound by the ALU pipe, and a good part of its ALU instructions is per-iteration rather than per-byte work: the ballots
// that carry escape / scalar / in-string state from lane to lane, the warp reductions of the summary, the mask and
// summary stores, the ring bookkeeping and the bulk-copy issue.  Here a warp takes a UNIT of two consecutive chunks
// (4 KiB) per iteration and a lane 128 consecutive bytes -- two mask words, A and B, chained inside the lane -- so all
// of that happens once per 4 KiB.  What leaves the kernel is unchanged: the two mask planes and one 16-byte summary PER
// CHUNK (lanes 0..15 hold chunk 2u, lanes 16..31 chunk 2u + 1; the second chunk's planes / counts are swapped when the
// first one flips the string state), so the scan and flatten launches do not know the difference.
//
// 128-byte lane strides would make the shared-memory reads of the input 8-way bank conflicted, so the unit is fetched with
// a tensor-map copy (cp.async.bulk.tensor.2d, 32 rows of 128 bytes, CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte piece k of
// row r lands at piece k ^ (r & 7), tools/ubench_swizzle.cu) into a 1024-byte aligned buffer: lane t owns row t, and the
// eight lanes a 16-byte load serves together read eight different pieces.  The 32 bytes before the unit come with a plain
// bulk copy.  Needs the document to start on a 128-byte boundary (the last row of a copy then stays inside the last
// 128-byte line of the data); capi.cu falls back to the 64-byte-lane kernel otherwise.
// Reference: json_structural_indexer.mojo:83-186 (step / next / finish), restated in oracle/stage1_oracle.c.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched from the driver at run time, capi.cu)

#include "stage1_stream.cuh"

#ifndef SJ_WIDE_NW
#define SJ_WIDE_NW 4
#endif
#ifndef SJ_WIDEREG
#define SJ_WIDEREG 96
#endif
#ifndef SJ_WIDE_RUN
#define SJ_WIDE_RUN 2   // units per draw from the ticket counter
#endif

namespace sjb200 {

#if defined(__CUDACC__)

template <int NW>
struct WideCfg {
    static constexpr int THREADS = NW * 32;
    static constexpr int DEPTH = 2;
    static constexpr int UNIT = 4096;
    static constexpr int HALO = 32;
    static constexpr int PARK = 32 * 80;                       // 32 parked UTF-8 words of 80 B per warp
    static constexpr int SIDE = DEPTH * HALO + PARK + 16;      // per warp: look-behind areas, parking slots, two mbarriers
    static constexpr int SMEM_BYTES = NW * DEPTH * UNIT + NW * SIDE;   // no static shared memory: the dynamic part starts 1024-byte aligned
    static constexpr int MAXREG = SJ_WIDEREG;
};

__device__ __forceinline__ void tensor_load_2d(uint32_t dst, const CUtensorMap *tmap, uint32_t x, uint32_t y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}

// one 64-byte word of a lane: planes, JSON classes; UTF-8: park the planes (or validate in place when too many lanes are flagged)
template <bool UTF8>
__device__ __forceinline__ void wide_word(const uint32_t w[16], uint32_t prev, uint32_t ends, uint32_t lt, uint4 *park, uint32_t &parked, LaneMasks &m,
                                          uint32_t &u8err) {
    uint32_t pl[8], ph[8];
    bitplanes32(w, pl);
    bitplanes32(w + 8, ph);
    Classes32 cl, ch;
    classify_json32(pl, cl);
    classify_json32(ph, ch);
    m.bs = join64(cl.bs, ch.bs);
    m.rq = join64(cl.rq, ch.rq);
    m.op = join64(cl.op, ch.op);
    m.ws = join64(cl.ws, ch.ws);
    m.ctl = join64(cl.ctl, ch.ctl);
    if (UTF8) {
        const bool any_hi = ((pl[7] | ph[7]) != 0) || ((prev & 0x80808080u) != 0);
        const uint32_t hi_lanes = __ballot_sync(0xFFFFFFFFu, any_hi);
        if (__popc(hi_lanes) <= SJ_U8_DEFER_MAX) {
            if (any_hi) {
                uint4 *s = park + 5 * (parked + __popc(hi_lanes & lt));   // 80 contiguous bytes per slot
                s[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                s[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                s[2] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                s[3] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                s[4] = make_uint4(prev, ends, 0u, 0u);
            }
            parked += (uint32_t)__popc(hi_lanes);
        } else {
            Utf8Pre32 ul, uh;
            utf8_pre32(pl, ul);
            utf8_pre32(ph, uh);
            const Utf8Carry uc = utf8_carry_from_prev_word(prev);
            uint32_t tail_must;
            const uint64_t ue = utf8_errors64(ul, uh, uc, &tail_must);
            u8err |= (ue != 0) || (ends && tail_must != 0);
        }
    }
}

template <int NW, bool UTF8>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(WideCfg<NW>::MAXREG) stage1_stream_classify2_kernel(const Stage1Params P, uint32_t nchunks,
                                                                                                           const uint8_t *src0 /* P.abase - HALO */,
                                                                                                           const __grid_constant__ CUtensorMap tmap) {
    using Cfg = WideCfg<NW>;
    constexpr int DEPTH = Cfg::DEPTH;
    static_assert(2 * SJ_U8_DEFER_MAX <= 32, "a unit must not park more words than the warp's slots hold");
    extern __shared__ __align__(1024) uint8_t smem_wide[];   // (its own name: the other kernels declare theirs with a smaller alignment)
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));   // read once (see stage1_stream.cuh)
    const uint32_t lt = lanemask_lt();
    const int warp = threadIdx.x >> 5;
    uint8_t *side = smem_wide + NW * DEPTH * Cfg::UNIT + warp * Cfg::SIDE;
    const uint32_t buf0 = smem_u32(smem_wide) + (uint32_t)warp * (DEPTH * Cfg::UNIT);   // unit buffers, 1024-byte aligned
    const uint32_t behind0 = smem_u32(side) + Cfg::HALO;                                // just past the look-behind bytes of buffer 0
    uint4 *park = reinterpret_cast<uint4 *>(side + DEPTH * Cfg::HALO);
    const uint32_t bar0 = smem_u32(side + DEPTH * Cfg::HALO + Cfg::PARK);
    uint32_t *ticket = P.ticket + 3, *exits = P.ticket + 5;
    const uint32_t nunits = (nchunks + 1u) >> 1;
    const uint32_t last = nchunks - 1u;
    const uint32_t last_bytes = (uint32_t)(P.alen - (uint64_t)last * 2048u);   // bytes of the last chunk, 1 .. 2048
    const bool last_partial = last_bytes < 2048u;
    // this lane's bytes: row `lane` of the buffer, piece k at (k ^ (lane & 7)); the 4 bytes before them: piece 7 of the row above
    const uint32_t swz_off = (uint32_t)lane * 128u + (((uint32_t)lane & 7u) << 4);
    const uint32_t prev_off = ((uint32_t)lane - 1u) * 128u + ((7u ^ (((uint32_t)lane - 1u) & 7u)) << 4) + 12u;

    grid_dependency_wait();   // launched while the previous kernel of the stream drains: nothing global is touched before this
    if (smem_u32(smem_wide) & 1023u) {   // the swizzle is a function of the shared-memory address: never expected, but then the
        if (threadIdx.x == 0) *P.spec_flag = P.gen;   // exact fallback (stage1_persistent.cuh) does this document
        return;
    }
    constexpr uint32_t RUN = SJ_WIDE_RUN;
    const uint32_t t_base = gridDim.x * NW * RUN;
    uint32_t t_cur = (blockIdx.x * NW + warp) * RUN, t_left = RUN;
    uint32_t drawn = lane == 0 ? ticket_draw(ticket, RUN) : 0u;      // lane 0: the run after this one (relative to t_base)
    auto next_unit = [&]() -> uint32_t {   // warp-uniform
        if (t_left == 0u) {
            t_cur = t_base + __shfl_sync(0xFFFFFFFFu, drawn, 0);
            t_left = RUN;
            if (lane == 0) drawn = ticket_draw(ticket, RUN);
        }
        t_left--;
        return t_cur++;
    };
    // lane 0: start the copies of unit u (< nunits) and its look-behind into buffer b; the box is always whole (rows past the end
    // of the data arrive as zeros)
    auto issue = [&](uint32_t b, uint32_t u) {
        const uint32_t bar = bar0 + 8u * b;
        const uint32_t halo = u > 0u ? (uint32_t)Cfg::HALO : 0u;
        mbar_expect_tx(bar, (uint32_t)Cfg::UNIT + halo);
        tensor_load_2d(buf0 + b * Cfg::UNIT, &tmap, 0u, 32u * u, bar);
        if (u > 0u) bulk_load(behind0 + b * Cfg::HALO - Cfg::HALO, src0 + (size_t)u * Cfg::UNIT, Cfg::HALO, bar);
    };
    uint32_t held0, held1;          // the unit each buffer holds (all lanes)
    if (lane == 0) {
        for (int b = 0; b < DEPTH; b++) mbar_init(bar0 + 8 * b, 1);
        fence_mbar_init();
    }
    __syncwarp();
    held0 = next_unit();
    if (held0 < nunits && lane == 0) issue(0u, held0);
    held1 = next_unit();
    if (held1 < nunits && lane == 0) issue(1u, held1);
    uint32_t b = 0, phase = 0;
    uint32_t parked = 0;
    bool u8_bad = false;
    uint32_t prev_u = NO_CHUNK - 1u, prev_tail = 0;   // the unit this warp processed last and its carries out
    while (true) {
        const uint32_t u = b ? held1 : held0;
        if (u >= nunits) break;
        mbar_wait(bar0 + 8u * b, phase);
        const uint32_t c = 2u * u;                     // first chunk of the unit
        const bool two = c + 1u < nchunks;             // the document's last unit may hold one chunk only
        const uint32_t chunk = buf0 + b * Cfg::UNIT, behind = behind0 + b * Cfg::HALO;
        uint32_t wa[16], wb[16];
        {
            const uint32_t src = chunk + swz_off;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 v = lds_u4(src ^ (16u * q));
                wa[4 * q + 0] = v.x; wa[4 * q + 1] = v.y; wa[4 * q + 2] = v.z; wa[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 v = lds_u4(src ^ (16u * (q + 4)));
                wb[4 * q + 0] = v.x; wb[4 * q + 1] = v.y; wb[4 * q + 2] = v.z; wb[4 * q + 3] = v.w;
            }
        }
        uint32_t prev = UTF8 ? lds_u32(lane ? chunk + prev_off : behind - 4u) : 0u;
        const int64_t g0 = (int64_t)c * 2048 + lane * 128;
        // the document may end exactly with a word's last byte
        const uint32_t lane_chunk = (uint32_t)lane >> 4;                                  // which chunk of the unit this lane belongs to
        const bool in_last = c + lane_chunk == last;
        const uint32_t off_in_chunk = ((uint32_t)lane & 15u) * 128u;
        const uint32_t ends_a = in_last && (off_in_chunk + 64u == last_bytes);
        const uint32_t ends_b = in_last && (off_in_chunk + 128u == last_bytes);
        const bool edge = (u == 0u) || (c + 1u >= last && (last_partial || !two));
        if (edge) {  // bytes outside [mis, alen) read as 0x20 (reference tail padding)
            const int64_t alen = (int64_t)P.alen;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                wa[k] = mask_word(wa[k], g0 + 4 * k, (int64_t)P.mis, alen);
                wb[k] = mask_word(wb[k], g0 + 64 + 4 * k, (int64_t)P.mis, alen);
            }
            if (UTF8) prev = (g0 == 0) ? 0x20202020u : mask_word(prev, g0 - 4, (int64_t)P.mis, alen);
        }
        // carries entering the unit
        PrevState wst = {0, 0, 0};
        if (u == prev_u + 1u) {
            wst.e = prev_tail & 1u;
            wst.p = (prev_tail >> 1) & 1u;
        } else if (u > 0u) {  // 32 bytes of look-behind, all inside the document
            const uint32_t bb = lds_u8(behind - 1u - (uint32_t)lane);
            const uint32_t bsm = __ballot_sync(0xFFFFFFFFu, bb == 0x5Cu);
            const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, bb, 0);
            wst = prev_state(bsm, 32, c1);
        }
        const uint32_t prev_b = wa[15];                // the 4 bytes before word B
        __syncwarp();  // every lane has its bytes in registers: the buffer can be refilled
        {
            const uint32_t nu = next_unit();
            if (b) held1 = nu; else held0 = nu;
            if (nu < nunits && lane == 0) issue(b, nu);
        }
        if (wst.unresolved && !resolve_long_runs(P, c, wst.unresolved, lane, wst) && lane == 0) *P.spec_flag = P.gen;

        // ---- the two words of the lane --------------------------------------------------------------------------------
        LaneMasks ma, mb;
        uint32_t u8err = 0;
        wide_word<UTF8>(wa, prev, ends_a, lt, park, parked, ma, u8err);
        wide_word<UTF8>(wb, prev_b, ends_b, lt, park, parked, mb, u8err);
        // escapes: a lane of 128 backslashes passes its carry through; otherwise the run the lane ends with decides
        uint32_t e_in = lane ? 0u : wst.e, e_out = 0u;
        if (__ballot_sync(0xFFFFFFFFu, (uint32_t)(mb.bs >> 32) >> 31)) {
            const bool b_all = lane_all_backslash(mb.bs);
            const uint32_t bA = __ballot_sync(0xFFFFFFFFu, b_all && lane_all_backslash(ma.bs));
            const uint32_t bO = __ballot_sync(0xFFFFFFFFu, b_all ? lane_trailing_run_parity(ma.bs) : lane_trailing_run_parity(mb.bs));
            e_in = warp_lane_e_in(bA, bO, lane, wst.e);
            e_out = warp_lane_e_in(bA, bO, 32, wst.e);
        }
        const uint64_t esc_a = escaped_mask(ma.bs, (uint64_t)(e_in & 1u));
        const uint64_t e_mid = (ma.bs & ~esc_a) >> 63;                 // an unescaped backslash at the end of A escapes B's first byte
        const uint64_t esc_b = escaped_mask(mb.bs, e_mid);
        const uint64_t quote_a = ma.rq & ~esc_a, quote_b = mb.rq & ~esc_b;
        const uint64_t ps_a = prefix_xor64(quote_a);
        const uint64_t ps_b = prefix_xor64(quote_b) ^ (0ull - (ps_a >> 63));      // in-string relative to the start of the lane
        const uint64_t scalar_a = ~(ma.op | ma.ws), scalar_b = ~(mb.op | mb.ws);
        const uint64_t nqs_a = scalar_a & ~quote_a, nqs_b = scalar_b & ~quote_b;
        const uint32_t bPB = __ballot_sync(0xFFFFFFFFu, (ps_b >> 63) != 0);
        const uint32_t bNQ = __ballot_sync(0xFFFFFFFFu, (nqs_b >> 63) != 0);
        const uint32_t rel = (uint32_t)__popc(bPB & lt) & 1u;          // parity of the unescaped quotes of the earlier lanes
        const uint32_t p_in = lane ? ((bNQ >> (lane - 1)) & 1u) : wst.p;
        const uint64_t relm = 0ull - (uint64_t)rel;
        const uint64_t in0_a = ps_a ^ relm, in0_b = ps_b ^ relm;       // in_string if the UNIT starts outside a string
        const uint64_t pot_a = ma.op | (scalar_a & ~((nqs_a << 1) | p_in));
        const uint64_t pot_b = mb.op | (scalar_b & ~((nqs_b << 1) | (nqs_a >> 63)));
        const uint64_t tail_a = in0_a ^ quote_a, tail_b = in0_b ^ quote_b;
        uint64_t m0a = pot_a & ~tail_a, m1a = pot_a & tail_a, m0b = pot_b & ~tail_b, m1b = pot_b & tail_b;
        uint32_t un0 = ((ma.ctl & in0_a) | (mb.ctl & in0_b)) != 0, un1 = ((ma.ctl & ~in0_a) | (mb.ctl & ~in0_b)) != 0;
        uint32_t c0 = (uint32_t)(__popcll(m0a) + __popcll(m0b)), c1 = (uint32_t)(__popcll(m1a) + __popcll(m1b));
        // ---- per-chunk results: lanes 0..15 are chunk c, lanes 16..31 chunk c + 1, whose own "starts outside a string" is the
        // unit's "starts inside" when chunk c holds an odd number of quotes
        const uint32_t par0 = (uint32_t)__popc(bPB & 0xFFFFu) & 1u;    // parity of chunk c's quotes
        const uint32_t par1 = ((uint32_t)__popc(bPB) & 1u) ^ par0;     // parity of chunk c + 1's
        if (lane_chunk & par0) {   // second chunk, string state flipped by the first one
            uint64_t t;
            t = m0a; m0a = m1a; m1a = t;
            t = m0b; m0b = m1b; m1b = t;
            uint32_t s;
            s = c0; c0 = c1; c1 = s;
            s = un0; un0 = un1; un1 = s;
        }
        const uint32_t pack = c0 | (c1 << 16);                         // a chunk holds at most 2048 indexes
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, pack);
        const uint32_t first = __reduce_add_sync(0xFFFFFFFFu, lane_chunk ? 0u : pack);
        const uint32_t second = tot - first;
        const uint32_t fl = (un0 << 1) | (un1 << 2) | (u8err << 3);
        const uint32_t flags_all = __reduce_or_sync(0xFFFFFFFFu, lane_chunk ? fl << 8 : fl);
        // stores: both planes of this lane's chunk ([chunk][parity][lane], this lane's two words are words 2 (lane & 15), + 1)
        if (lane_chunk == 0u || two) {
            uint64_t *mp = P.masks + (size_t)(c + lane_chunk) * 64 + 2u * ((uint32_t)lane & 15u);
            asm volatile("st.global.cs.v2.u64 [%0], {%1,%2};" ::"l"(mp), "l"(m0a), "l"(m0b) : "memory");
            asm volatile("st.global.cs.v2.u64 [%0], {%1,%2};" ::"l"(mp + 32), "l"(m1a), "l"(m1b) : "memory");
        }
        if (lane == 0) reinterpret_cast<uint4 *>(P.chunk_sum)[c] = make_uint4(first & 0xFFFFu, first >> 16, (flags_all & 0xFFu) | par0, 0u);
        if (lane == 16 && two) reinterpret_cast<uint4 *>(P.chunk_sum)[c + 1u] = make_uint4(second & 0xFFFFu, second >> 16, ((flags_all >> 8) & 0xFFu) | par1, 0u);
        prev_u = u;
        prev_tail = e_out | ((bNQ >> 31) << 1);
        if (UTF8 && parked > 32u - 2u * SJ_U8_DEFER_MAX) {   // the next unit might not find room
            __syncwarp();
            u8_bad |= validate_parked_lanes(reinterpret_cast<const uint8_t *>(park), parked, lane);
            parked = 0;
            __syncwarp();
        }
        b ^= 1u;
        if (b == 0u) phase ^= 1u;
    }
    if (UTF8 && parked) {
        __syncwarp();
        u8_bad |= validate_parked_lanes(reinterpret_cast<const uint8_t *>(park), parked, lane);
    }
    // a violation among the deferred words: the document's last launch folds it into the verdict (stage1_persistent.cuh)
    if (UTF8 && u8_bad && lane == 0) P.spec_flag[1] = P.gen;
    // the last CTA to leave frees the unit counter for the next launch
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(exits, 1u) == gridDim.x - 1u) {
        *ticket = 0;
        *exits = 0;
    }
}

#endif  // __CUDACC__

}  // namespace sjb200
