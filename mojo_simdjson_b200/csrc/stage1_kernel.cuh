// stage1_kernel.cuh -- the sm_100a stage-1 kernel: one launch turns input bytes into the ordered uint32
// structural index array + verdict.  Replaces JsonStructuralIndexer.index/step/next/finish and BitIndexer
// (reference generic/stage1/json_structural_indexer.mojo:33-58,81-186) and the scanners they call.
//
// Work decomposition
//   lane  : 64 consecutive bytes (one 64-bit word of every mask)
//   warp  : 2 KiB, carries between lanes resolved with ballots (no shuffles of data)
//   CTA   : WARPS x 2 KiB = one tile, loaded by ONE bulk async copy (cp.async.bulk -> UBLKCP) into shared
//           memory behind an mbarrier; the same shared memory is reused to stage the tile's indexes so the
//           global index write is coalesced 16-byte stores
//   grid  : one CTA per tile, tile ids handed out by an atomic ticket so that a tile only ever waits for
//           tiles that are already running (forward progress of the look-back)
// Cross-tile dependencies, both resolved by single-pass decoupled look-back over 8-byte descriptors that
// carry a per-call generation number (no reset pass between calls):
//   1. the (escaped, in-string, previous-scalar) carry: descriptors hold a SpanFn (stage1_core.cuh),
//      combined by function composition;
//   2. the output cursor + error flags: descriptors hold (count, flags), combined by (+, |).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stage1_core.cuh"

namespace sjb200 {

// simdjson error codes that stage 1 can produce (reference errors.mojo:2-34)
enum : int32_t {
    ERR_SUCCESS = 0,
    ERR_CAPACITY = 1,
    ERR_MEMALLOC = 2,
    ERR_UTF8_ERROR = 11,
    ERR_UNINITIALIZED = 12,
    ERR_EMPTY = 13,
    ERR_UNESCAPED_CHARS = 14,
    ERR_UNCLOSED_STRING = 15,
    ERR_UNEXPECTED_ERROR = 24,
};

struct Stage1Result {       // written by the last tile (mapped pinned host memory or device memory)
    int32_t error;          // verdict, reference priority order
    uint32_t n;             // n_structural_indexes (valid iff n_valid)
    uint32_t n_valid;       // 0 on UNCLOSED_STRING / UNESCAPED_CHARS / CAPACITY: the reference leaves n untouched
    uint32_t n_written;     // entries produced by the indexer (clipped to capacity in memory)
    int32_t utf8_error;     // 1 iff the input is not valid UTF-8 (always reported when validation is compiled in)
    uint32_t final_state;   // packed carry after the last byte (bit1 = still inside a string)
    uint32_t reserved[2];
};

struct Stage1Params {
    const uint8_t *abase;   // input pointer rounded down to 16 bytes
    uint64_t alen;          // mis + len: end of the data in aligned coordinates
    uint32_t mis;           // bytes between abase and the first input byte (0..15)
    uint32_t len;           // input length (< 2^32, reference base.mojo:2)
    uint32_t *out;          // device index array
    uint64_t cap;           // its capacity in entries
    uint64_t *desc1;        // per-tile carry descriptors
    uint64_t *desc2;        // per-tile count descriptors
    uint32_t *ticket;       // tile ticket counter, 0 at launch, reset by the last tile
    Stage1Result *result;
    int32_t *dev_status;    // optional device copy of {error, n} for on-device consumers (NCCL), may be null
    uint32_t gen;           // generation of this call (never 0)
    uint32_t ntiles;
    uint32_t flags;         // bit0: fold the UTF-8 verdict into the error code
};

#if defined(__CUDACC__)

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SJ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SJ_DONE;\n"
        "bra SJ_WAIT;\n"
        "SJ_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t ld_desc(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// descriptor encodings
static constexpr uint32_t ST_AGG = 1, ST_PREFIX = 2;
__device__ __forceinline__ uint64_t d1_make(uint32_t gen, uint32_t status, uint32_t fn) {
    return ((uint64_t)gen << 32) | (status << 7) | (fn & 0x7F);
}
__device__ __forceinline__ uint64_t d2_make(uint32_t gen, uint32_t status, uint32_t err, uint32_t count) {
    return ((uint64_t)(gen & 0x0FFFFFFFu) << 36) | ((uint64_t)status << 34) | ((uint64_t)(err & 3) << 32) | count;
}

__device__ __forceinline__ uint32_t mask_word(uint32_t w, int64_t g, int64_t vbeg, int64_t vend) {
    int64_t lo = vbeg - g, hi = vend - g;
    lo = lo < 0 ? 0 : (lo > 4 ? 4 : lo);
    hi = hi < 0 ? 0 : (hi > 4 ? 4 : hi);
    const uint32_t mh = hi >= 4 ? 0xFFFFFFFFu : ((1u << (8 * (int)hi)) - 1u);
    const uint32_t ml = lo >= 4 ? 0xFFFFFFFFu : ((1u << (8 * (int)lo)) - 1u);
    const uint32_t m = hi > lo ? (mh & ~ml) : 0u;
    return (w & m) | (0x20202020u & ~m);
}

static constexpr uint32_t EF_UNESCAPED = 1, EF_UTF8 = 2;

// ---------------------------------------------------------------------------------------------
// look-backs (executed by warp 0 of a tile)
// ---------------------------------------------------------------------------------------------
// returns the carry state entering `tile` (tile > 0)
__device__ __forceinline__ CarryState lookback_carry(const uint64_t *desc1, uint32_t gen, int tile, int lane) {
    SpanFn acc = SPAN_IDENT;
    int base = tile - 1;
    while (true) {
        const int j = base - lane;
        uint32_t st = ST_PREFIX, fn = span_const(0, 0, 0);
        if (j >= 0) {
            uint64_t d;
            do {
                d = ld_desc(desc1 + j);
            } while ((uint32_t)(d >> 32) != gen);
            st = ((uint32_t)d >> 7) & 3u;
            fn = (uint32_t)d & 0x7Fu;
        }
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, st == ST_PREFIX);
        const int k = pm ? (__ffs((int)pm) - 1) : 31;  // nearest tile whose inclusive state is known
        SpanFn f = lane <= k ? fn : SPAN_IDENT;
        // ordered reduction: lane k is the oldest span, lane 0 the newest
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const SpanFn older = __shfl_down_sync(0xFFFFFFFFu, f, d);
            if (lane + d < 32) f = span_compose(older, f);
        }
        const SpanFn window = __shfl_sync(0xFFFFFFFFu, f, 0);
        acc = span_compose(window, acc);
        if (pm) break;
        base -= 32;
    }
    CarryState zero = {0, 0, 0};
    return span_apply(acc, zero);  // acc starts with a constant function: the argument is irrelevant
}

// returns the number of indexes produced by all tiles before `tile` (tile > 0) and their OR-ed error flags
__device__ __forceinline__ void lookback_count(const uint64_t *desc2, uint32_t gen, int tile, int lane, uint32_t &sum,
                                               uint32_t &err) {
    const uint64_t want = (uint64_t)(gen & 0x0FFFFFFFu);
    sum = 0;
    err = 0;
    int base = tile - 1;
    while (true) {
        const int j = base - lane;
        uint32_t st = ST_PREFIX, cnt = 0, er = 0;
        if (j >= 0) {
            uint64_t d;
            do {
                d = ld_desc(desc2 + j);
            } while ((d >> 36) != want);
            st = (uint32_t)(d >> 34) & 3u;
            er = (uint32_t)(d >> 32) & 3u;
            cnt = (uint32_t)d;
        }
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, st == ST_PREFIX);
        const int k = pm ? (__ffs((int)pm) - 1) : 31;
        const bool take = lane <= k;
        sum += __reduce_add_sync(0xFFFFFFFFu, take ? cnt : 0u);
        err |= __reduce_or_sync(0xFFFFFFFFu, take ? er : 0u);
        if (pm) break;
        base -= 32;
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int WARPS>
struct TileCfg {
    static constexpr int THREADS = WARPS * 32;
    static constexpr int TILE = WARPS * 2048;                 // bytes per tile
    static constexpr int STAGE_CAP = TILE / 2;                // indexes staged in shared memory (density <= 0.5)
    static constexpr int SMEM_BYTES = (STAGE_CAP + 4) * 4;    // >= 16 + TILE
    static_assert(SMEM_BYTES >= 16 + TILE, "staging must cover the input tile");
};

template <int WARPS, bool UTF8>
__global__ void __launch_bounds__(WARPS * 32) stage1_kernel(const Stage1Params P) {
    using Cfg = TileCfg<WARPS>;
    constexpr int TILE = Cfg::TILE;
    extern __shared__ __align__(128) uint8_t smem_raw[];   // [0,16) halo, [16,16+TILE) tile; later: index staging
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_span[WARPS];        // SpanFn of each warp
    __shared__ uint32_t s_span_before[WARPS]; // composition of the warps before it
    __shared__ uint32_t s_carry_in;           // packed CarryState entering the tile
    __shared__ uint32_t s_wcnt[WARPS], s_werr[WARPS], s_woff[WARPS];
    __shared__ uint32_t s_total, s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = smem_u32(&s_mbar);

    // ---- ticket + bulk load -------------------------------------------------------------------
    if (tid == 0) {
        s_tile = atomicAdd(P.ticket, 1u);
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t tb = (int64_t)tile * TILE;               // aligned coordinate of the tile's first byte
    const int64_t alen = (int64_t)P.alen;
    if (tid == 0) {
        int64_t nbytes = alen - tb;
        nbytes = nbytes > TILE ? TILE : nbytes;
        nbytes = (nbytes + 15) & ~15ll;                     // stays inside the last 16-byte line of the data
        const uint32_t halo = tile > 0 ? 16u : 0u;          // 16 bytes of the previous tile (UTF-8 look-behind)
        mbar_expect_tx(bar, (uint32_t)nbytes + halo);
        bulk_load(smem_u32(smem_raw) + 16u - halo, P.abase + tb - halo, (uint32_t)nbytes + halo, bar);
    }
    mbar_wait(bar, 0);

    // ---- phase 1: this lane's 64 bytes -> masks ---------------------------------------------------
    const int off = warp * 2048 + lane * 64;
    const int64_t g0 = tb + off;
    uint32_t w[16];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(smem_raw + 16 + off);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = src[q];
            w[4 * q + 0] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
    }
    uint32_t prev = UTF8 ? *reinterpret_cast<const uint32_t *>(smem_raw + 16 + off - 4) : 0u;
    const bool edge = (tile == 0) || (tb + TILE > alen);
    if (edge) {  // only the first and last tile: bytes outside [mis, alen) read as 0x20 (reference tail padding)
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = mask_word(w[k], g0 + 4 * k, (int64_t)P.mis, alen);
        if (UTF8) prev = (g0 == 0) ? 0x20202020u : mask_word(prev, g0 - 4, (int64_t)P.mis, alen);
    }

    LaneMasks m;
    uint32_t u8err = 0;
    {
        uint32_t pl[8], ph[8];
        bitplanes32(w, pl);
        bitplanes32(w + 8, ph);
        Classes32 cl, ch;
        Utf8Pre32 ul, uh;
        classify32<UTF8>(pl, cl, ul);
        classify32<UTF8>(ph, ch, uh);
        m.bs = join64(cl.bs, ch.bs);
        m.rq = join64(cl.rq, ch.rq);
        m.op = join64(cl.op, ch.op);
        m.ws = join64(cl.ws, ch.ws);
        m.ctl = join64(cl.ctl, ch.ctl);
        if (UTF8) {
            // whole-warp fast path: nothing >= 0x80 in these 2 KiB nor in the 4 bytes before each chunk
            const bool any_hi = ((ul.hi | uh.hi) != 0) || ((prev & 0x80808080u) != 0);
            if (__any_sync(0xFFFFFFFFu, any_hi)) {
                const Utf8Carry uc = utf8_carry_from_prev_word(prev);
                uint32_t tail_must;
                const uint64_t ue = utf8_errors64(ul, uh, uc, &tail_must);
                u8err = (ue != 0) || (g0 + 64 == alen && tail_must != 0);
            }
        }
    }
    // in-warp escape resolution (assuming the warp itself starts unescaped)
    const uint32_t bA = __ballot_sync(0xFFFFFFFFu, lane_all_backslash(m.bs));
    const uint32_t bO = __ballot_sync(0xFFFFFFFFu, lane_trailing_run_parity(m.bs));
    {
        bool lead;
        const uint32_t e_in = warp_lane_e_in(bA, bO, lane, &lead);
        lane_resolve_quotes(m, e_in, lead);
    }
    uint32_t bPB = __ballot_sync(0xFFFFFFFFu, (m.ps >> 63) != 0);
    uint32_t bNQ = __ballot_sync(0xFFFFFFFFu, (lane_nonquote_scalar(m) >> 63) != 0);
    {
        const uint32_t bFQ = __ballot_sync(0xFFFFFFFFu, m.flipq != 0);
        const uint32_t bFQ63 = __ballot_sync(0xFFFFFFFFu, (m.flipq >> 63) != 0);
        if (lane == 0) s_span[warp] = warp_span(bA, bO, bPB, bNQ, bFQ, bFQ63);
    }
    __syncthreads();  // A: warp spans visible; every lane has its bytes in registers (shared input is dead)

    // ---- carry look-back (warp 0) -------------------------------------------------------------------
    if (warp == 0) {
        SpanFn f = lane < WARPS ? s_span[lane] : SPAN_IDENT;
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const SpanFn older = __shfl_up_sync(0xFFFFFFFFu, f, d);
            if (lane >= d) f = span_compose(older, f);
        }
        SpanFn before = __shfl_up_sync(0xFFFFFFFFu, f, 1);
        if (lane == 0) before = SPAN_IDENT;
        if (lane < WARPS) s_span_before[lane] = before;
        const SpanFn tile_fn = __shfl_sync(0xFFFFFFFFu, f, WARPS - 1);
        CarryState cin = {0, 0, 0};
        if (tile > 0) {
            if (lane == 0) st_desc(P.desc1 + tile, d1_make(P.gen, ST_AGG, tile_fn));
            cin = lookback_carry(P.desc1, P.gen, tile, lane);
        }
        const CarryState cout = span_apply(tile_fn, cin);
        if (lane == 0) {
            st_desc(P.desc1 + tile, d1_make(P.gen, ST_PREFIX, span_const(cout.e, cout.s, cout.p)));
            s_carry_in = carry_pack(cin) | (carry_pack(cout) << 8);
        }
    }
    __syncthreads();  // B: carry entering the tile known

    // ---- phase 2: exact carries -> structurals -> counts -------------------------------------------
    const CarryState tile_in = carry_unpack(s_carry_in);
    const CarryState cw = span_apply(s_span_before[warp], tile_in);
    if (cw.e) {  // rare: the warp's first byte is escaped by the previous warp/tile
        lane_apply_escape_carry(m);
        bPB = __ballot_sync(0xFFFFFFFFu, (m.ps >> 63) != 0);
        bNQ = __ballot_sync(0xFFFFFFFFu, (lane_nonquote_scalar(m) >> 63) != 0);
    }
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t s_in = cw.s ^ ((uint32_t)__popc(bPB & lt) & 1u);
    const uint32_t p_in = lane ? ((bNQ >> (lane - 1)) & 1u) : cw.p;
    const LaneOut lo = lane_structurals(m, s_in, p_in);
    const uint32_t st_lo = (uint32_t)lo.structural, st_hi = (uint32_t)(lo.structural >> 32);
    const uint32_t cnt = (uint32_t)(__popc(st_lo) + __popc(st_hi));
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - cnt;
    {
        const uint32_t e1 = __ballot_sync(0xFFFFFFFFu, lo.unescaped_err != 0);
        const uint32_t e2 = __ballot_sync(0xFFFFFFFFu, u8err != 0);
        if (lane == 31) {
            s_wcnt[warp] = incl;
            s_werr[warp] = (e1 ? EF_UNESCAPED : 0u) | (e2 ? EF_UTF8 : 0u);
        }
    }
    __syncthreads();  // C

    // ---- count look-back (warp 0), verdict (last tile) ------------------------------------------------
    if (warp == 0) {
        const uint32_t c = lane < WARPS ? s_wcnt[lane] : 0u;
        const uint32_t e = lane < WARPS ? s_werr[lane] : 0u;
        uint32_t ci = c;
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ci, d);
            if (lane >= d) ci += t;
        }
        if (lane < WARPS) s_woff[lane] = ci - c;
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, ci, WARPS - 1);
        uint32_t err = __reduce_or_sync(0xFFFFFFFFu, e);
        uint32_t base = 0;
        if (tile > 0) {
            if (lane == 0) st_desc(P.desc2 + tile, d2_make(P.gen, ST_AGG, err, total));
            uint32_t perr;
            lookback_count(P.desc2, P.gen, tile, lane, base, perr);
            err |= perr;
        }
        if (lane == 0) {
            st_desc(P.desc2 + tile, d2_make(P.gen, ST_PREFIX, err, base + total));
            s_total = total;
            s_base = base;
            if (tile == (int)P.ntiles - 1) {
                // finish(): reference json_structural_indexer.mojo:147-186, same priority order
                const CarryState cout = carry_unpack(s_carry_in >> 8);
                const uint64_t n = (uint64_t)base + total;
                Stage1Result r;
                r.n = (uint32_t)n;
                r.n_written = (uint32_t)n;
                r.n_valid = 0;
                r.utf8_error = (err & EF_UTF8) ? 1 : 0;
                r.final_state = carry_pack(cout);
                r.reserved[0] = r.reserved[1] = 0;
                if (cout.s) {
                    r.error = ERR_UNCLOSED_STRING;
                } else if (err & EF_UNESCAPED) {
                    r.error = ERR_UNESCAPED_CHARS;
                } else if (n + 3 > P.cap) {
                    r.error = ERR_CAPACITY;
                } else {
                    r.n_valid = 1;
                    P.out[n] = P.len;      // trailer: len, len, 0 (:167-173)
                    P.out[n + 1] = P.len;
                    P.out[n + 2] = 0;
                    if (n == 0) r.error = ERR_EMPTY;
                    else if ((P.flags & 1u) && (err & EF_UTF8)) r.error = ERR_UTF8_ERROR;
                    else r.error = ERR_SUCCESS;
                }
                *P.result = r;
                if (P.dev_status) {
                    P.dev_status[0] = r.error;
                    P.dev_status[1] = (int32_t)(r.n_valid ? r.n : 0u);
                }
                *P.ticket = 0;  // every tile has drawn its ticket by now
            }
        }
    }
    __syncthreads();  // D

    // ---- flatten: bitmask -> ascending uint32 indexes (BitIndexer.write, :46-58) -----------------------
    const uint32_t total = s_total, base = s_base;
    const uint32_t my = s_woff[warp] + excl;             // rank of this lane's first index inside the tile
    const uint32_t v0 = (uint32_t)(g0 - (int64_t)P.mis); // index value of bit 0 of the chunk
    if (total <= (uint32_t)Cfg::STAGE_CAP) {
        uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw);
        // keep shared and global 16-byte phases equal (the output pointer itself may be only 4-byte aligned)
        const uint32_t a = (base + (uint32_t)((reinterpret_cast<uintptr_t>(P.out) >> 2) & 3u)) & 3u;
        uint32_t o = a + my;
        uint32_t bits = st_lo;
        while (bits) {
            stage[o++] = v0 + (uint32_t)(__ffs((int)bits) - 1);
            bits &= bits - 1;
        }
        bits = st_hi;
        while (bits) {
            stage[o++] = v0 + 32u + (uint32_t)(__ffs((int)bits) - 1);
            bits &= bits - 1;
        }
        __syncthreads();  // E
        // coalesced copy-out: vector v holds staged entries [4v, 4v+4) = global entries gbase + 4v ..
        const int64_t gbase = (int64_t)base - (int64_t)a;  // out + gbase is 16-byte aligned; may be negative
        const uint32_t end = a + total;
        const uint32_t nvec = (end + 3u) >> 2;
        for (uint32_t v = tid; v < nvec; v += Cfg::THREADS) {
            const uint4 q = reinterpret_cast<const uint4 *>(stage)[v];
            const uint32_t j = 4u * v;
            const int64_t g = gbase + (int64_t)j;
            if (j >= a && j + 4u <= end && (uint64_t)(g + 4) <= P.cap) {
                *reinterpret_cast<uint4 *>(P.out + g) = q;
            } else {
                const uint32_t vals[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (j + i >= a && j + i < end && (uint64_t)(g + i) < P.cap) P.out[g + i] = vals[i];
            }
        }
    } else {
        // very dense tile (> 0.5 structurals per byte): write straight to global memory
        uint64_t o = (uint64_t)base + my;
        uint32_t bits = st_lo;
        while (bits) {
            if (o < P.cap) P.out[o] = v0 + (uint32_t)(__ffs((int)bits) - 1);
            o++;
            bits &= bits - 1;
        }
        bits = st_hi;
        while (bits) {
            if (o < P.cap) P.out[o] = v0 + 32u + (uint32_t)(__ffs((int)bits) - 1);
            o++;
            bits &= bits - 1;
        }
    }
}

#endif  // __CUDACC__

}  // namespace sjb200
