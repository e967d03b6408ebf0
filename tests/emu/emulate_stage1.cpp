// tests/emu/emulate_stage1.cpp -- TEST INFRASTRUCTURE.
//
// Host-side lane-by-lane emulation of the CUDA stage-1 kernel's phase structure, built from the very
// same header (mojo_simdjson_b200/csrc/stage1_core.cuh) the kernel compiles.  It lets the CPU test suite
// check every bit trick (byte transpose, bit-plane butterfly, plane classifier, bit-sliced UTF-8, escape
// carry algebra, span composition, edge masking, index arithmetic) against the oracle with no GPU.
// Warp collectives are emulated by looping over the 32 lanes between phases; the decoupled look-back is
// emulated by carrying the composed state sequentially from tile to tile.  It is NOT a CPU fallback: the
// product library never links it.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../mojo_simdjson_b200/csrc/stage1_core.cuh"

using namespace sjb200;

namespace {

uint32_t mask_word(uint32_t w, int64_t g, int64_t vbeg, int64_t vend) {
    // keep the bytes of w whose aligned coordinate g+k lies in [vbeg, vend); others become 0x20
    int64_t lo = vbeg - g, hi = vend - g;
    if (lo < 0) lo = 0;
    if (lo > 4) lo = 4;
    if (hi < 0) hi = 0;
    if (hi > 4) hi = 4;
    uint32_t mh = hi >= 4 ? 0xFFFFFFFFu : ((1u << (8 * hi)) - 1u);
    uint32_t ml = lo >= 4 ? 0xFFFFFFFFu : ((1u << (8 * lo)) - 1u);
    uint32_t m = hi > lo ? (mh & ~ml) : 0u;
    return (w & m) | (0x20202020u & ~m);
}

struct Lane {
    LaneMasks m;
    bool all;
    uint32_t tpar;
    uint32_t u8err;
    uint32_t e_in;
    bool lead;
    LaneOut out;
};

}  // namespace

extern "C" int32_t emu_stage1(const uint8_t *buf, uint64_t len, uint32_t mis, int warps, uint32_t *out, uint64_t cap,
                              uint32_t *n_out, uint64_t *n_written_out, int32_t *utf8_err_out, uint32_t flags) {
    if (n_written_out) *n_written_out = 0;
    if (utf8_err_out) *utf8_err_out = 0;
    if (len > 0xFFFFFFFFull) return 1;
    if (len == 0) return 13;
    const int64_t TILE = (int64_t)warps * 2048;
    const int64_t alen = (int64_t)mis + (int64_t)len;
    const int64_t ntiles = (alen + TILE - 1) / TILE;
    // aligned image of memory: hostile bytes wherever the kernel must not look
    std::vector<uint8_t> mem((size_t)(ntiles * TILE + 64), 0);
    for (size_t i = 0; i < mem.size(); i++) mem[i] = (i & 1) ? 0x22 : 0xF4;
    memcpy(mem.data() + mis, buf, (size_t)len);

    CarryState carry = {0, 0, 0};
    uint64_t total = 0;
    uint32_t err_unescaped = 0, err_utf8 = 0;
    std::vector<uint8_t> smem((size_t)TILE + 16);

    for (int64_t tile = 0; tile < ntiles; tile++) {
        const int64_t tb = tile * TILE;
        // what the bulk copy leaves in shared memory: 16 halo bytes (tile > 0) + the tile, stale bytes past
        // the rounded-up end of the data
        for (size_t i = 0; i < smem.size(); i++) smem[i] = 0xF4;  // stale
        int64_t nbytes = alen - tb;
        if (nbytes > TILE) nbytes = TILE;
        nbytes = (nbytes + 15) & ~15ll;
        memcpy(smem.data() + 16, mem.data() + tb, (size_t)nbytes);
        if (tile > 0) memcpy(smem.data(), mem.data() + tb - 16, 16);
        const bool edge = (tile == 0) || (tb + TILE > alen);

        std::vector<Lane> L((size_t)warps * 32);
        std::vector<SpanFn> wspan((size_t)warps);
        std::vector<uint32_t> bA(warps), bO(warps), bPB(warps), bNQ(warps);
        // ---- phase 1: local classification, in-warp escape resolution, warp span ----
        for (int w = 0; w < warps; w++) {
            uint32_t A = 0, O = 0;
            for (int l = 0; l < 32; l++) {
                Lane &ln = L[(size_t)w * 32 + l];
                const int64_t off = (int64_t)w * 2048 + (int64_t)l * 64;
                const int64_t g0 = tb + off;
                uint32_t words[16], prev;
                memcpy(words, smem.data() + 16 + off, 64);
                memcpy(&prev, smem.data() + 16 + off - 4, 4);
                if (edge) {
                    for (int k = 0; k < 16; k++) words[k] = mask_word(words[k], g0 + 4 * k, mis, alen);
                    prev = (g0 == 0) ? 0x20202020u : mask_word(prev, g0 - 4, mis, alen);
                }
                uint32_t pl[8], ph[8];
                bitplanes32(words, pl);
                bitplanes32(words + 8, ph);
                Classes32 cl, ch;
                Utf8Pre32 ul, uh;
                classify32<true>(pl, cl, ul);
                classify32<true>(ph, ch, uh);
                ln.m.bs = join64(cl.bs, ch.bs);
                ln.m.rq = join64(cl.rq, ch.rq);
                ln.m.op = join64(cl.op, ch.op);
                ln.m.ws = join64(cl.ws, ch.ws);
                ln.m.ctl = join64(cl.ctl, ch.ctl);
                Utf8Carry uc = utf8_carry_from_prev_word(prev);
                uint32_t tail_must;
                uint64_t ue = utf8_errors64(ul, uh, uc, &tail_must);
                ln.u8err = (ue != 0) || (g0 + 64 == alen && tail_must != 0);
                ln.all = lane_all_backslash(ln.m.bs);
                ln.tpar = lane_trailing_run_parity(ln.m.bs);
                A |= (uint32_t)ln.all << l;
                O |= ln.tpar << l;
            }
            uint32_t PB = 0, NQ = 0, FQ = 0, FQ63 = 0;
            for (int l = 0; l < 32; l++) {
                Lane &ln = L[(size_t)w * 32 + l];
                ln.e_in = warp_lane_e_in(A, O, l, &ln.lead);
                lane_resolve_quotes(ln.m, ln.e_in, ln.lead);
                PB |= (uint32_t)(ln.m.ps >> 63) << l;
                NQ |= (uint32_t)(lane_nonquote_scalar(ln.m) >> 63) << l;
                FQ |= (uint32_t)(ln.m.flipq != 0) << l;
                FQ63 |= (uint32_t)(ln.m.flipq >> 63) << l;
            }
            bA[w] = A; bO[w] = O; bPB[w] = PB; bNQ[w] = NQ;
            wspan[w] = warp_span(A, O, PB, NQ, FQ, FQ63);
        }
        // ---- tile span, look-back (sequential here) ----
        std::vector<SpanFn> wprefix((size_t)warps);
        SpanFn acc = SPAN_IDENT;
        for (int w = 0; w < warps; w++) {
            wprefix[w] = acc;
            acc = span_compose(acc, wspan[w]);
        }
        const CarryState tile_in = carry;
        carry = span_apply(acc, tile_in);
        // ---- phase 2: exact carries, structurals, counts, extraction ----
        for (int w = 0; w < warps; w++) {
            const CarryState cw = span_apply(wprefix[w], tile_in);
            uint32_t PB = bPB[w], NQ = bNQ[w];
            if (cw.e) {
                PB = 0; NQ = 0;
                for (int l = 0; l < 32; l++) {
                    Lane &ln = L[(size_t)w * 32 + l];
                    lane_apply_escape_carry(ln.m);
                    PB |= (uint32_t)(ln.m.ps >> 63) << l;
                    NQ |= (uint32_t)(lane_nonquote_scalar(ln.m) >> 63) << l;
                }
            }
            for (int l = 0; l < 32; l++) {
                Lane &ln = L[(size_t)w * 32 + l];
                const uint32_t lt = (1u << l) - 1u;
                const uint32_t s_in = cw.s ^ ((uint32_t)popc32(PB & lt) & 1u);
                const uint32_t p_in = l ? ((NQ >> (l - 1)) & 1u) : cw.p;
                ln.out = lane_structurals(ln.m, s_in, p_in);
                err_unescaped |= ln.out.unescaped_err;
                err_utf8 |= ln.u8err;
                uint64_t st = ln.out.structural;
                const int64_t g0 = tb + (int64_t)w * 2048 + (int64_t)l * 64;
                while (st) {
                    int bit = __builtin_ctzll(st);
                    st &= st - 1;
                    uint32_t v = (uint32_t)(g0 + bit - (int64_t)mis);
                    if (total < cap) out[total] = v;
                    total++;
                }
            }
        }
    }
    if (n_written_out) *n_written_out = total;
    if (utf8_err_out) *utf8_err_out = (int32_t)err_utf8;
    if (carry.s) return 15;
    if (err_unescaped) return 14;
    if (total + 3 > cap) return 1;
    if (n_out) *n_out = (uint32_t)total;
    out[total] = (uint32_t)len;
    out[total + 1] = (uint32_t)len;
    out[total + 2] = 0;
    if (total == 0) return 13;
    if ((flags & 1u) && err_utf8) return 11;
    return 0;
}

// direct checks of the transposition and classifier on arbitrary 32 bytes
extern "C" void emu_bitplanes32(const uint8_t *bytes32, uint32_t *planes8) {
    uint32_t w[8];
    memcpy(w, bytes32, 32);
    bitplanes32(w, planes8);
}

extern "C" uint32_t emu_span_compose(uint32_t older, uint32_t newer) { return span_compose(older, newer); }
