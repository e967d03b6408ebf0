/* stage1_simd.c -- TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * "cpu_simd": what a well-written CPU stage 1 costs -- the same specification as oracle/stage1_oracle.c (SURVEY.md
 * section 3.2, i.e. the reference's generic/stage1/json_structural_indexer.mojo:81-186 and callees) implemented the way
 * upstream simdjson implements it on x86: 64-byte blocks in one AVX-512 register (or two AVX2 registers), byte compares
 * into 64-bit masks, the reference's own nibble-table classifier as vpshufb lookups (haswell.mojo:22-74, including the
 * 0x0C / 0x1A artefact), the add-carry escape scanner (json_escape_scanner.mojo:18-45), a carry-less multiply for the
 * in-string prefix XOR (the reference's dead prefix_xor_old, stuff.mojo:12-17, done right), Keiser-Lemire lookup-table
 * UTF-8 validation and index extraction eight at a time.
 *
 * bench.py reports it in `cpu_baseline` beside the faithful port (oracle_stage1_ref), on one thread and on all cores, so
 * that the GPU numbers are not only compared with the reference's 64-iteration prefix_xor.  tests/test_oracle.py checks
 * it bit for bit against the oracle on the whole corpus.  Only tests/, smoke() and bench.py's CPU legs may load it.
 */
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

enum { SUCCESS = 0, CAPACITY = 1, UTF8_ERROR = 11, EMPTY = 13, UNESCAPED_CHARS = 14, UNCLOSED_STRING = 15 };

typedef struct {
    uint64_t bs, quote_raw, op, ws, ctl;
} BlockMasks;

typedef struct {
    uint64_t next_is_escaped, prev_in_string, prev_scalar, unescaped_err;
    int utf8_err;
} Carry;

static inline uint64_t escaped_mask(uint64_t bs, uint64_t *next_is_escaped) {
    /* json_escape_scanner.mojo:18-45 */
    const uint64_t ODD = 0xAAAAAAAAAAAAAAAAull;
    const uint64_t potential = bs & ~*next_is_escaped;
    const uint64_t maybe = potential << 1;
    const uint64_t maybe_odd = maybe | ODD;
    const uint64_t even_series = maybe_odd - potential;
    const uint64_t escape_and_terminal = even_series ^ ODD;
    const uint64_t escaped = escape_and_terminal ^ (bs | *next_is_escaped);
    const uint64_t escape = escape_and_terminal & bs;
    *next_is_escaped = escape >> 63;
    return escaped;
}

/* index extraction, eight per trip (BitIndexer.write, json_structural_indexer.mojo:46-58: same values, same order) */
static inline uint64_t flatten_bits(uint32_t *out, uint64_t pos, uint64_t cap, uint32_t base, uint64_t bits) {
    if (bits == 0) return pos;
    const int cnt = (int)__builtin_popcountll(bits);
    if (pos + (uint64_t)cnt + 8 <= cap) {
        uint32_t *o = out + pos;
        for (int i = 0; i < cnt; i += 8) {
            for (int k = 0; k < 8; k++) {
                o[i + k] = base + (uint32_t)__builtin_ctzll(bits);
                bits &= bits - 1;
            }
        }
    } else {
        uint64_t p = pos;
        while (bits) {
            if (p < cap) out[p] = base + (uint32_t)__builtin_ctzll(bits);
            p++;
            bits &= bits - 1;
        }
    }
    return pos + (uint64_t)cnt;
}

static inline uint64_t finish_block(const BlockMasks *m, Carry *c, uint64_t in_string_carryless /* prefix xor of quote */, uint64_t quote) {
    const uint64_t in_string = in_string_carryless ^ c->prev_in_string;
    c->prev_in_string = (uint64_t)((int64_t)in_string >> 63);
    const uint64_t scalar = ~(m->op | m->ws);
    const uint64_t nqs = scalar & ~quote;
    const uint64_t follows = (nqs << 1) | c->prev_scalar;
    c->prev_scalar = nqs >> 63;
    const uint64_t string_tail = in_string ^ quote;
    c->unescaped_err |= m->ctl & in_string;
    return (m->op | (scalar & ~follows)) & ~string_tail;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* AVX-512 (BW + VBMI not required) + (V)PCLMULQDQ                                                                      */
/* ------------------------------------------------------------------------------------------------------------------ */
#define T512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,pclmul,bmi,bmi2,popcnt,lzcnt")))

T512 static inline uint64_t prefix_xor_clmul(uint64_t x) {
    const __m128i all = _mm_set1_epi8((char)0xFF);
    const __m128i r = _mm_clmulepi64_si128(_mm_set_epi64x(0, (long long)x), all, 0);
    return (uint64_t)_mm_cvtsi128_si64(r);
}

T512 static inline void classify512(__m512i in, BlockMasks *m) {
    /* the reference's two 16-entry low-nibble tables (haswell.mojo:23-64), replicated per 128-bit lane */
    const __m512i ws_tab = _mm512_broadcast_i32x4(_mm_setr_epi8(' ', 100, 100, 100, 17, 100, 113, 2, 100, '\t', '\n', 112, 100, '\r', 100, 100));
    const __m512i op_tab = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, ':', '{', ',', '}', 0, 0));
    const __m512i lo = _mm512_and_si512(in, _mm512_set1_epi8(0x0F));
    m->ws = _mm512_cmpeq_epi8_mask(in, _mm512_shuffle_epi8(ws_tab, lo));
    const __m512i curlified = _mm512_or_si512(in, _mm512_set1_epi8(0x20));
    m->op = _mm512_cmpeq_epi8_mask(curlified, _mm512_shuffle_epi8(op_tab, lo));
    m->bs = _mm512_cmpeq_epi8_mask(in, _mm512_set1_epi8('\\'));
    m->quote_raw = _mm512_cmpeq_epi8_mask(in, _mm512_set1_epi8('"'));
    m->ctl = _mm512_cmple_epu8_mask(in, _mm512_set1_epi8(0x1F));
}

/* Keiser-Lemire: error bits of one 64-byte block given the previous block */
T512 static inline __m512i utf8_check512(__m512i in, __m512i prev_in) {
    const __m512i shifted = _mm512_permutex2var_epi64(in, _mm512_set_epi64(5, 4, 3, 2, 1, 0, 15, 14), prev_in);
    const __m512i prev1 = _mm512_alignr_epi8(in, shifted, 15);
    const __m512i prev2 = _mm512_alignr_epi8(in, shifted, 14);
    const __m512i prev3 = _mm512_alignr_epi8(in, shifted, 13);
    const __m512i b1h = _mm512_broadcast_i32x4(_mm_setr_epi8(0x02, 0x02, 0x02, 0x02, 0x02, 0x02, 0x02, 0x02, (char)0x80, (char)0x80, (char)0x80, (char)0x80, 0x21, 0x01, 0x15, 0x49));
    const __m512i b1l = _mm512_broadcast_i32x4(_mm_setr_epi8((char)0xE7, (char)0xA3, (char)0x83, (char)0x83, (char)0x8B, (char)0xCB, (char)0xCB, (char)0xCB, (char)0xCB, (char)0xCB, (char)0xCB,
                                                             (char)0xCB, (char)0xCB, (char)0xDB, (char)0xCB, (char)0xCB));
    const __m512i b2h = _mm512_broadcast_i32x4(_mm_setr_epi8(0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, (char)0xE6, (char)0xAE, (char)0xBA, (char)0xBA, 0x01, 0x01, 0x01, 0x01));
    const __m512i nib = _mm512_set1_epi8(0x0F);
    const __m512i p1hi = _mm512_and_si512(_mm512_srli_epi16(prev1, 4), nib);
    const __m512i p1lo = _mm512_and_si512(prev1, nib);
    const __m512i c_hi = _mm512_and_si512(_mm512_srli_epi16(in, 4), nib);
    const __m512i sc = _mm512_and_si512(_mm512_and_si512(_mm512_shuffle_epi8(b1h, p1hi), _mm512_shuffle_epi8(b1l, p1lo)), _mm512_shuffle_epi8(b2h, c_hi));
    /* must be a 2nd / 3rd continuation: prev2 >= 0xE0 or prev3 >= 0xF0 */
    const __mmask64 m23 = _mm512_cmpge_epu8_mask(prev2, _mm512_set1_epi8((char)0xE0)) | _mm512_cmpge_epu8_mask(prev3, _mm512_set1_epi8((char)0xF0));
    const __m512i must23 = _mm512_maskz_set1_epi8(m23, (char)0x80);
    return _mm512_xor_si512(sc, must23);
}

/* index extraction with vpcompressd: 16 candidate positions per instruction, no data-dependent loop (the icelake way) */
T512 static inline uint64_t flatten_bits512(uint32_t *out, uint64_t pos, uint64_t cap, uint32_t base, uint64_t bits) {
    if (bits == 0) return pos;
    if (pos + 64 > cap) return flatten_bits(out, pos, cap, base, bits);   /* near the end of the array: exact writes only */
    const __m512i lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    __m512i idx = _mm512_add_epi32(_mm512_set1_epi32((int)base), lane);
    const __m512i step = _mm512_set1_epi32(16);
    for (int q = 0; q < 4; q++) {
        const __mmask16 m = (__mmask16)(bits >> (16 * q));
        _mm512_storeu_si512((void *)(out + pos), _mm512_maskz_compress_epi32(m, idx));
        pos += (uint64_t)__builtin_popcount((unsigned)m);
        idx = _mm512_add_epi32(idx, step);
    }
    return pos;
}

T512 static int32_t stage1_avx512(const uint8_t *buf, uint64_t len, uint32_t *out, uint64_t cap, uint32_t *n_out, uint64_t *n_written, int32_t *utf8_err,
                                  uint32_t flags) {
    Carry c = {0, 0, 0, 0, 0};
    uint64_t pos = 0;
    __m512i prev_in = _mm512_setzero_si512();
    __m512i u8acc = _mm512_setzero_si512();
    const int want_u8 = utf8_err != NULL || (flags & 1u);
    const uint64_t nblocks = (len + 63) / 64;
    for (uint64_t b = 0; b < nblocks; b++) {
        __m512i in;
        if ((b + 1) * 64 <= len) {
            in = _mm512_loadu_si512((const void *)(buf + b * 64));
        } else { /* tail: 0x20 padding (json_structural_indexer.mojo:103, buf_block_reader.mojo:28-36) */
            uint8_t tmp[64];
            memset(tmp, 0x20, 64);
            memcpy(tmp, buf + b * 64, (size_t)(len - b * 64));
            in = _mm512_loadu_si512((const void *)tmp);
        }
        BlockMasks m;
        classify512(in, &m);
        const uint64_t escaped = escaped_mask(m.bs, &c.next_is_escaped);
        const uint64_t quote = m.quote_raw & ~escaped;
        const uint64_t structural = finish_block(&m, &c, prefix_xor_clmul(quote), quote);
        if (want_u8) {
            if (_mm512_movepi8_mask(_mm512_or_si512(in, prev_in)) != 0) u8acc = _mm512_or_si512(u8acc, utf8_check512(in, prev_in));
            prev_in = in;
        }
        pos = flatten_bits512(out, pos, cap, (uint32_t)(b * 64), structural);
    }
    int u8 = 0;
    if (want_u8) {
        u8 = _mm512_test_epi8_mask(u8acc, u8acc) != 0;
        if (len % 64 == 0 && len >= 1) { /* no padding byte follows: a sequence cut off by the end of input */
            const uint8_t b1 = buf[len - 1], b2 = len >= 2 ? buf[len - 2] : 0, b3 = len >= 3 ? buf[len - 3] : 0;
            if (b1 >= 0xC0 || b2 >= 0xE0 || b3 >= 0xF0) u8 = 1;
        }
    }
    if (utf8_err) *utf8_err = u8;
    if (n_written) *n_written = pos;
    /* finish(): json_structural_indexer.mojo:147-186, same order */
    if (c.prev_in_string) return UNCLOSED_STRING;
    if (c.unescaped_err) return UNESCAPED_CHARS;
    if (pos + 3 > cap) return CAPACITY;
    if (n_out) *n_out = (uint32_t)pos;
    out[pos] = (uint32_t)len;
    out[pos + 1] = (uint32_t)len;
    out[pos + 2] = 0;
    if (pos == 0) return EMPTY;
    if ((flags & 1u) && u8) return UTF8_ERROR;
    return SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* AVX2 + PCLMULQDQ (hosts without AVX-512)                                                                             */
/* ------------------------------------------------------------------------------------------------------------------ */
#define T256 __attribute__((target("avx2,pclmul,bmi,bmi2,popcnt,lzcnt")))

T256 static inline uint64_t prefix_xor_clmul2(uint64_t x) {
    const __m128i all = _mm_set1_epi8((char)0xFF);
    return (uint64_t)_mm_cvtsi128_si64(_mm_clmulepi64_si128(_mm_set_epi64x(0, (long long)x), all, 0));
}
T256 static inline uint64_t mm2(__m256i lo, __m256i hi) {
    return (uint64_t)(uint32_t)_mm256_movemask_epi8(lo) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(hi) << 32);
}
T256 static inline void classify256(__m256i a, __m256i b, BlockMasks *m) {
    const __m256i ws_tab = _mm256_broadcastsi128_si256(_mm_setr_epi8(' ', 100, 100, 100, 17, 100, 113, 2, 100, '\t', '\n', 112, 100, '\r', 100, 100));
    const __m256i op_tab = _mm256_broadcastsi128_si256(_mm_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, ':', '{', ',', '}', 0, 0));
    const __m256i nib = _mm256_set1_epi8(0x0F), c20 = _mm256_set1_epi8(0x20);
    const __m256i la = _mm256_and_si256(a, nib), lb = _mm256_and_si256(b, nib);
    m->ws = mm2(_mm256_cmpeq_epi8(a, _mm256_shuffle_epi8(ws_tab, la)), _mm256_cmpeq_epi8(b, _mm256_shuffle_epi8(ws_tab, lb)));
    m->op = mm2(_mm256_cmpeq_epi8(_mm256_or_si256(a, c20), _mm256_shuffle_epi8(op_tab, la)), _mm256_cmpeq_epi8(_mm256_or_si256(b, c20), _mm256_shuffle_epi8(op_tab, lb)));
    m->bs = mm2(_mm256_cmpeq_epi8(a, _mm256_set1_epi8('\\')), _mm256_cmpeq_epi8(b, _mm256_set1_epi8('\\')));
    m->quote_raw = mm2(_mm256_cmpeq_epi8(a, _mm256_set1_epi8('"')), _mm256_cmpeq_epi8(b, _mm256_set1_epi8('"')));
    const __m256i c1f = _mm256_set1_epi8(0x1F);
    m->ctl = mm2(_mm256_cmpeq_epi8(_mm256_max_epu8(a, c1f), c1f), _mm256_cmpeq_epi8(_mm256_max_epu8(b, c1f), c1f));
}
T256 static inline __m256i utf8_check256(__m256i in, __m256i prev_in) {
    const __m256i shifted = _mm256_permute2x128_si256(prev_in, in, 0x21);
    const __m256i prev1 = _mm256_alignr_epi8(in, shifted, 15), prev2 = _mm256_alignr_epi8(in, shifted, 14), prev3 = _mm256_alignr_epi8(in, shifted, 13);
    const __m256i b1h = _mm256_broadcastsi128_si256(_mm_setr_epi8(0x02, 0x02, 0x02, 0x02, 0x02, 0x02, 0x02, 0x02, (char)0x80, (char)0x80, (char)0x80, (char)0x80, 0x21, 0x01, 0x15, 0x49));
    const __m256i b1l = _mm256_broadcastsi128_si256(_mm_setr_epi8((char)0xE7, (char)0xA3, (char)0x83, (char)0x83, (char)0x8B, (char)0xCB, (char)0xCB, (char)0xCB, (char)0xCB, (char)0xCB,
                                                                 (char)0xCB, (char)0xCB, (char)0xCB, (char)0xDB, (char)0xCB, (char)0xCB));
    const __m256i b2h = _mm256_broadcastsi128_si256(_mm_setr_epi8(0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, (char)0xE6, (char)0xAE, (char)0xBA, (char)0xBA, 0x01, 0x01, 0x01, 0x01));
    const __m256i nib = _mm256_set1_epi8(0x0F);
    const __m256i sc = _mm256_and_si256(_mm256_and_si256(_mm256_shuffle_epi8(b1h, _mm256_and_si256(_mm256_srli_epi16(prev1, 4), nib)), _mm256_shuffle_epi8(b1l, _mm256_and_si256(prev1, nib))),
                                        _mm256_shuffle_epi8(b2h, _mm256_and_si256(_mm256_srli_epi16(in, 4), nib)));
    /* saturating subtraction: non-zero where prev2 >= 0xE0 or prev3 >= 0xF0 */
    const __m256i is3 = _mm256_subs_epu8(prev2, _mm256_set1_epi8((char)0xDF)), is4 = _mm256_subs_epu8(prev3, _mm256_set1_epi8((char)0xEF));
    const __m256i must23 = _mm256_and_si256(_mm256_cmpgt_epi8(_mm256_or_si256(is3, is4), _mm256_setzero_si256()), _mm256_set1_epi8((char)0x80));
    return _mm256_xor_si256(sc, must23);
}

T256 static int32_t stage1_avx2(const uint8_t *buf, uint64_t len, uint32_t *out, uint64_t cap, uint32_t *n_out, uint64_t *n_written, int32_t *utf8_err,
                                uint32_t flags) {
    Carry c = {0, 0, 0, 0, 0};
    uint64_t pos = 0;
    __m256i prev_hi = _mm256_setzero_si256(), u8acc = _mm256_setzero_si256();
    const int want_u8 = utf8_err != NULL || (flags & 1u);
    const uint64_t nblocks = (len + 63) / 64;
    for (uint64_t b = 0; b < nblocks; b++) {
        uint8_t tmp[64];
        const uint8_t *p = buf + b * 64;
        if ((b + 1) * 64 > len) {
            memset(tmp, 0x20, 64);
            memcpy(tmp, buf + b * 64, (size_t)(len - b * 64));
            p = tmp;
        }
        const __m256i lo = _mm256_loadu_si256((const __m256i *)p), hi = _mm256_loadu_si256((const __m256i *)(p + 32));
        BlockMasks m;
        classify256(lo, hi, &m);
        const uint64_t escaped = escaped_mask(m.bs, &c.next_is_escaped);
        const uint64_t quote = m.quote_raw & ~escaped;
        const uint64_t structural = finish_block(&m, &c, prefix_xor_clmul2(quote), quote);
        if (want_u8) {
            if (_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(lo, hi), prev_hi)) != 0) {
                u8acc = _mm256_or_si256(u8acc, utf8_check256(lo, prev_hi));
                u8acc = _mm256_or_si256(u8acc, utf8_check256(hi, lo));
            }
            prev_hi = hi;
        }
        pos = flatten_bits(out, pos, cap, (uint32_t)(b * 64), structural);
    }
    int u8 = 0;
    if (want_u8) {
        u8 = !_mm256_testz_si256(u8acc, u8acc);
        if (len % 64 == 0 && len >= 1) {
            const uint8_t b1 = buf[len - 1], b2 = len >= 2 ? buf[len - 2] : 0, b3 = len >= 3 ? buf[len - 3] : 0;
            if (b1 >= 0xC0 || b2 >= 0xE0 || b3 >= 0xF0) u8 = 1;
        }
    }
    if (utf8_err) *utf8_err = u8;
    if (n_written) *n_written = pos;
    if (c.prev_in_string) return UNCLOSED_STRING;
    if (c.unescaped_err) return UNESCAPED_CHARS;
    if (pos + 3 > cap) return CAPACITY;
    if (n_out) *n_out = (uint32_t)pos;
    out[pos] = (uint32_t)len;
    out[pos + 1] = (uint32_t)len;
    out[pos + 2] = 0;
    if (pos == 0) return EMPTY;
    if ((flags & 1u) && u8) return UTF8_ERROR;
    return SUCCESS;
}

/* 2 = AVX-512, 1 = AVX2, 0 = neither (then simd_stage1 returns -1: no scalar stand-in here, the oracle has those) */
EXPORT int simd_stage1_level(void) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") &&
        __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("bmi2"))
        return 2;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("bmi2")) return 1;
    return 0;
}

/* Same contract as oracle_stage1_ref (len == 0 -> EMPTY; n_out untouched on the early-return verdicts; utf8_err always
 * computed when the pointer is given).  `level`: 0 = best the CPU has, 1 = force AVX2, 2 = force AVX-512. */
EXPORT int32_t simd_stage1(const uint8_t *buf, uint64_t len, uint32_t *out, uint64_t cap, uint32_t *n_out, uint64_t *n_written, int32_t *utf8_err,
                           uint32_t flags, int level) {
    if (len == 0) {
        if (n_written) *n_written = 0;
        if (utf8_err) *utf8_err = 0;
        return EMPTY;
    }
    if (len > 0xFFFFFFFFull) return CAPACITY;
    const int have = simd_stage1_level();
    if (level == 0) level = have;
    if (level > have || level <= 0) return -1;
    return level == 2 ? stage1_avx512(buf, len, out, cap, n_out, n_written, utf8_err, flags)
                      : stage1_avx2(buf, len, out, cap, n_out, n_written, utf8_err, flags);
}
