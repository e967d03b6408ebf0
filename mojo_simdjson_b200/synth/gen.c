/*
 * gen.c -- deterministic synthetic JSON workloads (SplitMix64), host-only C.
 *
 * There is no network for datasets, so BASELINE.json's configs are made concrete here:
 *   sjb200_gen_twitter_pretty : {"statuses":[...],"search_metadata":{...}}, 2-space pretty printed, exact size
 *                               (config 2: 631,515 bytes, the size of upstream simdjson's twitter.json)
 *   sjb200_gen_status_array   : one top-level array of minified status objects, space padded before the final ']'
 *                               to an exact size (config 3: 2^30 bytes)
 *   sjb200_gen_ndjson         : one minified status object per '\n'-terminated line, exact size (config 4)
 * Each status has ~30 keys, a nested "user" object, "entities" arrays, text with CJK / emoji UTF-8
 * (about 10% of string bytes are multi-byte), and \/ \" \n \uXXXX escapes.
 * Every output is valid JSON / NDJSON and valid UTF-8.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct { uint64_t s; } rng_t;
static uint64_t rnd(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t rnd_below(rng_t *r, uint32_t n) { return (uint32_t)(rnd(r) % n); }

/* ---- growable writer with optional pretty printing ------------------------------------------- */
typedef struct {
    uint8_t *p;
    size_t n, cap;
    int pretty, depth;
    int need_comma[64];
} wr_t;

static void w_reserve(wr_t *w, size_t extra) {
    if (w->n + extra <= w->cap) return;
    size_t c = w->cap ? w->cap * 2 : 4096;
    while (c < w->n + extra) c *= 2;
    w->p = (uint8_t *)realloc(w->p, c);
    w->cap = c;
}
static void w_raw(wr_t *w, const void *s, size_t n) { w_reserve(w, n); memcpy(w->p + w->n, s, n); w->n += n; }
static void w_ch(wr_t *w, char c) { w_raw(w, &c, 1); }
static void w_cstr(wr_t *w, const char *s) { w_raw(w, s, strlen(s)); }
static void w_newline(wr_t *w) {
    if (!w->pretty) return;
    w_ch(w, '\n');
    for (int i = 0; i < w->depth; i++) w_cstr(w, "  ");
}
static void w_value_prefix(wr_t *w) {
    if (w->depth > 0) {
        if (w->need_comma[w->depth]) w_ch(w, ',');
        w->need_comma[w->depth] = 1;
        w_newline(w);
    }
}
static void w_open(wr_t *w, char c) { w_ch(w, c); w->depth++; w->need_comma[w->depth] = 0; }
static void w_close(wr_t *w, char c) {
    int had = w->need_comma[w->depth];
    w->depth--;
    if (had) w_newline(w);
    w_ch(w, c);
}
static void w_key(wr_t *w, const char *k) {
    w_value_prefix(w);
    w_ch(w, '"'); w_cstr(w, k); w_ch(w, '"'); w_ch(w, ':');
    if (w->pretty) w_ch(w, ' ');
}
/* after w_key the value must not emit its own prefix */
static void v_int(wr_t *w, long long v) { char b[32]; snprintf(b, sizeof b, "%lld", v); w_cstr(w, b); }
static void v_u64(wr_t *w, unsigned long long v) { char b[32]; snprintf(b, sizeof b, "%llu", v); w_cstr(w, b); }
static void v_lit(wr_t *w, const char *s) { w_cstr(w, s); }
static void v_qstr(wr_t *w, const char *s) { w_ch(w, '"'); w_cstr(w, s); w_ch(w, '"'); }

/* ---- text ---------------------------------------------------------------------------------- */
static const char *WORDS[] = {"the", "of", "and", "to", "in", "is", "that", "for", "it", "as", "was", "with", "be", "by",
    "on", "not", "he", "this", "are", "or", "his", "from", "at", "which", "but", "have", "an", "had", "they", "you",
    "were", "their", "one", "all", "we", "can", "her", "has", "there", "been", "if", "more", "when", "will", "would",
    "who", "so", "no", "json", "parser", "gpu", "kernel", "blackwell", "stream", "index", "quote", "string", "lol",
    "omg", "today", "tomorrow", "weather", "coffee", "music", "game", "happy", "new", "year", "love", "like", "RT"};
#define NWORDS (sizeof WORDS / sizeof WORDS[0])
/* 3-byte CJK / kana and 4-byte emoji, as UTF-8 */
static const char *CJK[] = {"\xe6\x97\xa5", "\xe6\x9c\xac", "\xe8\xaa\x9e", "\xe3\x81\x82", "\xe3\x81\x84", "\xe3\x81\x86",
    "\xe3\x82\xab", "\xe3\x82\xbf", "\xe4\xb8\xad", "\xe6\x96\x87", "\xe5\xa4\xa7", "\xe5\xad\xa6", "\xe2\x82\xac", "\xc3\xa9",
    "\xc3\xbc", "\xd0\xb6"};
#define NCJK (sizeof CJK / sizeof CJK[0])
static const char *EMOJI[] = {"\xf0\x9f\x98\x80", "\xf0\x9f\x98\x82", "\xf0\x9f\x91\x8d", "\xf0\x9f\x94\xa5", "\xf0\x9f\x8e\x89"};
#define NEMOJI (sizeof EMOJI / sizeof EMOJI[0])

static void v_text(wr_t *w, rng_t *r, int min_words, int max_words) {
    w_ch(w, '"');
    int n = min_words + (int)rnd_below(r, (uint32_t)(max_words - min_words + 1));
    int cjk_mode = rnd_below(r, 100) < 25;  /* some tweets are mostly CJK */
    for (int i = 0; i < n; i++) {
        if (i) w_ch(w, ' ');
        uint32_t k = rnd_below(r, 100);
        if (cjk_mode ? k < 70 : k < 3) {
            int m = 1 + (int)rnd_below(r, 6);
            for (int j = 0; j < m; j++) w_cstr(w, CJK[rnd_below(r, NCJK)]);
        } else if (k < 5) {
            w_cstr(w, EMOJI[rnd_below(r, NEMOJI)]);
        } else if (k < 7) {
            w_cstr(w, "\\\""); w_cstr(w, WORDS[rnd_below(r, NWORDS)]); w_cstr(w, "\\\"");
        } else if (k < 8) {
            w_cstr(w, "\\n");
        } else if (k < 9) {
            char b[16]; snprintf(b, sizeof b, "\\u%04x", 0x2600 + rnd_below(r, 0xff)); w_cstr(w, b);
        } else if (k < 11) {
            w_cstr(w, "http:\\/\\/t.co\\/");
            for (int j = 0; j < 10; j++) w_ch(w, (char)('a' + rnd_below(r, 26)));
        } else if (k < 13) {
            w_ch(w, '@'); w_cstr(w, WORDS[rnd_below(r, NWORDS)]); v_int(w, rnd_below(r, 1000));
        } else if (k < 14) {
            w_ch(w, '#'); w_cstr(w, WORDS[rnd_below(r, NWORDS)]);
        } else if (k < 15) {
            w_cstr(w, "C:\\\\"); w_cstr(w, WORDS[rnd_below(r, NWORDS)]);
        } else {
            w_cstr(w, WORDS[rnd_below(r, NWORDS)]);
        }
    }
    w_ch(w, '"');
}
static void v_url(wr_t *w, rng_t *r) {
    w_cstr(w, "\"http:\\/\\/");
    w_cstr(w, WORDS[rnd_below(r, NWORDS)]);
    w_cstr(w, ".example.com\\/");
    for (int j = 0; j < 8; j++) w_ch(w, (char)('a' + rnd_below(r, 26)));
    w_cstr(w, "\\/img_"); v_int(w, rnd_below(r, 100000)); w_cstr(w, ".png\"");
}
static void v_date(wr_t *w, rng_t *r) {
    static const char *D[] = {"Mon", "Tue", "Wed", "Thu", "Fri", "Sat", "Sun"};
    static const char *M[] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
    char b[64];
    snprintf(b, sizeof b, "\"%s %s %02u %02u:%02u:%02u +0000 2014\"", D[rnd_below(r, 7)], M[rnd_below(r, 12)],
             1 + rnd_below(r, 28), rnd_below(r, 24), rnd_below(r, 60), rnd_below(r, 60));
    w_cstr(w, b);
}

static void gen_user(wr_t *w, rng_t *r) {
    unsigned long long id = 100000000ull + rnd(r) % 2000000000ull;
    w_open(w, '{');
    w_key(w, "id"); v_u64(w, id);
    w_key(w, "id_str"); w_ch(w, '"'); v_u64(w, id); w_ch(w, '"');
    w_key(w, "name"); v_text(w, r, 1, 3);
    w_key(w, "screen_name"); w_ch(w, '"'); w_cstr(w, WORDS[rnd_below(r, NWORDS)]); w_ch(w, '_'); v_int(w, rnd_below(r, 10000)); w_ch(w, '"');
    w_key(w, "location"); v_text(w, r, 0, 3);
    w_key(w, "description"); v_text(w, r, 3, 18);
    w_key(w, "url"); if (rnd_below(r, 3)) v_url(w, r); else v_lit(w, "null");
    w_key(w, "protected"); v_lit(w, rnd_below(r, 10) ? "false" : "true");
    w_key(w, "followers_count"); v_int(w, rnd_below(r, 100000));
    w_key(w, "friends_count"); v_int(w, rnd_below(r, 5000));
    w_key(w, "listed_count"); v_int(w, rnd_below(r, 200));
    w_key(w, "created_at"); v_date(w, r);
    w_key(w, "favourites_count"); v_int(w, rnd_below(r, 30000));
    w_key(w, "utc_offset"); if (rnd_below(r, 2)) v_int(w, 3600 * ((int)rnd_below(r, 24) - 12)); else v_lit(w, "null");
    w_key(w, "time_zone"); if (rnd_below(r, 2)) v_qstr(w, "Tokyo"); else v_lit(w, "null");
    w_key(w, "geo_enabled"); v_lit(w, rnd_below(r, 2) ? "false" : "true");
    w_key(w, "verified"); v_lit(w, "false");
    w_key(w, "statuses_count"); v_int(w, rnd_below(r, 200000));
    w_key(w, "lang"); v_qstr(w, rnd_below(r, 3) ? "en" : "ja");
    w_key(w, "profile_background_color"); v_qstr(w, "C0DEED");
    w_key(w, "profile_image_url"); v_url(w, r);
    w_key(w, "profile_use_background_image"); v_lit(w, "true");
    w_key(w, "default_profile"); v_lit(w, rnd_below(r, 2) ? "false" : "true");
    w_key(w, "following"); v_lit(w, "false");
    w_key(w, "ratio"); { char b[32]; snprintf(b, sizeof b, "%u.%03u", rnd_below(r, 10), rnd_below(r, 1000)); w_cstr(w, b); }
    w_close(w, '}');
}

static void gen_entities(wr_t *w, rng_t *r) {
    w_open(w, '{');
    w_key(w, "hashtags"); w_open(w, '[');
    for (uint32_t i = 0, n = rnd_below(r, 3); i < n; i++) {
        w_value_prefix(w); w_open(w, '{');
        w_key(w, "text"); v_qstr(w, WORDS[rnd_below(r, NWORDS)]);
        w_key(w, "indices"); w_open(w, '['); w_value_prefix(w); v_int(w, rnd_below(r, 100)); w_value_prefix(w); v_int(w, 100 + rnd_below(r, 40)); w_close(w, ']');
        w_close(w, '}');
    }
    w_close(w, ']');
    w_key(w, "symbols"); w_open(w, '['); w_close(w, ']');
    w_key(w, "urls"); w_open(w, '[');
    for (uint32_t i = 0, n = rnd_below(r, 3); i < n; i++) {
        w_value_prefix(w); w_open(w, '{');
        w_key(w, "url"); v_url(w, r);
        w_key(w, "expanded_url"); v_url(w, r);
        w_key(w, "indices"); w_open(w, '['); w_value_prefix(w); v_int(w, rnd_below(r, 100)); w_value_prefix(w); v_int(w, 100 + rnd_below(r, 40)); w_close(w, ']');
        w_close(w, '}');
    }
    w_close(w, ']');
    w_key(w, "user_mentions"); w_open(w, '[');
    for (uint32_t i = 0, n = rnd_below(r, 3); i < n; i++) {
        w_value_prefix(w); w_open(w, '{');
        w_key(w, "screen_name"); v_qstr(w, WORDS[rnd_below(r, NWORDS)]);
        w_key(w, "name"); v_text(w, r, 1, 2);
        w_key(w, "id"); v_u64(w, rnd(r) % 3000000000ull);
        w_key(w, "indices"); w_open(w, '['); w_value_prefix(w); v_int(w, rnd_below(r, 100)); w_value_prefix(w); v_int(w, 100 + rnd_below(r, 40)); w_close(w, ']');
        w_close(w, '}');
    }
    w_close(w, ']');
    w_close(w, '}');
}

/* the opening brace is written by the caller-side prefix logic: this emits a complete object value */
static void gen_status(wr_t *w, rng_t *r) {
    unsigned long long id = 500000000000000000ull + rnd(r) % 99999999999999999ull;
    w_open(w, '{');
    w_key(w, "metadata"); w_open(w, '{'); w_key(w, "result_type"); v_qstr(w, "recent"); w_key(w, "iso_language_code"); v_qstr(w, "ja"); w_close(w, '}');
    w_key(w, "created_at"); v_date(w, r);
    w_key(w, "id"); v_u64(w, id);
    w_key(w, "id_str"); w_ch(w, '"'); v_u64(w, id); w_ch(w, '"');
    w_key(w, "text"); v_text(w, r, 4, 24);
    w_key(w, "source"); w_cstr(w, "\"<a href=\\\"http:\\/\\/twitter.com\\/download\\/iphone\\\" rel=\\\"nofollow\\\">Twitter for iPhone<\\/a>\"");
    w_key(w, "truncated"); v_lit(w, "false");
    w_key(w, "in_reply_to_status_id"); if (rnd_below(r, 4) == 0) v_u64(w, id - 12345); else v_lit(w, "null");
    w_key(w, "in_reply_to_user_id"); if (rnd_below(r, 4) == 0) v_u64(w, rnd(r) % 3000000000ull); else v_lit(w, "null");
    w_key(w, "in_reply_to_screen_name"); v_lit(w, "null");
    w_key(w, "user"); gen_user(w, r);
    w_key(w, "geo"); v_lit(w, "null");
    w_key(w, "coordinates");
    if (rnd_below(r, 8) == 0) {
        w_open(w, '{'); w_key(w, "type"); v_qstr(w, "Point"); w_key(w, "coordinates"); w_open(w, '[');
        char b[48];
        w_value_prefix(w); snprintf(b, sizeof b, "%d.%06u", (int)rnd_below(r, 360) - 180, rnd_below(r, 1000000)); w_cstr(w, b);
        w_value_prefix(w); snprintf(b, sizeof b, "%d.%06ue-1", (int)rnd_below(r, 180) - 90, rnd_below(r, 1000000)); w_cstr(w, b);
        w_close(w, ']'); w_close(w, '}');
    } else v_lit(w, "null");
    w_key(w, "place"); v_lit(w, "null");
    w_key(w, "contributors"); v_lit(w, "null");
    w_key(w, "retweet_count"); v_int(w, rnd_below(r, 1000));
    w_key(w, "favorite_count"); v_int(w, rnd_below(r, 100));
    w_key(w, "entities"); gen_entities(w, r);
    w_key(w, "favorited"); v_lit(w, "false");
    w_key(w, "retweeted"); v_lit(w, rnd_below(r, 5) ? "false" : "true");
    w_key(w, "possibly_sensitive"); v_lit(w, "false");
    w_key(w, "lang"); v_qstr(w, rnd_below(r, 3) ? "en" : "ja");
    w_close(w, '}');
}

/* ---- public generators ---------------------------------------------------------------------- */

EXPORT int sjb200_gen_status_array(uint8_t *out, uint64_t size, uint64_t seed) {
    if (size < 2) return -1;
    rng_t r = {seed};
    wr_t w;
    memset(&w, 0, sizeof w);
    uint64_t n = 0;
    out[n++] = '[';
    int first = 1;
    for (;;) {
        w.n = 0; w.depth = 0; w.pretty = 0;
        gen_status(&w, &r);
        uint64_t need = w.n + (first ? 0 : 1);
        if (n + need + 1 > size) break;
        if (!first) out[n++] = ',';
        memcpy(out + n, w.p, w.n);
        n += w.n;
        first = 0;
    }
    while (n + 1 < size) out[n++] = ' ';
    out[n++] = ']';
    free(w.p);
    return 0;
}

EXPORT int sjb200_gen_ndjson(uint8_t *out, uint64_t size, uint64_t seed) {
    rng_t r = {seed};
    wr_t w;
    memset(&w, 0, sizeof w);
    uint64_t n = 0, last_nl = 0;
    for (;;) {
        w.n = 0; w.depth = 0; w.pretty = 0;
        gen_status(&w, &r);
        if (n + w.n + 1 > size) break;
        memcpy(out + n, w.p, w.n);
        n += w.n;
        last_nl = n;
        out[n++] = '\n';
    }
    free(w.p);
    if (n == 0) return -1;  /* not even one line fits */
    /* pad the last line with spaces before its newline so the batch is exactly `size` bytes of whole lines */
    for (uint64_t i = last_nl; i + 1 < size; i++) out[i] = ' ';
    out[size - 1] = '\n';
    return 0;
}

EXPORT int sjb200_gen_twitter_pretty(uint8_t *out, uint64_t size, uint64_t seed) {
    rng_t r = {seed};
    /* tail that closes the document, pretty printed like the head */
    static const char *TAIL =
        "\n  ],\n  \"search_metadata\": {\n    \"completed_in\": 0.087,\n    \"max_id\": 505874924095815700,\n"
        "    \"max_id_str\": \"505874924095815681\",\n    \"next_results\": \"?max_id=505874847260352512&q=%E4%B8%80&count=100\",\n"
        "    \"query\": \"%E4%B8%80\",\n    \"count\": 100,\n    \"since_id\": 0,\n    \"since_id_str\": \"0\"\n  }\n}";
    static const char *HEAD = "{\n  \"statuses\": [";
    const uint64_t tail_n = strlen(TAIL), head_n = strlen(HEAD);
    if (size < head_n + tail_n + 1) return -1;
    uint64_t n = 0;
    memcpy(out, HEAD, head_n);
    n = head_n;
    wr_t w;
    memset(&w, 0, sizeof w);
    int first = 1;
    for (;;) {
        w.n = 0; w.pretty = 1; w.depth = 2; w.need_comma[2] = 0;
        /* emit "\n    {...}" at depth 2 */
        w_newline(&w);
        gen_status(&w, &r);
        uint64_t need = w.n + (first ? 0 : 1);
        if (n + need + tail_n > size) break;
        if (!first) out[n++] = ',';
        memcpy(out + n, w.p, w.n);
        n += w.n;
        first = 0;
    }
    free(w.p);
    /* absorb the slack as trailing spaces after the last status (whitespace is legal there) */
    while (n + tail_n < size) out[n++] = ' ';
    memcpy(out + n, TAIL, tail_n);
    return 0;
}
