/*
 * gbwt_synth.c -- TEST / BENCHMARK INPUT GENERATOR (not part of the product, not the oracle).
 *
 * Writes synthetic bubble-chain pangenome GBWT indexes in the Simple-SDS GBWT file format
 *  and samples query patterns from their haplotypes
 * (SURVEY.md 8(d)). The reference has no construction code beyond the test-only BWTBuilder
 * (gbwt-rs src/bwt.rs:211-254), so this encoder is written from the format description:
 *   record   = varint sigma, sigma x (varint delta-node, varint offset), RLE body  (bwt.rs:241-253)
 *   ByteCode = 7 data bits per byte, bit 7 = continue                              (support.rs:1068-1075)
 *   RLE      = sigma < 255: byte value + sigma*(len-1), escape when len >= 256/sigma (support.rs:1238-1248)
 * tests/test_synth.py pins it: a brute-force GBWT builder (tests/gbwt_builder.py) that reproduces the
 * reference's C++-built fixtures byte for byte is compared with this closed-form generator.
 *
 * Graph: sites s = 0..S-1 with nodes A_s = 3s+1, B_s = 3s+2, C_s = 3s+3 and a final anchor
 * A_S = 3S+1. Haplotype h visits A_0, x_0, A_1, ..., x_{S-1}, A_S with x_s = C_s if allele(h, s)
 * else B_s; allele(h, s) = mix64(seed + h*S + s) >> 63. GBWT node = 2*id + orientation; sequence 2h is
 * the forward path, 2h+1 its reverse (support.rs:155, 310-314). offset = 1, alphabet_size = 2(3S+2).
 *
 * Variant model (the *_v entry points, alt_ppm > 0): nodes A_s = 4s+1, B_s, C_s, D_s = 4s+2..4, final anchor 4S+1;
 * allele 1 (C_s) with probability alt_ppm / 10^6, and at every site with mix64(31 seed + 0x5BD1E995 + s) % tri_mod == 0
 * a third allele (D_s) with the same probability: low-frequency variants give the anchor records long runs, the
 * tri-allelic sites give them outdegree 3 (run-length bodies on the device). D_s has an empty record elsewhere.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

uint64_t synth_mix64(uint64_t x) { return mix64(x); }

/* The allele model. Legacy (thr == 0): two equally likely alleles per site, three node ids per site. Variant model
 * (thr > 0): the alternative allele has probability thr / 2^32; every tri_mod-th site (chosen by a hash of the site) has a
 * second alternative allele of the same probability; four node ids per site (the fourth has an empty record at a
 * bi-allelic site). */
typedef struct { uint64_t seed, S; uint32_t thr, tri_mod, stride; } model_t;

static inline model_t make_model(uint64_t seed, uint64_t S, uint64_t alt_ppm, uint64_t tri_mod) {
    model_t m;
    m.seed = seed; m.S = S;
    m.thr = alt_ppm == 0 ? 0 : (uint32_t)((alt_ppm * 4294967296.0) / 1e6);
    if (alt_ppm != 0 && m.thr == 0) m.thr = 1;
    m.tri_mod = alt_ppm == 0 ? 0 : (uint32_t)tri_mod;
    m.stride = alt_ppm == 0 ? 3 : 4;
    return m;
}

static inline int site_is_tri(const model_t* m, uint64_t s) {
    return m->tri_mod != 0 && mix64(m->seed * 31 + 0x5BD1E995ULL + s) % m->tri_mod == 0;
}

static inline unsigned allele(const model_t* m, uint64_t h, uint64_t s) {
    uint64_t r = mix64(m->seed + h * m->S + s);
    if (m->thr == 0) return (unsigned)(r >> 63);
    uint64_t u = r >> 32;
    if (u < m->thr) return 1;
    if (u < 2 * (uint64_t)m->thr && site_is_tri(m, s)) return 2;
    return 0;
}

unsigned synth_allele(uint64_t seed, uint64_t S, uint64_t h, uint64_t s) { model_t m = make_model(seed, S, 0, 0); return allele(&m, h, s); }
unsigned synth_allele_v(uint64_t seed, uint64_t S, uint64_t alt_ppm, uint64_t tri_mod, uint64_t h, uint64_t s) {
    model_t m = make_model(seed, S, alt_ppm, tri_mod);
    return allele(&m, h, s);
}

static inline size_t put_varint(uint8_t* p, uint64_t v) {
    size_t n = 0;
    while (v > 0x7F) { p[n++] = (uint8_t)((v & 0x7F) | 0x80); v >>= 7; }
    p[n++] = (uint8_t)v;
    return n;
}

/* One run with alphabet size sigma (1 or 2 here, general for sigma < 255). */
static inline size_t put_run(uint8_t* p, uint64_t sigma, uint64_t value, uint64_t len) {
    uint64_t threshold = 256 / sigma;
    if (len < threshold) { p[0] = (uint8_t)(value + sigma * (len - 1)); return 1; }
    p[0] = (uint8_t)(value + sigma * (threshold - 1));
    return 1 + put_varint(p + 1, len - threshold);
}

/* ---- Simple-SDS writer -------------------------------------------------------------------------- */

typedef struct { uint8_t* p; size_t len, cap; } wbuf;

static void wb_reserve(wbuf* w, size_t extra) {
    if (w->len + extra <= w->cap) return;
    size_t cap = w->cap ? w->cap : 4096;
    while (cap < w->len + extra) cap *= 2;
    w->p = (uint8_t*)realloc(w->p, cap);
    w->cap = cap;
}
static void wb_u64(wbuf* w, uint64_t v) { wb_reserve(w, 8); memcpy(w->p + w->len, &v, 8); w->len += 8; }
static void wb_words(wbuf* w, const uint64_t* v, size_t n) { wb_reserve(w, n * 8); memcpy(w->p + w->len, v, n * 8); w->len += n * 8; }
static void wb_bytes_padded(wbuf* w, const uint8_t* b, size_t n) {
    size_t padded = (n + 7) / 8 * 8;
    wb_u64(w, n);
    wb_reserve(w, padded);
    memcpy(w->p + w->len, b, n);
    memset(w->p + w->len + n, 0, padded - n);
    w->len += padded;
}

/* SparseVector (Elias-Fano) of strictly increasing values; parameters as in SURVEY.md App. A. */
static void wb_sparse(wbuf* w, uint64_t universe, const uint64_t* values, uint64_t ones) {
    uint64_t width = 1;
    if (ones > 0 && ones <= universe) {
        double r = round(log2(((double)universe * log(2.0)) / (double)ones));
        width = r < 1.0 ? 1 : (uint64_t)r;
    }
    uint64_t mask = width < 64 ? ((1ULL << width) - 1) : ~0ULL;
    uint64_t buckets = (width < 64 ? (universe >> width) : 0) + ((universe & mask) ? 1 : 0);
    uint64_t high_len = ones + buckets;
    uint64_t high_words = (high_len + 63) / 64;
    uint64_t low_bits = ones * width;
    uint64_t low_words = (low_bits + 63) / 64;
    uint64_t* high = (uint64_t*)calloc(high_words + 1, 8);
    uint64_t* low = (uint64_t*)calloc(low_words + 1, 8);
    for (uint64_t j = 0; j < ones; j++) {
        uint64_t v = values[j];
        uint64_t hp = (v >> width) + j;
        high[hp / 64] |= 1ULL << (hp % 64);
        uint64_t lv = v & mask, bit = j * width;
        low[bit / 64] |= lv << (bit % 64);
        if ((bit % 64) + width > 64) low[bit / 64 + 1] |= lv >> (64 - bit % 64);
    }
    wb_u64(w, universe);
    /* BitVector: ones, RawVector, three absent support structures */
    wb_u64(w, ones); wb_u64(w, high_len); wb_u64(w, high_words); wb_words(w, high, high_words);
    wb_u64(w, 0); wb_u64(w, 0); wb_u64(w, 0);
    /* IntVector: len, width, RawVector */
    wb_u64(w, ones); wb_u64(w, width); wb_u64(w, low_bits); wb_u64(w, low_words); wb_words(w, low, low_words);
    free(high); free(low);
}

/* Tags = StringArray of [key, value, ...] (support.rs:610-626): starts, alphabet, packed chars. */
static void wb_tags(wbuf* w, const char** strings, size_t n) {
    uint64_t starts[16];
    uint8_t all[256];
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        starts[i] = total;
        size_t l = strlen(strings[i]);
        memcpy(all + total, strings[i], l);
        total += l;
    }
    wb_sparse(w, n ? starts[n - 1] + 1 : 0, starts, n);
    int present[256] = {0};
    for (size_t i = 0; i < total; i++) present[all[i]] = 1;
    uint8_t alphabet[256]; uint8_t pack[256]; size_t sigma = 0;
    for (int c = 0; c < 256; c++) if (present[c]) { pack[c] = (uint8_t)sigma; alphabet[sigma++] = (uint8_t)c; }
    wb_bytes_padded(w, alphabet, sigma);
    uint64_t width = 1;
    while (sigma > 1 && ((sigma - 1) >> width) != 0) width++;
    uint64_t bits = total * width, words = (bits + 63) / 64;
    uint64_t* packed = (uint64_t*)calloc(words + 1, 8);
    for (size_t i = 0; i < total; i++) {
        uint64_t v = pack[all[i]], bit = i * width;
        packed[bit / 64] |= v << (bit % 64);
        if ((bit % 64) + width > 64) packed[bit / 64 + 1] |= v >> (64 - bit % 64);
    }
    wb_u64(w, total); wb_u64(w, width); wb_u64(w, bits); wb_u64(w, words); wb_words(w, packed, words);
    free(packed);
}

#define GBWT_TAG 0x6B376B37ULL
#define GBWT_VERSION 5ULL

/* Assemble a Simple-SDS GBWT image from raw parts (gbwt-rs src/gbwt.rs:389-400 field order):
 * header, tags, BWT (SparseVector index + Vec<u8> data), empty DA samples, no metadata.
 * `data` may be NULL, in which case room for data_len bytes is left and *data_at receives its offset. */
static uint8_t* assemble_image(uint64_t sequences, uint64_t size, uint64_t offset, uint64_t alphabet_size, uint64_t flags,
                               const uint64_t* rec_starts, uint64_t records, const uint8_t* data, uint64_t data_len,
                               uint64_t* out_len, uint64_t* data_at) {
    wbuf w = {0, 0, 0};
    wb_u64(&w, GBWT_TAG | (GBWT_VERSION << 32));
    wb_u64(&w, sequences); wb_u64(&w, size); wb_u64(&w, offset); wb_u64(&w, alphabet_size); wb_u64(&w, flags);
    const char* tags[2] = {"source", "jltsiren/gbwt"};
    wb_tags(&w, tags, 2);
    wb_sparse(&w, data_len, rec_starts, records);
    size_t padded = (size_t)((data_len + 7) / 8 * 8);
    wb_u64(&w, data_len);
    wb_reserve(&w, padded + 16);
    if (data_at) *data_at = w.len;
    if (data) memcpy(w.p + w.len, data, data_len);
    memset(w.p + w.len + data_len, 0, padded - data_len);
    w.len += padded;
    wb_u64(&w, 0); /* DA samples: empty Vec<u64> */
    wb_u64(&w, 0); /* Option<Metadata>: None */
    *out_len = w.len;
    return w.p;
}

uint8_t* synth_gbwt_image(uint64_t sequences, uint64_t size, uint64_t offset, uint64_t alphabet_size, uint64_t flags,
                          const uint64_t* rec_starts, uint64_t records, const uint8_t* data, uint64_t data_len,
                          uint64_t* out_len) {
    return assemble_image(sequences, size, offset, alphabet_size, flags, rec_starts, records, data, data_len, out_len, NULL);
}

/* Only the BWT section (SparseVector + Vec<u8>), for byte-exact comparison with the fixtures. */
uint8_t* synth_bwt_section(const uint64_t* rec_starts, uint64_t records, const uint8_t* data, uint64_t data_len, uint64_t* out_len) {
    wbuf w = {0, 0, 0};
    wb_sparse(&w, data_len, rec_starts, records);
    wb_bytes_padded(&w, data, data_len);
    *out_len = w.len;
    return w.p;
}

/* ---- GBZ container with node labels (gbwt-rs src/gbz.rs:650-696, src/graph.rs:284-294) ------------ */

#include <dlfcn.h>

/* StringArray::serialize body after the index (src/support.rs:592-610): alphabet + packed characters. */
static void wb_packed_strings(wbuf* w, const uint8_t* bytes, uint64_t total) {
    int present[256] = {0};
    for (uint64_t i = 0; i < total; i++) present[bytes[i]] = 1;
    uint8_t alphabet[256] = {0}; uint8_t pack[256] = {0}; size_t sigma = 0;
    for (int c = 0; c < 256; c++) if (present[c]) { pack[c] = (uint8_t)sigma; alphabet[sigma++] = (uint8_t)c; }
    wb_bytes_padded(w, alphabet, sigma);
    uint64_t width = 1;
    while (sigma > 1 && ((sigma - 1) >> width) != 0) width++;
    uint64_t bits = total * width, words = (bits + 63) / 64;
    uint64_t* packed = (uint64_t*)calloc(words + 1, 8);
    for (uint64_t i = 0; i < total; i++) {
        uint64_t v = pack[bytes[i]], bit = i * width;
        packed[bit / 64] |= v << (bit % 64);
        if ((bit % 64) + width > 64) packed[bit / 64 + 1] |= v >> (64 - bit % 64);
    }
    wb_u64(w, total); wb_u64(w, width); wb_u64(w, bits); wb_u64(w, words); wb_words(w, packed, words);
    free(packed);
}

/* A GBZ image around an existing GBWT image: header, tags, GBWT, Graph {header, node labels, no segment
 * names, empty node-to-segment mapping}. graph_version 3 writes the labels as a packed StringArray
 * (GBZ version 1 files), 4 as a Zstandard frame (GBZ version 2; libzstd is resolved with dlopen and NULL is
 * returned without it). starts has n_labels + 1 entries. */
uint8_t* synth_gbz_image(const uint8_t* gbwt, uint64_t gbwt_len, uint64_t nodes, uint64_t n_labels, const uint64_t* starts,
                         const uint8_t* labels, int graph_version, uint64_t* out_len) {
    if (graph_version != 3 && graph_version != 4) return NULL;
    uint64_t total = starts[n_labels];
    wbuf w = {0, 0, 0};
    wb_u64(&w, 0x205A4247ULL | ((uint64_t)(graph_version == 4 ? 2 : 1) << 32));
    wb_u64(&w, 0);
    const char* tags[2] = {"source", "jltsiren/gbwt-rs"};
    wb_tags(&w, tags, 2);
    wb_reserve(&w, gbwt_len);
    memcpy(w.p + w.len, gbwt, gbwt_len);
    w.len += gbwt_len;
    wb_u64(&w, 0x6B3764AFULL | ((uint64_t)graph_version << 32));
    wb_u64(&w, nodes);
    wb_u64(&w, 2); /* FLAG_SIMPLE_SDS, no translation */
    wb_sparse(&w, n_labels ? starts[n_labels - 1] + 1 : 0, starts, n_labels);
    if (graph_version == 4) {
        typedef size_t (*bound_fn)(size_t);
        typedef size_t (*compress_fn)(void*, size_t, const void*, size_t, int);
        typedef unsigned (*iserror_fn)(size_t);
        void* h = dlopen("libzstd.so.1", RTLD_NOW);
        bound_fn bound = h ? (bound_fn)dlsym(h, "ZSTD_compressBound") : NULL;
        compress_fn compress = h ? (compress_fn)dlsym(h, "ZSTD_compress") : NULL;
        iserror_fn iserror = h ? (iserror_fn)dlsym(h, "ZSTD_isError") : NULL;
        if (!bound || !compress || !iserror) { free(w.p); return NULL; }
        size_t cap = bound((size_t)total);
        uint8_t* tmp = (uint8_t*)malloc(cap + 8);
        size_t n = compress(tmp, cap, labels, (size_t)total, 3);
        if (iserror(n)) { free(tmp); free(w.p); return NULL; }
        wb_u64(&w, total);
        wb_bytes_padded(&w, tmp, n);
        free(tmp);
    } else {
        wb_packed_strings(&w, labels, total);
    }
    wb_sparse(&w, 0, NULL, 0);          /* segments: empty StringArray */
    wb_packed_strings(&w, NULL, 0);
    wb_sparse(&w, 0, NULL, 0);          /* mapping */
    *out_len = w.len;
    return w.p;
}

void synth_free(void* p) { free(p); }

/* ---- bubble chain ------------------------------------------------------------------------------- */

/* Order of the haplotypes when they arrive at the anchor of site `s` on one strand: sorted by the
 * alleles of the previously visited sites (most recent first), ties by haplotype id. Computed exactly
 * by replaying the stable partitions of the last K sites and extending K while ties remain. */
static void order_at(const model_t* m, uint64_t H, int reverse, uint64_t s, uint32_t* order, uint32_t* tmp, uint64_t* key) {
    uint64_t S = m->S;
    uint64_t avail = reverse ? (S - 1 - s) : s; /* number of sites already visited */
    uint64_t K = m->thr == 0 ? 64 : 32;
    const unsigned classes = m->thr == 0 ? 2 : 3, key_bits = m->thr == 0 ? 1 : 2;
    for (;;) {
        if (K > avail) K = avail;
        for (uint64_t i = 0; i < H; i++) order[i] = (uint32_t)i;
        /* visited sites, oldest of the window first */
        for (uint64_t j = K; j >= 1; j--) {
            uint64_t site = reverse ? (s + j) : (s - j);
            uint64_t nb = 0;
            for (unsigned c = 0; c < classes; c++)
                for (uint64_t i = 0; i < H; i++) if (allele(m, order[i], site) == c) tmp[nb++] = order[i];
            memcpy(order, tmp, H * sizeof(uint32_t));
        }
        if (K == avail) return;
        /* exact iff no two neighbours share all K window alleles */
        int tie = 0;
        if (K * key_bits <= 64) {
            for (uint64_t i = 0; i < H; i++) {
                uint64_t k = 0;
                for (uint64_t j = 1; j <= K; j++) k = (k << key_bits) | allele(m, order[i], reverse ? (s + j) : (s - j));
                key[i] = k;
            }
            for (uint64_t i = 1; i < H && !tie; i++) tie = (key[i] == key[i - 1]);
        } else {
            for (uint64_t i = 1; i < H && !tie; i++) {
                int same = 1;
                for (uint64_t j = 1; j <= K && same; j++) {
                    uint64_t site = reverse ? (s + j) : (s - j);
                    same = allele(m, order[i], site) == allele(m, order[i - 1], site);
                }
                tie = same;
            }
        }
        if (!tie) return;
        K *= 4;
    }
}

typedef struct {
    uint8_t* buf; size_t len, cap;
} chunk_buf;

static inline uint8_t* cb_reserve(chunk_buf* c, size_t extra) {
    if (c->len + extra > c->cap) {
        size_t cap = c->cap ? c->cap : (1 << 16);
        while (cap < c->len + extra) cap *= 2;
        c->buf = (uint8_t*)realloc(c->buf, cap);
        c->cap = cap;
    }
    return c->buf + c->len;
}

/* Emits the records of one site on one strand (anchor, then one per allele node) and advances the order.
 * alt_nodes[] / next_anchor are GBWT node ids; rec_len[0] = anchor, rec_len[1 + a] = node of allele a. */
static void emit_site(chunk_buf* cb, const model_t* m, uint64_t H, uint64_t site, const uint64_t* alt_nodes, uint64_t next_anchor,
                      uint32_t* order, uint32_t* tmp, uint8_t* bits, uint32_t* rec_len) {
    const unsigned alleles = m->stride - 1;
    uint64_t cnt[3] = {0, 0, 0};
    unsigned rank_of[3] = {0, 0, 0};
    for (uint64_t i = 0; i < H; i++) {
        unsigned a = allele(m, order[i], site);
        bits[i] = (uint8_t)a;
        cnt[a]++;
    }
    /* anchor record */
    uint8_t* p = cb_reserve(cb, 64 + 3 * H);
    size_t n = 0;
    uint64_t sigma = 0;
    for (unsigned a = 0; a < alleles; a++) { rank_of[a] = (unsigned)sigma; sigma += cnt[a] > 0; }
    n += put_varint(p + n, sigma);
    uint64_t prev = 0;
    for (unsigned a = 0; a < alleles; a++)
        if (cnt[a] > 0) { n += put_varint(p + n, alt_nodes[a] - prev); n += put_varint(p + n, 0); prev = alt_nodes[a]; }
    uint64_t i = 0;
    while (i < H) {
        uint64_t j = i + 1;
        while (j < H && bits[j] == bits[i]) j++;
        n += put_run(p + n, sigma, rank_of[bits[i]], j - i);
        i = j;
    }
    cb->len += n; rec_len[0] = (uint32_t)n;
    /* allele records: the takers of allele a continue to the next anchor, after the takers of the smaller alleles */
    uint64_t before = 0;
    for (unsigned a = 0; a < alleles; a++) {
        p = cb_reserve(cb, 64);
        n = 0;
        if (cnt[a] > 0) {
            n += put_varint(p + n, 1); n += put_varint(p + n, next_anchor); n += put_varint(p + n, before);
            n += put_run(p + n, 1, 0, cnt[a]);
        } else {
            p[n++] = 0;
        }
        cb->len += n; rec_len[1 + a] = (uint32_t)n;
        before += cnt[a];
    }
    /* stable partition */
    uint64_t k = 0;
    for (unsigned a = 0; a < alleles; a++)
        for (uint64_t q = 0; q < H; q++) if (bits[q] == a) tmp[k++] = order[q];
    memcpy(order, tmp, H * sizeof(uint32_t));
}

/* Returns a malloc'd Simple-SDS GBWT image (free with synth_free). */
static uint8_t* bubble_chain_image(const model_t* m, uint64_t H, int threads, uint64_t* out_len) {
    const uint64_t S = m->S, T = m->stride;
    if (S == 0 || H == 0 || H > 0xFFFFFFFFULL) return NULL;
#ifdef _OPENMP
    /* all processors unless told otherwise: launchers such as torchrun export OMP_NUM_THREADS=1 */
    if (threads <= 0) threads = omp_get_num_procs();
#else
    threads = 1;
#endif
    const uint64_t CH = 2048; /* sites per chunk */
    uint64_t n_chunks = (S + CH - 1) / CH;
    uint64_t n_ids = T * S + 1; /* original node ids 1..T*S+1 */
    chunk_buf* fwd = (chunk_buf*)calloc(n_chunks, sizeof(chunk_buf));
    chunk_buf* rev = (chunk_buf*)calloc(n_chunks, sizeof(chunk_buf));
    /* per-record lengths, indexed by original node id (1-based) */
    uint32_t* len_f = (uint32_t*)calloc(n_ids + 2, sizeof(uint32_t));
    uint32_t* len_r = (uint32_t*)calloc(n_ids + 2, sizeof(uint32_t));
    uint64_t NEXT_LAST_F = 2 * (T * S + 1); /* A_S forward */

#pragma omp parallel num_threads(threads)
    {
        uint32_t* order = (uint32_t*)malloc(H * sizeof(uint32_t));
        uint32_t* tmp = (uint32_t*)malloc(H * sizeof(uint32_t));
        uint64_t* key = (uint64_t*)malloc(H * sizeof(uint64_t));
        uint8_t* bits = (uint8_t*)malloc(H);
#pragma omp for schedule(dynamic, 1)
        for (int64_t job = 0; job < (int64_t)(2 * n_chunks); job++) {
            uint64_t c = (uint64_t)job / 2;
            int reverse = (int)(job & 1);
            uint64_t s0 = c * CH, s1 = s0 + CH < S ? s0 + CH : S;
            uint32_t rl[4];
            uint64_t alts[3];
            if (!reverse) {
                order_at(m, H, 0, s0, order, tmp, key);
                for (uint64_t s = s0; s < s1; s++) {
                    uint64_t a = T * s + 1;
                    uint64_t next = (s + 1 == S) ? NEXT_LAST_F : 2 * (T * (s + 1) + 1);
                    for (uint64_t d = 1; d < T; d++) alts[d - 1] = 2 * (a + d);
                    emit_site(&fwd[c], m, H, s, alts, next, order, tmp, bits, rl);
                    for (uint64_t d = 0; d < T; d++) len_f[a + d] = rl[d];
                }
            } else {
                /* reverse strand visits sites in descending order: anchor A_{s+1} rev -> allele nodes of s rev -> A_s rev */
                order_at(m, H, 1, s1 - 1, order, tmp, key);
                for (uint64_t s = s1; s-- > s0;) {
                    uint64_t a = T * s + 1;
                    for (uint64_t d = 1; d < T; d++) alts[d - 1] = 2 * (a + d) + 1;
                    emit_site(&rev[c], m, H, s, alts, 2 * a + 1, order, tmp, bits, rl);
                    len_r[a + T] = rl[0];
                    for (uint64_t d = 1; d < T; d++) len_r[a + d] = rl[d];
                }
            }
        }
        free(order); free(tmp); free(key); free(bits);
    }

    /* terminal anchors: A_S forward and A_0 reverse end every sequence: edge (ENDMARKER, 0), run (0, H) */
    uint8_t term[32];
    size_t term_len = 0;
    term_len += put_varint(term + term_len, 1); term_len += put_varint(term + term_len, 0); term_len += put_varint(term + term_len, 0);
    term_len += put_run(term + term_len, 1, 0, H);
    len_f[T * S + 1] = (uint32_t)term_len;
    len_r[1] = (uint32_t)term_len;

    /* endmarker record: sequences alternate forward (starts at A_0 fwd = 2) and reverse (A_S rev) */
    size_t em_cap = 64 + 2 * (size_t)H;
    uint8_t* em = (uint8_t*)malloc(em_cap);
    size_t em_len = 0;
    em_len += put_varint(em + em_len, 2);
    em_len += put_varint(em + em_len, 2); em_len += put_varint(em + em_len, 0);
    em_len += put_varint(em + em_len, (2 * (T * S + 1) + 1) - 2); em_len += put_varint(em + em_len, 0);
    for (uint64_t h = 0; h < H; h++) { em[em_len++] = 0; em[em_len++] = 1; } /* runs (0,1),(1,1) with sigma 2 */

    /* record starts: record 0 = endmarker, then for id = 1..T*S+1: forward, reverse */
    uint64_t records = 2 * n_ids + 1;
    uint64_t* starts = (uint64_t*)malloc(records * sizeof(uint64_t));
    uint64_t pos = 0;
    starts[0] = 0; pos = em_len;
    for (uint64_t id = 1; id <= n_ids; id++) {
        starts[2 * id - 1] = pos; pos += len_f[id];
        starts[2 * id] = pos; pos += len_r[id];
    }
    uint64_t data_len = pos;
    uint64_t sequences = 2 * H;
    uint64_t size = sequences * (2 * S + 2);
    uint64_t data_at = 0;
    uint8_t* image = assemble_image(sequences, size, 1, 2 * (T * S + 2), 1 | 4, starts, records, NULL, data_len, out_len, &data_at);
    uint8_t* data = image + data_at;
    memcpy(data, em, em_len);
    memcpy(data + starts[2 * (T * S + 1) - 1], term, term_len);
    memcpy(data + starts[2], term, term_len);

#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t job = 0; job < (int64_t)(2 * n_chunks); job++) {
        uint64_t c = (uint64_t)job / 2;
        int reverse = (int)(job & 1);
        uint64_t s0 = c * CH, s1 = s0 + CH < S ? s0 + CH : S;
        if (!reverse) {
            const uint8_t* src = fwd[c].buf;
            for (uint64_t s = s0; s < s1; s++) {
                uint64_t a = T * s + 1;
                for (uint64_t d = 0; d < T; d++) {
                    memcpy(data + starts[2 * (a + d) - 1], src, len_f[a + d]);
                    src += len_f[a + d];
                }
            }
        } else {
            const uint8_t* src = rev[c].buf;
            for (uint64_t s = s1; s-- > s0;) {
                uint64_t a = T * s + 1;
                for (uint64_t d = 0; d < T; d++) {
                    uint64_t id = d == 0 ? a + T : a + d;
                    memcpy(data + starts[2 * id], src, len_r[id]);
                    src += len_r[id];
                }
            }
        }
        free(reverse ? rev[c].buf : fwd[c].buf);
    }
    free(fwd); free(rev); free(len_f); free(len_r); free(em); free(starts);
    return image;
}

uint8_t* synth_bubble_chain_gbwt(uint64_t S, uint64_t H, uint64_t seed, int threads, uint64_t* out_len) {
    model_t m = make_model(seed, S, 0, 0);
    return bubble_chain_image(&m, H, threads, out_len);
}

/* The variant model: alternative alleles with probability alt_ppm / 10^6, every tri_mod-th site tri-allelic (0: none),
 * N = 4 S + 1 node ids. */
uint8_t* synth_bubble_chain_gbwt_v(uint64_t S, uint64_t H, uint64_t seed, uint64_t alt_ppm, uint64_t tri_mod, int threads,
                                   uint64_t* out_len) {
    model_t m = make_model(seed, S, alt_ppm, tri_mod);
    return bubble_chain_image(&m, H, threads, out_len);
}

/* ---- haplotype paths and query patterns ----------------------------------------- */

/* Node at position p (0..2S) of forward sequence 2h. */
static inline uint64_t fwd_node(const model_t* m, uint64_t h, uint64_t p) {
    uint64_t s = p >> 1;
    if ((p & 1) == 0) return 2 * (m->stride * s + 1);
    return 2 * (m->stride * s + 2 + allele(m, h, s));
}

static inline uint64_t seq_node(const model_t* m, uint64_t seq, uint64_t p) {
    uint64_t h = seq >> 1;
    if ((seq & 1) == 0) return fwd_node(m, h, p);
    return fwd_node(m, h, 2 * m->S - p) ^ 1;
}

/* Full sequence `seq` (2S+1 nodes). */
void synth_sequence_v(uint64_t S, uint64_t H, uint64_t seed, uint64_t alt_ppm, uint64_t tri_mod, uint64_t seq, uint64_t* out) {
    (void)H;
    model_t m = make_model(seed, S, alt_ppm, tri_mod);
    for (uint64_t p = 0; p <= 2 * S; p++) out[p] = seq_node(&m, seq, p);
}
void synth_sequence(uint64_t S, uint64_t H, uint64_t seed, uint64_t seq, uint64_t* out) { synth_sequence_v(S, H, seed, 0, 0, seq, out); }

/* Queries q0 .. q0+n: h = mix64(seed_q + 3q) % H, o = mix64(seed_q + 3q + 1) & 1,
 * t = mix64(seed_q + 3q + 2) % (2S + 1 - (k - 1)); pattern = sequence 2h+o positions [t, t+k). */
void synth_patterns_v(uint64_t S, uint64_t H, uint64_t seed, uint64_t alt_ppm, uint64_t tri_mod, uint64_t seed_q, uint64_t q0,
                      uint64_t n, uint64_t k, uint64_t* out, int threads) {
#ifdef _OPENMP
    /* all processors unless told otherwise: launchers such as torchrun export OMP_NUM_THREADS=1 */
    if (threads <= 0) threads = omp_get_num_procs();
#else
    threads = 1;
#endif
    const model_t m = make_model(seed, S, alt_ppm, tri_mod);
    uint64_t span = 2 * S + 1 - (k - 1);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t q = q0 + (uint64_t)i;
        uint64_t h = mix64(seed_q + 3 * q) % H;
        uint64_t o = mix64(seed_q + 3 * q + 1) & 1;
        uint64_t t = mix64(seed_q + 3 * q + 2) % span;
        uint64_t* dst = out + (uint64_t)i * k;
        for (uint64_t j = 0; j < k; j++) dst[j] = seq_node(&m, 2 * h + o, t + j);
    }
}
void synth_patterns(uint64_t S, uint64_t H, uint64_t seed, uint64_t seed_q, uint64_t q0, uint64_t n, uint64_t k,
                    uint64_t* out, int threads) {
    synth_patterns_v(S, H, seed, 0, 0, seed_q, q0, n, k, out, threads);
}
