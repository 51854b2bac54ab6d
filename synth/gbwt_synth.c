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

static inline unsigned allele(uint64_t seed, uint64_t S, uint64_t h, uint64_t s) {
    return (unsigned)(mix64(seed + h * S + s) >> 63);
}

unsigned synth_allele(uint64_t seed, uint64_t S, uint64_t h, uint64_t s) { return allele(seed, S, h, s); }

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
static void order_at(uint64_t seed, uint64_t S, uint64_t H, int reverse, uint64_t s, uint32_t* order, uint32_t* tmp, uint64_t* key) {
    uint64_t avail = reverse ? (S - 1 - s) : s; /* number of sites already visited */
    uint64_t K = 64;
    for (;;) {
        if (K > avail) K = avail;
        for (uint64_t i = 0; i < H; i++) order[i] = (uint32_t)i;
        /* visited sites, oldest of the window first */
        for (uint64_t j = K; j >= 1; j--) {
            uint64_t site = reverse ? (s + j) : (s - j);
            uint64_t nb = 0;
            for (uint64_t i = 0; i < H; i++) if (!allele(seed, S, order[i], site)) tmp[nb++] = order[i];
            for (uint64_t i = 0; i < H; i++) if (allele(seed, S, order[i], site)) tmp[nb++] = order[i];
            memcpy(order, tmp, H * sizeof(uint32_t));
        }
        if (K == avail) return;
        /* exact iff no two neighbours share all K window alleles */
        int tie = 0;
        if (K <= 64) {
            for (uint64_t i = 0; i < H; i++) {
                uint64_t k = 0;
                for (uint64_t j = 1; j <= K; j++) k = (k << 1) | allele(seed, S, order[i], reverse ? (s + j) : (s - j));
                key[i] = k;
            }
            for (uint64_t i = 1; i < H && !tie; i++) tie = (key[i] == key[i - 1]);
        } else {
            for (uint64_t i = 1; i < H && !tie; i++) {
                int same = 1;
                for (uint64_t j = 1; j <= K && same; j++) {
                    uint64_t site = reverse ? (s + j) : (s - j);
                    same = allele(seed, S, order[i], site) == allele(seed, S, order[i - 1], site);
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

/* Emits the three records of one site on one strand and advances the order.
 * b_node / c_node / next_anchor are GBWT node ids. */
static void emit_site(chunk_buf* cb, uint64_t seed, uint64_t S, uint64_t H, uint64_t site,
                      uint64_t b_node, uint64_t c_node, uint64_t next_anchor,
                      uint32_t* order, uint32_t* tmp, uint8_t* bits, uint32_t* rec_len /* [3]: anchor, B, C */) {
    uint64_t nb = 0, nc = 0;
    for (uint64_t i = 0; i < H; i++) {
        unsigned a = allele(seed, S, order[i], site);
        bits[i] = (uint8_t)a;
        nc += a;
    }
    nb = H - nc;
    /* anchor record */
    uint8_t* p = cb_reserve(cb, 64 + 3 * H);
    size_t n = 0;
    uint64_t sigma = (nb > 0) + (nc > 0);
    n += put_varint(p + n, sigma);
    uint64_t prev = 0;
    if (nb > 0) { n += put_varint(p + n, b_node - prev); n += put_varint(p + n, 0); prev = b_node; }
    if (nc > 0) { n += put_varint(p + n, c_node - prev); n += put_varint(p + n, 0); prev = c_node; }
    uint64_t i = 0;
    while (i < H) {
        uint64_t j = i + 1;
        while (j < H && bits[j] == bits[i]) j++;
        uint64_t value = (sigma == 2) ? bits[i] : 0;
        n += put_run(p + n, sigma, value, j - i);
        i = j;
    }
    cb->len += n; rec_len[0] = (uint32_t)n;
    /* B record: all B-takers continue to the next anchor at offsets 0.. */
    p = cb_reserve(cb, 64);
    n = 0;
    if (nb > 0) {
        n += put_varint(p + n, 1); n += put_varint(p + n, next_anchor); n += put_varint(p + n, 0);
        n += put_run(p + n, 1, 0, nb);
    } else {
        p[n++] = 0;
    }
    cb->len += n; rec_len[1] = (uint32_t)n;
    /* C record: C-takers follow the B-takers in the next anchor */
    p = cb_reserve(cb, 64);
    n = 0;
    if (nc > 0) {
        n += put_varint(p + n, 1); n += put_varint(p + n, next_anchor); n += put_varint(p + n, nb);
        n += put_run(p + n, 1, 0, nc);
    } else {
        p[n++] = 0;
    }
    cb->len += n; rec_len[2] = (uint32_t)n;
    /* stable partition */
    uint64_t k = 0;
    for (uint64_t q = 0; q < H; q++) if (!bits[q]) tmp[k++] = order[q];
    for (uint64_t q = 0; q < H; q++) if (bits[q]) tmp[k++] = order[q];
    memcpy(order, tmp, H * sizeof(uint32_t));
}

/* Returns a malloc'd Simple-SDS GBWT image (free with synth_free). */
uint8_t* synth_bubble_chain_gbwt(uint64_t S, uint64_t H, uint64_t seed, int threads, uint64_t* out_len) {
    if (S == 0 || H == 0 || H > 0xFFFFFFFFULL) return NULL;
#ifdef _OPENMP
    /* all processors unless told otherwise: launchers such as torchrun export OMP_NUM_THREADS=1 */
    if (threads <= 0) threads = omp_get_num_procs();
#else
    threads = 1;
#endif
    const uint64_t CH = 2048; /* sites per chunk */
    uint64_t n_chunks = (S + CH - 1) / CH;
    uint64_t n_ids = 3 * S + 1; /* original node ids 1..3S+1 */
    chunk_buf* fwd = (chunk_buf*)calloc(n_chunks, sizeof(chunk_buf));
    chunk_buf* rev = (chunk_buf*)calloc(n_chunks, sizeof(chunk_buf));
    /* per-record lengths, indexed by original node id (1-based) */
    uint32_t* len_f = (uint32_t*)calloc(n_ids + 2, sizeof(uint32_t));
    uint32_t* len_r = (uint32_t*)calloc(n_ids + 2, sizeof(uint32_t));
    uint64_t NEXT_LAST_F = 2 * (3 * S + 1); /* A_S forward */

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
            uint32_t rl[3];
            if (!reverse) {
                order_at(seed, S, H, 0, s0, order, tmp, key);
                for (uint64_t s = s0; s < s1; s++) {
                    uint64_t a = 3 * s + 1;
                    uint64_t next = (s + 1 == S) ? NEXT_LAST_F : 2 * (3 * (s + 1) + 1);
                    emit_site(&fwd[c], seed, S, H, s, 2 * (a + 1), 2 * (a + 2), next, order, tmp, bits, rl);
                    len_f[a] = rl[0]; len_f[a + 1] = rl[1]; len_f[a + 2] = rl[2];
                }
            } else {
                /* reverse strand visits sites in descending order: anchor A_{s+1} rev -> B_s/C_s rev -> A_s rev */
                order_at(seed, S, H, 1, s1 - 1, order, tmp, key);
                for (uint64_t s = s1; s-- > s0;) {
                    uint64_t a = 3 * s + 1;
                    emit_site(&rev[c], seed, S, H, s, 2 * (a + 1) + 1, 2 * (a + 2) + 1, 2 * a + 1,
                              order, tmp, bits, rl);
                    len_r[a + 3] = rl[0]; len_r[a + 1] = rl[1]; len_r[a + 2] = rl[2];
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
    len_f[3 * S + 1] = (uint32_t)term_len;
    len_r[1] = (uint32_t)term_len;

    /* endmarker record: sequences alternate forward (starts at A_0 fwd = 2) and reverse (A_S rev) */
    size_t em_cap = 64 + 2 * (size_t)H;
    uint8_t* em = (uint8_t*)malloc(em_cap);
    size_t em_len = 0;
    em_len += put_varint(em + em_len, 2);
    em_len += put_varint(em + em_len, 2); em_len += put_varint(em + em_len, 0);
    em_len += put_varint(em + em_len, (2 * (3 * S + 1) + 1) - 2); em_len += put_varint(em + em_len, 0);
    for (uint64_t h = 0; h < H; h++) { em[em_len++] = 0; em[em_len++] = 1; } /* runs (0,1),(1,1) with sigma 2 */

    /* record starts: record 0 = endmarker, then for id = 1..3S+1: forward, reverse */
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
    uint8_t* image = assemble_image(sequences, size, 1, 2 * (3 * S + 2), 1 | 4, starts, records, NULL, data_len, out_len, &data_at);
    uint8_t* data = image + data_at;
    memcpy(data, em, em_len);
    memcpy(data + starts[2 * (3 * S + 1) - 1], term, term_len);
    memcpy(data + starts[2], term, term_len);

#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t job = 0; job < (int64_t)(2 * n_chunks); job++) {
        uint64_t c = (uint64_t)job / 2;
        int reverse = (int)(job & 1);
        uint64_t s0 = c * CH, s1 = s0 + CH < S ? s0 + CH : S;
        if (!reverse) {
            const uint8_t* src = fwd[c].buf;
            for (uint64_t s = s0; s < s1; s++) {
                uint64_t a = 3 * s + 1;
                for (uint64_t d = 0; d < 3; d++) {
                    memcpy(data + starts[2 * (a + d) - 1], src, len_f[a + d]);
                    src += len_f[a + d];
                }
            }
        } else {
            const uint8_t* src = rev[c].buf;
            for (uint64_t s = s1; s-- > s0;) {
                uint64_t a = 3 * s + 1;
                uint64_t ids[3] = {a + 3, a + 1, a + 2};
                for (uint64_t d = 0; d < 3; d++) {
                    memcpy(data + starts[2 * ids[d]], src, len_r[ids[d]]);
                    src += len_r[ids[d]];
                }
            }
        }
        free(reverse ? rev[c].buf : fwd[c].buf);
    }
    free(fwd); free(rev); free(len_f); free(len_r); free(em); free(starts);
    return image;
}

/* ---- haplotype paths and query patterns ----------------------------------------- */

/* Node at position p (0..2S) of forward sequence 2h. */
static inline uint64_t fwd_node(uint64_t seed, uint64_t S, uint64_t h, uint64_t p) {
    uint64_t s = p >> 1;
    if ((p & 1) == 0) return 2 * (3 * s + 1);
    return 2 * (3 * s + 2 + allele(seed, S, h, s));
}

static inline uint64_t seq_node(uint64_t seed, uint64_t S, uint64_t seq, uint64_t p) {
    uint64_t h = seq >> 1;
    if ((seq & 1) == 0) return fwd_node(seed, S, h, p);
    return fwd_node(seed, S, h, 2 * S - p) ^ 1;
}

/* Full sequence `seq` (2S+1 nodes). */
void synth_sequence(uint64_t S, uint64_t H, uint64_t seed, uint64_t seq, uint64_t* out) {
    (void)H;
    for (uint64_t p = 0; p <= 2 * S; p++) out[p] = seq_node(seed, S, seq, p);
}

/* Queries q0 .. q0+n: h = mix64(seed_q + 3q) % H, o = mix64(seed_q + 3q + 1) & 1,
 * t = mix64(seed_q + 3q + 2) % (2S + 1 - (k - 1)); pattern = sequence 2h+o positions [t, t+k). */
void synth_patterns(uint64_t S, uint64_t H, uint64_t seed, uint64_t seed_q, uint64_t q0, uint64_t n, uint64_t k,
                    uint64_t* out, int threads) {
#ifdef _OPENMP
    /* all processors unless told otherwise: launchers such as torchrun export OMP_NUM_THREADS=1 */
    if (threads <= 0) threads = omp_get_num_procs();
#else
    threads = 1;
#endif
    uint64_t span = 2 * S + 1 - (k - 1);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t q = q0 + (uint64_t)i;
        uint64_t h = mix64(seed_q + 3 * q) % H;
        uint64_t o = mix64(seed_q + 3 * q + 1) & 1;
        uint64_t t = mix64(seed_q + 3 * q + 2) % span;
        uint64_t* dst = out + (uint64_t)i * k;
        for (uint64_t j = 0; j < k; j++) dst[j] = seq_node(seed, S, 2 * h + o, t + j);
    }
}
