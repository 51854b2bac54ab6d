/*
 * gbwt_oracle.c -- TEST INFRASTRUCTURE ONLY (see gbwt_oracle.h).
 *
 * Plain-C restatement of the gbwt-rs CPU path. It deliberately keeps the reference's
 * per-call structure (Elias-Fano select per record fetch, a heap-allocated edge vector
 * per Record, a second allocation in lf(), byte-serial run decoding with the
 * reference's early exits) so that it can double as the "C++/C restatement of the
 * gbwt-rs CPU path" baseline of BASELINE.md section 3. It is never linked into the product.
 *
 * Third-party piece: simple-sds 0.4 (un-vendored crates.io dependency, no lock file).
 * Only its *container format* (SparseVector / BitVector / IntVector / Vec / Option
 * serialization) and SparseVector::select_iter are touched by the path; both are
 * restated here from the format rules in SURVEY.md App. A and checked against the
 * reference's fixture files by tests/test_oracle_fixtures.py.
 */
#include "gbwt_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_ENDMARKER 0ULL /* src/lib.rs:59 */

/* ------------------------------------------------------------------------------------------ */
/* ByteCode (src/support.rs:1048-1164)                                                        */
/* ------------------------------------------------------------------------------------------ */

/* ByteCode::write, support.rs:1068-1075: 7 data bits per byte, high bit = "continues". */
size_t orc_bytecode_write(uint8_t* buf, uint64_t value) {
    size_t n = 0;
    while (value > 0x7F) {
        buf[n++] = (uint8_t)((value & 0x7F) | 0x80);
        value >>= 7;
    }
    buf[n++] = (uint8_t)value;
    return n;
}

/* ByteCodeIter::next, support.rs:1151-1164. A truncated integer at the end yields None. */
int orc_bytecode_next(const uint8_t* bytes, size_t len, size_t* pos, uint64_t* out) {
    unsigned shift = 0;
    uint64_t result = 0;
    while (*pos < len) {
        uint8_t value = bytes[*pos];
        *pos += 1;
        /* Rust `<<` on usize panics/wraps past 63 bits; valid encodings never get there. */
        if (shift < 64) result += ((uint64_t)(value & 0x7F)) << shift;
        shift += 7;
        if ((value & 0x80) == 0) { *out = result; return 1; }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* RLE (src/support.rs:1166-1433)                                                             */
/* ------------------------------------------------------------------------------------------ */

#define ORC_RLE_THRESHOLD 255ULL /* support.rs:1197 */
#define ORC_RLE_UNIVERSE 256ULL  /* support.rs:1198 */

/* RLE::sanitize, support.rs:1292-1296. */
static void rle_sanitize(uint64_t sigma, uint64_t* eff_sigma, uint64_t* threshold) {
    uint64_t s = (sigma == 0) ? UINT64_MAX : sigma;
    *eff_sigma = s;
    *threshold = (s < ORC_RLE_THRESHOLD) ? (ORC_RLE_UNIVERSE / s) : 0;
}

/* RLE::write_unchecked + write_basic, support.rs:1229-1248, 1286-1289. */
size_t orc_rle_write(uint8_t* buf, uint64_t sigma, orc_run run) {
    uint64_t s, threshold;
    rle_sanitize(sigma, &s, &threshold);
    size_t n = 0;
    if (run.len == 0) return 0; /* RLE::write, support.rs:1221-1223 */
    if (s >= ORC_RLE_THRESHOLD) {
        n += orc_bytecode_write(buf + n, run.value);
        n += orc_bytecode_write(buf + n, run.len - 1);
    } else if (run.len < threshold) {
        buf[n++] = (uint8_t)(run.value + s * (run.len - 1));
    } else {
        buf[n++] = (uint8_t)(run.value + s * (threshold - 1));
        n += orc_bytecode_write(buf + n, run.len - threshold);
    }
    return n;
}

/* RLEIter::next, support.rs:1413-1430. */
int orc_rle_next(const uint8_t* bytes, size_t len, size_t* pos, uint64_t sigma, orc_run* out) {
    uint64_t s, threshold;
    rle_sanitize(sigma, &s, &threshold);
    orc_run run = {0, 0};
    if (s >= ORC_RLE_THRESHOLD) {
        uint64_t v;
        if (!orc_bytecode_next(bytes, len, pos, &v)) return 0;
        run.value = v;
        if (!orc_bytecode_next(bytes, len, pos, &v)) return 0;
        run.len = v + 1;
    } else {
        if (*pos >= len) return 0;
        uint8_t byte = bytes[*pos];
        *pos += 1;
        run.value = (uint64_t)byte % s;
        run.len = (uint64_t)byte / s + 1;
        if (run.len == threshold) {
            uint64_t v;
            if (!orc_bytecode_next(bytes, len, pos, &v)) return 0;
            run.len += v;
        }
    }
    *out = run;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* simple-sds SparseVector (Elias-Fano), format per SURVEY.md App. A                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t universe;    /* SparseVector::len() */
    uint64_t ones;        /* count_ones() */
    uint64_t high_len;    /* bits in high */
    uint64_t* high;       /* unary-coded high parts */
    uint64_t high_words;
    uint64_t low_width;
    uint64_t* low;        /* packed low parts */
    uint64_t low_words;
    uint64_t* samples;    /* bit position in `high` of every 64th one (select support) */
    uint64_t n_samples;
} orc_sparse;

static void sparse_free(orc_sparse* sv) {
    free(sv->high); free(sv->low); free(sv->samples);
    memset(sv, 0, sizeof(*sv));
}

static int sparse_build_select(orc_sparse* sv) {
    sv->n_samples = sv->ones / 64 + 1;
    sv->samples = (uint64_t*)malloc(sv->n_samples * sizeof(uint64_t));
    if (!sv->samples) return 0;
    uint64_t seen = 0;
    for (uint64_t w = 0; w < sv->high_words; w++) {
        uint64_t word = sv->high[w];
        uint64_t pc = (uint64_t)__builtin_popcountll(word);
        /* Does a multiple of 64 fall in (seen, seen + pc]? At most one per word. */
        uint64_t next_multiple = ((seen + 63) / 64) * 64;
        if (pc > 0 && next_multiple < seen + pc) {
            uint64_t k = next_multiple - seen; /* k-th one (0-based) of this word */
            uint64_t tmp = word;
            for (uint64_t j = 0; j < k; j++) tmp &= tmp - 1;
            sv->samples[next_multiple / 64] = w * 64 + (uint64_t)__builtin_ctzll(tmp);
        }
        seen += pc;
    }
    if (seen != sv->ones) return 0;
    if (sv->ones % 64 == 0) sv->samples[sv->ones / 64] = sv->high_len; /* sentinel */
    return 1;
}

/* Position in `high` of the i-th one (i < ones). */
static uint64_t sparse_select_high(const orc_sparse* sv, uint64_t i) {
    uint64_t pos = sv->samples[i / 64];
    uint64_t remaining = i % 64;
    uint64_t w = pos / 64;
    uint64_t word = sv->high[w] & (~0ULL << (pos % 64));
    for (;;) {
        uint64_t pc = (uint64_t)__builtin_popcountll(word);
        if (remaining < pc) break;
        remaining -= pc;
        w++;
        word = sv->high[w];
    }
    for (uint64_t j = 0; j < remaining; j++) word &= word - 1;
    return w * 64 + (uint64_t)__builtin_ctzll(word);
}

static inline uint64_t sparse_low(const orc_sparse* sv, uint64_t i) {
    uint64_t width = sv->low_width;
    if (width == 0) return 0;
    uint64_t bit = i * width;
    uint64_t w = bit / 64, off = bit % 64;
    uint64_t v = sv->low[w] >> off;
    if (off + width > 64) v |= sv->low[w + 1] << (64 - off);
    if (width < 64) v &= (1ULL << width) - 1;
    return v;
}

/* SparseVector::select_iter(i).next() twice (src/bwt.rs:117-119): value of the i-th one and,
 * if it exists, of the (i+1)-th (found by scanning `high` forward like OneIter does). */
static void sparse_select_pair(const orc_sparse* sv, uint64_t i, uint64_t* first, int want_second, uint64_t* second) {
    uint64_t pos = sparse_select_high(sv, i);
    *first = ((pos - i) << sv->low_width) | sparse_low(sv, i);
    if (want_second) {
        uint64_t p = pos + 1;
        uint64_t w = p / 64;
        uint64_t word = (w < sv->high_words) ? (sv->high[w] & (~0ULL << (p % 64))) : 0;
        while (word == 0) { w++; word = sv->high[w]; }
        uint64_t pos2 = w * 64 + (uint64_t)__builtin_ctzll(word);
        *second = ((pos2 - (i + 1)) << sv->low_width) | sparse_low(sv, i + 1);
    }
}

/* SparseBuilder parameters (simple-sds; restated in SURVEY.md App. A). */
static void sparse_params(uint64_t universe, uint64_t ones, uint64_t* low_width, uint64_t* buckets) {
    uint64_t w = 1;
    if (ones > 0 && ones <= universe) {
        double ideal = log2(((double)universe * log(2.0)) / (double)ones);
        double r = round(ideal);
        w = (r < 1.0) ? 1 : (uint64_t)r;
    }
    uint64_t b = (w < 64) ? (universe >> w) : 0;
    uint64_t mask = (w < 64) ? ((1ULL << w) - 1) : ~0ULL;
    if (universe & mask) b++;
    *low_width = w; *buckets = b;
}

/* Build from strictly increasing values < universe. */
static int sparse_from_values(orc_sparse* sv, uint64_t universe, const uint64_t* values, uint64_t ones) {
    memset(sv, 0, sizeof(*sv));
    uint64_t w, buckets;
    sparse_params(universe, ones, &w, &buckets);
    sv->universe = universe; sv->ones = ones; sv->low_width = w;
    sv->high_len = ones + buckets;
    sv->high_words = (sv->high_len + 63) / 64 + 1;
    sv->high = (uint64_t*)calloc(sv->high_words, 8);
    sv->low_words = (ones * w + 63) / 64 + 1;
    sv->low = (uint64_t*)calloc(sv->low_words, 8);
    if (!sv->high || !sv->low) return 0;
    for (uint64_t j = 0; j < ones; j++) {
        uint64_t v = values[j];
        uint64_t hp = (v >> w) + j;
        sv->high[hp / 64] |= 1ULL << (hp % 64);
        uint64_t lv = v & ((w < 64) ? ((1ULL << w) - 1) : ~0ULL);
        uint64_t bit = j * w;
        sv->low[bit / 64] |= lv << (bit % 64);
        if ((bit % 64) + w > 64) sv->low[bit / 64 + 1] |= lv >> (64 - bit % 64);
    }
    return sparse_build_select(sv);
}

/* ------------------------------------------------------------------------------------------ */
/* Simple-SDS reader                                                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct { const uint8_t* p; size_t len; size_t pos; int ok; } orc_reader;

static uint64_t rd_u64(orc_reader* r) {
    if (!r->ok || r->pos + 8 > r->len) { r->ok = 0; return 0; }
    uint64_t v; memcpy(&v, r->p + r->pos, 8); r->pos += 8;
    return v;
}
static void rd_skip(orc_reader* r, uint64_t elements) {
    if (!r->ok || elements > (r->len - r->pos) / 8) { r->ok = 0; return; }
    r->pos += (size_t)elements * 8;
}
static uint64_t* rd_words(orc_reader* r, uint64_t n, uint64_t extra) {
    if (!r->ok || n > (r->len - r->pos) / 8) { r->ok = 0; return NULL; }
    uint64_t* out = (uint64_t*)calloc((size_t)(n + extra), 8);
    if (!out) { r->ok = 0; return NULL; }
    memcpy(out, r->p + r->pos, (size_t)n * 8);
    r->pos += (size_t)n * 8;
    return out;
}
/* Option<T>: size in elements, then T (0 = None). Skipped by declared size. */
static void rd_skip_option(orc_reader* r) { uint64_t n = rd_u64(r); rd_skip(r, n); }
/* RawVector: len_bits, n_words, words. */
static uint64_t* rd_rawvector(orc_reader* r, uint64_t* len_bits, uint64_t* n_words) {
    *len_bits = rd_u64(r);
    *n_words = rd_u64(r);
    if (r->ok && *n_words != (*len_bits + 63) / 64) { r->ok = 0; return NULL; }
    return rd_words(r, *n_words, 1);
}
static void rd_skip_rawvector(orc_reader* r) { (void)rd_u64(r); uint64_t n = rd_u64(r); rd_skip(r, n); }
static void rd_skip_intvector(orc_reader* r) { (void)rd_u64(r); (void)rd_u64(r); rd_skip_rawvector(r); }
static void rd_skip_bitvector(orc_reader* r) {
    (void)rd_u64(r); rd_skip_rawvector(r);
    rd_skip_option(r); rd_skip_option(r); rd_skip_option(r);
}
static void rd_skip_sparse(orc_reader* r) { (void)rd_u64(r); rd_skip_bitvector(r); rd_skip_intvector(r); }
static void rd_skip_bytes(orc_reader* r) { uint64_t n = rd_u64(r); rd_skip(r, (n + 7) / 8); }
/* Tags = StringArray (src/support.rs:600-643, 981-1007): SparseVector, Vec<u8>, IntVector. */
static void rd_skip_tags(orc_reader* r) { rd_skip_sparse(r); rd_skip_bytes(r); rd_skip_intvector(r); }

static int rd_sparse(orc_reader* r, orc_sparse* sv) {
    memset(sv, 0, sizeof(*sv));
    sv->universe = rd_u64(r);
    sv->ones = rd_u64(r); /* BitVector: ones, data, 3 optional supports */
    sv->high = rd_rawvector(r, &sv->high_len, &sv->high_words);
    rd_skip_option(r); rd_skip_option(r); rd_skip_option(r);
    uint64_t low_len = rd_u64(r);
    sv->low_width = rd_u64(r);
    uint64_t low_bits = 0;
    sv->low = rd_rawvector(r, &low_bits, &sv->low_words);
    if (!r->ok) return 0;
    if (low_len != sv->ones || sv->low_width > 64 || low_bits != low_len * sv->low_width) return 0;
    return sparse_build_select(sv);
}

/* ------------------------------------------------------------------------------------------ */
/* BWT and GBWT containers                                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    orc_sparse index;   /* src/bwt.rs:98 */
    uint8_t* data;      /* src/bwt.rs:99 */
    uint64_t data_len;
} orc_bwt;

struct orc_gbwt {
    /* Header<GBWTPayload>, src/headers.rs:57-62, 190-203 */
    uint32_t tag, version;
    uint64_t sequences, size, offset, alphabet_size, flags;
    orc_bwt bwt;
    orc_pos* endmarker; /* src/gbwt.rs:99, 413-414 */
    uint64_t endmarker_len;
    /* Graph::sequences (src/graph.rs:84-89) when loaded from a GBZ file: concatenated node labels. */
    int has_graph;
    uint64_t graph_nodes, n_seq;
    uint64_t* seq_starts; /* n_seq + 1 */
    uint8_t* seq_bytes;
};

#define GBWT_TAG 0x6B376B37u
#define GBWT_VERSION 5u
#define GBWT_FLAG_BIDIRECTIONAL 1ULL
#define GBWT_FLAG_METADATA 2ULL
#define GBWT_FLAG_SIMPLE_SDS 4ULL
#define GBZ_TAG 0x205A4247u

/* BWT::len, src/bwt.rs:105-107. */
static inline uint64_t bwt_len(const orc_bwt* b) { return b->index.ones; }

/* BWT::record_bytes, src/bwt.rs:116-121. */
static void bwt_record_bytes(const orc_bwt* b, uint64_t i, uint64_t* start, uint64_t* limit) {
    int has_next = (i + 1 < bwt_len(b));
    uint64_t next = 0;
    sparse_select_pair(&b->index, i, start, has_next, &next);
    *limit = has_next ? next : b->data_len;
}

/* Record, src/bwt.rs:333-337. `edges` is heap-allocated per record like the reference's Vec. */
typedef struct {
    uint64_t id;
    orc_pos* edges;
    uint64_t sigma;
    const uint8_t* bwt;
    uint64_t bwt_len;
} orc_record;

static void record_drop(orc_record* rec) { free(rec->edges); rec->edges = NULL; }

/* Record::new + decompress_edges, src/bwt.rs:341-351, 378-395. Returns 0 for None. */
static int record_new(uint64_t id, const uint8_t* bytes, uint64_t len, orc_record* rec) {
    if (len == 0) return 0;
    size_t pos = 0;
    uint64_t sigma;
    if (!orc_bytecode_next(bytes, len, &pos, &sigma)) return 0;
    if (sigma == 0) return 0;
    orc_pos* edges = (orc_pos*)malloc((size_t)sigma * sizeof(orc_pos));
    if (!edges) return 0;
    uint64_t prev = 0;
    for (uint64_t i = 0; i < sigma; i++) {
        uint64_t delta, offset;
        if (!orc_bytecode_next(bytes, len, &pos, &delta) || !orc_bytecode_next(bytes, len, &pos, &offset)) {
            free(edges); return 0;
        }
        uint64_t node = delta + prev;
        prev = node;
        edges[i].node = node; edges[i].offset = offset;
    }
    rec->id = id; rec->edges = edges; rec->sigma = sigma;
    rec->bwt = bytes + pos; rec->bwt_len = len - pos;
    return 1;
}

/* BWT::record, src/bwt.rs:124-130. */
static int bwt_record(const orc_bwt* b, uint64_t i, orc_record* rec) {
    if (i >= bwt_len(b)) return 0;
    uint64_t start, limit;
    bwt_record_bytes(b, i, &start, &limit);
    return record_new(i, b->data + start, limit - start, rec);
}

/* Record::len, src/bwt.rs:449-455. */
static uint64_t record_len(const orc_record* rec) {
    uint64_t result = 0;
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) result += run.len;
    return result;
}

/* Record::decompress, src/bwt.rs:465-475. Returns the length; writes at most cap items. */
static uint64_t record_decompress(const orc_record* rec, orc_pos* out, uint64_t cap) {
    orc_pos* edges = (orc_pos*)malloc((size_t)rec->sigma * sizeof(orc_pos));
    memcpy(edges, rec->edges, (size_t)rec->sigma * sizeof(orc_pos));
    uint64_t n = 0;
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) {
        for (uint64_t j = 0; j < run.len; j++) {
            if (n < cap) out[n] = edges[run.value];
            n++;
            edges[run.value].offset += 1;
        }
    }
    free(edges);
    return n;
}

/* Record::lf, src/bwt.rs:480-496 (including the clone of the edge vector at :481). */
static int record_lf(const orc_record* rec, uint64_t i, orc_pos* out) {
    orc_pos* edges = (orc_pos*)malloc((size_t)rec->sigma * sizeof(orc_pos));
    memcpy(edges, rec->edges, (size_t)rec->sigma * sizeof(orc_pos));
    uint64_t offset = 0;
    size_t pos = 0; orc_run run;
    int found = 0;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) {
        if (offset + run.len > i) {
            if (rec->edges[run.value].node == ORC_ENDMARKER) {
                found = 0;
            } else {
                edges[run.value].offset += i - offset;
                *out = edges[run.value];
                found = 1;
            }
            break;
        }
        edges[run.value].offset += run.len;
        offset += run.len;
    }
    free(edges);
    return found;
}

/* support::node_id / flip_node, src/support.rs:163-165, 188-190. */
static inline uint64_t node_id(uint64_t id) { return id / 2; }
static inline uint64_t flip_node(uint64_t id) { return id ^ 1; }

/* Record::predecessor_at, src/bwt.rs:502-540. */
static int record_predecessor_at(const orc_record* rec, uint64_t i, uint64_t* out) {
    orc_pos* edges = (orc_pos*)malloc((size_t)rec->sigma * sizeof(orc_pos));
    for (uint64_t rank = 0; rank < rec->sigma; rank++) { edges[rank].node = rec->edges[rank].node; edges[rank].offset = 0; }
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) edges[run.value].offset += run.len;
    for (uint64_t rank = 0; rank < rec->sigma; rank++)
        if (edges[rank].node != ORC_ENDMARKER) edges[rank].node = flip_node(edges[rank].node);
    for (uint64_t rank = 1; rank < rec->sigma; rank++) {
        if (node_id(edges[rank - 1].node) == node_id(edges[rank].node)) {
            orc_pos t = edges[rank - 1]; edges[rank - 1] = edges[rank]; edges[rank] = t;
        }
    }
    uint64_t offset = 0;
    int found = 0;
    for (uint64_t rank = 0; rank < rec->sigma; rank++) {
        offset += edges[rank].offset;
        if (offset > i) {
            if (edges[rank].node != ORC_ENDMARKER) { *out = edges[rank].node; found = 1; }
            break;
        }
    }
    free(edges);
    return found;
}

/* Record::edge_to, src/bwt.rs:543-555 (binary search). Returns rank or -1. */
static int64_t record_edge_to(const orc_record* rec, uint64_t node) {
    uint64_t low = 0, high = rec->sigma;
    while (low < high) {
        uint64_t mid = low + (high - low) / 2;
        uint64_t m = rec->edges[mid].node;
        if (node < m) high = mid;
        else if (node == m) return (int64_t)mid;
        else low = mid + 1;
    }
    return -1;
}

/* Record::offset_to, src/bwt.rs:558-584. */
static int record_offset_to(const orc_record* rec, orc_pos p, uint64_t* out) {
    if (p.node == ORC_ENDMARKER) return 0;
    int64_t outrank = record_edge_to(rec, p.node);
    if (outrank < 0) return 0;
    uint64_t succ_rank = rec->edges[outrank].offset;
    if (succ_rank > p.offset) return 0;
    uint64_t offset = 0;
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) {
        offset += run.len;
        if (run.value != (uint64_t)outrank) continue;
        succ_rank += run.len;
        if (succ_rank > p.offset) { *out = offset - (succ_rank - p.offset); return 1; }
    }
    return 0;
}

/* support::intersect(...).len(), src/support.rs:332-334 with Range::len (empty if start >= end). */
static inline uint64_t intersect_len(uint64_t a_start, uint64_t a_end, uint64_t b_start, uint64_t b_end) {
    uint64_t s = a_start > b_start ? a_start : b_start;
    uint64_t e = a_end < b_end ? a_end : b_end;
    return e > s ? e - s : 0;
}

/* Record::follow, src/bwt.rs:595-616. */
static int record_follow(const orc_record* rec, uint64_t start, uint64_t end, uint64_t node,
                         uint64_t* out_start, uint64_t* out_end) {
    if (start >= end || node == ORC_ENDMARKER) return 0;
    int64_t rank = record_edge_to(rec, node);
    if (rank < 0) return 0;
    uint64_t rs = rec->edges[rank].offset, re = rs;
    uint64_t offset = 0;
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) {
        if (run.value == (uint64_t)rank) {
            rs += intersect_len(offset, offset + run.len, 0, start);
            re += intersect_len(offset, offset + run.len, 0, end);
        }
        offset += run.len;
        if (offset >= end) break;
    }
    if (rs >= re) return 0;
    *out_start = rs; *out_end = re;
    return 1;
}

/* Record::bd_follow, src/bwt.rs:630-656. */
static int record_bd_follow(const orc_record* rec, uint64_t start, uint64_t end, uint64_t node,
                            uint64_t* out_start, uint64_t* out_end, uint64_t* out_count) {
    if (start >= end || node == ORC_ENDMARKER) return 0;
    int64_t rank = record_edge_to(rec, node);
    if (rank < 0) return 0;
    uint64_t reverse = flip_node(node);
    uint64_t rs = rec->edges[rank].offset, re = rs;
    uint64_t count = 0, offset = 0;
    size_t pos = 0; orc_run run;
    while (orc_rle_next(rec->bwt, rec->bwt_len, &pos, rec->sigma, &run)) {
        if (run.value == (uint64_t)rank) {
            rs += intersect_len(offset, offset + run.len, 0, start);
            re += intersect_len(offset, offset + run.len, 0, end);
        }
        if (flip_node(rec->edges[run.value].node) < reverse)
            count += intersect_len(offset, offset + run.len, start, end);
        offset += run.len;
        if (offset >= end) break;
    }
    if (rs >= re) return 0;
    *out_start = rs; *out_end = re; *out_count = count;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Loading                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static void set_err(char* err, size_t errlen, const char* msg) {
    if (err && errlen) { snprintf(err, errlen, "%s", msg); }
}

static void bwt_free(orc_bwt* b) { sparse_free(&b->index); free(b->data); b->data = NULL; }

void orc_free(orc_gbwt* g) {
    if (!g) return;
    bwt_free(&g->bwt);
    free(g->endmarker);
    free(g->seq_starts);
    free(g->seq_bytes);
    free(g);
}

/* The endmarker is decompressed at load (src/gbwt.rs:413-414). */
static int gbwt_finish(orc_gbwt* g) {
    g->endmarker = NULL; g->endmarker_len = 0;
    if (bwt_len(&g->bwt) == 0) return 1;
    orc_record rec;
    if (!bwt_record(&g->bwt, ORC_ENDMARKER, &rec)) return 0; /* the reference unwrap()s here */
    uint64_t n = record_len(&rec);
    g->endmarker = (orc_pos*)malloc((size_t)(n ? n : 1) * sizeof(orc_pos));
    g->endmarker_len = record_decompress(&rec, g->endmarker, n);
    record_drop(&rec);
    return 1;
}

/* GBWT::load, src/gbwt.rs:402-438; Header::validate, src/headers.rs:102-115, 226-231. */
static orc_gbwt* load_gbwt(orc_reader* r, char* err, size_t errlen) {
    orc_gbwt* g = (orc_gbwt*)calloc(1, sizeof(orc_gbwt));
    uint64_t tv = rd_u64(r);
    g->tag = (uint32_t)(tv & 0xFFFFFFFFu); g->version = (uint32_t)(tv >> 32);
    g->sequences = rd_u64(r); g->size = rd_u64(r); g->offset = rd_u64(r);
    g->alphabet_size = rd_u64(r); g->flags = rd_u64(r);
    if (!r->ok) { set_err(err, errlen, "GBWTHeader: unexpected end of data"); goto fail; }
    if (g->tag != GBWT_TAG) { set_err(err, errlen, "GBWTHeader: Invalid tag"); goto fail; }
    if (g->version != GBWT_VERSION) { set_err(err, errlen, "GBWTHeader: Invalid version"); goto fail; }
    if (g->flags & ~(GBWT_FLAG_BIDIRECTIONAL | GBWT_FLAG_METADATA | GBWT_FLAG_SIMPLE_SDS)) {
        set_err(err, errlen, "GBWTHeader: Invalid flags"); goto fail;
    }
    if (!(g->flags & GBWT_FLAG_SIMPLE_SDS)) { set_err(err, errlen, "GBWTHeader: SDSL format is not supported"); goto fail; }
    rd_skip_tags(r);
    if (!r->ok) { set_err(err, errlen, "Tags: invalid data"); goto fail; }
    /* BWT::load, src/bwt.rs:176-185 */
    if (!rd_sparse(r, &g->bwt.index)) { set_err(err, errlen, "BWT: invalid index"); goto fail; }
    g->bwt.data_len = rd_u64(r);
    if (!r->ok || (g->bwt.data_len + 7) / 8 > (r->len - r->pos) / 8) { set_err(err, errlen, "BWT: invalid data"); goto fail; }
    g->bwt.data = (uint8_t*)malloc((size_t)g->bwt.data_len + 8);
    memcpy(g->bwt.data, r->p + r->pos, (size_t)g->bwt.data_len);
    r->pos += (size_t)((g->bwt.data_len + 7) / 8) * 8;
    if (g->bwt.index.universe != g->bwt.data_len) { set_err(err, errlen, "BWT: Index / data length mismatch"); goto fail; }
    if (!gbwt_finish(g)) { set_err(err, errlen, "GBWT: missing endmarker record"); goto fail; }
    /* DA samples: Vec<u64> passed through (src/gbwt.rs:417). */
    { uint64_t n = rd_u64(r); rd_skip(r, n); }
    /* Option<Metadata> (src/gbwt.rs:420-423). */
    {
        uint64_t n = rd_u64(r);
        rd_skip(r, n);
        if (!r->ok) { set_err(err, errlen, "GBWT: unexpected end of data"); goto fail; }
        int has_meta = (n != 0);
        if (((g->flags & GBWT_FLAG_METADATA) != 0) != has_meta) {
            set_err(err, errlen, "GBWT: Invalid metadata flag in the header"); goto fail;
        }
    }
    return g;
fail:
    orc_free(g);
    return NULL;
}

static int load_graph(orc_reader* r, orc_gbwt* g, char* err, size_t errlen);

orc_gbwt* orc_load_bytes(const uint8_t* bytes, size_t len, char* err, size_t errlen) {
    orc_reader r = { bytes, len, 0, 1 };
    if (len < 8) { set_err(err, errlen, "file too short"); return NULL; }
    uint32_t tag; memcpy(&tag, bytes, 4);
    if (tag == GBZ_TAG) {
        /* GBZ::load, src/gbz.rs:678-690: header (tag|version, flags), Tags, then the GBWT. */
        uint64_t tv = rd_u64(&r);
        uint32_t version = (uint32_t)(tv >> 32);
        uint64_t flags = rd_u64(&r);
        if (version < 1 || version > 2) { set_err(err, errlen, "GBZHeader: Invalid version"); return NULL; }
        if (flags != 0) { set_err(err, errlen, "GBZHeader: Invalid flags"); return NULL; }
        rd_skip_tags(&r);
        orc_gbwt* g = load_gbwt(&r, err, errlen);
        if (g && !(g->flags & GBWT_FLAG_BIDIRECTIONAL)) {
            set_err(err, errlen, "GBZ: The GBWT index is not bidirectional");
            orc_free(g); return NULL;
        }
        if (g && !load_graph(&r, g, err, errlen)) { orc_free(g); return NULL; }
        return g;
    }
    return load_gbwt(&r, err, errlen);
}

orc_gbwt* orc_load_file(const char* path, char* err, size_t errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { set_err(err, errlen, "cannot open file"); return NULL; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc((size_t)n + 8);
    size_t got = fread(buf, 1, (size_t)n, f);
    fclose(f);
    orc_gbwt* g = (got == (size_t)n) ? orc_load_bytes(buf, (size_t)n, err, errlen) : NULL;
    free(buf);
    return g;
}

/* BWTBuilder::append (src/bwt.rs:241-253) for every record, then BWT::from (192-203). */
orc_gbwt* orc_from_records(uint64_t n_records, const uint64_t* edge_counts, const orc_pos* edges,
                           const uint64_t* run_counts, const orc_run* runs,
                           uint64_t sequences, uint64_t size, uint64_t offset, uint64_t alphabet_size,
                           int bidirectional) {
    orc_gbwt* g = (orc_gbwt*)calloc(1, sizeof(orc_gbwt));
    g->tag = GBWT_TAG; g->version = GBWT_VERSION;
    g->sequences = sequences; g->size = size; g->offset = offset; g->alphabet_size = alphabet_size;
    g->flags = GBWT_FLAG_SIMPLE_SDS | (bidirectional ? GBWT_FLAG_BIDIRECTIONAL : 0);
    uint64_t total_edges = 0, total_runs = 0;
    for (uint64_t i = 0; i < n_records; i++) { total_edges += edge_counts[i]; total_runs += run_counts[i]; }
    size_t cap = (size_t)(10 * n_records + 20 * total_edges + 20 * total_runs + 16);
    uint8_t* data = (uint8_t*)malloc(cap);
    uint64_t* offsets = (uint64_t*)malloc((size_t)(n_records ? n_records : 1) * 8);
    size_t n = 0;
    const orc_pos* e = edges; const orc_run* rn = runs;
    for (uint64_t i = 0; i < n_records; i++) {
        offsets[i] = n;
        n += orc_bytecode_write(data + n, edge_counts[i]);
        uint64_t prev = 0;
        for (uint64_t j = 0; j < edge_counts[i]; j++, e++) {
            n += orc_bytecode_write(data + n, e->node - prev);
            n += orc_bytecode_write(data + n, e->offset);
            prev = e->node;
        }
        for (uint64_t j = 0; j < run_counts[i]; j++, rn++) n += orc_rle_write(data + n, edge_counts[i], *rn);
    }
    g->bwt.data = data; g->bwt.data_len = n;
    int ok = sparse_from_values(&g->bwt.index, n, offsets, n_records);
    free(offsets);
    if (!ok || !gbwt_finish(g)) { orc_free(g); return NULL; }
    return g;
}

/* ------------------------------------------------------------------------------------------ */
/* Statistics, src/gbwt.rs:105-175                                                            */
/* ------------------------------------------------------------------------------------------ */

uint64_t orc_len(const orc_gbwt* g) { return g->size; }
uint64_t orc_sequences(const orc_gbwt* g) { return g->sequences; }
uint64_t orc_alphabet_size(const orc_gbwt* g) { return g->alphabet_size; }
uint64_t orc_alphabet_offset(const orc_gbwt* g) { return g->offset; }
uint64_t orc_effective_size(const orc_gbwt* g) { return g->alphabet_size - g->offset; }
uint64_t orc_first_node(const orc_gbwt* g) { return g->offset + 1; }
int orc_has_node(const orc_gbwt* g, uint64_t id) { return id > g->offset && id < g->alphabet_size; }
int orc_is_bidirectional(const orc_gbwt* g) { return (g->flags & GBWT_FLAG_BIDIRECTIONAL) != 0; }
uint64_t orc_flags(const orc_gbwt* g) { return g->flags; }
uint64_t orc_bwt_records(const orc_gbwt* g) { return bwt_len(&g->bwt); }
uint64_t orc_bwt_data_len(const orc_gbwt* g) { return g->bwt.data_len; }
const uint8_t* orc_bwt_data(const orc_gbwt* g) { return g->bwt.data; }

int orc_record_bytes(const orc_gbwt* g, uint64_t i, uint64_t* start, uint64_t* limit) {
    if (i >= bwt_len(&g->bwt)) return 0;
    bwt_record_bytes(&g->bwt, i, start, limit);
    return 1;
}

/* node_to_record, src/gbwt.rs:150-152. Rust's usize subtraction wraps in release builds, which
 * makes BWT::record return None for node < offset; restated as an explicit failure. */
static inline int node_to_record(const orc_gbwt* g, uint64_t node, uint64_t* rec) {
    if (node < g->offset) return 0;
    *rec = node - g->offset;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Record-level entry points                                                                  */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_record_outdegree(const orc_gbwt* g, uint64_t rec_id) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return -1;
    int64_t s = (int64_t)rec.sigma; record_drop(&rec); return s;
}
int orc_record_edge(const orc_gbwt* g, uint64_t rec_id, uint64_t rank, orc_pos* out) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = rank < rec.sigma; if (ok) *out = rec.edges[rank];
    record_drop(&rec); return ok;
}
int64_t orc_record_len(const orc_gbwt* g, uint64_t rec_id) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return -1;
    int64_t n = (int64_t)record_len(&rec); record_drop(&rec); return n;
}
int orc_record_lf(const orc_gbwt* g, uint64_t rec_id, uint64_t i, orc_pos* out) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = record_lf(&rec, i, out); record_drop(&rec); return ok;
}
int orc_record_follow(const orc_gbwt* g, uint64_t rec_id, uint64_t start, uint64_t end, uint64_t node,
                      uint64_t* out_start, uint64_t* out_end) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = record_follow(&rec, start, end, node, out_start, out_end); record_drop(&rec); return ok;
}
int orc_record_bd_follow(const orc_gbwt* g, uint64_t rec_id, uint64_t start, uint64_t end, uint64_t node,
                         uint64_t* out_start, uint64_t* out_end, uint64_t* out_count) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = record_bd_follow(&rec, start, end, node, out_start, out_end, out_count); record_drop(&rec); return ok;
}
int64_t orc_record_decompress(const orc_gbwt* g, uint64_t rec_id, orc_pos* out, uint64_t cap) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return -1;
    int64_t n = (int64_t)record_decompress(&rec, out, cap); record_drop(&rec); return n;
}
int orc_record_predecessor_at(const orc_gbwt* g, uint64_t rec_id, uint64_t i, uint64_t* out_node) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = record_predecessor_at(&rec, i, out_node); record_drop(&rec); return ok;
}
int orc_record_offset_to(const orc_gbwt* g, uint64_t rec_id, orc_pos pos, uint64_t* out_offset) {
    orc_record rec; if (!bwt_record(&g->bwt, rec_id, &rec)) return 0;
    int ok = record_offset_to(&rec, pos, out_offset); record_drop(&rec); return ok;
}

/* ------------------------------------------------------------------------------------------ */
/* GBWT-level queries                                                                         */
/* ------------------------------------------------------------------------------------------ */

static const orc_state STATE_NONE = {0, 0, 0};
static const orc_pos POS_NONE = {0, 0};

/* GBWT::find, src/gbwt.rs:269-281. */
int orc_find(const orc_gbwt* g, uint64_t node, orc_state* out) {
    *out = STATE_NONE;
    if (node < orc_first_node(g)) return 0;
    orc_record rec;
    uint64_t rid;
    if (!node_to_record(g, node, &rid) || !bwt_record(&g->bwt, rid, &rec)) return 0;
    out->node = node; out->start = 0; out->end = record_len(&rec);
    record_drop(&rec);
    return 1;
}

/* GBWT::extend, src/gbwt.rs:292-304. */
int orc_extend(const orc_gbwt* g, const orc_state* state, uint64_t node, orc_state* out) {
    orc_state st = *state; /* out may alias state */
    *out = STATE_NONE;
    if (node < orc_first_node(g)) return 0;
    orc_record rec;
    uint64_t rid;
    if (!node_to_record(g, st.node, &rid) || !bwt_record(&g->bwt, rid, &rec)) return 0;
    uint64_t s, e;
    int ok = record_follow(&rec, st.start, st.end, node, &s, &e);
    record_drop(&rec);
    if (!ok) return 0;
    out->node = node; out->start = s; out->end = e;
    return 1;
}

static void bd_none(orc_bdstate* out) { out->forward = STATE_NONE; out->reverse = STATE_NONE; }

/* GBWT::bd_find, src/gbwt.rs:311-324. */
int orc_bd_find(const orc_gbwt* g, uint64_t node, orc_bdstate* out) {
    bd_none(out);
    if (!orc_is_bidirectional(g)) return -1; /* assert!, gbwt.rs:312 */
    orc_state st;
    if (!orc_find(g, node, &st)) return 0;
    out->forward = st;
    out->reverse.node = flip_node(st.node); out->reverse.start = st.start; out->reverse.end = st.end;
    return 1;
}

/* GBWT::bd_internal, src/gbwt.rs:370-384. */
static int bd_internal(const orc_record* rec, const orc_bdstate* state, uint64_t node, orc_bdstate* out) {
    uint64_t s, e, count;
    if (!record_bd_follow(rec, state->forward.start, state->forward.end, node, &s, &e, &count)) return 0;
    uint64_t pos = state->reverse.start + count;
    out->forward.node = node; out->forward.start = s; out->forward.end = e;
    out->reverse.node = state->reverse.node; out->reverse.start = pos; out->reverse.end = pos + (e - s);
    return 1;
}

/* GBWT::extend_forward, src/gbwt.rs:339-347. */
int orc_extend_forward(const orc_gbwt* g, const orc_bdstate* state, uint64_t node, orc_bdstate* out) {
    orc_bdstate st = *state;
    bd_none(out);
    if (!orc_is_bidirectional(g)) return -1; /* assert!, gbwt.rs:340 */
    if (node < orc_first_node(g)) return 0;
    orc_record rec;
    uint64_t rid;
    if (!node_to_record(g, st.forward.node, &rid) || !bwt_record(&g->bwt, rid, &rec)) return 0;
    orc_bdstate res;
    int ok = bd_internal(&rec, &st, node, &res);
    record_drop(&rec);
    if (ok) *out = res;
    return ok;
}

/* GBWT::extend_backward, src/gbwt.rs:362-367 with BidirectionalState::flip (506-511). */
int orc_extend_backward(const orc_gbwt* g, const orc_bdstate* state, uint64_t node, orc_bdstate* out) {
    orc_bdstate flipped = { state->reverse, state->forward };
    orc_bdstate res;
    int ok = orc_extend_forward(g, &flipped, flip_node(node), &res);
    bd_none(out);
    if (ok <= 0) return ok;
    out->forward = res.reverse; out->reverse = res.forward;
    return 1;
}

/* GBWT::start, src/gbwt.rs:213-219. */
int orc_start(const orc_gbwt* g, uint64_t id, orc_pos* out) {
    *out = POS_NONE;
    if (id < g->endmarker_len && g->endmarker[id].node != ORC_ENDMARKER) { *out = g->endmarker[id]; return 1; }
    return 0;
}

/* GBWT::forward, src/gbwt.rs:222-229. */
int orc_forward(const orc_gbwt* g, orc_pos pos, orc_pos* out) {
    *out = POS_NONE;
    if (pos.node < orc_first_node(g)) return 0;
    orc_record rec;
    uint64_t rid;
    if (!node_to_record(g, pos.node, &rid) || !bwt_record(&g->bwt, rid, &rec)) return 0;
    orc_pos res;
    int ok = record_lf(&rec, pos.offset, &res);
    record_drop(&rec);
    if (ok) *out = res;
    return ok;
}

/* GBWT::backward, src/gbwt.rs:236-250. Returns -1 where the reference panics. */
int orc_backward(const orc_gbwt* g, orc_pos pos, orc_pos* out) {
    *out = POS_NONE;
    if (!orc_is_bidirectional(g)) return -1; /* assert!, gbwt.rs:237 */
    if (pos.node <= orc_first_node(g)) return 0;
    uint64_t reverse_id;
    orc_record rec;
    if (!node_to_record(g, flip_node(pos.node), &reverse_id) || !bwt_record(&g->bwt, reverse_id, &rec)) return 0;
    uint64_t predecessor;
    int ok = record_predecessor_at(&rec, pos.offset, &predecessor);
    record_drop(&rec);
    if (!ok) return 0;
    uint64_t pred_id;
    orc_record pred;
    if (!node_to_record(g, predecessor, &pred_id) || !bwt_record(&g->bwt, pred_id, &pred)) return 0;
    uint64_t offset;
    ok = record_offset_to(&pred, pos, &offset);
    record_drop(&pred);
    if (!ok) return 0;
    out->node = predecessor; out->offset = offset;
    return 1;
}

/* GBWT::sequence + SequenceIter::next, src/gbwt.rs:253-261, 557-568. */
int64_t orc_sequence(const orc_gbwt* g, uint64_t id, uint64_t* out, uint64_t cap) {
    if (id >= g->sequences) return -1;
    uint64_t n = 0;
    orc_pos pos;
    int some = orc_start(g, id, &pos);
    while (some) {
        if (out && n < cap) out[n] = pos.node;
        n++;
        orc_pos next;
        some = orc_forward(g, pos, &next);
        pos = next;
    }
    return (int64_t)n;
}

/* ------------------------------------------------------------------------------------------ */
/* Batch drivers: dynamic scheduling over queries = the rayon par_iter analogue               */
/* ------------------------------------------------------------------------------------------ */

/* All processors of the host, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1). */
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

static int pick_threads(int threads) {
    if (threads <= 0) return orc_max_threads();
    return threads;
}

/* benchmark.rs:161-167: find(query[0]) then extend over the rest. */
static void search_one(const orc_gbwt* g, const uint64_t* pat, uint64_t k, orc_state* out) {
    *out = STATE_NONE;
    if (k == 0) return;
    orc_state st;
    if (!orc_find(g, pat[0], &st)) return;
    for (uint64_t i = 1; i < k; i++) {
        orc_state nx;
        if (!orc_extend(g, &st, pat[i], &nx)) return;
        st = nx;
    }
    *out = st;
}

void orc_find_extend_batch(const orc_gbwt* g, const uint64_t* patterns, uint64_t n, uint64_t k,
                           orc_state* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 256) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) search_one(g, patterns + (uint64_t)q * k, k, &out[q]);
}

void orc_find_extend_ragged(const orc_gbwt* g, const uint64_t* nodes, const uint64_t* offsets,
                            uint64_t n, orc_state* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 256) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++)
        search_one(g, nodes + offsets[q], offsets[q + 1] - offsets[q], &out[q]);
}

void orc_find_batch(const orc_gbwt* g, const uint64_t* nodes, uint64_t n, orc_state* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) orc_find(g, nodes[q], &out[q]);
}

void orc_extend_batch(const orc_gbwt* g, const orc_state* in, const uint64_t* nodes, uint64_t n,
                      orc_state* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) orc_extend(g, &in[q], nodes[q], &out[q]);
}

int orc_bd_find_batch(const orc_gbwt* g, const uint64_t* nodes, uint64_t n, orc_bdstate* out, int threads) {
    if (!orc_is_bidirectional(g)) return -1;
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) orc_bd_find(g, nodes[q], &out[q]);
    return 0;
}

int orc_bd_extend_batch(const orc_gbwt* g, const orc_bdstate* in, const uint64_t* nodes, uint64_t n,
                        int backward, orc_bdstate* out, int threads) {
    if (!orc_is_bidirectional(g)) return -1;
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) {
        if (backward) orc_extend_backward(g, &in[q], nodes[q], &out[q]);
        else orc_extend_forward(g, &in[q], nodes[q], &out[q]);
    }
    return 0;
}

/* bd_search, src/gbwt/tests.rs:352-361. */
static void bd_search_one(const orc_gbwt* g, const uint64_t* path, uint64_t first, uint64_t start, uint64_t end,
                          orc_bdstate* out) {
    bd_none(out);
    orc_bdstate st;
    if (orc_bd_find(g, path[first], &st) <= 0) return;
    for (uint64_t i = first + 1; i < end; i++) {
        orc_bdstate nx;
        if (orc_extend_forward(g, &st, path[i], &nx) <= 0) return;
        st = nx;
    }
    for (uint64_t i = first; i > start; i--) {
        orc_bdstate nx;
        if (orc_extend_backward(g, &st, path[i - 1], &nx) <= 0) return;
        st = nx;
    }
    *out = st;
}

int orc_bd_search_batch(const orc_gbwt* g, const uint64_t* nodes, const uint64_t* offsets,
                        const uint64_t* first, const uint64_t* start, const uint64_t* end,
                        uint64_t n, orc_bdstate* out, int threads) {
    if (!orc_is_bidirectional(g)) return -1;
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 64) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++)
        bd_search_one(g, nodes + offsets[q], first[q], start[q], end[q], &out[q]);
    return 0;
}

void orc_forward_batch(const orc_gbwt* g, const orc_pos* in, uint64_t n, orc_pos* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(t)
    for (int64_t q = 0; q < (int64_t)n; q++) orc_forward(g, in[q], &out[q]);
}

void orc_sequence_lengths(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint64_t* lengths, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1) num_threads(t)
    for (int64_t i = 0; i < (int64_t)m; i++) {
        int64_t n = orc_sequence(g, ids[i], NULL, 0);
        lengths[i] = (n < 0) ? UINT64_MAX : (uint64_t)n;
    }
}

void orc_extract_batch(const orc_gbwt* g, const uint64_t* ids, uint64_t m, const uint64_t* out_offsets,
                       uint64_t* nodes, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1) num_threads(t)
    for (int64_t i = 0; i < (int64_t)m; i++)
        (void)orc_sequence(g, ids[i], nodes + out_offsets[i], out_offsets[i + 1] - out_offsets[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Algorithmic-byte accounting: measurement helper, not a query path          */
/* ------------------------------------------------------------------------------------------ */

/* Bytes of record `rid` the reference's loops consume: the header (decompress_edges decodes all of it,
 * src/bwt.rs:378-395) plus the run bytes read until the loop exits. mode 0 = Record::len (whole record,
 * src/bwt.rs:449-455); mode 1 = follow with early exit at offset >= end (src/bwt.rs:610-612), only when the
 * edge exists; mode 2 = lf, through the run containing i (src/bwt.rs:484). */
static uint64_t consumed_bytes(const orc_gbwt* g, uint64_t rid, int mode, uint64_t end_or_i, uint64_t node) {
    orc_record rec;
    if (!bwt_record(&g->bwt, rid, &rec)) {
        uint64_t s, l;
        if (rid >= bwt_len(&g->bwt)) return 0;
        bwt_record_bytes(&g->bwt, rid, &s, &l);
        return l - s; /* an empty record is still fetched and its sigma decoded */
    }
    uint64_t s, l;
    bwt_record_bytes(&g->bwt, rid, &s, &l);
    uint64_t header = (l - s) - rec.bwt_len;
    uint64_t used = header;
    if (mode == 1 && record_edge_to(&rec, node) < 0) { record_drop(&rec); return used; }
    size_t pos = 0; orc_run run; uint64_t offset = 0;
    while (orc_rle_next(rec.bwt, rec.bwt_len, &pos, rec.sigma, &run)) {
        offset += run.len;
        if (mode == 1 && offset >= end_or_i) break;
        if (mode == 2 && offset > end_or_i) break;
    }
    used += pos;
    record_drop(&rec);
    return used;
}

/* Sum over queries of: per step 16 (two record boundaries) + 8 (pattern symbol) + consumed record bytes,
 * plus 24 bytes of result per query. */
uint64_t orc_find_extend_bytes(const orc_gbwt* g, const uint64_t* patterns, uint64_t n, uint64_t k, int threads) {
    int t = pick_threads(threads); (void)t;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(t) reduction(+:total)
    for (int64_t q = 0; q < (int64_t)n; q++) {
        const uint64_t* pat = patterns + (uint64_t)q * k;
        uint64_t bytes = 24;
        orc_state st;
        uint64_t rid;
        if (k > 0) {
            bytes += 8;
            if (pat[0] >= orc_first_node(g) && node_to_record(g, pat[0], &rid) && rid < bwt_len(&g->bwt))
                bytes += 16 + consumed_bytes(g, rid, 0, 0, 0);
            int some = orc_find(g, pat[0], &st);
            for (uint64_t i = 1; i < k && some; i++) {
                bytes += 8;
                if (pat[i] >= orc_first_node(g) && node_to_record(g, st.node, &rid) && rid < bwt_len(&g->bwt))
                    bytes += 16 + consumed_bytes(g, rid, 1, st.end, pat[i]);
                orc_state nx;
                some = orc_extend(g, &st, pat[i], &nx);
                st = nx;
            }
        }
        total += bytes;
    }
    return total;
}

/* Sum over sequences of: per forward() call 16 + consumed record bytes (lf), plus 8 bytes per extracted node. */
uint64_t orc_extract_bytes(const orc_gbwt* g, const uint64_t* ids, uint64_t m, int threads) {
    int t = pick_threads(threads); (void)t;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(t) reduction(+:total)
    for (int64_t j = 0; j < (int64_t)m; j++) {
        uint64_t bytes = 0;
        if (ids[j] < g->sequences) {
            orc_pos pos;
            int some = orc_start(g, ids[j], &pos);
            while (some) {
                bytes += 8;
                uint64_t rid;
                if (pos.node >= orc_first_node(g) && node_to_record(g, pos.node, &rid) && rid < bwt_len(&g->bwt))
                    bytes += 16 + consumed_bytes(g, rid, 2, pos.offset, 0);
                orc_pos next;
                some = orc_forward(g, pos, &next);
                pos = next;
            }
        }
        total += bytes;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* All extensions of a state: GBZ::follow_forward / follow_backward    */
/* ------------------------------------------------------------------------------------------ */

/* GBZ::has_node, src/gbz.rs:286-289: the forward GBWT node is in the alphabet and its record is non-empty
 * (the real_nodes bit set at load from BWT::id_iter, src/gbz.rs:695-704, src/bwt.rs:310-316). */
static int gbz_has_node(const orc_gbwt* g, uint64_t node_id) {
    uint64_t gbwt_node = 2 * node_id;
    if (!orc_has_node(g, gbwt_node)) return 0;
    uint64_t start, limit;
    bwt_record_bytes(&g->bwt, gbwt_node - g->offset, &start, &limit);
    return limit > start && g->bwt.data[start] != 0;
}

/* GBZ::follow_forward / follow_backward + StateIter::next (src/gbz.rs:519-544, 1223-1231) over
 * EdgeIter (src/gbz.rs:835-861): every successor of the last node in edge order, skipping an edge to the
 * endmarker, extended with GBWT::bd_internal; backward = the same on the flipped state, results flipped back.
 * Writes at most cap states; returns their number, or -1 where the reference returns None. */
int64_t orc_follow(const orc_gbwt* g, const orc_bdstate* state, int backward, orc_bdstate* out, uint64_t cap) {
    orc_bdstate st = *state;
    if (backward) { st.forward = state->reverse; st.reverse = state->forward; }
    if (!gbz_has_node(g, st.forward.node / 2)) return -1;        /* GBZ::successors, src/gbz.rs:327-335 */
    orc_record rec;
    uint64_t rid;
    if (!node_to_record(g, st.forward.node, &rid) || !bwt_record(&g->bwt, rid, &rec)) return -1;
    uint64_t n = 0;
    uint64_t rank = (rec.sigma > 0 && rec.edges[0].node == ORC_ENDMARKER) ? 1 : 0; /* EdgeIter::new */
    for (; rank < rec.sigma; rank++) {
        orc_bdstate next;
        if (!bd_internal(&rec, &st, rec.edges[rank].node, &next)) continue;
        if (backward) { orc_state t = next.forward; next.forward = next.reverse; next.reverse = t; }
        if (n < cap) out[n] = next;
        n++;
    }
    record_drop(&rec);
    return (int64_t)n;
}

void orc_follow_counts(const orc_gbwt* g, const orc_bdstate* states, uint64_t n, int backward, uint64_t* counts, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 256) num_threads(t)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        int64_t c = orc_follow(g, &states[i], backward, NULL, 0);
        counts[i] = c < 0 ? UINT64_MAX : (uint64_t)c;
    }
}

void orc_follow_batch(const orc_gbwt* g, const orc_bdstate* states, uint64_t n, int backward, const uint64_t* offsets,
                      orc_bdstate* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 256) num_threads(t)
    for (int64_t i = 0; i < (int64_t)n; i++)
        (void)orc_follow(g, &states[i], backward, out + offsets[i], offsets[i + 1] - offsets[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Graph: node sequences of a GBZ file                                 */
/* ------------------------------------------------------------------------------------------ */

#include <dlfcn.h>

#define GRAPH_TAG 0x6B3764AFu

/* zstd 0.13 crate = libzstd; the shared library is present without headers, so the two stable entry points
 * are declared here and resolved with dlopen. */
static int zstd_decompress(uint8_t* dst, uint64_t dst_len, const uint8_t* src, uint64_t src_len, uint64_t* got) {
    typedef size_t (*decompress_fn)(void*, size_t, const void*, size_t);
    typedef unsigned (*iserror_fn)(size_t);
    static decompress_fn decompress = NULL;
    static iserror_fn iserror = NULL;
    if (!decompress) {
        void* h = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h) return 0;
        decompress = (decompress_fn)dlsym(h, "ZSTD_decompress");
        iserror = (iserror_fn)dlsym(h, "ZSTD_isError");
        if (!decompress || !iserror) return 0;
    }
    size_t n = decompress(dst, (size_t)dst_len, src, (size_t)src_len);
    if (iserror(n)) return 0;
    *got = n;
    return 1;
}

/* All values of a serialized SparseVector (string start offsets). */
static uint64_t* rd_sparse_values(orc_reader* r, uint64_t* count) {
    orc_sparse sv;
    if (!rd_sparse(r, &sv)) { sparse_free(&sv); return NULL; }
    uint64_t* out = (uint64_t*)malloc((size_t)(sv.ones + 1) * 8);
    for (uint64_t i = 0; i < sv.ones; i++) {
        uint64_t v, dummy;
        sparse_select_pair(&sv, i, &v, 0, &dummy);
        out[i] = v;
    }
    *count = sv.ones;
    sparse_free(&sv);
    return out;
}

/* Graph::load, src/graph.rs:295-330: header, node sequences (StringArray::decompress for version >= 4,
 * src/support.rs:545-571; StringArray::load otherwise, src/support.rs:619-643). The segment names and the
 * node-to-segment mapping that follow are not on the path and are left unread. */
static int load_graph(orc_reader* r, orc_gbwt* g, char* err, size_t errlen) {
    uint64_t tv = rd_u64(r);
    uint32_t tag = (uint32_t)tv, version = (uint32_t)(tv >> 32);
    g->graph_nodes = rd_u64(r);
    uint64_t flags = rd_u64(r);
    if (!r->ok || tag != GRAPH_TAG) { set_err(err, errlen, "GraphHeader: Invalid tag"); return 0; }
    if (version < 3 || version > 4) { set_err(err, errlen, "GraphHeader: Invalid version"); return 0; }
    if (flags & ~3ULL) { set_err(err, errlen, "GraphHeader: Invalid flags"); return 0; }
    if (!(flags & 2ULL)) { set_err(err, errlen, "GraphHeader: SDSL format is not supported"); return 0; }
    uint64_t n = 0;
    uint64_t* starts = rd_sparse_values(r, &n);
    if (!starts) { set_err(err, errlen, "StringArray: invalid index"); return 0; }
    uint64_t total = 0;
    uint8_t* bytes = NULL;
    if (version >= 4) {
        total = rd_u64(r);
        uint64_t clen = rd_u64(r);
        if (!r->ok || (clen + 7) / 8 > (r->len - r->pos) / 8) { free(starts); set_err(err, errlen, "StringArray: invalid data"); return 0; }
        bytes = (uint8_t*)malloc((size_t)total + 8);
        uint64_t got = 0;
        int ok = zstd_decompress(bytes, total, r->p + r->pos, clen, &got);
        r->pos += (size_t)((clen + 7) / 8) * 8;
        if (!ok || got != total) {
            free(starts); free(bytes);
            set_err(err, errlen, "StringArray: Decompressed string length does not match the expected length");
            return 0;
        }
    } else {
        uint64_t alen = rd_u64(r);
        if (!r->ok || (alen + 7) / 8 > (r->len - r->pos) / 8) { free(starts); set_err(err, errlen, "StringArray: invalid alphabet"); return 0; }
        const uint8_t* alphabet = r->p + r->pos;
        r->pos += (size_t)((alen + 7) / 8) * 8;
        total = rd_u64(r);
        uint64_t width = rd_u64(r), bits = 0, nwords = 0;
        uint64_t* packed = rd_rawvector(r, &bits, &nwords);
        if (!r->ok || !packed || width == 0 || width > 64 || bits != total * width) { free(starts); free(packed); set_err(err, errlen, "StringArray: invalid strings"); return 0; }
        bytes = (uint8_t*)malloc((size_t)total + 8);
        for (uint64_t i = 0; i < total; i++) {
            uint64_t bit = i * width, w = bit / 64, off = bit % 64;
            uint64_t v = packed[w] >> off;
            if (off + width > 64) v |= packed[w + 1] << (64 - off);
            if (width < 64) v &= (1ULL << width) - 1;
            bytes[i] = v < alen ? alphabet[v] : 0;
        }
        free(packed);
    }
    if (n > 0 && starts[0] != 0) { free(starts); free(bytes); set_err(err, errlen, "StringArray: First string does not start at offset 0"); return 0; }
    starts[n] = total;
    /* GBZ::load, src/gbz.rs:690-694 */
    if (n != (g->alphabet_size - (g->offset + 1)) / 2) {
        free(starts); free(bytes);
        set_err(err, errlen, "GBZ: Mismatch between GBWT alphabet size and Graph sequence count");
        return 0;
    }
    g->has_graph = 1; g->n_seq = n; g->seq_starts = starts; g->seq_bytes = bytes;
    /* Segment names (StringArray::load) and the node-to-segment mapping (SparseVector) are not used on the
     * path; they are walked so that a truncated file fails like it does in the reference (src/graph.rs:309-328). */
    uint64_t segments = 0;
    {
        (void)rd_u64(r);                    /* universe */
        segments = rd_u64(r);               /* BitVector::ones = number of segment names */
        rd_skip_rawvector(r); rd_skip_option(r); rd_skip_option(r); rd_skip_option(r);
        rd_skip_intvector(r);
        rd_skip_bytes(r); rd_skip_intvector(r);
    }
    uint64_t mapping_len = rd_u64(r), mapping_ones = rd_u64(r);
    rd_skip_rawvector(r); rd_skip_option(r); rd_skip_option(r); rd_skip_option(r);
    rd_skip_intvector(r);
    if (!r->ok) { set_err(err, errlen, "Graph: unexpected end of data"); return 0; }
    if (((flags & 1ULL) != 0) == (segments == 0)) {
        set_err(err, errlen, "Graph: Translation flag does not match the presence of segment names"); return 0;
    }
    if (flags & 1ULL) {
        if (mapping_len <= g->graph_nodes) { set_err(err, errlen, "Graph: Node-to-segment mapping does not match the number of nodes"); return 0; }
        if (mapping_len != n + 1) { set_err(err, errlen, "Graph: Node-to-segment mapping does not match the number of sequences"); return 0; }
        if (mapping_ones != segments) { set_err(err, errlen, "Graph: Node-to-segment mapping does not match the number of segments"); return 0; }
    }
    return 1;
}

int orc_has_graph(const orc_gbwt* g) { return g->has_graph; }
uint64_t orc_graph_sequences(const orc_gbwt* g) { return g->n_seq; }

/* GBZ::sequence(node_id), src/gbz.rs:292-298 with Graph::sequence (src/graph.rs:124-126). Returns the length
 * and sets *seq, or -1 for None. */
int64_t orc_node_sequence(const orc_gbwt* g, uint64_t node_id, const uint8_t** seq) {
    if (!g->has_graph || !gbz_has_node(g, node_id)) return -1;
    uint64_t sequence_id = (2 * node_id - (g->offset + 1)) / 2; /* gbwt_node_to_sequence, src/gbz.rs:253-255 */
    if (sequence_id >= g->n_seq) return -1;
    *seq = g->seq_bytes + g->seq_starts[sequence_id];
    return (int64_t)(g->seq_starts[sequence_id + 1] - g->seq_starts[sequence_id]);
}

/* support::COMPLEMENT, src/support.rs:87-98: upper-case complement, anything else -> 'N'. */
static uint8_t complement(uint8_t c) {
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': return 'A';
    default: return 'N';
    }
}

/* support::reverse_complement, src/support.rs:104-110. */
void orc_reverse_complement(const uint8_t* seq, uint64_t len, uint8_t* out) {
    for (uint64_t i = 0; i < len; i++) out[i] = complement(seq[len - 1 - i]);
}

/* extract_sequence, src/bin/gbz-extract.rs:173-189, for GBWT sequence `seq_id` (= encode_path(path, orientation)):
 * the labels of the nodes on the path, reverse-complemented for reverse-oriented nodes, then the endmarker
 * byte. Writes at most cap bytes and returns the full length, or -1 where gbz.path() is None. */
int64_t orc_extract_dna(const orc_gbwt* g, uint64_t seq_id, uint8_t endmarker, uint8_t* out, uint64_t cap) {
    if (!g->has_graph || seq_id >= g->sequences) return -1;
    uint64_t n = 0;
    orc_pos pos;
    int some = orc_start(g, seq_id, &pos);
    while (some) {
        const uint8_t* seq = NULL;
        int64_t len = orc_node_sequence(g, pos.node / 2, &seq); /* the reference unwrap()s */
        for (int64_t j = 0; j < len; j++) {
            uint8_t c = (pos.node & 1) ? complement(seq[len - 1 - j]) : seq[j];
            if (out && n < cap) out[n] = c;
            n++;
        }
        orc_pos next;
        some = orc_forward(g, pos, &next);
        pos = next;
    }
    if (out && n < cap) out[n] = endmarker;
    n++;
    return (int64_t)n;
}

void orc_dna_lengths(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint64_t* lengths, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1) num_threads(t)
    for (int64_t i = 0; i < (int64_t)m; i++) {
        int64_t n = orc_extract_dna(g, ids[i], 0, NULL, 0);
        lengths[i] = n < 0 ? UINT64_MAX : (uint64_t)n;
    }
}

void orc_extract_dna_batch(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint8_t endmarker, const uint64_t* offsets,
                           uint8_t* out, int threads) {
    int t = pick_threads(threads); (void)t;
#pragma omp parallel for schedule(dynamic, 1) num_threads(t)
    for (int64_t i = 0; i < (int64_t)m; i++)
        (void)orc_extract_dna(g, ids[i], endmarker, out + offsets[i], offsets[i + 1] - offsets[i]);
}
