/*
 * gbwt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the gbwt-rs hot path: ByteCode/RLE codecs, BWT record
 * store, Record::{len,lf,follow,bd_follow,decompress,predecessor_at,offset_to} and
 * GBWT::{find,extend,bd_find,extend_forward,extend_backward,start,forward,backward,sequence}.
 * Every function cites the reference file:line it follows (paths relative to the
 * gbwt-rs source tree, crate `gbz` 0.5.1).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the CPU baseline. The product
 * library (libgbwt_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle_*.py check this code against every golden vector
 * the reference's own tests hold for the path and against the
 * shipped fixture files (tests/golden/).
 */
#ifndef GBWT_ORACLE_H
#define GBWT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/bwt.rs:63-69 (Pos), src/support.rs:115-121 (Run). */
typedef struct { uint64_t node, offset; } orc_pos;
typedef struct { uint64_t value, len; } orc_run;

/* src/gbwt.rs:454-474 (SearchState). None is returned as 0 and `out` set to {0,0,0}. */
typedef struct { uint64_t node, start, end; } orc_state;
/* src/gbwt.rs:484-528 (BidirectionalState). */
typedef struct { orc_state forward, reverse; } orc_bdstate;

typedef struct orc_gbwt orc_gbwt;

/* ---- codecs (src/support.rs:1048-1433) -------------------------------------------- */
/* ByteCode::write (support.rs:1068-1075). Appends to buf, returns bytes written (<= 10). */
size_t orc_bytecode_write(uint8_t* buf, uint64_t value);
/* ByteCodeIter::next (support.rs:1151-1164). Returns 1 and advances *pos, or 0 at end. */
int orc_bytecode_next(const uint8_t* bytes, size_t len, size_t* pos, uint64_t* out);
/* RLE::write_unchecked (support.rs:1229-1248) with sanitize (1292-1296). */
size_t orc_rle_write(uint8_t* buf, uint64_t sigma, orc_run run);
/* RLEIter::next (support.rs:1413-1430). */
int orc_rle_next(const uint8_t* bytes, size_t len, size_t* pos, uint64_t sigma, orc_run* out);

/* ---- loading (src/gbwt.rs:402-438, src/bwt.rs:176-185, src/gbz.rs:678-690) ---------- */
/* Accepts a Simple-SDS GBWT image or a GBZ image (the embedded GBWT is used).
 * Returns NULL and writes a message to err on InvalidData. */
orc_gbwt* orc_load_bytes(const uint8_t* bytes, size_t len, char* err, size_t errlen);
orc_gbwt* orc_load_file(const char* path, char* err, size_t errlen);
/* BWTBuilder::append + BWT::from (src/bwt.rs:192-203, 241-253): records given as
 * flattened edge / run arrays with per-record counts. Header fields are supplied. */
orc_gbwt* orc_from_records(uint64_t n_records, const uint64_t* edge_counts, const orc_pos* edges,
                           const uint64_t* run_counts, const orc_run* runs,
                           uint64_t sequences, uint64_t size, uint64_t offset, uint64_t alphabet_size,
                           int bidirectional);
void orc_free(orc_gbwt* g);

/* ---- statistics (src/gbwt.rs:105-175) ---------------------------------------------- */
uint64_t orc_len(const orc_gbwt* g);
uint64_t orc_sequences(const orc_gbwt* g);
uint64_t orc_alphabet_size(const orc_gbwt* g);
uint64_t orc_alphabet_offset(const orc_gbwt* g);
uint64_t orc_effective_size(const orc_gbwt* g);
uint64_t orc_first_node(const orc_gbwt* g);
int orc_has_node(const orc_gbwt* g, uint64_t id);
int orc_is_bidirectional(const orc_gbwt* g);
uint64_t orc_flags(const orc_gbwt* g);
/* raw BWT access (for layout / generator checks) */
uint64_t orc_bwt_records(const orc_gbwt* g);
uint64_t orc_bwt_data_len(const orc_gbwt* g);
const uint8_t* orc_bwt_data(const orc_gbwt* g);
/* record_bytes (src/bwt.rs:116-121): [start, limit) of record i. */
int orc_record_bytes(const orc_gbwt* g, uint64_t i, uint64_t* start, uint64_t* limit);

/* ---- record-level queries (src/bwt.rs:329-657); record id, not node id -------------- */
int64_t orc_record_outdegree(const orc_gbwt* g, uint64_t rec);               /* -1 = None */
int orc_record_edge(const orc_gbwt* g, uint64_t rec, uint64_t rank, orc_pos* out);
int64_t orc_record_len(const orc_gbwt* g, uint64_t rec);                     /* -1 = None */
int orc_record_lf(const orc_gbwt* g, uint64_t rec, uint64_t i, orc_pos* out);
int orc_record_follow(const orc_gbwt* g, uint64_t rec, uint64_t start, uint64_t end, uint64_t node,
                      uint64_t* out_start, uint64_t* out_end);
int orc_record_bd_follow(const orc_gbwt* g, uint64_t rec, uint64_t start, uint64_t end, uint64_t node,
                         uint64_t* out_start, uint64_t* out_end, uint64_t* out_count);
/* decompress: writes up to cap positions, returns the full length (or -1 for None). */
int64_t orc_record_decompress(const orc_gbwt* g, uint64_t rec, orc_pos* out, uint64_t cap);
int orc_record_predecessor_at(const orc_gbwt* g, uint64_t rec, uint64_t i, uint64_t* out_node);
int orc_record_offset_to(const orc_gbwt* g, uint64_t rec, orc_pos pos, uint64_t* out_offset);

/* ---- GBWT-level queries (src/gbwt.rs:208-385) -------------------------------------- */
int orc_find(const orc_gbwt* g, uint64_t node, orc_state* out);
int orc_extend(const orc_gbwt* g, const orc_state* state, uint64_t node, orc_state* out);
/* The bd_* functions return -1 where the reference panics (index not bidirectional). */
int orc_bd_find(const orc_gbwt* g, uint64_t node, orc_bdstate* out);
int orc_extend_forward(const orc_gbwt* g, const orc_bdstate* state, uint64_t node, orc_bdstate* out);
int orc_extend_backward(const orc_gbwt* g, const orc_bdstate* state, uint64_t node, orc_bdstate* out);
int orc_start(const orc_gbwt* g, uint64_t id, orc_pos* out);
int orc_forward(const orc_gbwt* g, orc_pos pos, orc_pos* out);
int orc_backward(const orc_gbwt* g, orc_pos pos, orc_pos* out);
/* sequence(id).collect(): writes up to cap nodes, returns the length, -1 if id >= sequences. */
int64_t orc_sequence(const orc_gbwt* g, uint64_t id, uint64_t* out, uint64_t cap);

/* ---- batch drivers (analogue of benchmark.rs:155-169 under rayon par_iter) ---------- */
/* n fixed-length patterns (k nodes each): find + (k-1) extends. threads<=0: all cores. */
void orc_find_extend_batch(const orc_gbwt* g, const uint64_t* patterns, uint64_t n, uint64_t k,
                           orc_state* out, int threads);
/* ragged patterns: pattern q = nodes[offsets[q] .. offsets[q+1]). Empty pattern -> None. */
void orc_find_extend_ragged(const orc_gbwt* g, const uint64_t* nodes, const uint64_t* offsets,
                            uint64_t n, orc_state* out, int threads);
void orc_find_batch(const orc_gbwt* g, const uint64_t* nodes, uint64_t n, orc_state* out, int threads);
void orc_extend_batch(const orc_gbwt* g, const orc_state* in, const uint64_t* nodes, uint64_t n,
                      orc_state* out, int threads);
int orc_bd_find_batch(const orc_gbwt* g, const uint64_t* nodes, uint64_t n, orc_bdstate* out, int threads);
int orc_bd_extend_batch(const orc_gbwt* g, const orc_bdstate* in, const uint64_t* nodes, uint64_t n,
                        int backward, orc_bdstate* out, int threads);
/* bd_search of src/gbwt/tests.rs:352-361 for ragged paths with (first, start, end). */
int orc_bd_search_batch(const orc_gbwt* g, const uint64_t* nodes, const uint64_t* offsets,
                        const uint64_t* first, const uint64_t* start, const uint64_t* end,
                        uint64_t n, orc_bdstate* out, int threads);
void orc_forward_batch(const orc_gbwt* g, const orc_pos* in, uint64_t n, orc_pos* out, int threads);
/* lengths[i] = length of sequence ids[i] (UINT64_MAX if id >= sequences). */
void orc_sequence_lengths(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint64_t* lengths, int threads);
/* Extract sequences ids[0..m) into nodes[out_offsets[i] .. out_offsets[i+1]). */
void orc_extract_batch(const orc_gbwt* g, const uint64_t* ids, uint64_t m, const uint64_t* out_offsets,
                       uint64_t* nodes, int threads);
int orc_max_threads(void);

/* ---- all extensions of a state: GBZ::follow_forward / follow_backward + StateIter (src/gbz.rs:519-544, 1223-1231) */
int64_t orc_follow(const orc_gbwt* g, const orc_bdstate* state, int backward, orc_bdstate* out, uint64_t cap);
void orc_follow_counts(const orc_gbwt* g, const orc_bdstate* states, uint64_t n, int backward, uint64_t* counts, int threads);
void orc_follow_batch(const orc_gbwt* g, const orc_bdstate* states, uint64_t n, int backward, const uint64_t* offsets,
                      orc_bdstate* out, int threads);

/* ---- node sequences and DNA-level extraction (GBZ files; src/graph.rs:295-330, src/gbz.rs:286-298,
 *      src/support.rs:87-110, src/bin/gbz-extract.rs:173-189) ------------------------------------------- */
int orc_has_graph(const orc_gbwt* g);
uint64_t orc_graph_sequences(const orc_gbwt* g);
int64_t orc_node_sequence(const orc_gbwt* g, uint64_t node_id, const uint8_t** seq);
void orc_reverse_complement(const uint8_t* seq, uint64_t len, uint8_t* out);
int64_t orc_extract_dna(const orc_gbwt* g, uint64_t seq_id, uint8_t endmarker, uint8_t* out, uint64_t cap);
void orc_dna_lengths(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint64_t* lengths, int threads);
void orc_extract_dna_batch(const orc_gbwt* g, const uint64_t* ids, uint64_t m, uint8_t endmarker, const uint64_t* offsets,
                           uint8_t* out, int threads);

/* ---- algorithmic-byte accounting of SURVEY.md 8(d) (measurement helper for bench.py) ------------ */
uint64_t orc_find_extend_bytes(const orc_gbwt* g, const uint64_t* patterns, uint64_t n, uint64_t k, int threads);
uint64_t orc_extract_bytes(const orc_gbwt* g, const uint64_t* ids, uint64_t m, int threads);

#ifdef __cplusplus
}
#endif
#endif
