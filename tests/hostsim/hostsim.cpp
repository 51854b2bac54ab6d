// hostsim.cpp -- TEST-ONLY harness: runs the product's layout builder (K0) and the one-lane query code of
// record_scan.cuh on the CPU so that the device layout and the scan logic can be checked against the
// oracle in `-m "not gpu"` tests, on a machine without a GPU.
//
// This is NOT a CPU fallback: nothing in libgbwt_b200.so links or calls it, and it is built only by the
// tests (tests/hostsim_build.py) into tests/hostsim/libgbwt_hostsim.so.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../gbwt-rs_b200/csrc/layout_builder.h"
#include "../../gbwt-rs_b200/csrc/layout_writer.h"
#include "../../gbwt-rs_b200/csrc/record_scan.cuh"
#include "../../gbwt-rs_b200/csrc/sds_loader.h"

using namespace gbwt_b200;

struct HostSim {
    HostLayout layout;
    IndexView view;
    ParsedGBWT parsed;
};

extern "C" {

HostSim* hs_load(const uint8_t* bytes, size_t len, int policy, char* err, size_t errlen) {
    HostSim* h = new HostSim();
    std::string msg;
    int rc;
    try {
        rc = parse_gbwt_image(bytes, len, h->parsed, msg);
        if (rc == GBWT_B200_OK) rc = build_layout(h->parsed, policy, h->layout, msg);
    } catch (const std::exception& e) {  // like the C ABI: sizes in a damaged image can be absurd
        rc = GBWT_B200_E_INVALID_DATA;
        msg = std::string("invalid data (") + e.what() + ")";
    }
    if (rc != GBWT_B200_OK) {
        if (err && errlen) std::snprintf(err, errlen, "%d: %s", rc, msg.c_str());
        delete h;
        return nullptr;
    }
    IndexView& v = h->view;
    v.desc = h->layout.desc.data();
    v.bodies = reinterpret_cast<const Unit16*>(h->layout.bodies.data());
    v.edges = h->layout.edges.data();
    v.endmarker = h->layout.endmarker.data();
    v.records = h->layout.desc.size();
    v.offset = h->parsed.offset;
    v.alphabet_size = h->parsed.alphabet_size;
    v.sequences = h->parsed.sequences;
    v.endmarker_len = h->layout.endmarker.size();
    v.bidirectional = (h->parsed.flags & GBWT_FLAG_BIDIRECTIONAL) != 0;
    v.skips = reinterpret_cast<const Unit16*>(h->layout.skips.data());
    v.edges_valid = h->layout.edges_valid ? 1 : 0;
    v.walk_limit = h->layout.total_length;
    h->parsed.bwt = nullptr;  // the image may go away
    return h;
}

void hs_free(HostSim* h) { delete h; }

// The layout written back as a GBWT image (layout_writer.cpp): returns the size; fills `out` when it is large enough.
uint64_t hs_serialize(const HostSim* h, int gbz, uint8_t* out, uint64_t cap) {
    LayoutArrays in;
    in.desc = h->layout.desc.data(); in.records = h->layout.desc.size();
    in.bodies = h->layout.bodies.data(); in.edges = h->layout.edges.data();
    GBWTHeaderFields header;
    header.sequences = h->parsed.sequences; header.size = h->parsed.size; header.offset = h->parsed.offset;
    header.alphabet_size = h->parsed.alphabet_size; header.flags = h->parsed.flags;
    std::vector<uint8_t> image;
    std::string err;
    Carried carried;
    carried.tags = h->parsed.tags; carried.gbz_tags = h->parsed.gbz_tags; carried.da_samples = h->parsed.da_samples;
    carried.metadata = h->parsed.metadata; carried.graph_section = h->parsed.graph_section;
    const int rc = gbz ? write_gbz_image(header, in, carried, h->parsed.has_graph ? h->parsed.label_starts.data() : nullptr,
                                         h->parsed.has_graph ? h->parsed.label_starts.size() - 1 : 0, h->parsed.label_bytes.data(), image, err)
                       : write_gbwt_image(header, in, carried, image, err);
    if (rc != GBWT_B200_OK) return 0;
    if (out != nullptr && cap >= image.size()) std::memcpy(out, image.data(), image.size());
    return image.size();
}

// Forgets the Graph section the image came with, as if the node labels had been attached by hand: the GBZ writer then
// writes a version-3 Graph from the labels.
void hs_drop_graph_section(HostSim* h) { h->parsed.graph_section.clear(); }

// Path-walk shortcuts of K0 pass 3 (layout.h, IndexView::skips): 4 x u32 per record; and the descriptor words the
// shortcuts are derived from: out[0..7] = {total_len, meta, w0, w1, body, body_len, w2, w3}.
const uint32_t* hs_skips(const HostSim* h) { return reinterpret_cast<const uint32_t*>(h->layout.skips.data()); }
int hs_edges_valid(const HostSim* h) { return h->layout.edges_valid ? 1 : 0; }
uint64_t hs_records(const HostSim* h) { return h->layout.desc.size(); }
void hs_desc_words(const HostSim* h, uint64_t rec, uint32_t* out) { std::memcpy(out, &h->layout.desc[rec], 32); }

// Node labels as the product's loader parsed them (sds_loader.cpp parse_graph): has_graph, count, starts, bytes.
int hs_has_graph(const HostSim* h) { return h->parsed.has_graph ? 1 : 0; }
uint64_t hs_label_count(const HostSim* h) { return h->parsed.has_graph ? h->parsed.label_starts.size() - 1 : 0; }
const uint64_t* hs_label_starts(const HostSim* h) { return h->parsed.label_starts.data(); }
const uint8_t* hs_label_bytes(const HostSim* h) { return h->parsed.label_bytes.data(); }

void hs_format_counts(const HostSim* h, uint64_t* out) { std::memcpy(out, h->layout.format_counts, sizeof(h->layout.format_counts)); }
uint64_t hs_body_bytes(const HostSim* h) { return h->layout.bodies.size() * 8; }
uint64_t hs_checkpointed_records(const HostSim* h) { return h->layout.checkpointed_records; }
int hs_record_format(const HostSim* h, uint64_t rec) { return h->layout.desc[rec].fmt; }

void hs_find(const HostSim* h, const uint64_t* nodes, size_t n, gbwt_b200_state* out) {
    for (size_t i = 0; i < n; i++) gbwt_find(h->view, nodes[i], out[i]);
}
void hs_extend(const HostSim* h, const gbwt_b200_state* in, const uint64_t* nodes, size_t n, gbwt_b200_state* out) {
    for (size_t i = 0; i < n; i++) { gbwt_b200_state s = in[i]; gbwt_extend(h->view, s, nodes[i], out[i]); }
}
void hs_find_extend(const HostSim* h, const uint64_t* patterns, size_t n, size_t k, gbwt_b200_state* out) {
    for (size_t i = 0; i < n; i++) query_find_extend(h->view, patterns + i * k, k, out[i]);
}
void hs_bd_find(const HostSim* h, const uint64_t* nodes, size_t n, gbwt_b200_bdstate* out) {
    for (size_t i = 0; i < n; i++) gbwt_bd_find(h->view, nodes[i], out[i]);
}
void hs_bd_extend(const HostSim* h, const gbwt_b200_bdstate* in, const uint64_t* nodes, size_t n, int backward,
                  gbwt_b200_bdstate* out) {
    for (size_t i = 0; i < n; i++) {
        gbwt_b200_bdstate s = in[i];
        if (backward) gbwt_extend_backward(h->view, s, nodes[i], out[i]);
        else gbwt_extend_forward(h->view, s, nodes[i], out[i]);
    }
}
void hs_bd_search(const HostSim* h, const uint64_t* nodes, const uint64_t* offsets, const uint64_t* first,
                  const uint64_t* start, const uint64_t* end, size_t n, gbwt_b200_bdstate* out) {
    for (size_t i = 0; i < n; i++)
        query_bd_search(h->view, nodes + offsets[i], offsets[i + 1] - offsets[i], first[i], start[i], end[i], out[i]);
}
void hs_follow(const HostSim* h, const gbwt_b200_bdstate* states, size_t n, int backward, const uint64_t* offsets,
               gbwt_b200_bdstate* out, uint64_t* counts) {
    for (size_t i = 0; i < n; i++) {
        gbwt_b200_bdstate* dst = out ? out + offsets[i] : nullptr;
        uint64_t cap = out ? offsets[i + 1] - offsets[i] : 0;
        counts[i] = gbwt_follow_all(h->view, states[i], backward != 0, dst, cap);
    }
}
void hs_start(const HostSim* h, const uint64_t* ids, size_t n, gbwt_b200_pos* out) {
    for (size_t i = 0; i < n; i++) gbwt_start(h->view, ids[i], out[i]);
}
void hs_forward(const HostSim* h, const gbwt_b200_pos* in, size_t n, gbwt_b200_pos* out) {
    for (size_t i = 0; i < n; i++) { gbwt_b200_pos p = in[i]; gbwt_forward(h->view, p, out[i]); }
}
int hs_backward(const HostSim* h, const gbwt_b200_pos* in, size_t n, gbwt_b200_pos* out) {
    if (!h->view.bidirectional) return GBWT_B200_E_NOT_BIDIRECTIONAL;
    for (size_t i = 0; i < n; i++) { gbwt_b200_pos p = in[i]; gbwt_backward(h->view, p, out[i]); }
    return 0;
}
void hs_sequence_lengths(const HostSim* h, const uint64_t* ids, size_t m, uint64_t* lengths) {
    for (size_t i = 0; i < m; i++) lengths[i] = walk_sequence(h->view, ids[i], nullptr, 0);
}
void hs_extract(const HostSim* h, const uint64_t* ids, size_t m, const uint64_t* out_offsets, uint64_t* nodes,
                uint64_t* lengths) {
    for (size_t i = 0; i < m; i++)
        lengths[i] = walk_sequence(h->view, ids[i], nodes + out_offsets[i], out_offsets[i + 1] - out_offsets[i]);
}

}  // extern "C"
