// sds_loader.h -- host-side reader for Simple-SDS GBWT / GBZ images (boundary input, not accelerated).
// Format: SURVEY.md App. A; field order from gbwt-rs src/gbwt.rs:402-438, src/bwt.rs:176-185,
// src/gbz.rs:678-696, src/graph.rs:295-330, src/support.rs:545-571, 619-643, src/headers.rs:57-62, 190-280.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace gbwt_b200 {

struct ParsedGBWT {
    uint64_t sequences = 0, size = 0, offset = 0, alphabet_size = 0, flags = 0;
    std::vector<uint64_t> record_starts;  // start of every record in `bwt` (what the Elias-Fano index selects)
    const uint8_t* bwt = nullptr;         // points into the caller's image
    uint64_t bwt_len = 0;
    // Node labels of a GBZ image (Graph::sequences, src/graph.rs:84-89): label i is
    // label_bytes[label_starts[i] .. label_starts[i + 1]); sequence id = node id - first node id.
    bool has_graph = false;
    std::vector<uint64_t> label_starts;
    std::vector<uint8_t> label_bytes;
    // What is not on the accelerated path is carried through as it is, so that an index can be written back whole
    // (GBWT::serialize, src/gbwt.rs:388-400; GBZ::serialize, src/gbz.rs:662-671): the tags as key / value strings, and
    // the serialized elements of the document-array samples (Vec<u64>, length word included), of Option<Metadata>
    // (size word included) and, for a GBZ image, of the whole Graph section (header to end of image).
    std::vector<std::pair<std::string, std::string>> tags, gbz_tags;
    std::vector<uint8_t> da_samples, metadata, graph_section;
    bool from_gbz = false;
};

constexpr uint64_t GBWT_FLAG_BIDIRECTIONAL = 1, GBWT_FLAG_METADATA = 2, GBWT_FLAG_SIMPLE_SDS = 4;

// Returns a GBWT_B200_* status; on failure `err` holds the reference's error text where it has one.
int parse_gbwt_image(const uint8_t* bytes, size_t len, ParsedGBWT& out, std::string& err);

}  // namespace gbwt_b200
