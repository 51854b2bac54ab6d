// layout_writer.h -- the way back: turns the device layout of layout.h into a Simple-SDS GBWT image that the
// reference (and this library) can load again (GBWT::serialize, gbwt-rs src/gbwt.rs:388-400;
// BWT::serialize, src/bwt.rs:170-174; the record encoding of BWTBuilder::append, src/bwt.rs:241-253, with
// RLE::write, src/support.rs:1225-1248, and ByteCode::write, src/support.rs:1063-1070). Host code, run on demand.
//
// Records are re-encoded from the layout, not kept from the input: edges from the descriptor / edge lists, runs from
// the body with adjacent runs of one value merged (the layout may have split them; maximal runs are what the
// reference's builders write). Tags are reduced to the `source` tag, document-array samples and metadata are not
// carried (the flag is cleared): they are outside the accelerated path and never reach the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "layout.h"

namespace gbwt_b200 {

struct LayoutArrays {
    const RecordDesc* desc = nullptr;
    uint64_t records = 0;
    const uint64_t* bodies = nullptr;  // 16-byte units as pairs of words
    const Edge* edges = nullptr;
};

struct GBWTHeaderFields {
    uint64_t sequences = 0, size = 0, offset = 0, alphabet_size = 0, flags = 0;
};

// The concatenated records in the reference encoding and the start offset of each.
int encode_bwt(const LayoutArrays& in, std::vector<uint8_t>& data, std::vector<uint64_t>& record_starts, std::string& err);

// A complete GBWT file image: header, tags, BWT (Elias-Fano index + data), no DA samples, no metadata.
int write_gbwt_image(const GBWTHeaderFields& header, const LayoutArrays& in, std::vector<uint8_t>& image, std::string& err);

}  // namespace gbwt_b200
