// sds_loader.cpp -- see sds_loader.h.
#include "sds_loader.h"

#include <dlfcn.h>

#include <cstring>

#include "../../include/gbwt_b200.h"

namespace gbwt_b200 {
namespace {

constexpr uint32_t TAG_GBWT = 0x6B376B37u, TAG_GBZ = 0x205A4247u, TAG_GRAPH = 0x6B3764AFu;

// Sequential reader over 64-bit little-endian "elements"; every structure in the format is a whole
// number of elements and most can be skipped by their declared size.
class Cursor {
public:
    Cursor(const uint8_t* p, size_t n) : p_(p), n_(n) {}
    bool ok() const { return ok_; }
    size_t at() const { return at_; }
    const uint8_t* here() const { return p_ + at_; }
    uint64_t word() {
        if (!ok_ || n_ - at_ < 8) { ok_ = false; return 0; }
        uint64_t v;
        std::memcpy(&v, p_ + at_, 8);
        at_ += 8;
        return v;
    }
    void skip_words(uint64_t k) {
        if (!ok_ || k > (n_ - at_) / 8) { ok_ = false; return; }
        at_ += static_cast<size_t>(k) * 8;
    }
    void skip_option() { skip_words(word()); }  // Option<T>: size in elements (0 = None), then T
    void skip_raw_vector() { word(); skip_words(word()); }
    void skip_int_vector() { word(); word(); skip_raw_vector(); }
    void skip_bit_vector() { word(); skip_raw_vector(); skip_option(); skip_option(); skip_option(); }
    void skip_sparse_vector() { word(); skip_bit_vector(); skip_int_vector(); }
    void skip_byte_vector() { uint64_t n = word(); skip_words((n + 7) / 8); }
    // Tags are a StringArray: SparseVector of starts, Vec<u8> alphabet, IntVector of packed characters.
    void skip_tags() { skip_sparse_vector(); skip_byte_vector(); skip_int_vector(); }

private:
    const uint8_t* p_;
    size_t n_, at_ = 0;
    bool ok_ = true;
};

inline uint64_t load_word(const uint8_t* base, uint64_t i) {
    uint64_t v;
    std::memcpy(&v, base + 8 * i, 8);
    return v;
}

// Elias-Fano SparseVector -> all values. value_j = ((pos_j - j) << width) | low[j], pos_j = position of the
// j-th set bit of `high`.
bool read_sparse_values(Cursor& c, uint64_t& universe, std::vector<uint64_t>& values) {
    universe = c.word();
    uint64_t ones = c.word();
    uint64_t high_bits = c.word(), high_words = c.word();
    const uint8_t* high = c.here();
    c.skip_words(high_words);
    c.skip_option(); c.skip_option(); c.skip_option();
    uint64_t low_len = c.word(), width = c.word();
    uint64_t low_bits = c.word(), low_words = c.word();
    const uint8_t* low = c.here();
    c.skip_words(low_words);
    if (!c.ok()) return false;
    if (high_words != (high_bits + 63) / 64 || low_words != (low_bits + 63) / 64) return false;
    if (low_len != ones || width > 64 || low_bits != ones * width) return false;
    values.clear();
    values.reserve(ones);
    const uint64_t mask = width >= 64 ? ~0ULL : ((1ULL << width) - 1);
    uint64_t j = 0;
    for (uint64_t w = 0; w < high_words && j < ones; w++) {
        uint64_t bits = load_word(high, w);
        while (bits != 0 && j < ones) {
            uint64_t pos = w * 64 + static_cast<uint64_t>(__builtin_ctzll(bits));
            bits &= bits - 1;
            uint64_t lo = 0;
            if (width > 0) {
                uint64_t bit = j * width, lw = bit / 64, sh = bit % 64;
                lo = load_word(low, lw) >> sh;
                if (sh + width > 64) lo |= load_word(low, lw + 1) << (64 - sh);
                lo &= mask;
            }
            uint64_t hi = pos - j;
            values.push_back(width >= 64 ? lo : ((hi << width) | lo));
            j++;
        }
    }
    return j == ones;
}

// StringArray::load (src/support.rs:624-655) for small arrays (tags): start offsets, alphabet, packed characters.
bool read_string_array(Cursor& c, std::vector<std::string>& strings) {
    uint64_t universe = 0;
    std::vector<uint64_t> starts;
    if (!read_sparse_values(c, universe, starts)) return false;
    const uint64_t alphabet_len = c.word();
    const uint8_t* alphabet = c.here();
    c.skip_words((alphabet_len + 7) / 8);
    const uint64_t total = c.word(), width = c.word(), bits = c.word(), words = c.word();
    const uint8_t* packed = c.here();
    c.skip_words(words);
    if (!c.ok() || width == 0 || width > 64 || words != (bits + 63) / 64 || bits != total * width || total > (uint64_t(1) << 32)) return false;
    std::string all(total, '\0');
    const uint64_t mask = width >= 64 ? ~0ull : ((1ull << width) - 1);
    for (uint64_t i = 0; i < total; i++) {
        const uint64_t bit = i * width, w = bit / 64, sh = bit % 64;
        uint64_t v = load_word(packed, w) >> sh;
        if (sh + width > 64) v |= load_word(packed, w + 1) << (64 - sh);
        v &= mask;
        if (v >= alphabet_len) return false;
        all[i] = static_cast<char>(alphabet[v]);
    }
    strings.clear();
    for (size_t i = 0; i < starts.size(); i++) {
        const uint64_t lo = starts[i], hi = i + 1 < starts.size() ? starts[i + 1] : total;
        if (lo > hi || hi > total || (i == 0 && lo != 0)) return false;
        strings.push_back(all.substr(lo, hi - lo));
    }
    return true;
}

// Tags::load (src/support.rs:992-1007): [key, value, ...], keys lower-cased, no duplicates.
bool read_tags(Cursor& c, std::vector<std::pair<std::string, std::string>>& tags, std::string& err) {
    std::vector<std::string> strings;
    if (!read_string_array(c, strings)) { err = "Tags: invalid data"; return false; }
    if (strings.size() % 2 != 0) { err = "Tags: Key without a value"; return false; }
    tags.clear();
    for (size_t i = 0; i < strings.size(); i += 2) {
        std::string key = strings[i];
        for (char& ch : key) if (ch >= 'A' && ch <= 'Z') ch = static_cast<char>(ch - 'A' + 'a');
        for (const auto& kv : tags) if (kv.first == key) { err = "Tags: Duplicate keys"; return false; }
        tags.emplace_back(key, strings[i + 1]);
    }
    return true;
}

int parse_gbwt(Cursor& c, ParsedGBWT& out, std::string& err) {
    uint64_t tv = c.word();
    out.sequences = c.word(); out.size = c.word(); out.offset = c.word();
    out.alphabet_size = c.word(); out.flags = c.word();
    if (!c.ok()) { err = "GBWTHeader: unexpected end of data"; return GBWT_B200_E_INVALID_DATA; }
    uint32_t tag = static_cast<uint32_t>(tv), version = static_cast<uint32_t>(tv >> 32);
    // Header::validate, src/headers.rs:102-115; GBWTPayload, src/headers.rs:211-231
    if (tag != TAG_GBWT) { err = "GBWTHeader: Invalid tag"; return GBWT_B200_E_INVALID_DATA; }
    if (version != 5) { err = "GBWTHeader: Invalid version (expected 5 to 5)"; return GBWT_B200_E_INVALID_DATA; }
    if (out.flags & ~(GBWT_FLAG_BIDIRECTIONAL | GBWT_FLAG_METADATA | GBWT_FLAG_SIMPLE_SDS)) {
        err = "GBWTHeader: Invalid flags"; return GBWT_B200_E_INVALID_DATA;
    }
    if (!(out.flags & GBWT_FLAG_SIMPLE_SDS)) { err = "GBWTHeader: SDSL format is not supported"; return GBWT_B200_E_INVALID_DATA; }
    if (out.alphabet_size < out.offset) { err = "GBWTHeader: alphabet size below offset"; return GBWT_B200_E_INVALID_DATA; }
    if (!read_tags(c, out.tags, err)) return GBWT_B200_E_INVALID_DATA;
    // BWT::load, src/bwt.rs:176-185
    uint64_t universe = 0;
    if (!read_sparse_values(c, universe, out.record_starts)) { err = "BWT: invalid index"; return GBWT_B200_E_INVALID_DATA; }
    out.bwt_len = c.word();
    out.bwt = c.here();
    c.skip_words((out.bwt_len + 7) / 8);
    if (!c.ok()) { err = "BWT: invalid data"; return GBWT_B200_E_INVALID_DATA; }
    if (universe != out.bwt_len) { err = "BWT: Index / data length mismatch"; return GBWT_B200_E_INVALID_DATA; }
    for (size_t i = 0; i < out.record_starts.size(); i++) {
        uint64_t s = out.record_starts[i];
        if (s >= out.bwt_len || (i > 0 && s <= out.record_starts[i - 1])) { err = "BWT: invalid index"; return GBWT_B200_E_INVALID_DATA; }
    }
    // DA samples pass through as Vec<u64> (src/gbwt.rs:417); Option<Metadata> must agree with the flag (:420-423).
    const uint8_t* da_at = c.here();
    c.skip_words(c.word());
    const uint8_t* meta_at = c.here();
    uint64_t meta_words = c.word();
    c.skip_words(meta_words);
    if (!c.ok()) { err = "GBWT: unexpected end of data"; return GBWT_B200_E_INVALID_DATA; }
    out.da_samples.assign(da_at, meta_at);
    out.metadata.assign(meta_at, c.here());
    if (((out.flags & GBWT_FLAG_METADATA) != 0) != (meta_words != 0)) {
        err = "GBWT: Invalid metadata flag in the header"; return GBWT_B200_E_INVALID_DATA;
    }
    return GBWT_B200_OK;
}

// The reference decompresses node labels with the zstd crate (= libzstd). The image has the shared library
// but not its headers, so the two stable entry points are resolved at run time; a GBZ file with compressed
// labels fails to load with a clear message when the library is missing.
struct ZstdApi {
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    unsigned long long (*frame_content_size)(const void*, size_t) = nullptr;
    ZstdApi() {
        void* lib = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (lib == nullptr) lib = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
        if (lib == nullptr) return;
        decompress = reinterpret_cast<decltype(decompress)>(dlsym(lib, "ZSTD_decompress"));
        is_error = reinterpret_cast<decltype(is_error)>(dlsym(lib, "ZSTD_isError"));
        frame_content_size = reinterpret_cast<decltype(frame_content_size)>(dlsym(lib, "ZSTD_getFrameContentSize"));
    }
};

const ZstdApi& zstd_api() {
    static const ZstdApi api;  // resolved once (thread-safe initialisation)
    return api;
}

// The decompressed size the frame itself declares, when it does (frames written by one-shot compression, which is
// what the reference uses, always do): lets the loader reject a damaged `total` before it allocates that much.
// Returns false when the frame declares a different size (or is not a frame at all).
bool zstd_size_plausible(const uint8_t* src, size_t src_len, uint64_t total) {
    const ZstdApi& api = zstd_api();
    if (api.frame_content_size == nullptr) return true;
    const unsigned long long declared = api.frame_content_size(src, src_len);
    if (declared == 0ULL - 2) return total == 0 && src_len == 0;  // ZSTD_CONTENTSIZE_ERROR
    if (declared == 0ULL - 1) return true;                        // ZSTD_CONTENTSIZE_UNKNOWN
    return declared == total;
}

bool zstd_decompress(uint8_t* dst, size_t dst_len, const uint8_t* src, size_t src_len, size_t& produced, std::string& err) {
    const ZstdApi& api = zstd_api();
    if (api.decompress == nullptr || api.is_error == nullptr) { err = "StringArray: libzstd is not available"; return false; }
    produced = api.decompress(dst, dst_len, src, src_len);
    if (api.is_error(produced)) { err = "StringArray: Zstandard decompression failed"; return false; }
    return true;
}

// Graph::load, src/graph.rs:295-330, as far as the node labels: header, then the sequences as a compressed
// (version 4, StringArray::decompress) or packed (version 3, StringArray::load) string array. Segment names and
// the node-to-segment mapping follow in the image and are not needed on the path.
int parse_graph(Cursor& c, ParsedGBWT& out, std::string& err) {
    const uint64_t tv = c.word();
    const uint64_t nodes = c.word();
    const uint64_t flags = c.word();
    if (!c.ok()) { err = "GraphHeader: unexpected end of data"; return GBWT_B200_E_INVALID_DATA; }
    const uint32_t tag = static_cast<uint32_t>(tv), version = static_cast<uint32_t>(tv >> 32);
    if (tag != TAG_GRAPH) { err = "GraphHeader: Invalid tag"; return GBWT_B200_E_INVALID_DATA; }
    if (version < 3 || version > 4) { err = "GraphHeader: Invalid version (expected 3 to 4)"; return GBWT_B200_E_INVALID_DATA; }
    if (flags & ~3ull) { err = "GraphHeader: Invalid flags"; return GBWT_B200_E_INVALID_DATA; }
    if (!(flags & 2ull)) { err = "GraphHeader: SDSL format is not supported"; return GBWT_B200_E_INVALID_DATA; }
    uint64_t universe = 0;
    if (!read_sparse_values(c, universe, out.label_starts)) { err = "StringArray: invalid index"; return GBWT_B200_E_INVALID_DATA; }
    uint64_t total = 0;
    if (version >= 4) {
        total = c.word();
        const uint64_t compressed_len = c.word();
        const uint8_t* compressed = c.here();
        c.skip_words((compressed_len + 7) / 8);
        if (!c.ok() || total > (uint64_t(1) << 48)) { err = "StringArray: invalid data"; return GBWT_B200_E_INVALID_DATA; }
        if (!zstd_size_plausible(compressed, compressed_len, total)) {
            err = "StringArray: Decompressed string length does not match the expected length";
            return GBWT_B200_E_INVALID_DATA;
        }
        out.label_bytes.resize(total);
        size_t produced = 0;
        uint8_t scratch = 0;
        if (!zstd_decompress(total ? out.label_bytes.data() : &scratch, total, compressed, compressed_len, produced, err))
            return GBWT_B200_E_INVALID_DATA;
        if (produced != total) {
            err = "StringArray: Decompressed string length does not match the expected length";
            return GBWT_B200_E_INVALID_DATA;
        }
    } else {
        const uint64_t alphabet_len = c.word();
        const uint8_t* alphabet = c.here();
        c.skip_words((alphabet_len + 7) / 8);
        total = c.word();
        const uint64_t width = c.word(), bits = c.word(), words = c.word();
        const uint8_t* packed = c.here();
        c.skip_words(words);
        if (!c.ok() || width == 0 || width > 64 || words != (bits + 63) / 64 || total > bits || bits != total * width) {
            err = "StringArray: invalid strings"; return GBWT_B200_E_INVALID_DATA;
        }
        out.label_bytes.resize(total);
        const uint64_t mask = width >= 64 ? ~0ull : ((1ull << width) - 1);
        for (uint64_t i = 0; i < total; i++) {
            const uint64_t bit = i * width, w = bit / 64, sh = bit % 64;
            uint64_t v = load_word(packed, w) >> sh;
            if (sh + width > 64) v |= load_word(packed, w + 1) << (64 - sh);
            v &= mask;
            if (v >= alphabet_len) { err = "StringArray: invalid strings"; return GBWT_B200_E_INVALID_DATA; }
            out.label_bytes[i] = alphabet[v];
        }
    }
    if (!out.label_starts.empty() && out.label_starts[0] != 0) {
        err = "StringArray: First string does not start at offset 0"; return GBWT_B200_E_INVALID_DATA;
    }
    for (size_t i = 0; i < out.label_starts.size(); i++) {
        if (out.label_starts[i] > total || (i > 0 && out.label_starts[i] < out.label_starts[i - 1])) {
            err = "StringArray: invalid index"; return GBWT_B200_E_INVALID_DATA;
        }
    }
    // GBZ::load, src/gbz.rs:690-694
    if (out.label_starts.size() != (out.alphabet_size - (out.offset + 1)) / 2) {
        err = "GBZ: Mismatch between GBWT alphabet size and Graph sequence count"; return GBWT_B200_E_INVALID_DATA;
    }
    // Segment names and the node-to-segment mapping are not needed on the path; they are stepped over with the
    // reference's consistency checks (src/graph.rs:309-328) so that a damaged file is rejected like it is there.
    c.word();  // universe of the segment-name index
    const uint64_t segments = c.word();
    c.skip_raw_vector(); c.skip_option(); c.skip_option(); c.skip_option();
    c.skip_int_vector();
    c.skip_byte_vector(); c.skip_int_vector();
    const uint64_t mapping_len = c.word(), mapping_ones = c.word();
    c.skip_raw_vector(); c.skip_option(); c.skip_option(); c.skip_option();
    c.skip_int_vector();
    if (!c.ok()) { err = "Graph: unexpected end of data"; return GBWT_B200_E_INVALID_DATA; }
    const bool translation = (flags & 1ull) != 0;
    if (translation == (segments == 0)) {
        err = "Graph: Translation flag does not match the presence of segment names"; return GBWT_B200_E_INVALID_DATA;
    }
    if (translation) {
        if (mapping_len <= nodes) { err = "Graph: Node-to-segment mapping does not match the number of nodes"; return GBWT_B200_E_INVALID_DATA; }
        if (mapping_len != out.label_starts.size() + 1) { err = "Graph: Node-to-segment mapping does not match the number of sequences"; return GBWT_B200_E_INVALID_DATA; }
        if (mapping_ones != segments) { err = "Graph: Node-to-segment mapping does not match the number of segments"; return GBWT_B200_E_INVALID_DATA; }
    }
    out.label_starts.push_back(total);
    out.has_graph = true;
    return GBWT_B200_OK;
}

}  // namespace

int parse_gbwt_image(const uint8_t* bytes, size_t len, ParsedGBWT& out, std::string& err) {
    if (bytes == nullptr || len < 16) { err = "image too short"; return GBWT_B200_E_INVALID_DATA; }
    Cursor c(bytes, len);
    uint32_t tag;
    std::memcpy(&tag, bytes, 4);
    if (tag == TAG_GBZ) {
        // GBZ::load, src/gbz.rs:678-696: header {tag|version, flags}, tags, GBWT, graph.
        uint64_t tv = c.word(), flags = c.word();
        uint32_t version = static_cast<uint32_t>(tv >> 32);
        if (version < 1 || version > 2) { err = "GBZHeader: Invalid version (expected 1 to 2)"; return GBWT_B200_E_INVALID_DATA; }
        if (flags != 0) { err = "GBZHeader: Invalid flags"; return GBWT_B200_E_INVALID_DATA; }
        if (!read_tags(c, out.gbz_tags, err)) return GBWT_B200_E_INVALID_DATA;
        int rc = parse_gbwt(c, out, err);
        if (rc != GBWT_B200_OK) return rc;
        if (!(out.flags & GBWT_FLAG_BIDIRECTIONAL)) { err = "GBZ: The GBWT index is not bidirectional"; return GBWT_B200_E_INVALID_DATA; }
        const uint8_t* graph_at = c.here();
        rc = parse_graph(c, out, err);
        if (rc != GBWT_B200_OK) return rc;
        out.graph_section.assign(graph_at, c.here());
        out.from_gbz = true;
        return GBWT_B200_OK;
    }
    return parse_gbwt(c, out, err);
}

}  // namespace gbwt_b200
