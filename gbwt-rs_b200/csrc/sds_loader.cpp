// sds_loader.cpp -- see sds_loader.h.
#include "sds_loader.h"

#include <cstring>

#include "../../include/gbwt_b200.h"

namespace gbwt_b200 {
namespace {

constexpr uint32_t TAG_GBWT = 0x6B376B37u, TAG_GBZ = 0x205A4247u;

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
    c.skip_tags();
    if (!c.ok()) { err = "Tags: invalid data"; return GBWT_B200_E_INVALID_DATA; }
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
    c.skip_words(c.word());
    uint64_t meta_words = c.word();
    c.skip_words(meta_words);
    if (!c.ok()) { err = "GBWT: unexpected end of data"; return GBWT_B200_E_INVALID_DATA; }
    if (((out.flags & GBWT_FLAG_METADATA) != 0) != (meta_words != 0)) {
        err = "GBWT: Invalid metadata flag in the header"; return GBWT_B200_E_INVALID_DATA;
    }
    return GBWT_B200_OK;
}

}  // namespace

int parse_gbwt_image(const uint8_t* bytes, size_t len, ParsedGBWT& out, std::string& err) {
    if (bytes == nullptr || len < 16) { err = "image too short"; return GBWT_B200_E_INVALID_DATA; }
    Cursor c(bytes, len);
    uint32_t tag;
    std::memcpy(&tag, bytes, 4);
    if (tag == TAG_GBZ) {
        // GBZ::load, src/gbz.rs:678-690: header {tag|version, flags}, tags, GBWT, graph (ignored).
        uint64_t tv = c.word(), flags = c.word();
        uint32_t version = static_cast<uint32_t>(tv >> 32);
        if (version < 1 || version > 2) { err = "GBZHeader: Invalid version (expected 1 to 2)"; return GBWT_B200_E_INVALID_DATA; }
        if (flags != 0) { err = "GBZHeader: Invalid flags"; return GBWT_B200_E_INVALID_DATA; }
        c.skip_tags();
        int rc = parse_gbwt(c, out, err);
        if (rc != GBWT_B200_OK) return rc;
        if (!(out.flags & GBWT_FLAG_BIDIRECTIONAL)) { err = "GBZ: The GBWT index is not bidirectional"; return GBWT_B200_E_INVALID_DATA; }
        return GBWT_B200_OK;
    }
    return parse_gbwt(c, out, err);
}

}  // namespace gbwt_b200
