// Exercises gbwt-rs_b200/host/gbwt.hpp like the reference's GBWT doc-test (gbwt-rs src/gbwt.rs:46-92) and
// its extract / bd_extend tests (src/gbwt/tests.rs:164-189, 393-462). Exit code 0 = all assertions hold,
// 77 = no CUDA device (the mirror refuses to run; there is no CPU fallback), anything else = failure.
#include <cstdio>
#include <string>
#include <vector>

#include "../../gbwt-rs_b200/host/gbwt.hpp"

using namespace gbwt_b200_host;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static std::size_t encode_node(std::size_t id, bool reverse) { return 2 * id + (reverse ? 1 : 0); }  // support.rs:155

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    try {
        GBWT index = GBWT::load(dir + "/example.gbwt");
        // Statistics (gbwt.rs:54-58)
        CHECK(index.len() == 68 && index.sequences() == 12 && index.alphabet_size() == 52 && index.is_bidirectional());
        CHECK(index.alphabet_offset() == 21 && index.first_node() == 22 && index.effective_size() == 31);
        CHECK(!index.has_node(21) && index.has_node(22) && !index.has_node(52));
        // second-to-last node of path 2 forward (gbwt.rs:60-68)
        auto pos = index.start(4);
        std::optional<Pos> last;
        while (pos) { last = pos; pos = index.forward(*pos); }
        CHECK(last.has_value());
        auto prev = index.backward(*last);
        CHECK(prev && prev->node == encode_node(15, false));
        // unidirectional search (gbwt.rs:70-75)
        auto state = index.find(encode_node(12, false));
        CHECK(state.has_value());
        state = index.extend(*state, encode_node(14, false));
        CHECK(state.has_value());
        state = index.extend(*state, encode_node(15, false));
        CHECK(state && state->node == encode_node(15, false) && state->len() == 2);
        CHECK(!index.extend(*state, encode_node(11, false)) && !index.find(0) && !index.find(36));
        // bidirectional search (gbwt.rs:77-83)
        auto bd = index.bd_find(encode_node(14, false));
        CHECK(bd.has_value());
        bd = index.extend_backward(*bd, encode_node(12, false));
        CHECK(bd.has_value());
        bd = index.extend_forward(*bd, encode_node(15, false));
        CHECK(bd && bd->forward.node == encode_node(15, false) && bd->reverse.node == encode_node(12, true) && bd->len() == 2);
        CHECK(bd->from() == std::make_pair(std::size_t(12), false) && bd->to() == std::make_pair(std::size_t(15), false));
        // all extensions of a state (StateIter doc-test, gbz.rs:1189-1208)
        auto st14 = index.bd_find(encode_node(14, false));
        auto successors = index.follow_forward(*st14);
        CHECK(successors && successors->size() == 2);
        std::size_t preds = 0;
        for (const auto& s : *successors) {
            auto p = index.follow_backward(s);
            CHECK(p && p->size() == 1);
            preds += p->size();
            if (s.forward.node == encode_node(15, false)) CHECK((*p)[0].from() == std::make_pair(std::size_t(12), false) && (*p)[0].len() == 2);
            else CHECK((*p)[0].from() == std::make_pair(std::size_t(13), false) && (*p)[0].to() == std::make_pair(std::size_t(16), false) && (*p)[0].len() == 1);
        }
        CHECK(preds == 2);
        // sequence iterator (gbwt.rs:546-548)
        auto path = index.sequence(7);
        CHECK(path && *path == (std::vector<std::size_t>{35, 33, 29, 27, 23}));
        CHECK(!index.sequence(12).has_value() && !index.start(12).has_value());
        // batch: every length-4 window of every sequence (BASELINE.json configs[0]): 20 queries, 32 occurrences
        std::vector<uint64_t> patterns;
        for (std::size_t i = 0; i < index.sequences(); i++) {
            auto p = *index.sequence(i);
            for (std::size_t j = 0; j + 4 <= p.size(); j++) patterns.insert(patterns.end(), p.begin() + j, p.begin() + j + 4);
        }
        auto out = index.find_extend_batch(patterns, 4);
        std::size_t occ = 0;
        for (auto& s : out) occ += s.end - s.start;
        CHECK(out.size() == 20 && occ == 32);
        // the same batch with 32-bit nodes and as ragged patterns; a ragged batch with an empty and a one-node pattern
        std::vector<uint32_t> narrow(patterns.begin(), patterns.end());
        auto out32 = index.find_extend_batch(narrow, 4);
        std::vector<uint64_t> offsets;
        for (std::size_t q = 0; q <= out.size(); q++) offsets.push_back(4 * q);
        auto ragged = index.find_extend_ragged(patterns, offsets);
        CHECK(out32.size() == out.size() && ragged.size() == out.size());
        for (std::size_t q = 0; q < out.size(); q++) {
            CHECK(out32[q].node == out[q].node && out32[q].start == out[q].start && out32[q].end == out[q].end);
            CHECK(ragged[q].node == out[q].node && ragged[q].start == out[q].start && ragged[q].end == out[q].end);
        }
        auto odd = index.find_extend_ragged({encode_node(12, false), encode_node(12, false), encode_node(14, false)}, {0, 0, 1, 3});
        CHECK(odd.size() == 3 && odd[0].end <= odd[0].start && odd[1].node == encode_node(12, false) && odd[2].node == encode_node(14, false));
        CHECK(index.device() == 0 && index.device_bytes() > 0 && !GBWT::version().empty());
        // serialize::test (src/gbwt/tests.rs:72-84): the image loads again and answers alike
        auto image = index.serialize();
        GBWT again = GBWT::from_bytes(image.data(), image.size());
        CHECK(again.len() == index.len() && again.sequences() == index.sequences() && again.sequence(7) == index.sequence(7));
        auto out_again = again.find_extend_batch(patterns, 4);
        for (std::size_t q = 0; q < out.size(); q++) CHECK(out_again[q].start == out[q].start && out_again[q].end == out[q].end);
        // node sequences and DNA of a GBZ (src/gbz/tests.rs:237-247; extract_sequence, src/bin/gbz-extract.rs:173-189)
        CHECK(!index.has_graph());
        GBWT gbz = GBWT::load(dir + "/example.gbz");
        CHECK(gbz.has_graph() && gbz.node_sequence(11) == std::string("G") && gbz.node_sequence(16) == std::string("C"));
        CHECK(gbz.sequence_len(13) == std::size_t(1) && !gbz.node_sequence(18) && !gbz.sequence_len(10) && !gbz.node_sequence(26));
        CHECK(gbz.path_dna(0, '$') == std::string("GATAA$") && gbz.path_dna(1, '$') == std::string("TTATC$") && !gbz.path_dna(12));
        auto dna = gbz.extract_dna({0, 1, 2, 3}, '$');
        CHECK(dna.second == "GATAA$TTATC$GATA$TATC$" && dna.first == (std::vector<uint64_t>{0, 6, 12, 17, 22}));
        bool refused = false;
        try { index.path_dna(0); } catch (const std::runtime_error&) { refused = true; }
        CHECK(refused);
        // GBZ::serialize (src/gbz.rs:662-671): round trip with the node labels; a GBWT alone has no GBZ form
        auto gbz_image = gbz.serialize_gbz();
        GBWT gbz_again = GBWT::from_bytes(gbz_image.data(), gbz_image.size());
        CHECK(gbz_again.has_graph() && gbz_again.path_dna(0, '$') == std::string("GATAA$") && gbz_again.node_sequence(16) == std::string("C"));
        refused = false;
        try { index.serialize_gbz(); } catch (const std::runtime_error&) { refused = true; }
        CHECK(refused);
    } catch (const std::runtime_error& e) {
        if (std::string(e.what()).find("error 7") != std::string::npos) { std::fprintf(stderr, "%s\n", e.what()); return 77; }
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 3;
    }
    std::puts("host mirror ok");
    return 0;
}
