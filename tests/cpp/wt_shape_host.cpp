// wt_shape_host.cpp — CPU harness for sdsl-lite_b200/csrc/wt_shape.h (Huffman shape with the reference's tie-breaking
// and BFS numbering + the host fill of the bit planes) and sdsl_pack.h's byte_tree writer: produces m_bv and the
// serialised tree of wt_huff<>(text) so that tests/test_wt_shape_cpu.py can compare both with the reference's blob.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sdsl-lite_b200/csrc/sdsl_pack.h"
#include "../../sdsl-lite_b200/csrc/wt_shape.h"

using namespace sdslgpu;

// returns the number of bits of m_bv; bv_out receives ceil(bits/64) words (cap_words), tree_out the byte_tree bytes
extern "C" uint64_t wt_shape_host(uint8_t const * text, uint64_t n, uint64_t * bv_out, uint64_t cap_words, uint8_t * tree_out, uint64_t cap_tree,
                                  uint64_t * tree_bytes, uint64_t * sigma_out)
{
    uint64_t C[256] = {0};
    for (uint64_t k = 0; k < n; ++k)
        ++C[text[k]];
    WtTree tree;
    uint64_t sigma = 0;
    uint64_t bits = build_huff_tree(C, tree, sigma);
    std::vector<uint64_t> bv(((bits + 63) >> 6) + 1, 0);
    fill_bit_planes(text, n, tree, bv);
    // inner nodes: bv_pos_rank = rank1(m_bv, bv_pos) (wt_helper.hpp:319-327; the library asks the device for these)
    for (uint32_t v = 0; v < tree.nnodes; ++v)
        if (tree.child[v][0] != kWtUndef)
        {
            uint64_t p = tree.bv_pos[v], r = 0;
            for (uint64_t w = 0; w < (p >> 6); ++w)
                r += (uint64_t)__builtin_popcountll(bv[w]);
            if (p & 63)
                r += (uint64_t)__builtin_popcountll(bv[p >> 6] & ((1ull << (p & 63)) - 1));
            tree.bv_pos_rank[v] = r;
        }
    std::vector<uint8_t> blob;
    pack::Sink sink{blob};
    pack::write_byte_tree(tree, sink);
    *tree_bytes = blob.size();
    *sigma_out = sigma;
    if (blob.size() <= cap_tree)
        std::memcpy(tree_out, blob.data(), blob.size());
    uint64_t nw = (bits + 63) >> 6;
    if (nw <= cap_words)
        std::memcpy(bv_out, bv.data(), nw * 8);
    return bits;
}

// the shape alone from a histogram: number of bits of m_bv, or ~0 when a code would be deeper than 56 levels
// (the reference throws there, wt_helper.hpp:304-307); max_depth_out = longest code length
extern "C" uint64_t wt_shape_from_histogram(uint64_t const * C256, uint32_t * max_depth_out)
{
    uint64_t C[256];
    std::memcpy(C, C256, sizeof(C));
    WtTree tree;
    uint64_t sigma = 0;
    uint64_t bits = build_huff_tree(C, tree, sigma);
    uint32_t d = 0;
    for (int c = 0; c < 256; ++c)
        if (tree.c_to_leaf[c] != kWtUndef && (uint32_t)(tree.path[c] >> 56) > d)
            d = (uint32_t)(tree.path[c] >> 56);
    *max_depth_out = d;
    return bits;
}
