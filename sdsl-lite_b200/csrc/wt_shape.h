// wt_shape.h — host side of the wt_huff<> constructor (wt_pc.hpp:194-248, wt_huff.hpp:82-115, wt_helper.hpp:230-327):
// the Huffman shape with the reference's tie-breaking and BFS numbering, and the multi-threaded fill of the bit
// planes (the fallback of wt_build.cu's device fill).  Plain C++ (no CUDA) so that tests/test_wt_shape_cpu.py can
// check tree and bits against the reference's serialised tree on a box without a GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include "wt_tree.h"

namespace sdslgpu
{

// ------------------------------------------------------------------------------------------------
// host: Huffman shape + BFS layout (tiny), then the bit planes in parallel over text chunks
// ------------------------------------------------------------------------------------------------
static constexpr uint64_t kWtTooDeep = ~0ull; // build_huff_tree: a code longer than 56 bits (the reference throws)

// Fills tree (bv_pos, children, parents, c_to_leaf, path) and returns the number of bits of m_bv, or kWtTooDeep when
// some symbol's code would be longer than the 56 bits a path word holds — the reference refuses such an input with
// std::logic_error("Code depth greater than 56!!!") (wt_helper.hpp:304-307).
//
// The shape must be THE tree of the reference, because m_bv and the serialised node table are compared byte for byte:
//  * merge order (wt_huff.hpp:82-115): always the two lightest subtrees, ties broken by node number; the lighter one
//    becomes child 0.  Leaves are numbered by symbol, merged nodes in order of creation, so the classic two-queue form
//    applies: leaves sorted by (weight, symbol) in one queue, merged nodes — created with non-decreasing weight — in a
//    second; the next lightest subtree is at the head of one of them, and on equal weight the leaf wins (smaller number).
//  * numbering (wt_helper.hpp:236-271): breadth first from the root, the two children of a node adjacent.
// The code word of a symbol falls out of the same breadth-first pass: a child's word is its parent's with the branch
// bit appended at position depth(parent) (bit j of a path = the branch taken at depth j, wt_pc.hpp:384-395).
inline uint64_t build_huff_tree(uint64_t const (&C)[256], WtTree & tree, uint64_t & sigma)
{
    struct Sub // a subtree during the merge
    {
        uint64_t weight;
        uint32_t kid[2]; // kUndef32 for a leaf
        uint32_t symbol;
    };
    constexpr uint32_t kUndef32 = 0xFFFFFFFFu;
    std::vector<Sub> sub;
    for (uint32_t c = 0; c < 256; ++c)
        if (C[c])
            sub.push_back(Sub{C[c], {kUndef32, kUndef32}, c});
    sigma = sub.size();
    std::vector<uint32_t> leaves(sub.size());
    for (uint32_t k = 0; k < leaves.size(); ++k)
        leaves[k] = k;
    std::stable_sort(leaves.begin(), leaves.end(), [&](uint32_t a, uint32_t b) { return sub[a].weight < sub[b].weight; });
    size_t lh = 0, mh = sigma; // heads of the leaf queue (`leaves`) and of the merged queue (sub[mh ...])
    auto lightest = [&]() -> uint32_t {
        bool const leaf_left = lh < leaves.size(), merged_left = mh < sub.size();
        if (leaf_left && (!merged_left || sub[leaves[lh]].weight <= sub[mh].weight))
            return leaves[lh++];
        return (uint32_t)mh++;
    };
    for (uint64_t merges = sigma > 1 ? sigma - 1 : 0; merges > 0; --merges)
    {
        uint32_t const a = lightest(), b = lightest();
        sub.push_back(Sub{sub[a].weight + sub[b].weight, {a, b}, 0});
    }
    std::memset(&tree, 0, sizeof(tree));
    for (int c = 0; c < 256; ++c)
    {
        tree.c_to_leaf[c] = kWtUndef;
        tree.path[c] = 0;
    }
    tree.nnodes = (uint32_t)sub.size();
    if (sub.empty())
        return 0;
    // breadth-first numbering; origin[v] = which subtree became node v, word / depth = its code so far
    std::vector<uint32_t> origin(sub.size());
    std::vector<uint64_t> word(sub.size(), 0);
    std::vector<uint32_t> depth(sub.size(), 0);
    uint64_t bits = 0;
    uint32_t placed = 1;
    bool too_deep = false;
    origin[0] = (uint32_t)sub.size() - 1; // the last merge is the root
    tree.parent[0] = kWtUndef;
    for (uint32_t v = 0; v < placed; ++v)
    {
        Sub const & me = sub[origin[v]];
        tree.bv_pos[v] = bits;
        if (me.kid[0] == kUndef32)
        { // a leaf keeps its symbol in the rank field (wt_helper.hpp:119-137) and ends a code word
            tree.child[v][0] = tree.child[v][1] = kWtUndef;
            tree.bv_pos_rank[v] = me.symbol;
            tree.c_to_leaf[me.symbol] = (uint16_t)v;
            too_deep |= depth[v] > 56;
            tree.path[me.symbol] = word[v] | ((uint64_t)depth[v] << 56);
            continue;
        }
        bits += me.weight; // an inner node owns one bit per symbol below it
        for (uint32_t k = 0; k < 2; ++k, ++placed)
        {
            origin[placed] = me.kid[k];
            tree.parent[placed] = (uint16_t)v;
            tree.child[v][k] = (uint16_t)placed;
            depth[placed] = depth[v] + 1;
            word[placed] = depth[v] < 64 ? (word[v] | ((uint64_t)k << depth[v])) : word[v];
        }
    }
    // a symbol that does not occur has an empty code; its path word names the last occurring symbol before it
    // (wt_helper.hpp:311-315) — part of the serialised tree
    uint64_t last_present = 0;
    for (uint64_t c = 0; c < 256; ++c)
    {
        if (tree.c_to_leaf[c] != kWtUndef)
            last_present = c;
        else
            tree.path[c] = last_present;
    }
    return too_deep ? kWtTooDeep : bits;
}

// The bit planes: chunk the text over T threads.  Per chunk and node the start offset is the node's
// bv_pos plus the number of symbols of that node's subtree in earlier chunks, so every thread writes
// disjoint bit ranges; words shared between two ranges are merged with an atomic OR.
inline void fill_bit_planes(uint8_t const * text, uint64_t n, WtTree const & tree, std::vector<uint64_t> & bv)
{
    unsigned T = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    if (n < (1u << 16))
        T = 1;
    uint64_t chunk = (n + T - 1) / T;
    uint32_t const nn = tree.nnodes;
    // per chunk: symbols under each node
    std::vector<std::vector<uint64_t>> cnt(T, std::vector<uint64_t>(nn, 0));
    auto count_chunk = [&](unsigned t) {
        uint64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        uint64_t h[256] = {0};
        for (uint64_t k = lo; k < hi; ++k)
            ++h[text[k]];
        for (int c = 0; c < 256; ++c)
        {
            if (!h[c])
                continue;
            uint16_t v = tree.c_to_leaf[c];
            while (v != 0)
            {
                v = tree.parent[v];
                cnt[t][v] += h[c];
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back(count_chunk, t);
        for (auto & x : th)
            x.join();
    }
    std::vector<std::vector<uint64_t>> start(T, std::vector<uint64_t>(nn, 0));
    for (uint32_t v = 0; v < nn; ++v)
    {
        uint64_t p = tree.bv_pos[v];
        for (unsigned t = 0; t < T; ++t)
        {
            start[t][v] = p;
            p += cnt[t][v];
        }
    }
    std::atomic<uint64_t> * words = reinterpret_cast<std::atomic<uint64_t> *>(bv.data());
    auto fill_chunk = [&](unsigned t) {
        uint64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        std::vector<uint64_t> pos(start[t]);
        // per node: the word being assembled
        std::vector<uint64_t> cur_idx(nn, ~0ull), cur_bits(nn, 0);
        auto flush = [&](uint32_t v) {
            if (cur_idx[v] != ~0ull && cur_bits[v])
                words[cur_idx[v]].fetch_or(cur_bits[v], std::memory_order_relaxed);
        };
        for (uint64_t k = lo; k < hi; ++k)
        {
            uint64_t p = tree.path[text[k]];
            uint32_t len = (uint32_t)(p >> 56);
            uint16_t v = 0;
            for (uint32_t l = 0; l < len; ++l, p >>= 1)
            {
                uint64_t q = pos[v]++;
                uint64_t wi = q >> 6;
                if (wi != cur_idx[v])
                {
                    flush(v);
                    cur_idx[v] = wi;
                    cur_bits[v] = 0;
                }
                cur_bits[v] |= (p & 1) << (q & 63);
                v = tree.child[v][p & 1];
            }
        }
        for (uint32_t v = 0; v < nn; ++v)
            flush(v);
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back(fill_chunk, t);
    for (auto & x : th)
        x.join();
}

} // namespace sdslgpu
