// wt_shape.h — host side of the wt_huff<> constructor (wt_pc.hpp:194-248, wt_huff.hpp:82-115, wt_helper.hpp:230-327):
// the Huffman shape with the reference's tie-breaking and BFS numbering, and the multi-threaded fill of the bit
// planes (the fallback of wt_build.cu's device fill).  Plain C++ (no CUDA) so that tests/test_wt_shape_cpu.py can
// check tree and bits against the reference's serialised tree on a box without a GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <queue>
#include <thread>
#include <utility>
#include <vector>

#include "wt_tree.h"

namespace sdslgpu
{

// ------------------------------------------------------------------------------------------------
// host: Huffman shape + BFS layout (tiny), then the bit planes in parallel over text chunks
// ------------------------------------------------------------------------------------------------
namespace
{
struct PcNode
{
    uint64_t freq, sym, parent, child[2];
};
} // namespace

// Fills tree (bv_pos, children, parents, c_to_leaf, path) and returns the number of bits of m_bv.
// Tie-breaking follows the reference exactly: min-heap ordered by (frequency, node number); the first
// node popped becomes child 0 (wt_huff.hpp:102-114); nodes are renumbered in BFS order with the two
// children of a node adjacent (wt_helper.hpp:236-271).
inline uint64_t build_huff_tree(uint64_t const (&C)[256], WtTree & tree, uint64_t & sigma)
{
    std::vector<PcNode> t;
    typedef std::pair<uint64_t, uint64_t> P;
    std::priority_queue<P, std::vector<P>, std::greater<P>> pq;
    sigma = 0;
    for (uint64_t c = 0; c < 256; ++c)
        if (C[c] > 0)
        {
            pq.push(P(C[c], t.size()));
            t.push_back(PcNode{C[c], c, ~0ull, {~0ull, ~0ull}});
            ++sigma;
        }
    while (pq.size() > 1)
    {
        P a = pq.top();
        pq.pop();
        P b = pq.top();
        pq.pop();
        t[a.second].parent = t.size();
        t[b.second].parent = t.size();
        pq.push(P(a.first + b.first, t.size()));
        t.push_back(PcNode{a.first + b.first, 0, ~0ull, {a.second, b.second}});
    }
    std::memset(&tree, 0, sizeof(tree));
    tree.nnodes = (uint32_t)t.size();
    // BFS relabel
    std::vector<uint64_t> src(t.size()); // BFS id -> index in t
    std::vector<uint64_t> freq(t.size());
    uint64_t bv_size = 0, node_cnt = 1, head = 0;
    src[0] = t.size() - 1;
    tree.parent[0] = kWtUndef;
    while (head < node_cnt)
    {
        uint64_t idx = head++;
        PcNode const & p = t[src[idx]];
        tree.bv_pos[idx] = bv_size;
        if (p.child[0] != ~0ull)
        {
            bv_size += p.freq;
            for (int k = 0; k < 2; ++k)
            {
                src[node_cnt] = p.child[k];
                tree.parent[node_cnt] = (uint16_t)idx;
                tree.child[idx][k] = (uint16_t)node_cnt++;
            }
        }
        else
        {
            tree.child[idx][0] = tree.child[idx][1] = kWtUndef;
            tree.bv_pos_rank[idx] = p.sym; // leaves keep the symbol here (wt_helper.hpp:119-137)
        }
    }
    for (int c = 0; c < 256; ++c)
        tree.c_to_leaf[c] = kWtUndef;
    for (uint64_t v = 0; v < t.size(); ++v)
        if (tree.child[v][0] == kWtUndef)
            tree.c_to_leaf[(uint8_t)tree.bv_pos_rank[v]] = (uint16_t)v;
    uint64_t prev_c = 0;
    for (uint64_t c = 0; c < 256; ++c)
    {
        if (tree.c_to_leaf[c] != kWtUndef)
        {
            uint16_t v = tree.c_to_leaf[c];
            uint64_t pw = 0, pl = 0;
            while (v != 0)
            {
                pw <<= 1;
                if (tree.child[tree.parent[v]][1] == v)
                    pw |= 1;
                ++pl;
                v = tree.parent[v];
            }
            tree.path[c] = pw | (pl << 56);
            prev_c = c;
        }
        else
            tree.path[c] = prev_c; // length 0 (wt_helper.hpp:311-315 stores the previous symbol here)
    }
    return bv_size;
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
