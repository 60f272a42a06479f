// sdsl_pack.h — host-side writers for the reference's serialised byte format (egress; SURVEY.md §8(f)-1 "and back").
//
// Pure host C++ (no CUDA): sdsl_egress.cu feeds it with argument positions computed by the device select kernels,
// tests/cpp/pack_host.cpp feeds it from a naive scan so the packing rules are checked on a CPU-only box.
//
// Formats written (all little-endian):
//   int_vector<w>        u64 (width << 56 | bit_size), ceil(bit_size / 64) words      int_vector.hpp:904-916, 1995-2004
//   select_support_mcl   u64 arg_cnt; if != 0: m_superblock, mini_or_long, then one int_vector<0> per superblock
//                        (select_support_mcl.hpp:474-518), with the CONTENT rules of init_slow (:207-266, vectors
//                        shorter than 100000 bits) and init_fast (:269-381) — stated here over the argument positions
//                        P[0..m) instead of the reference's word scan:
//     superblock j holds the arguments k in [4096 j, min(4096 (j+1), m)), cnt_j of them;
//     init_slow:  first = P[4096 j], last = P[4096 j + cnt_j - 1];
//     init_fast:  a block with cnt_j >= 4033 is closed when its 4033rd argument is met, and the scan for "the last
//                 argument of the block" (:302-309) runs one argument too far: last = P[4096 (j+1)] when that argument
//                 exists, else P[m-1]; a trailing block with cnt_j <= 4032 becomes a LONG block of width
//                 hi(size-1)+1 whose m_superblock entry is never written (stays 0) (:366-380);
//     last - first > logn^4  => long block: 4096 entries of hi(last)+1 bits, entry k = P[4096 j + k] for k < cnt_j;
//     else                      mini block: 64 entries of hi(last-first)+1 bits, entry t = P[4096 j + 64 t] - first;
//     mini_or_long is an EMPTY bit_vector when no long block exists, else one bit per superblock (1 = mini).
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace sdslgpu
{
namespace pack
{

// bits::hi (bits.hpp:653-684): index of the most significant set bit, hi(0) = 0
inline uint32_t hi(uint64_t x)
{
    return x ? 63u - (uint32_t)__builtin_clzll(x) : 0u;
}

struct Sink
{
    std::vector<uint8_t> & b;
    void raw(void const * p, size_t n)
    {
        uint8_t const * c = static_cast<uint8_t const *>(p);
        b.insert(b.end(), c, c + n);
    }
    void u64(uint64_t x)
    {
        raw(&x, 8);
    }
    void u32(uint32_t x)
    {
        raw(&x, 4);
    }
    void u16(uint16_t x)
    {
        raw(&x, 2);
    }
    void u8(uint8_t x)
    {
        b.push_back(x);
    }
    // int_vector<width> given as ready-made words
    void int_vector(uint32_t width, uint64_t bit_size, uint64_t const * words)
    {
        u64(((uint64_t)width << 56) | bit_size);
        raw(words, ((bit_size + 63) >> 6) * 8);
    }
};

// int_vector<0>(count, 0, width): zero-initialised, entries set once
struct PackedInts
{
    uint32_t width;
    uint64_t count;
    std::vector<uint64_t> w;
    PackedInts(uint64_t count_, uint32_t width_) : width(width_), count(count_), w(((count_ * width_ + 63) >> 6) + 1, 0)
    {}
    void set(uint64_t i, uint64_t v)
    {
        uint64_t pos = i * width;
        uint32_t off = (uint32_t)(pos & 63);
        if (width < 64)
            v &= (1ull << width) - 1;
        w[pos >> 6] |= v << off;
        if (off + width > 64)
            w[(pos >> 6) + 1] |= v >> (64 - off);
    }
    void write(Sink & s) const
    {
        s.int_vector(width, count * width, w.data());
    }
};

// pos[j] = position of argument number keys[j] (0-based: P[keys[j]]), keys[j] < m.  Returns 0 or an error status.
typedef std::function<int(uint64_t const * keys, uint64_t n, uint64_t * pos)> SelectFn;

// select_support_mcl<b,1>::serialize for a vector of nbits bits holding m arguments
inline int write_select_mcl(uint64_t nbits, uint64_t m, SelectFn const & sel, Sink & out)
{
    out.u64(m);
    if (m == 0)
        return 0;
    uint32_t const logn = hi(((nbits + 63) >> 6) << 6) + 1; // initData, :456-458
    uint64_t const logn4 = (uint64_t)(logn * logn) * (uint64_t)(logn * logn);
    bool const fast = nbits >= 100000; // ctor dispatch, :121-128
    uint64_t const sb = (m + 4095) >> 12;
    PackedInts super(sb, logn);
    std::vector<uint8_t> is_mini(sb, 1);
    bool any_long = false;
    std::vector<uint8_t> body;
    Sink bs{body};
    uint64_t const kChunk = 2048, kSlots = 65; // superblocks per device round trip; 64 mini keys + the "last" key
    std::vector<uint64_t> keys, pos, lkeys, lpos;
    for (uint64_t j0 = 0; j0 < sb; j0 += kChunk)
    {
        uint64_t j1 = j0 + kChunk < sb ? j0 + kChunk : sb;
        keys.assign((j1 - j0) * kSlots, 0);
        for (uint64_t j = j0; j < j1; ++j)
        {
            uint64_t base = j << 12, cnt = m - base < 4096 ? m - base : 4096;
            uint64_t * k = keys.data() + (j - j0) * kSlots;
            for (uint64_t t = 0; t < 64; ++t)
                k[t] = base + 64 * t < m ? base + 64 * t : m - 1;
            if (fast && cnt >= 4033)
                k[64] = base + 4096 < m ? base + 4096 : m - 1;
            else
                k[64] = base + cnt - 1;
        }
        pos.resize(keys.size());
        if (int st = sel(keys.data(), keys.size(), pos.data()))
            return st;
        // which blocks of this chunk are long, and their arguments
        lkeys.clear();
        for (uint64_t j = j0; j < j1; ++j)
        {
            uint64_t base = j << 12, cnt = m - base < 4096 ? m - base : 4096;
            uint64_t const * p = pos.data() + (j - j0) * kSlots;
            bool trailing = fast && cnt <= 4032;
            if (trailing || p[64] - p[0] > logn4)
            {
                is_mini[j] = 0;
                any_long = true;
                for (uint64_t k = 0; k < cnt; ++k)
                    lkeys.push_back(base + k);
            }
        }
        lpos.resize(lkeys.size());
        if (!lkeys.empty())
            if (int st = sel(lkeys.data(), lkeys.size(), lpos.data()))
                return st;
        uint64_t lcur = 0;
        for (uint64_t j = j0; j < j1; ++j)
        {
            uint64_t base = j << 12, cnt = m - base < 4096 ? m - base : 4096;
            uint64_t const * p = pos.data() + (j - j0) * kSlots;
            bool trailing = fast && cnt <= 4032;
            if (!trailing)
                super.set(j, p[0]);
            if (!is_mini[j])
            {
                PackedInts lb(4096, trailing ? hi(nbits - 1) + 1 : hi(p[64]) + 1);
                for (uint64_t k = 0; k < cnt; ++k)
                    lb.set(k, lpos[lcur + k]);
                lcur += cnt;
                lb.write(bs);
            }
            else
            {
                PackedInts mb(64, hi(p[64] - p[0]) + 1);
                for (uint64_t t = 0; 64 * t < cnt; ++t)
                    mb.set(t, p[t] - p[0]);
                mb.write(bs);
            }
        }
    }
    super.write(out);
    if (any_long)
    {
        std::vector<uint64_t> mol(((sb + 63) >> 6) + 1, 0);
        for (uint64_t j = 0; j < sb; ++j)
            if (is_mini[j])
                mol[j >> 6] |= 1ull << (j & 63);
        out.int_vector(1, sb, mol.data());
    }
    else
        out.int_vector(1, 0, nullptr);
    out.raw(body.data(), body.size());
    return 0;
}

// byte_tree::serialize (wt_helper.hpp:362-375; node = bv_pos, bv_pos_rank, parent, child[0], child[1], :139-150)
template <class Tree>
inline void write_byte_tree(Tree const & t, Sink & out)
{
    out.u64(t.nnodes);
    for (uint32_t v = 0; v < t.nnodes; ++v)
    {
        out.u64(t.bv_pos[v]);
        out.u64(t.bv_pos_rank[v]);
        out.u16(t.parent[v]);
        out.u16(t.child[v][0]);
        out.u16(t.child[v][1]);
    }
    for (int c = 0; c < 256; ++c)
        out.u16(t.c_to_leaf[c]);
    for (int c = 0; c < 256; ++c)
        out.u64(t.path[c]);
}

} // namespace pack
} // namespace sdslgpu
