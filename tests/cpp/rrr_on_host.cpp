// rrr_on_host.cpp — the per-query DEVICE functions of sdsl-lite_b200/csrc/rrr_device.cuh (rrr_rank1_one,
// rrr_rank1_and_bit, rrr_select_one<B>, rrr_decode, the record accessors) compiled as plain C++ (-DSDSLGPU_HOST_EMU)
// over an image built on the host from a serialised rrr_vector<63> (rrr_vector.hpp:366-378) with the product's own
// rrr_records_host / host_tables — what sdslgpu_load_sdsl does before uploading.  tests/test_device_logic_cpu.py
// checks the results against the oracle.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sdsl-lite_b200/csrc/rrr_device.cuh"

using namespace sdslgpu;

namespace
{
struct IntVec
{
    uint32_t width = 0;
    uint64_t bits = 0;
    std::vector<uint64_t> w;
    uint64_t size() const
    {
        return width ? bits / width : 0;
    }
    uint64_t get(uint64_t i) const
    {
        uint64_t pos = i * width;
        uint32_t off = (uint32_t)(pos & 63);
        uint64_t lo = w[pos >> 6] >> off;
        if (off + width > 64)
            lo |= w[(pos >> 6) + 1] << (64 - off);
        return width == 64 ? lo : (lo & ((1ull << width) - 1));
    }
};

uint8_t const * read_iv(uint8_t const * p, IntVec & v)
{
    uint64_t h;
    std::memcpy(&h, p, 8);
    v.width = (uint32_t)(h >> 56);
    v.bits = h & ((1ull << 56) - 1);
    uint64_t nw = (v.bits + 63) >> 6;
    v.w.assign(nw + 2, 0);
    std::memcpy(v.w.data(), p + 8, nw * 8);
    return p + 8 + nw * 8;
}

struct Emu
{
    std::vector<uint64_t> btnr, rec;
    std::vector<uint32_t> hint[2];
    RrrView v;
};
} // namespace

extern "C"
{
    void * rrr_emu_load(uint8_t const * blob)
    {
        Emu * e = new Emu;
        uint64_t size;
        std::memcpy(&size, blob, 8);
        IntVec bt, btnr, btnrp, rank, inv;
        uint8_t const * p = read_iv(blob + 8, bt);
        p = read_iv(p, btnr);
        p = read_iv(p, btnrp);
        p = read_iv(p, rank);
        read_iv(p, inv);
        uint64_t nblocks = bt.size(), nsuper = btnrp.size(), ones = rank.get(rank.size() - 1);
        std::vector<uint64_t> rk(nsuper), bp(nsuper);
        std::vector<uint8_t> iv(nsuper);
        for (uint64_t g = 0; g < nsuper; ++g)
        {
            rk[g] = rank.get(g);
            bp[g] = btnrp.get(g);
            iv[g] = (uint8_t)inv.get(g);
        }
        rrr_records_host(bt.w.data(), nblocks, nsuper, ones, rk, bp, iv, 0, e->rec);
        e->btnr = btnr.w;
        for (int b = 0; b < 2; ++b)
        { // rrr.cu rrr_hint_kernel: superblock holding the (j * 2^shift + 1)-th b-bit, two sentinels
            uint64_t args = b ? ones : nsuper * kBs * kK - ones;
            uint32_t const kHintShift = rrr_hint_shift(args, nsuper);
            e->v.hint_shift[b] = kHintShift;
            uint64_t nhint = args ? ((args - 1) >> kHintShift) + 1 : 0;
            e->hint[b].assign(nhint + 2, (uint32_t)(nsuper ? nsuper - 1 : 0));
            for (uint64_t g = 0; g < nsuper; ++g)
            {
                uint64_t r0 = e->rec[g * kRecWords], r1 = e->rec[(g + 1) * kRecWords];
                uint64_t a = b ? r0 : g * kBs * kK - r0, en = b ? r1 : (g + 1) * kBs * kK - r1;
                for (uint64_t j = (a + (1ull << kHintShift) - 1) >> kHintShift; en > a && j < nhint && (j << kHintShift) + 1 <= en; ++j)
                    e->hint[b][j] = (uint32_t)g;
            }
            e->v.hint[b] = e->hint[b].data();
        }
        e->v.size = size;
        e->v.nblocks = nblocks;
        e->v.nsuper = nsuper;
        e->v.ones = ones;
        e->v.try_sparse = 1; // the one-at-a-time search wherever a block qualifies, whatever the density of the vector
        e->v.btnr = e->btnr.data();
        e->v.records = e->rec.data();
        e->v.tables = &host_tables();
        return e;
    }
    void rrr_emu_free(void * h)
    {
        delete static_cast<Emu *>(h);
    }
    void rrr_emu_rank(void * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out)
    {
        Emu * e = static_cast<Emu *>(h);
        for (uint64_t k = 0; k < n; ++k)
        {
            uint64_t r = rrr_rank1_one(e->v, e->v.tables, idx[k]);
            out[k] = b ? r : idx[k] - r;
        }
    }
    void rrr_emu_access(void * h, uint64_t const * idx, uint64_t n, uint64_t * out, uint64_t * rank_out)
    {
        Emu * e = static_cast<Emu *>(h);
        for (uint64_t k = 0; k < n; ++k)
        {
            uint32_t bit = 0;
            rank_out[k] = rrr_rank1_and_bit(e->v, e->v.tables, idx[k], bit);
            out[k] = bit;
        }
    }
    void rrr_emu_select(void * h, int b, uint64_t const * i, uint64_t n, uint64_t * out)
    {
        Emu * e = static_cast<Emu *>(h);
        for (uint64_t k = 0; k < n; ++k)
            out[k] = b ? rrr_select_one<1>(e->v, e->v.tables, i[k]) : rrr_select_one<0>(e->v, e->v.tables, i[k]);
    }
}
