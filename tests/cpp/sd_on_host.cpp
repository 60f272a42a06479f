// sd_on_host.cpp — the per-query DEVICE functions of sdsl-lite_b200/csrc/sd_device.cuh (sd_rank1_one, sd_select1_one,
// sd_low) compiled as plain C++ (-DSDSLGPU_HOST_EMU) over an image built on the host from a serialised sd_vector<>
// (sd_vector.hpp:426-438: size, wl, m_low, m_high, ...): m_high becomes a sector-block image (host_image.h), m_low is
// used as it is — what sdslgpu_load_sdsl does before uploading.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sdsl-lite_b200/csrc/sd_device.cuh"
#include "host_image.h"

using namespace sdslgpu;

namespace
{
struct Sd
{
    std::vector<uint64_t> low, high;
    hostimg::HostImage img;
    std::vector<uint32_t> samp0;
    SdView v;
};
} // namespace

extern "C"
{
    void * sd_emu_load(uint8_t const * blob)
    {
        Sd * s = new Sd;
        uint64_t size, h;
        std::memcpy(&size, blob, 8);
        uint32_t wl = blob[8];
        uint8_t const * p = blob + 9;
        std::memcpy(&h, p, 8);
        uint64_t lbits = h & ((1ull << 56) - 1), lw = (lbits + 63) >> 6;
        s->low.assign(lw + 2, 0);
        std::memcpy(s->low.data(), p + 8, lw * 8);
        p += 8 + lw * 8;
        std::memcpy(&h, p, 8);
        uint64_t hbits = h & ((1ull << 56) - 1), hw = (hbits + 63) >> 6;
        s->high.assign(hw + 2, 0);
        std::memcpy(s->high.data(), p + 8, hw * 8);
        hostimg::build(s->img, s->high.data(), hbits, 6, 0);
        s->v.size = size;
        s->v.m = wl ? lbits / wl : 0;
        s->v.wl = wl;
        s->v.high = s->img.view;
        s->v.low = s->low.data();
        s->v.samp0 = nullptr;
        s->v.log_s0 = 0;
        return s;
    }
    void sd_emu_free(void * h)
    {
        delete static_cast<Sd *>(h);
    }
    void sd_emu_rank(void * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out)
    {
        Sd * s = static_cast<Sd *>(h);
        for (uint64_t k = 0; k < n; ++k)
        {
            uint64_t r = sd_rank1_one(s->v, idx[k]);
            out[k] = b ? r : idx[k] - r;
        }
    }
    void sd_emu_select1(void * h, uint64_t const * i, uint64_t n, uint64_t * out)
    {
        Sd * s = static_cast<Sd *>(h);
        for (uint64_t k = 0; k < n; ++k)
            out[k] = sd_select1_one(s->v, i[k]);
    }
    // select_0 through the sample table, built here exactly as sd.cu's sd_samp0_kernel does (one "thread" per block);
    // log_s < 0: the library's choice (about one sample per block).  *fallbacks = queries the samples gave up on.
    void sd_emu_select0(void * h, int log_s, uint64_t const * i, uint64_t n, uint64_t * out, uint64_t * fallbacks)
    {
        Sd * s = static_cast<Sd *>(h);
        uint64_t const zeros = s->v.size - s->v.m, nblocks = s->v.high.nbits / kBlockBits + 1;
        *fallbacks = 0;
        if (zeros == 0)
            return;
        uint32_t ls = 0;
        if (log_s < 0)
            while ((zeros >> ls) > nblocks)
                ++ls;
        else
            ls = (uint32_t)log_s;
        uint64_t const nsamp = ((zeros - 1) >> ls) + 1, S = 1ull << ls;
        s->samp0.assign(nsamp + 1, 0);
        s->v.samp0 = nullptr;
        for (uint64_t g = 0; g < nblocks; ++g)
        {
            int64_t v_prev, v_last;
            if (!sd_block_zero_span(s->v, g, v_prev, v_last))
                continue;
            for (uint64_t q = ((uint64_t)v_prev + S - 1) >> ls; q < nsamp && (q << ls) + 1 <= (uint64_t)v_last; ++q)
                s->samp0[q] = (uint32_t)g;
        }
        s->v.samp0 = s->samp0.data();
        s->v.log_s0 = ls;
        for (uint64_t k = 0; k < n; ++k)
        {
            uint64_t r = sd_select0_one(s->v, i[k]);
            if (r == ~0ull)
            {
                ++*fallbacks;
                r = sd_select0_bsearch(s->v, i[k]);
            }
            out[k] = r;
        }
    }
}
