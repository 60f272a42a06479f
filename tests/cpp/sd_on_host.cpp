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
}
