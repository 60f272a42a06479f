// host_image.h — TEST code: a sector-block image (DESIGN.md §3.2 / bv.cu) of a plain bit vector built on the host,
// for the harnesses that run the device query functions under SDSLGPU_HOST_EMU.
#pragma once
#include <cstdint>
#include <vector>

#include "../../sdsl-lite_b200/csrc/bv_device.cuh"

namespace hostimg
{
using namespace sdslgpu;

struct HostImage
{
    std::vector<bvblock> blocks;
    std::vector<uint64_t> top;
    std::vector<uint32_t> samp[2];
    BvView view;
};

// the layout of DESIGN.md §3.2 / bv.cu: 224 payload bits per block, cnt = ones since the start of the superblock of
// 2^24 blocks, top[] = absolute count per superblock; samp[b][j] = block of the (j * S + 1)-th b-bit, two sentinels
inline void build(HostImage & im, uint64_t const * words, uint64_t nbits, uint32_t log_s, uint32_t interp, uint32_t pos_mode = 0)
{
    uint64_t nblocks = nbits / kBlockBits + 1, n32 = (nbits + 31) >> 5;
    uint32_t const * w32 = reinterpret_cast<uint32_t const *>(words);
    im.blocks.assign(nblocks, bvblock{});
    im.top.assign(((nblocks - 1) >> kSuperShift) + 1, 0);
    uint64_t ones = 0, base = 0;
    for (int b = 0; b < 2; ++b)
        im.samp[b].clear();
    uint64_t seen[2] = {0, 0};
    for (uint64_t k = 0; k < nblocks; ++k)
    {
        if ((k & ((1ull << kSuperShift) - 1)) == 0)
        {
            base = ones;
            im.top[k >> kSuperShift] = base;
        }
        im.blocks[k].cnt = (uint32_t)(ones - base);
        for (int j = 0; j < 7; ++j)
        {
            uint64_t c = k * 7 + j;
            uint32_t x = c < n32 ? w32[c] : 0u;
            if (c + 1 == n32 && (nbits & 31))
                x &= (1u << (nbits & 31)) - 1u;
            im.blocks[k].d[j] = x;
        }
        uint64_t first = k * kBlockBits, valid = first >= nbits ? 0 : (nbits - first < kBlockBits ? nbits - first : kBlockBits);
        for (uint64_t o = 0; o < valid; ++o)
        {
            int bit = (im.blocks[k].d[o >> 5] >> (o & 31)) & 1;
            if ((seen[bit] & ((1ull << log_s) - 1)) == 0)
                im.samp[bit].push_back(pos_mode ? (uint32_t)((first + o) >> 5) : (uint32_t)k);
            ++seen[bit];
            ones += bit;
        }
    }
    for (int b = 0; b < 2; ++b)
    {
        im.samp[b].push_back((uint32_t)(pos_mode ? (nblocks - 1) * 7 + 6 : nblocks - 1));
        im.samp[b].push_back((uint32_t)(pos_mode ? (nblocks - 1) * 7 + 6 : nblocks - 1));
        im.view.sect[b] = nullptr; // select sectors: built by the harnesses that test them
        im.view.sect_stride[b] = 0;
        im.view.sect_magic[b] = 0;
        im.view.samp_pos[b] = pos_mode;
        im.view.samp[b] = im.samp[b].data();
        im.view.log_s[b] = log_s;
        im.view.interp[b] = interp;
    }
    im.view.blocks = im.blocks.data();
    im.view.top = im.top.data();
    im.view.nbits = nbits;
    im.view.ones = ones;
}

} // namespace hostimg
