// device_on_host.cpp — the per-query DEVICE functions of sdsl-lite_b200/csrc/bv_device.cuh (bv_rank1, bv_rank1_and_bit,
// bv_bit, bv_select<B> with its hint / walk / bisect logic) compiled as plain C++ (-DSDSLGPU_HOST_EMU, common.cuh)
// over a sector-block image built here on the host, so that tests/test_device_logic_cpu.py can check the very source
// the kernels run against the oracle on a box without a GPU, for every sample stride and with interpolation on / off.
#include <cstdint>
#include <vector>

#include "../../sdsl-lite_b200/csrc/bv_device.cuh"
#include "host_image.h"

using namespace sdslgpu;
using hostimg::HostImage;
using hostimg::build;


extern "C"
{
    // out[k] = rank1(idx[k]); bit_out[k] (nullable) = the bit at idx[k] via bv_rank1_and_bit / bv_bit (idx < nbits)
    void emu_rank1(uint64_t const * words, uint64_t nbits, uint64_t const * idx, uint64_t n, uint64_t * out, uint64_t * bit_out)
    {
        HostImage im;
        build(im, words, nbits, 6, 0);
        for (uint64_t k = 0; k < n; ++k)
        {
            out[k] = bv_rank1(im.view, idx[k]);
            if (bit_out && idx[k] < nbits)
            {
                uint32_t bit = 0;
                uint64_t r = bv_rank1_and_bit(im.view, idx[k], bit);
                bit_out[k] = (r == out[k] && bit == bv_bit(im.view, idx[k])) ? bit : 99;
            }
        }
    }
    // out[k] = select_b(i[k]), 1 <= i[k] <= #b-bits, with samples every 2^log_s b-bits
    void emu_select(uint64_t const * words, uint64_t nbits, int b, uint32_t log_s, uint32_t interp, uint32_t pos_mode, uint64_t const * i, uint64_t n, uint64_t * out)
    {
        HostImage im;
        build(im, words, nbits, log_s, interp, pos_mode);
        for (uint64_t k = 0; k < n; ++k)
            out[k] = b ? bv_select<1>(im.view, i[k]) : bv_select<0>(im.view, i[k]);
    }
    // the same through select sectors (bv_device.cuh): built with bv_make_sector as bv.cu's kernel does, answered with
    // bv_select_sector, marked sectors by the sampled select.  stride = 0: the library's choice (bv_sect_stride).
    // Returns the number of queries that met a marked sector, or -1 if the density rules sectors out.
    int64_t emu_select_sectors(uint64_t const * words, uint64_t nbits, int b, uint32_t stride, uint64_t const * i, uint64_t n, uint64_t * out)
    {
        HostImage im;
        build(im, words, nbits, 9, 1, 1);
        uint64_t const args = b ? im.view.ones : nbits - im.view.ones;
        if (args == 0)
            return -1;
        uint32_t const ls = stride ? stride : bv_sect_stride(args, nbits);
        if (ls == 0)
            return -1;
        uint64_t const nsect = (args - 1) / ls + 1, nblocks = nbits / kBlockBits + 1;
        std::vector<bvblock> sect(nsect + 1);
        for (uint64_t j = 0; j < nsect; ++j)
        {
            if (b)
                bv_make_sector<1>(im.view, nblocks, args, ls, j, sect[j].cnt, sect[j].d);
            else
                bv_make_sector<0>(im.view, nblocks, args, ls, j, sect[j].cnt, sect[j].d);
        }
        im.view.sect[b] = sect.data();
        im.view.sect_stride[b] = ls;
        im.view.sect_magic[b] = bv_sect_magic(ls);
        int64_t marked = 0;
        for (uint64_t k = 0; k < n; ++k)
        {
            bool fits = false;
            uint64_t r = b ? bv_select_sector<1>(im.view, i[k] - 1, fits) : bv_select_sector<0>(im.view, i[k] - 1, fits);
            if (!fits)
            {
                ++marked;
                r = b ? bv_select<1>(im.view, i[k]) : bv_select<0>(im.view, i[k]);
            }
            out[k] = r;
        }
        return marked;
    }
    // key / stride the way bv_select_sector does it: multiply-high by bv_sect_magic(stride)
    uint64_t emu_sect_div(uint32_t stride, uint64_t key)
    {
        return __umul64hi(key, bv_sect_magic(stride));
    }
    uint32_t emu_sect_stride(uint64_t args, uint64_t nbits)
    {
        return bv_sect_stride(args, nbits);
    }
    uint32_t emu_sel64(uint64_t x, uint32_t k)
    {
        return sel64(x, k);
    }
}
