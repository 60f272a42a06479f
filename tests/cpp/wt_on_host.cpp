// wt_on_host.cpp — the per-query DEVICE functions of sdsl-lite_b200/csrc/wt_device.cuh (wt_rank_one,
// wt_inverse_select_one) compiled as plain C++ (-DSDSLGPU_HOST_EMU) over a wavelet tree whose shape and bit planes come
// from the product's own host code (wt_shape.h) and whose bit vector is served by the PlainBits policy of
// bits_access.cuh over a host-built sector-block image (host_image.h).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sdsl-lite_b200/csrc/wt_device.cuh"
#include "../../sdsl-lite_b200/csrc/wt_shape.h"
#include "host_image.h"

using namespace sdslgpu;

namespace
{
struct Wt
{
    WtTree tree;
    uint64_t size = 0, sigma = 0, bits = 0;
    std::vector<uint64_t> bv;
    hostimg::HostImage plain;
    PlainBits pb;
};
} // namespace

extern "C"
{
    void * wt_emu_create(uint8_t const * text, uint64_t n)
    {
        Wt * w = new Wt;
        uint64_t C[256] = {0};
        for (uint64_t k = 0; k < n; ++k)
            ++C[text[k]];
        w->size = n;
        w->bits = build_huff_tree(C, w->tree, w->sigma);
        w->bv.assign(((w->bits + 63) >> 6) + 2, 0);
        fill_bit_planes(text, n, w->tree, w->bv);
        hostimg::build(w->plain, w->bv.data(), w->bits, 6, 0);
        w->pb.v = w->plain.view;
        for (uint32_t v = 0; v < w->tree.nnodes; ++v) // wt_helper.hpp:319-327 (the library asks the device for these)
            if (w->tree.child[v][0] != kWtUndef)
                w->tree.bv_pos_rank[v] = bv_rank1(w->plain.view, w->tree.bv_pos[v]);
        return w;
    }
    void wt_emu_free(void * h)
    {
        delete static_cast<Wt *>(h);
    }
    void wt_emu_rank(void * h, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out)
    {
        Wt * w = static_cast<Wt *>(h);
        for (uint64_t k = 0; k < n; ++k)
            out[k] = wt_rank_one(w->pb, &w->tree, w->sigma, i[k], c[k]);
    }
    void wt_emu_inverse_select(void * h, uint64_t const * i, uint64_t n, uint64_t * rank_out, uint64_t * sym_out)
    {
        Wt * w = static_cast<Wt *>(h);
        for (uint64_t k = 0; k < n; ++k)
        {
            uint32_t sym = 0;
            rank_out[k] = wt_inverse_select_one(w->pb, &w->tree, i[k], sym);
            sym_out[k] = sym;
        }
    }
}
