// sdsl_egress.cu — complete reference-format blobs out of the device images (SURVEY.md §8(f)-1 "... and back"):
// what serialize() / store_to_file of the reference writes for
//   select_support_mcl<1>/<0>                          select_support_mcl.hpp:474-518
//   sd_vector<>          incl. its two select supports  sd_vector.hpp:426-438
//   wt_huff<> / wt_huff<rrr_vector<63>>                 wt_pc.hpp:713-726, wt_helper.hpp:362-375
//   wt_int<>                                            wt_int.hpp:792-805
//   csa_wt<wt_huff<>, t_dens, t_inv_dens>               csa_wt.hpp:389-402, csa_alphabet_strategy.hpp:258-268
// so that an index built on the GPU (seconds for a 2^30-byte text) can be stored and loaded by the reference.
// The heavy parts stay on the device: the bit vector is unpacked from the sector blocks, the reference's
// m_basic_block table is built by bv.cu's table kernels, and every argument position the select supports store
// (every 64th argument, the block ends, all arguments of long blocks) comes from the batched select kernel; the
// host only packs the variable-width fields (sdsl_pack.h).
#include "internal.h"
#include "sdsl_pack.h"

namespace sdslgpu
{

namespace
{

struct DBuf
{
    void * p = nullptr;
    ~DBuf()
    {
        reset();
    }
    void reset()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
    }
    int alloc(uint64_t bytes)
    {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
        if (e != cudaSuccess)
        {
            p = nullptr;
            return cuda_fail(e, "cudaMalloc (serialisation scratch)", __FILE__, __LINE__);
        }
        return SDSLGPU_OK;
    }
    template <class T>
    T * as() const
    {
        return static_cast<T *>(p);
    }
};

// sector blocks -> the plain LSB-first words of int_vector<1> (224 = 7 * 32: chunk c is payload word c % 7 of block c / 7)
__global__ void __launch_bounds__(kThreads) bv_unpack_kernel(bvblock const * __restrict__ blocks, uint64_t nblocks, uint64_t nchunks, uint32_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride)
        out[c] = (c / 7 < nblocks) ? ld_nc_u32(&blocks[c / 7].d[c % 7]) : 0u;
}

__global__ void __launch_bounds__(kThreads) add_one_kernel(uint64_t * __restrict__ k, uint64_t n)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
        k[q] += 1;
}

// device words (nwords + 2, zero padded) of a bit-vector image
int unpack_words(BvImage const & v, DBuf & words)
{
    uint64_t nwords = (v.nbits + 63) >> 6;
    SG_TRY(words.alloc((nwords + 2) * 8));
    bv_unpack_kernel<<<grid_for(2 * (nwords + 2)), kThreads>>>(v.blocks, v.nblocks, 2 * (nwords + 2), words.as<uint32_t>());
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace

// what 0 = bit_vector, 1 = rank_support_v<1>, 2 = rank_support_v<0>, 5 = rank_support_v5<1>, 6 = rank_support_v5<0>
// of an image that keeps the sector blocks only (handles created with SDSLGPU_F_SDSL_LAYOUT serve 0..2 from their
// resident words / tables, api.cu)
int egress_bv_part(BvImage const & v, int what, std::vector<uint8_t> & blob)
{
    pack::Sink out{blob};
    uint64_t nwords = (v.nbits + 63) >> 6;
    DBuf words;
    SG_TRY(unpack_words(v, words));
    std::vector<uint64_t> host;
    if (what == 0)
    {
        host.assign(nwords + 1, 0);
        if (nwords)
            SG_CUDA(cudaMemcpy(host.data(), words.p, nwords * 8, cudaMemcpyDeviceToHost));
        out.int_vector(1, v.nbits, host.data());
        return SDSLGPU_OK;
    }
    int const b = (what == 1 || what == 5) ? 1 : 0;
    DevicePool scratch;
    BvImage tmp;
    tmp.nbits = v.nbits;
    tmp.nwords = nwords;
    tmp.words = words.as<uint64_t>();
    int st = bv_build_sdsl_rank_table(scratch, tmp, b, nullptr, what >= 5);
    if (st == SDSLGPU_OK)
    {
        host.assign(tmp.table_words + 1, 0);
        cudaError_t e = cudaMemcpy(host.data(), tmp.rank_table[b], tmp.table_words * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
            st = cuda_fail(e, "D2H (rank table)", __FILE__, __LINE__);
        else
            out.int_vector(64, tmp.table_words * 64, host.data());
    }
    scratch.release_all();
    return st;
}

// select_support_mcl<b,1>::serialize over the image's bit vector
int egress_select_mcl(BvImage const & v, int b, std::vector<uint8_t> & blob)
{
    pack::Sink out{blob};
    uint64_t m = b ? v.ones : v.nbits - v.ones;
    if (m && v.samp[b] == nullptr)
    {
        set_error("serialising a select support needs a handle created without SDSLGPU_F_NO_SELECT");
        return SDSLGPU_ENOTSUP;
    }
    DBuf dk, dp;
    uint64_t cap = 0;
    pack::SelectFn sel = [&](uint64_t const * keys, uint64_t n, uint64_t * pos) -> int {
        if (n > cap)
        {
            dk.reset();
            dp.reset();
            cap = n + n / 4 + 1024;
            SG_TRY(dk.alloc(cap * 8));
            SG_TRY(dp.alloc(cap * 8));
        }
        SG_CUDA(cudaMemcpy(dk.p, keys, n * 8, cudaMemcpyHostToDevice));
        add_one_kernel<<<grid_for(n), kThreads>>>(dk.as<uint64_t>(), n); // argument k is select(k + 1)
        SG_CUDA(cudaGetLastError());
        SG_TRY(bv_select_device(v, b, dk.as<uint64_t>(), n, dp.as<uint64_t>(), nullptr));
        SG_CUDA(cudaMemcpy(pos, dp.p, n * 8, cudaMemcpyDeviceToHost));
        return SDSLGPU_OK;
    };
    return pack::write_select_mcl(v.nbits, m, sel, out);
}

// sd_vector<>::serialize: size, wl, m_low, m_high, select_support_mcl<1> and <0> over m_high (sd_vector.hpp:426-438)
int egress_sd(sdslgpu_handle const * h, std::vector<uint8_t> & blob)
{
    SdImage const & d = h->sd;
    pack::Sink out{blob};
    out.u64(d.size);
    out.u8((uint8_t)d.wl);
    uint64_t lbits = d.m * d.wl, lw = (lbits + 63) >> 6, hw = (d.high_bits + 63) >> 6;
    std::vector<uint64_t> host(lw + 1, 0);
    if (lw)
        SG_CUDA(cudaMemcpy(host.data(), d.low, lw * 8, cudaMemcpyDeviceToHost));
    out.int_vector(d.wl, lbits, host.data());
    DBuf words;
    SG_TRY(unpack_words(d.high, words));
    host.assign(hw + 1, 0);
    if (hw)
        SG_CUDA(cudaMemcpy(host.data(), words.p, hw * 8, cudaMemcpyDeviceToHost));
    out.int_vector(1, d.high_bits, host.data());
    SG_TRY(egress_select_mcl(d.high, 1, blob));
    return egress_select_mcl(d.high, 0, blob);
}

// v5_scan: the wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>> form of the
// reference's count-benchmark index (benchmark/indexing_count/index.config:8): the rank slot holds rank_support_v5's
// table and the two scanning select supports serialise to nothing (select_support_scan.hpp:63-66)
int egress_wt_huff(sdslgpu_handle const * h, std::vector<uint8_t> & blob, bool v5_scan)
{
    WtHuffImage const & w = h->wt;
    pack::Sink out{blob};
    out.u64(w.size);
    out.u64(w.sigma);
    if (w.use_rrr)
        SG_TRY(rrr_serialize_image(w.rrr, blob)); // rank_support_rrr / select_support_rrr serialise to nothing (rrr_vector.hpp:580-585)
    else
    {
        SG_TRY(egress_bv_part(w.bv, 0, blob));
        SG_TRY(egress_bv_part(w.bv, v5_scan ? 5 : 1, blob)); // rank_support_v<1> (rank_support_v.hpp:151-158) / rank_support_v5<1>
        if (!v5_scan)
        {
            SG_TRY(egress_select_mcl(w.bv, 1, blob));
            SG_TRY(egress_select_mcl(w.bv, 0, blob));
        }
    }
    pack::write_byte_tree(w.host_tree, out);
    return SDSLGPU_OK;
}

int egress_wt_int(sdslgpu_handle const * h, std::vector<uint8_t> & blob)
{
    WtIntImage const & w = h->wti;
    pack::Sink out{blob};
    out.u64(w.size);
    out.u64(w.sigma);
    SG_TRY(egress_bv_part(w.tree, 0, blob));
    SG_TRY(egress_bv_part(w.tree, 1, blob));
    SG_TRY(egress_select_mcl(w.tree, 1, blob));
    SG_TRY(egress_select_mcl(w.tree, 0, blob));
    out.u32(w.max_level);
    return SDSLGPU_OK;
}

int egress_csa(sdslgpu_handle const * h, std::vector<uint8_t> & blob, bool v5_scan)
{
    CsaImage const & c = h->csa;
    SG_TRY(egress_wt_huff(h, blob, v5_scan));
    pack::Sink out{blob};
    uint32_t const width = pack::hi(c.n) + 1; // csa_sampling_strategy.hpp:103, 762
    std::vector<uint64_t> host;
    auto samples = [&](uint64_t const * dev, uint64_t count) -> int {
        host.assign(count + 1, 0);
        if (count)
            SG_CUDA(cudaMemcpy(host.data(), dev, count * 8, cudaMemcpyDeviceToHost));
        pack::PackedInts iv(count, width);
        for (uint64_t k = 0; k < count; ++k)
            iv.set(k, host[k]);
        iv.write(out);
        return SDSLGPU_OK;
    };
    SG_TRY(samples(c.samples, c.nsamples));
    SG_TRY(samples(c.isa_samples, c.nisa));
    // byte_alphabet: char2comp[256], comp2char[sigma], C[sigma + 1], sigma (u16)
    FmTables const & t = c.host_tab;
    uint64_t w8[33] = {0};
    std::memcpy(w8, t.char2comp, 256);
    out.int_vector(8, 256 * 8, w8);
    std::memset(w8, 0, sizeof(w8));
    std::memcpy(w8, t.comp2char, t.sigma);
    out.int_vector(8, (uint64_t)t.sigma * 8, w8);
    out.int_vector(64, ((uint64_t)t.sigma + 1) * 64, t.C);
    out.u16((uint16_t)t.sigma);
    return SDSLGPU_OK;
}

} // namespace sdslgpu
