// sdsl_format.cu — ingest of the reference's own serialised byte format (what store_to_file / serialize()
// write), so that indexes built by SDSL on the CPU can be served by this engine without re-construction.
//
// Formats (verified byte for byte by tests/test_oracle_*.py):
//   int_vector<w>      u64 (width << 56 | bit_size), ceil(bit_size/64) u64 words      int_vector.hpp:904-916,1995-2004
//   rank_support_v     one int_vector<64>                                              rank_support_v.hpp:151-158
//   select_support_mcl u64 arg_cnt; if != 0: superblock, mini_or_long, per-superblock vectors  select_support_mcl.hpp:474-518
//   rrr_vector<63>     u64 size, bt, btnr, btnrp, rank, invert                         rrr_vector.hpp:366-378
//   sd_vector<>        u64 size, u8 wl, low, high, select_1, select_0                  sd_vector.hpp:426-438
//   wt_huff<>          u64 size, u64 sigma, bv, rank, select_1, select_0, tree        wt_pc.hpp:713-726
//   byte_tree          u64 n, n x 22-byte nodes, u16 c_to_leaf[256], u64 path[256]    wt_helper.hpp:362-375,139-150
//   wt_int<>           u64 size, u64 sigma, tree, rank, select_1, select_0, u32 max_level   wt_int.hpp:792-805
//   csa_wt<wt_huff<>>  wavelet tree, SA samples, ISA samples, byte_alphabet            csa_wt.hpp:389-402
//   byte_alphabet      int_vector<8> char2comp, int_vector<8> comp2char, int_vector<64> C, u16 sigma   csa_alphabet_strategy.hpp:258-268
// The rank/select supports inside a blob are skipped: this engine rebuilds its own (sector blocks + samples)
// on the device from the bit vector, which is what makes the answers identical by construction.
#include "internal.h"

namespace sdslgpu
{

namespace
{

struct Reader
{
    uint8_t const * p;
    uint64_t n, pos = 0;
    bool ok = true;
    bool need(uint64_t k)
    {
        if (!ok || n - pos < k)
        {
            ok = false;
            return false;
        }
        return true;
    }
    uint64_t u64()
    {
        uint64_t v = 0;
        if (need(8))
        {
            std::memcpy(&v, p + pos, 8);
            pos += 8;
        }
        return v;
    }
    uint32_t u32()
    {
        uint32_t v = 0;
        if (need(4))
        {
            std::memcpy(&v, p + pos, 4);
            pos += 4;
        }
        return v;
    }
    uint16_t u16()
    {
        uint16_t v = 0;
        if (need(2))
        {
            std::memcpy(&v, p + pos, 2);
            pos += 2;
        }
        return v;
    }
    uint8_t u8()
    {
        uint8_t v = 0;
        if (need(1))
            v = p[pos++];
        return v;
    }
};

struct IntVec
{
    uint32_t width = 0;
    uint64_t bits = 0;
    std::vector<uint64_t> words; // + one pad word
    uint64_t size() const
    {
        return width ? bits / width : 0;
    }
    uint64_t get(uint64_t i) const
    {
        uint64_t pos = i * width;
        uint32_t off = (uint32_t)(pos & 63);
        uint64_t lo = words[pos >> 6] >> off;
        if (off + width > 64)
            lo |= words[(pos >> 6) + 1] << (64 - off);
        return width == 64 ? lo : (lo & ((1ull << width) - 1));
    }
};

bool read_iv(Reader & r, IntVec & v, bool keep = true)
{
    uint64_t h = r.u64();
    v.width = (uint32_t)(h >> 56);
    v.bits = h & ((1ull << 56) - 1);
    uint64_t nw = (v.bits + 63) >> 6;
    if (!r.need(nw * 8))
        return false;
    if (keep)
    {
        v.words.assign(nw + 1, 0);
        if (nw)
            std::memcpy(v.words.data(), r.p + r.pos, nw * 8);
    }
    r.pos += nw * 8;
    return r.ok;
}

bool skip_iv(Reader & r)
{
    IntVec tmp;
    return read_iv(r, tmp, false);
}

// select_support_mcl.hpp:474-518
bool skip_select_mcl(Reader & r)
{
    uint64_t arg_cnt = r.u64();
    if (!r.ok)
        return false;
    if (arg_cnt == 0)
        return true;
    uint64_t sb = (arg_cnt + 4095) >> 12;
    if (!skip_iv(r))
        return false;
    IntVec mol;
    if (!read_iv(r, mol))
        return false;
    for (uint64_t k = 0; k < sb; ++k)
        if (!skip_iv(r))
            return false;
    return true;
}

int malformed(char const * what)
{
    set_error("sdslgpu_load_sdsl: malformed or truncated %s blob", what);
    return SDSLGPU_EINVAL;
}

int load_bv(sdslgpu_handle * h, Reader & r, cudaStream_t s)
{
    IntVec bv;
    if (!read_iv(r, bv) || bv.width != 1)
        return malformed("bit_vector");
    return bv_build(h->pool, h->bv, h->flags, bv.words.data(), false, bv.bits, s);
}

int load_rrr_image(DevicePool & pool, RrrImage & im, Reader & r, cudaStream_t s)
{
    im.size = r.u64();
    IntVec bt, btnr, btnrp, rank, inv;
    if (!read_iv(r, bt) || !read_iv(r, btnr) || !read_iv(r, btnrp) || !read_iv(r, rank) || !read_iv(r, inv) || bt.width != 6 || btnr.width != 1 ||
        inv.width != 1 || rank.width == 0 || rank.width > 64 || btnrp.width == 0 || btnrp.width > 64 || rank.size() == 0)
        return malformed("rrr_vector<63>");
    im.nblocks = bt.size();
    im.nsuper = btnrp.size();
    if (im.nblocks != (im.size + 63) / 63 || im.nsuper != (im.nblocks + 31) / 32 || rank.size() < im.nsuper || inv.bits < im.nsuper)
        return malformed("rrr_vector<63> (inconsistent sizes; only t_bs = 63, t_k = 32 is supported)");
    im.ones = rank.get(rank.size() - 1);
    im.btnr_bits = btnr.bits;
    SG_TRY(rrr_upload_tables(pool, im, s));
    std::vector<uint64_t> rk(im.nsuper), bp(im.nsuper);
    std::vector<uint8_t> iv(im.nsuper);
    for (uint64_t g = 0; g < im.nsuper; ++g)
    {
        rk[g] = rank.get(g);
        bp[g] = btnrp.get(g);
        iv[g] = (uint8_t)inv.get(g);
        // offsets into m_btnr must be non-decreasing and inside it; the sampled ranks non-decreasing and <= size
        if (bp[g] > btnr.bits || (g && bp[g] < bp[g - 1]) || rk[g] > im.size || (g && rk[g] < rk[g - 1]))
            return malformed("rrr_vector<63> (superblock samples out of range)");
    }
    if (im.ones > im.size || (im.nsuper && im.ones < rk[im.nsuper - 1]))
        return malformed("rrr_vector<63> (rank samples)");
    // the blob does not store the exact number of offset bits (m_btnr is padded to >= 64 bits); only its top bit
    // matters (it fixes the width of m_btnrp when serialising back), and m_btnrp's width preserves that
    uint64_t total_bits_hint = btnrp.width ? (1ull << (btnrp.width - 1)) : 0;
    bt.words.resize(bt.words.size() + 2, 0);
    SG_TRY(rrr_records_from_sdsl(pool, im, bt.words.data(), im.nblocks, rk, bp, iv, total_bits_hint, s));
    uint64_t nrw = ((btnr.bits + 63) >> 6) + 2;
    std::vector<uint64_t> nrp(nrw, 0);
    std::memcpy(nrp.data(), btnr.words.data(), std::min<uint64_t>(btnr.words.size(), nrw) * 8);
    SG_TRY(pool.alloc_t(&im.btnr, nrw));
    SG_CUDA(cudaMemcpyAsync(im.btnr, nrp.data(), nrw * 8, cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return rrr_build_hints(pool, im, s);
}

int load_rrr(sdslgpu_handle * h, Reader & r, cudaStream_t s)
{
    return load_rrr_image(h->pool, h->rrr, r, s);
}

int load_sd(sdslgpu_handle * h, Reader & r, cudaStream_t s)
{
    SdImage & d = h->sd;
    d.size = r.u64();
    d.wl = r.u8();
    IntVec low, high;
    if (!read_iv(r, low) || !read_iv(r, high) || high.width != 1 || low.width != d.wl || d.wl == 0)
        return malformed("sd_vector");
    d.m = low.size();
    d.high_bits = high.bits;
    {
        uint64_t ones = 0;
        for (uint64_t k = 0; k < (high.bits >> 6); ++k)
            ones += (uint64_t)__builtin_popcountll(high.words[k]);
        if (high.bits & 63)
            ones += (uint64_t)__builtin_popcountll(high.words[high.bits >> 6] & ((1ull << (high.bits & 63)) - 1));
        // sd_vector.hpp:236-253: one 1 per element plus one 0 per bucket of 2^wl positions
        if (d.wl > 63 || ones != d.m || high.bits < d.m || ((high.bits - d.m) << d.wl) < d.size)
            return malformed("sd_vector (m_high does not match m_low / size)");
    }
    d.low_words = low.words.size() + 1;
    std::vector<uint64_t> lw(d.low_words, 0);
    std::memcpy(lw.data(), low.words.data(), low.words.size() * 8);
    SG_TRY(h->pool.alloc_t(&d.low, d.low_words));
    SG_CUDA(cudaMemcpyAsync(d.low, lw.data(), d.low_words * 8, cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    // the two select_support_mcl blobs that follow are not needed
    SG_TRY(bv_build(h->pool, d.high, h->flags & SDSLGPU_F_SDSL_LAYOUT, high.words.data(), false, high.bits, s));
    return sd_build_select0_samples(h, s);
}

int load_wt_huff(sdslgpu_handle * h, Reader & r, cudaStream_t s)
{
    uint64_t size = r.u64(), sigma = r.u64();
    IntVec bv;
    bool const over_rrr = (h->flags & SDSLGPU_F_RRR_BV) != 0;
    if (over_rrr)
    { // wt_huff<rrr_vector<63>>: the rrr_vector, then its rank/select supports which serialise to nothing
        h->wt.use_rrr = true;
        SG_TRY(load_rrr_image(h->pool, h->wt.rrr, r, s));
    }
    else if (!read_iv(r, bv) || bv.width != 1 || !skip_iv(r))
        return malformed("wt_huff");
    else if (!(h->flags & SDSLGPU_F_V5_SCAN) && (!skip_select_mcl(r) || !skip_select_mcl(r)))
        return malformed("wt_huff (select supports)"); // with rank_support_v5 + select_support_scan nothing follows the table
    uint64_t nn = r.u64();
    if (!r.ok || nn > 511 || !r.need(nn * 22 + 512 + 2048))
        return malformed("wt_huff (byte_tree)");
    WtTree tree;
    std::memset(&tree, 0, sizeof(tree));
    tree.nnodes = (uint32_t)nn;
    for (uint64_t v = 0; v < nn; ++v)
    {
        tree.bv_pos[v] = r.u64();
        tree.bv_pos_rank[v] = r.u64();
        tree.parent[v] = r.u16();
        tree.child[v][0] = r.u16();
        tree.child[v][1] = r.u16();
    }
    for (int c = 0; c < 256; ++c)
        tree.c_to_leaf[c] = r.u16();
    for (int c = 0; c < 256; ++c)
        tree.path[c] = r.u64();
    if (!r.ok)
        return malformed("wt_huff (byte_tree)");
    {
        uint64_t const bits = over_rrr ? h->wt.rrr.size : bv.bits;
        for (uint64_t v = 0; v < nn; ++v)
        {
            bool const idx_ok = (tree.parent[v] == 0xFFFF || tree.parent[v] < nn) && (tree.child[v][0] == 0xFFFF || tree.child[v][0] < nn) &&
                                (tree.child[v][1] == 0xFFFF || tree.child[v][1] < nn);
            bool const inner = tree.child[v][0] != 0xFFFF; // a leaf keeps its symbol in bv_pos_rank (wt_helper.hpp:279-283)
            if (!idx_ok || tree.bv_pos[v] > bits || (inner && tree.bv_pos_rank[v] > tree.bv_pos[v]) || (v && tree.bv_pos[v] < tree.bv_pos[v - 1]))
                return malformed("wt_huff (byte_tree node out of range)");
        }
        for (int c = 0; c < 256 && size; ++c)
            if (tree.c_to_leaf[c] != 0xFFFF && tree.c_to_leaf[c] >= nn)
                return malformed("wt_huff (c_to_leaf out of range)");
        for (int c = 0; c < 256 && size; ++c)
            if ((tree.path[c] >> 56) > 56)
                return malformed("wt_huff (code longer than 56 bits)");
    }
    if (size == 0)
        for (int c = 0; c < 256; ++c)
            tree.c_to_leaf[c] = 0xFFFF; // an empty reference tree serialises uninitialised tables
    if (over_rrr)
        return wt_huff_finish(h, size, sigma, tree, s);
    return wt_huff_upload(h, size, sigma, tree, bv.words.data(), bv.bits, s);
}

int load_wt_int(sdslgpu_handle * h, Reader & r, cudaStream_t s)
{
    WtIntImage & w = h->wti;
    w.size = r.u64();
    w.sigma = r.u64();
    IntVec tree;
    if (!read_iv(r, tree) || tree.width != 1 || !skip_iv(r) || !skip_select_mcl(r) || !skip_select_mcl(r))
        return malformed("wt_int");
    w.max_level = r.u32();
    if (!r.ok || w.max_level > 64 || (uint64_t)w.max_level * w.size != tree.bits || (w.size && w.max_level == 0))
        return malformed("wt_int (level count)");
    return bv_build(h->pool, w.tree, h->flags & ~SDSLGPU_F_NO_SELECT, tree.words.data(), false, tree.bits, s);
}

int load_csa(sdslgpu_handle * h, Reader & r, uint32_t sa_dens, uint32_t isa_dens, cudaStream_t s)
{
    SG_TRY(load_wt_huff(h, r, s));
    IntVec sa, isa, c2c, comp2char, C;
    if (!read_iv(r, sa) || !read_iv(r, isa) || !read_iv(r, c2c) || !read_iv(r, comp2char) || !read_iv(r, C) || c2c.width != 8 || c2c.size() != 256 ||
        C.width != 64 || comp2char.width != 8 || sa.width == 0 || sa.width > 64 || (isa.bits && (isa.width == 0 || isa.width > 64)))
        return malformed("csa_wt");
    uint16_t sigma = r.u16();
    if (!r.ok || C.size() != (uint64_t)sigma + 1 || sigma > 256 || comp2char.size() < sigma)
        return malformed("csa_wt (alphabet)");
    CsaImage & c = h->csa;
    c.n = h->wt.size;
    c.sa_dens = sa_dens;
    if (sa.size() != (c.n + sa_dens - 1) / sa_dens)
    {
        set_error("sdslgpu_load_sdsl: %llu SA samples do not match size %llu at density %u (pass the index's t_dens)", (unsigned long long)sa.size(),
                  (unsigned long long)c.n, sa_dens);
        return SDSLGPU_EINVAL;
    }
    FmTables & tab = c.host_tab;
    std::memset(&tab, 0, sizeof(tab));
    for (int k = 0; k < 256; ++k)
        tab.char2comp[k] = (uint8_t)c2c.get(k);
    for (uint32_t k = 0; k < sigma && k < 256; ++k)
        tab.comp2char[k] = (uint8_t)comp2char.get(k);
    for (uint32_t k = 0; k <= sigma; ++k)
        tab.C[k] = C.get(k);
    tab.sigma = sigma;
    std::vector<uint64_t> samples(sa.size() + 1, 0);
    for (uint64_t k = 0; k < sa.size(); ++k)
        samples[k] = sa.get(k);
    c.nsamples = sa.size();
    SG_TRY(h->pool.alloc_t(&c.samples, c.nsamples + 1));
    SG_CUDA(cudaMemcpyAsync(c.samples, samples.data(), c.nsamples * 8, cudaMemcpyHostToDevice, s));
    SG_TRY(h->pool.alloc_t(&c.tab, 1));
    SG_CUDA(cudaMemcpyAsync(c.tab, &c.host_tab, sizeof(FmTables), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    // ISA samples: isa.size() == (n - 1) / t_inv_dens + 1 (csa_sampling_strategy.hpp:762-763).  t_inv_dens is a
    // template parameter of the reference and is not stored: sdslgpu_load_sdsl_ex takes it; without it (0) the default
    // 64, then the powers of two are tried, and a count no candidate reproduces is an error, never a guess.
    std::vector<uint64_t> isav(isa.size() + 1, 0);
    for (uint64_t k = 0; k < isa.size(); ++k)
        isav[k] = isa.get(k);
    if (c.n == 0 || isa.size() == 0)
        return malformed("csa_wt (no ISA samples)");
    auto isa_count = [&](uint64_t d) { return (c.n - 1) / d + 1; };
    if (isa_dens)
    {
        if (isa_count(isa_dens) != isa.size())
        {
            set_error("sdslgpu_load_sdsl: %llu ISA samples do not match size %llu at t_inv_dens %u", (unsigned long long)isa.size(), (unsigned long long)c.n,
                      isa_dens);
            return SDSLGPU_EINVAL;
        }
        c.isa_dens = isa_dens;
    }
    else
    {
        uint64_t d = 64;
        if (isa_count(d) != isa.size())
            for (d = 1; d <= (1ull << 32) && isa_count(d) != isa.size(); d <<= 1)
            {}
        if (d > (1ull << 32))
        {
            set_error("sdslgpu_load_sdsl: cannot infer t_inv_dens from %llu ISA samples (size %llu): not a power of two — pass it to sdslgpu_load_sdsl_ex",
                      (unsigned long long)isa.size(), (unsigned long long)c.n);
            return SDSLGPU_EINVAL;
        }
        // several densities can give the same count when n is small; any of them indexes the samples that exist
        c.isa_dens = (uint32_t)(d > 0xFFFFFFFFull ? 0x80000000u : d);
    }
    SG_TRY(csa_upload_isa(h, isav.data(), isa.size(), s));
    // the searches' one-hot occurrence bitmaps, decoded from the ingested tree (fm16.cu) unless a compact index was asked for
    if (!(h->flags & SDSLGPU_F_COMPACT) && !h->wt.use_rrr)
        SG_TRY(occ16_build_from_wt(h, s));
    return SDSLGPU_OK;
}

} // namespace

// the select supports behind sd_vector's m_high (sd_vector.hpp:434-435) and the end of the other structures
static int finish_tail(sdslgpu_handle * h, Reader & r)
{
    if (h->kind == SDSLGPU_KIND_SD && (!skip_select_mcl(r) || !skip_select_mcl(r)))
        r.ok = true, r.pos = r.n; // older callers pass only size, wl, low, high: the supports are optional on ingest
    return SDSLGPU_OK;
}

int load_sdsl_blob(sdslgpu_handle * h, uint8_t const * blob, uint64_t nbytes, uint32_t sa_dens, uint32_t isa_dens, uint64_t * consumed, cudaStream_t s)
{
    Reader r{blob, nbytes};
    int st = SDSLGPU_EINVAL;
    switch (h->kind)
    {
    case SDSLGPU_KIND_BV:
        st = load_bv(h, r, s);
        break;
    case SDSLGPU_KIND_RRR63:
        st = load_rrr(h, r, s);
        break;
    case SDSLGPU_KIND_SD:
        st = load_sd(h, r, s);
        if (st == SDSLGPU_OK)
            st = finish_tail(h, r);
        break;
    case SDSLGPU_KIND_WT_HUFF:
        st = load_wt_huff(h, r, s);
        break;
    case SDSLGPU_KIND_WT_INT:
        st = load_wt_int(h, r, s);
        break;
    case SDSLGPU_KIND_CSA_WT:
        st = load_csa(h, r, sa_dens ? sa_dens : 32, isa_dens, s);
        break;
    default:
        set_error("sdslgpu_load_sdsl: unknown kind %d", h->kind);
        return SDSLGPU_EINVAL;
    }
    if (st == SDSLGPU_OK && consumed)
        *consumed = r.pos; // bytes of the structure itself: a stream holding several structures continues here
    return st;
}

} // namespace sdslgpu
