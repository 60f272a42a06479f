// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" wrappers around the UNMODIFIED reference (xxsds/sdsl-lite 3.0.5) headers,
// compiled from where they lie (-I/root/reference/include) into oracle/_ref/libsdslref.so by
// oracle/Makefile.  No reference source is copied: this file only *calls* the reference's public
// API (bit_vector, rank_support_v, select_support_mcl, rrr_vector<63>, sd_vector<>, wt_huff<>,
// wt_int<>, csa_wt<wt_huff<>>, count, locate, serialize).
//
// Used by: tests/ (as the parity pin for oracle/oracle.c and the checker for the CUDA path) and by
// bench.py's cpu_baseline / `--impl reference` arm.  Never linked or loaded by the product library.
//
// Every batch entry point takes `threads` (>=1): queries are split statically across std::threads
// (legal: SDSL query methods are const and stateless).

#include <sdsl/bit_vectors.hpp>
#include <sdsl/suffix_arrays.hpp>
#include <sdsl/wavelet_trees.hpp>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

using namespace sdsl;

namespace
{

template <class F>
void parallel_for(uint64_t n, int threads, F f)
{
    if (threads <= 1 || n < 1024)
    {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    uint64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t)
    {
        uint64_t lo = std::min<uint64_t>(n, t * chunk), hi = std::min<uint64_t>(n, lo + chunk);
        if (lo < hi)
            pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto & th : pool)
        th.join();
}

template <class T>
uint64_t serialize_to(T const & obj, uint8_t * buf, uint64_t cap)
{
    std::ostringstream os(std::ios::binary);
    obj.serialize(os);
    std::string s = os.str();
    if (buf != nullptr && s.size() <= cap)
        std::memcpy(buf, s.data(), s.size());
    return s.size();
}

bit_vector make_bv(uint64_t const * words, uint64_t nbits)
{
    bit_vector bv(nbits, 0);
    uint64_t nw = (nbits + 63) >> 6;
    if (nw)
        std::memcpy(bv.data(), words, nw * 8);
    return bv;
}

struct ref_bv
{
    bit_vector bv;
    rank_support_v<1> r1;
    rank_support_v<0> r0;
    select_support_mcl<1> s1;
    select_support_mcl<0> s0;
    bool has_select = false;
    // two-bit patterns, built on first use: codes 2 = "10", 3 = "01", 4 = "00", 5 = "11"
    // (rank_support_test.cpp:44-47, select_support_test.cpp:39-42 instantiate exactly these)
    std::unique_ptr<rank_support_v<10, 2>> r10;
    std::unique_ptr<rank_support_v<01, 2>> r01;
    std::unique_ptr<rank_support_v<00, 2>> r00;
    std::unique_ptr<rank_support_v<11, 2>> r11;
    std::unique_ptr<select_support_mcl<10, 2>> s10;
    std::unique_ptr<select_support_mcl<01, 2>> s01;
    std::unique_ptr<select_support_mcl<00, 2>> s00;
    std::unique_ptr<select_support_mcl<11, 2>> s11;
};

template <class S, class F>
void run_pat(std::unique_ptr<S> & sup, bit_vector const * bv, uint64_t n, int threads, F f)
{
    if (!sup)
        sup = std::make_unique<S>(bv);
    S const * p = sup.get();
    parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k)
            f(p, k);
    });
}

struct ref_rrr
{
    rrr_vector<63> v;
    rrr_vector<63>::rank_1_type r1;
    rrr_vector<63>::rank_0_type r0;
    rrr_vector<63>::select_1_type s1;
    rrr_vector<63>::select_0_type s0;
};

struct ref_sd
{
    sd_vector<> v;
    sd_vector<>::rank_1_type r1;
    sd_vector<>::rank_0_type r0;
    sd_vector<>::select_1_type s1;
    sd_vector<>::select_0_type s0;
};

struct ref_wt_huff
{
    wt_huff<> wt;
};
struct ref_wt_int
{
    wt_int<> wt;
};
struct ref_csa
{
    csa_wt<wt_huff<>> csa; // defaults: t_dens=32, t_inv_dens=64, sa_order_sa_sampling, isa_sampling, byte_alphabet
};

} // namespace

extern "C"
{

    // ---------------------------------------------------------------- plain bit_vector
    void * ref_bv_create(uint64_t const * words, uint64_t nbits, int with_select)
    {
        auto * h = new ref_bv;
        h->bv = make_bv(words, nbits);
        h->r1 = rank_support_v<1>(&h->bv);
        h->r0 = rank_support_v<0>(&h->bv);
        if (with_select)
        {
            h->s1 = select_support_mcl<1>(&h->bv);
            h->s0 = select_support_mcl<0>(&h->bv);
            h->has_select = true;
        }
        return h;
    }
    void ref_bv_free(void * p)
    {
        delete static_cast<ref_bv *>(p);
    }
    void ref_bv_rank(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_bv *>(p);
        auto rk = [=](auto const * s, uint64_t k) {
            out[k] = s->rank(idx[k]);
        };
        if (pattern == 2)
            run_pat(h->r10, &h->bv, n, threads, rk);
        else if (pattern == 3)
            run_pat(h->r01, &h->bv, n, threads, rk);
        else if (pattern == 4)
            run_pat(h->r00, &h->bv, n, threads, rk);
        else if (pattern == 5)
            run_pat(h->r11, &h->bv, n, threads, rk);
        else if (pattern)
            parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
                for (uint64_t k = lo; k < hi; ++k)
                    out[k] = h->r1.rank(idx[k]);
            });
        else
            parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
                for (uint64_t k = lo; k < hi; ++k)
                    out[k] = h->r0.rank(idx[k]);
            });
    }
    void ref_bv_select(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_bv *>(p);
        auto sl = [=](auto const * s, uint64_t k) {
            out[k] = s->select(idx[k]);
        };
        if (pattern == 2)
            run_pat(h->s10, &h->bv, n, threads, sl);
        else if (pattern == 3)
            run_pat(h->s01, &h->bv, n, threads, sl);
        else if (pattern == 4)
            run_pat(h->s00, &h->bv, n, threads, sl);
        else if (pattern == 5)
            run_pat(h->s11, &h->bv, n, threads, sl);
        else if (pattern)
            parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
                for (uint64_t k = lo; k < hi; ++k)
                    out[k] = h->s1.select(idx[k]);
            });
        else
            parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
                for (uint64_t k = lo; k < hi; ++k)
                    out[k] = h->s0.select(idx[k]);
            });
    }
    // what: 0 = bit_vector, 1 = rank_support_v<1>, 2 = rank_support_v<0>, 3 = select_support_mcl<1>, 4 = <0>,
    //       5 = rank_support_v5<1>, 6 = rank_support_v5<0>
    uint64_t ref_bv_serialize(void * p, int what, uint8_t * buf, uint64_t cap)
    {
        auto * h = static_cast<ref_bv *>(p);
        switch (what)
        {
        case 5:
            return serialize_to(rank_support_v5<1>(&h->bv), buf, cap);
        case 6:
            return serialize_to(rank_support_v5<0>(&h->bv), buf, cap);
        case 0:
            return serialize_to(h->bv, buf, cap);
        case 1:
            return serialize_to(h->r1, buf, cap);
        case 2:
            return serialize_to(h->r0, buf, cap);
        case 3:
            return serialize_to(h->s1, buf, cap);
        case 4:
            return serialize_to(h->s0, buf, cap);
        }
        return 0;
    }

    // ---------------------------------------------------------------- rrr_vector<63>
    void * ref_rrr_create(uint64_t const * words, uint64_t nbits)
    {
        auto * h = new ref_rrr;
        bit_vector bv = make_bv(words, nbits);
        h->v = rrr_vector<63>(bv);
        h->r1.set_vector(&h->v);
        h->r0.set_vector(&h->v);
        h->s1.set_vector(&h->v);
        h->s0.set_vector(&h->v);
        return h;
    }
    void ref_rrr_free(void * p)
    {
        delete static_cast<ref_rrr *>(p);
    }
    void ref_rrr_rank(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_rrr *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = pattern ? h->r1.rank(idx[k]) : h->r0.rank(idx[k]);
        });
    }
    void ref_rrr_select(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_rrr *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = pattern ? h->s1.select(idx[k]) : h->s0.select(idx[k]);
        });
    }
    void ref_rrr_access(void * p, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_rrr *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->v[idx[k]];
        });
    }
    uint64_t ref_rrr_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_rrr *>(p)->v, buf, cap);
    }

    // ---------------------------------------------------------------- sd_vector<>
    void * ref_sd_create(uint64_t const * words, uint64_t nbits)
    {
        auto * h = new ref_sd;
        bit_vector bv = make_bv(words, nbits);
        h->v = sd_vector<>(bv);
        h->r1.set_vector(&h->v);
        h->r0.set_vector(&h->v);
        h->s1.set_vector(&h->v);
        h->s0.set_vector(&h->v);
        return h;
    }
    void ref_sd_free(void * p)
    {
        delete static_cast<ref_sd *>(p);
    }
    void ref_sd_rank(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_sd *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = pattern ? h->r1.rank(idx[k]) : h->r0.rank(idx[k]);
        });
    }
    void ref_sd_select(void * p, int pattern, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_sd *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = pattern ? h->s1.select(idx[k]) : h->s0.select(idx[k]);
        });
    }
    void ref_sd_access(void * p, uint64_t const * idx, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_sd *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->v[idx[k]];
        });
    }
    uint64_t ref_sd_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_sd *>(p)->v, buf, cap);
    }

    // ---------------------------------------------------------------- wt_huff<>
    void * ref_wt_huff_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new ref_wt_huff;
        int_vector<8> t(n);
        if (n)
            std::memcpy(t.data(), text, n);
        construct_im(h->wt, t, 0); // 0: the ram file is a serialized int_vector<8>
        return h;
    }
    void ref_wt_huff_free(void * p)
    {
        delete static_cast<ref_wt_huff *>(p);
    }
    uint64_t ref_wt_huff_size(void * p)
    {
        return static_cast<ref_wt_huff *>(p)->wt.size();
    }
    uint64_t ref_wt_huff_sigma(void * p)
    {
        return static_cast<ref_wt_huff *>(p)->wt.sigma;
    }
    void ref_wt_huff_rank(void * p, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_wt_huff *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->wt.rank(i[k], c[k]);
        });
    }
    void ref_wt_huff_select(void * p, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_wt_huff *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->wt.select(i[k], c[k]);
        });
    }
    // inverse_select: sym_out[k] = wt[i], rank_out[k] = rank(i, wt[i]); rank_out may be null (= operator[])
    void ref_wt_huff_access(void * p, uint64_t const * i, uint64_t n, uint64_t * sym_out, uint64_t * rank_out, int threads)
    {
        auto * h = static_cast<ref_wt_huff *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
            {
                if (rank_out)
                {
                    auto rc = h->wt.inverse_select(i[k]);
                    rank_out[k] = rc.first;
                    sym_out[k] = rc.second;
                }
                else
                    sym_out[k] = h->wt[i[k]];
            }
        });
    }
    uint64_t ref_wt_huff_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_wt_huff *>(p)->wt, buf, cap);
    }

    // ---------------------------------------------------------------- wt_int<>
    void * ref_wt_int_create(uint64_t const * seq, uint64_t n)
    {
        auto * h = new ref_wt_int;
        int_vector<> t(n, 0, 64);
        for (uint64_t k = 0; k < n; ++k)
            t[k] = seq[k];
        util::bit_compress(t);
        construct_im(h->wt, t, 0);
        return h;
    }
    void ref_wt_int_free(void * p)
    {
        delete static_cast<ref_wt_int *>(p);
    }
    uint64_t ref_wt_int_sigma(void * p)
    {
        return static_cast<ref_wt_int *>(p)->wt.sigma;
    }
    uint64_t ref_wt_int_max_level(void * p)
    {
        return static_cast<ref_wt_int *>(p)->wt.max_level;
    }
    void ref_wt_int_rank(void * p, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_wt_int *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->wt.rank(i[k], c[k]);
        });
    }
    void ref_wt_int_select(void * p, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_wt_int *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->wt.select(i[k], c[k]);
        });
    }
    void ref_wt_int_access(void * p, uint64_t const * i, uint64_t n, uint64_t * sym_out, uint64_t * rank_out, int threads)
    {
        auto * h = static_cast<ref_wt_int *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
            {
                if (rank_out)
                {
                    auto rc = h->wt.inverse_select(i[k]);
                    rank_out[k] = rc.first;
                    sym_out[k] = rc.second;
                }
                else
                    sym_out[k] = h->wt[i[k]];
            }
        });
    }
    uint64_t ref_wt_int_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_wt_int *>(p)->wt, buf, cap);
    }

    // ---------------------------------------------------------------- csa_wt<wt_huff<>>
    // text must be zero-free (construct.hpp:34-46); the sentinel is appended by the reference.
    void * ref_csa_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new ref_csa;
        std::string s(reinterpret_cast<char const *>(text), n);
        try
        {
            construct_im(h->csa, s, 1);
        }
        catch (...)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    void * ref_csa_load(uint8_t const * blob, uint64_t nbytes)
    {
        auto * h = new ref_csa;
        std::istringstream is(std::string(reinterpret_cast<char const *>(blob), nbytes), std::ios::binary);
        std::istream & in = is; // bind to the istream overload, not the cereal archive template
        h->csa.load(in);
        return h;
    }
    void ref_csa_free(void * p)
    {
        delete static_cast<ref_csa *>(p);
    }
    uint64_t ref_csa_size(void * p)
    {
        return static_cast<ref_csa *>(p)->csa.size();
    }
    uint64_t ref_csa_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_csa *>(p)->csa, buf, cap);
    }
    // patterns in CSR form: pattern k = pats[off[k] .. off[k+1])
    void ref_csa_count(void * p,
                       uint8_t const * pats,
                       uint64_t const * off,
                       uint64_t n,
                       uint64_t * cnt_out,
                       uint64_t * l_out,
                       int threads)
    {
        auto * h = static_cast<ref_csa *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
            {
                if (l_out == nullptr)
                    cnt_out[k] = count(h->csa, pats + off[k], pats + off[k + 1]);
                else
                {
                    uint64_t l = 0, r = 0;
                    uint64_t m = off[k + 1] - off[k];
                    if (m > h->csa.size())
                    {
                        cnt_out[k] = 0;
                        l_out[k] = 0;
                    }
                    else
                    {
                        cnt_out[k] =
                            backward_search(h->csa, 0, h->csa.size() - 1, pats + off[k], pats + off[k + 1], l, r);
                        l_out[k] = l;
                    }
                }
            }
        });
    }
    // locate: two calls. First with occ_out == nullptr fills occ_off[0..n] (exclusive prefix sums of counts);
    // then with a buffer of occ_off[n] entries fills the occurrences in SA order per pattern.
    void ref_csa_locate(void * p,
                        uint8_t const * pats,
                        uint64_t const * off,
                        uint64_t n,
                        uint64_t * occ_off,
                        uint64_t * occ_out,
                        int threads)
    {
        auto * h = static_cast<ref_csa *>(p);
        if (occ_out == nullptr)
        {
            std::vector<uint64_t> cnt(n);
            parallel_for(n, threads, [&](uint64_t lo, uint64_t hi) {
                for (uint64_t k = lo; k < hi; ++k)
                    cnt[k] = count(h->csa, pats + off[k], pats + off[k + 1]);
            });
            occ_off[0] = 0;
            for (uint64_t k = 0; k < n; ++k)
                occ_off[k + 1] = occ_off[k] + cnt[k];
            return;
        }
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
            {
                auto occ = locate(h->csa, pats + off[k], pats + off[k + 1]);
                for (uint64_t j = 0; j < occ.size(); ++j)
                    occ_out[occ_off[k] + j] = occ[j];
            }
        });
    }
    // SA access csa[i]
    void ref_csa_sa(void * p, uint64_t const * i, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_csa *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->csa[i[k]];
        });
    }
    // bwt.rank(i, c) on the CSA's wavelet tree (raw chars)
    void ref_csa_bwt_rank(void * p, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_csa *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->csa.bwt.rank(i[k], c[k]);
        });
    }
    // extract text[lo..hi] (inclusive), out has hi-lo+1 bytes
    void ref_csa_extract(void * p, uint64_t lo, uint64_t hi, uint8_t * out)
    {
        auto * h = static_cast<ref_csa *>(p);
        auto s = extract(h->csa, lo, hi);
        std::memcpy(out, s.data(), s.size());
    }

    // ---------------------------------------------------------------- wt_huff<rrr_vector<63>> and the CSA over it
    // (the reference's FM_HUFF_RRR63 benchmark index, benchmark/indexing_count/index.config:10)
    struct ref_wt_huff_rrr
    {
        wt_huff<rrr_vector<63>> wt;
    };
    struct ref_csa_rrr
    {
        csa_wt<wt_huff<rrr_vector<63>>> csa;
    };
    void * ref_wt_huff_rrr_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new ref_wt_huff_rrr;
        int_vector<8> t(n);
        if (n)
            std::memcpy(t.data(), text, n);
        construct_im(h->wt, t, 0);
        return h;
    }
    void ref_wt_huff_rrr_free(void * p)
    {
        delete static_cast<ref_wt_huff_rrr *>(p);
    }
    void ref_wt_huff_rrr_rank(void * p, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, int threads)
    {
        auto * h = static_cast<ref_wt_huff_rrr *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                out[k] = h->wt.rank(i[k], c[k]);
        });
    }
    uint64_t ref_wt_huff_rrr_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_wt_huff_rrr *>(p)->wt, buf, cap);
    }
    void * ref_csa_rrr_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new ref_csa_rrr;
        std::string s(reinterpret_cast<char const *>(text), n);
        try
        {
            construct_im(h->csa, s, 1);
        }
        catch (...)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    void ref_csa_rrr_free(void * p)
    {
        delete static_cast<ref_csa_rrr *>(p);
    }
    uint64_t ref_csa_rrr_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(static_cast<ref_csa_rrr *>(p)->csa, buf, cap);
    }
    void ref_csa_rrr_count(void * p, uint8_t const * pats, uint64_t const * off, uint64_t n, uint64_t * cnt_out, int threads)
    {
        auto * h = static_cast<ref_csa_rrr *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                cnt_out[k] = count(h->csa, pats + off[k], pats + off[k + 1]);
        });
    }

    // ---------------------------------------------------------------- the reference's own count-benchmark index
    // FM_HUFF of benchmark/indexing_count/index.config:8 — rank_support_v5 + select_support_scan inside the tree,
    // one SA / ISA sample per 2^20 entries
    typedef wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>> wt_huff_v5;
    typedef csa_wt<wt_huff_v5, 1 << 20, 1 << 20> fm_huff;
    void * ref_wt_huff_v5_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new wt_huff_v5;
        int_vector<8> t(n);
        if (n)
            std::memcpy(t.data(), text, n);
        construct_im(*h, t, 0);
        return h;
    }
    void ref_wt_huff_v5_free(void * p)
    {
        delete static_cast<wt_huff_v5 *>(p);
    }
    uint64_t ref_wt_huff_v5_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(*static_cast<wt_huff_v5 *>(p), buf, cap);
    }
    void * ref_fm_huff_create(uint8_t const * text, uint64_t n)
    {
        auto * h = new fm_huff;
        std::string s(reinterpret_cast<char const *>(text), n);
        try
        {
            construct_im(*h, s, 1);
        }
        catch (...)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    void * ref_fm_huff_load(uint8_t const * blob, uint64_t nbytes)
    {
        auto * h = new fm_huff;
        std::istringstream is(std::string(reinterpret_cast<char const *>(blob), nbytes), std::ios::binary);
        std::istream & in = is;
        h->load(in);
        return h;
    }
    void ref_fm_huff_free(void * p)
    {
        delete static_cast<fm_huff *>(p);
    }
    uint64_t ref_fm_huff_serialize(void * p, uint8_t * buf, uint64_t cap)
    {
        return serialize_to(*static_cast<fm_huff *>(p), buf, cap);
    }
    void ref_fm_huff_count(void * p, uint8_t const * pats, uint64_t const * off, uint64_t n, uint64_t * cnt_out, int threads)
    {
        auto * h = static_cast<fm_huff *>(p);
        parallel_for(n, threads, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t k = lo; k < hi; ++k)
                cnt_out[k] = count(*h, pats + off[k], pats + off[k + 1]);
        });
    }

    char const * ref_version()
    {
        return "sdsl-lite 3.0.5 (reference headers, unmodified) via oracle/ref_driver.cpp";
    }

} // extern "C"
