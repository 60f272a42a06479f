// tests/cpp/shim_test.cpp — the SDSL-shaped C++ classes (sdsl-lite_b200/include/sdsl_b200.hpp) exercised the way
// the reference's own typed tests exercise sdsl:: (test/rank_support_test.cpp:109-126, select_support_test.cpp:85-104,
// wt_byte_test.cpp:134-203, csa_byte_test.cpp:84-110): every result is compared with a naive scan.
// Built by tests/test_cpp_shim.py:  g++ -std=c++17 shim_test.cpp -L<pkg> -lsdslgpu ; needs a GPU to run.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>
#include <string>

#include "../../sdsl-lite_b200/include/sdsl_b200.hpp"

using namespace sdsl_b200;

#define EXPECT(cond)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);                                             \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

template <class t_vec, class t_rank1, class t_rank0, class t_sel1, class t_sel0>
void check_bitvector(bit_vector const & plain, t_vec const & v)
{
    t_rank1 r1(&v);
    t_rank0 r0(&v);
    t_sel1 s1(&v);
    t_sel0 s0(&v);
    std::vector<uint64_t> idx(plain.size() + 1);
    for (uint64_t j = 0; j <= plain.size(); ++j)
        idx[j] = j;
    auto got1 = r1.rank(idx), got0 = r0.rank(idx); // batch overloads
    uint64_t ones = 0;
    std::vector<uint64_t> pos1, pos0;
    for (uint64_t j = 0; j < plain.size(); ++j)
    {
        EXPECT(got1[j] == ones && got0[j] == j - ones);
        (plain[j] ? pos1 : pos0).push_back(j);
        ones += plain[j];
    }
    EXPECT(got1[plain.size()] == ones);
    EXPECT(r1.rank(plain.size() / 2) == got1[plain.size() / 2] && r1(plain.size()) == ones); // scalar drop-in calls
    std::vector<uint64_t> k1(pos1.size()), k0(pos0.size());
    for (size_t k = 0; k < k1.size(); ++k)
        k1[k] = k + 1;
    for (size_t k = 0; k < k0.size(); ++k)
        k0[k] = k + 1;
    EXPECT(s1.select(k1) == pos1);
    EXPECT(s0.select(k0) == pos0);
    if (!pos1.empty())
        EXPECT(s1.select(1) == pos1[0] && s1(pos1.size()) == pos1.back());
}

// rank_support_v<pat, 2> / select_support_mcl<pat, 2> the way rank_support_test.cpp:44-47 and
// select_support_test.cpp:39-42 instantiate them; an occurrence sits at the position of its second bit
template <uint8_t t_b>
void check_pattern(bit_vector const & bv, bool prev, bool cur)
{
    rank_support_v<t_b, 2> rs(&bv);
    select_support_mcl<t_b, 2> ss(&bv);
    std::vector<uint64_t> idx(bv.size() + 1), pos;
    for (uint64_t j = 0; j <= bv.size(); ++j)
        idx[j] = j;
    auto got = rs.rank(idx);
    for (uint64_t j = 0; j < bv.size(); ++j)
    {
        EXPECT(got[j] == pos.size());
        if (j > 0 && bv[j - 1] == prev && bv[j] == cur)
            pos.push_back(j);
    }
    EXPECT(got[bv.size()] == pos.size());
    std::vector<uint64_t> k(pos.size());
    for (size_t q = 0; q < k.size(); ++q)
        k[q] = q + 1;
    EXPECT(ss.select(k) == pos);
}

static std::ostream & none_sink()
{
    static std::stringstream s;
    s.str("");
    return s;
}

int main(int argc, char ** argv)
{
    if (argc > 1 && std::string(argv[1]) == "--host-only")
    { // the parts of the shim that never touch the device: bit_vector storage and its serialised form
        bit_vector bv(130);
        bv.set(0, true);
        bv.set(64, true);
        bv.set(129, true);
        std::stringstream ss;
        EXPECT(bv.serialize(ss) == 8 + 3 * 8);
        std::string bytes = ss.str();
        uint64_t header;
        std::memcpy(&header, bytes.data(), 8);
        EXPECT(header == ((1ull << 56) | 130)); // int_vector.hpp:904-916
        bit_vector back;
        back.load(ss);
        EXPECT(back == bv && back.size() == 130 && back[64] && !back[65]);
        back.set(1, true);
        EXPECT(back != bv);
        std::printf("shim_test host-only ok\n");
        return 0;
    }
    if (argc > 1)
        set_device(std::atoi(argv[1]));
    std::mt19937_64 rng(4711);
    // ---- bit vectors
    for (uint64_t n : {1ull, 63ull, 64ull, 1000ull, 100000ull})
        for (double d : {0.5, 0.03})
        {
            bit_vector bv(n);
            for (uint64_t j = 0; j < n; ++j)
                if ((rng() % 10000) < d * 10000)
                    bv.set(j, true);
            check_bitvector<bit_vector, rank_support_v<1>, rank_support_v<0>, select_support_mcl<1>, select_support_mcl<0>>(bv, bv);
            check_pattern<10>(bv, true, false);
            check_pattern<01>(bv, false, true);
            check_pattern<00>(bv, false, false);
            check_pattern<11>(bv, true, true);
            rrr_vector<63> rrr(bv);
            check_bitvector<rrr_vector<63>, rrr_vector<63>::rank_1_type, rrr_vector<63>::rank_0_type, rrr_vector<63>::select_1_type,
                            rrr_vector<63>::select_0_type>(bv, rrr);
            sd_vector<> sd(bv);
            check_bitvector<sd_vector<>, sd_vector<>::rank_1_type, sd_vector<>::rank_0_type, sd_vector<>::select_1_type,
                            sd_vector<>::select_0_type>(bv, sd);
            EXPECT(rrr[n / 2] == bv[n / 2] && sd[n / 2] == bv[n / 2]);
        }
    // ---- wt_huff
    std::string text;
    for (int j = 0; j < 20000; ++j)
        text += (char)("abracadabra_$%&XYZ"[rng() % 18]);
    wt_huff<> wt(text);
    {
        std::string u(text);
        std::sort(u.begin(), u.end());
        EXPECT(wt.size() == text.size() && wt.sigma == (uint64_t)(std::unique(u.begin(), u.end()) - u.begin()));
    }
    {
        std::vector<uint64_t> cnt(256, 0);
        for (uint64_t j = 0; j < text.size(); ++j)
        {
            uint8_t c = (uint8_t)text[j];
            if (j % 37 == 0)
            {
                EXPECT(wt.rank(j, c) == cnt[c]);                       // wt_byte_test.cpp:134-168
                EXPECT(wt[j] == c);                                    // :106-131
                auto rc = wt.inverse_select(j);                        // :187-203
                EXPECT(rc.first == cnt[c] && rc.second == c);
                EXPECT(wt.select(cnt[c] + 1, c) == j);                 // :171-184
            }
            ++cnt[c];
        }
        EXPECT(wt.rank(text.size(), 'q') == 0 && wt.select(1, 'q') == text.size()); // absent symbol
    }
    // ---- wt_int
    std::vector<uint64_t> seq(5000);
    for (auto & x : seq)
        x = rng() % 1000;
    wt_int<> wi(seq);
    EXPECT(wi.size() == seq.size());
    for (uint64_t j = 0; j < seq.size(); j += 97)
    {
        uint64_t c = seq[j], r = std::count(seq.begin(), seq.begin() + j, c);
        EXPECT(wi.rank(j, c) == r && wi[j] == c && wi.select(r + 1, c) == j);
    }
    // ---- csa_wt + count / locate (examples/fm-index.cpp:64-69)
    std::string t2 = "abracadabra abracadabra simsalabim abracadabra";
    csa_wt<> csa(t2);
    EXPECT(csa.size() == t2.size() + 1);
    EXPECT(count(csa, std::string("abra")) == 6);
    EXPECT(count(csa, std::string("")) == t2.size() + 1 && count(csa, std::string("xyz")) == 0);
    auto occ = locate(csa, std::string("abracadabra"));
    std::sort(occ.begin(), occ.end());
    EXPECT((occ == std::vector<uint64_t>{0, 12, 35}));
    std::vector<uint64_t> sa(t2.size() + 1);
    for (uint64_t j = 0; j <= t2.size(); ++j)
        sa[j] = csa[j];
    std::vector<uint64_t> sorted(sa);
    std::sort(sorted.begin(), sorted.end());
    for (uint64_t j = 0; j <= t2.size(); ++j)
        EXPECT(sorted[j] == j); // SA is a permutation
    for (uint64_t j = 1; j <= t2.size(); ++j)
        EXPECT(t2.compare(sa[j - 1], std::string::npos, t2, sa[j], std::string::npos) < 0); // and sorted (csa_byte_test.cpp:162-175)
    EXPECT(extract(csa, 12, 22) == "abracadabra" && extract(csa, 0, t2.size() - 1) == t2); // examples/fm-index.cpp:83
    auto cnts = count(csa, std::vector<std::string>{"a", "abra", "sim", "zzz"});
    EXPECT((cnts == std::vector<uint64_t>{(uint64_t)std::count(t2.begin(), t2.end(), 'a'), 6, 1, 0}));
    std::vector<uint64_t> occ_off, occs;
    locate(csa, std::vector<std::string>{"sim", "cad"}, occ_off, occs);
    EXPECT(occ_off.size() == 3 && occ_off[2] == 4 && occs[0] == 24);
    // ---- serialize / load / store_to_file / load_from_file (io.hpp:877-896, 992-1011): round trips through the
    //      reference's byte format
    {
        std::stringstream ss;
        uint64_t written = wt.serialize(ss);
        EXPECT(written == ss.str().size() && written == size_in_bytes(wt));
        wt_huff<> wt2;
        wt2.load(ss);
        EXPECT(wt2.size() == wt.size() && wt2.sigma == wt.sigma);
        for (uint64_t j = 0; j < text.size(); j += 131)
            EXPECT(wt2[j] == wt[j] && wt2.rank(j, (uint8_t)text[j]) == wt.rank(j, (uint8_t)text[j]));
        std::string file = "/tmp/sdsl_b200_shim_test.csa";
        EXPECT(store_to_file(csa, file));
        csa_wt<> csa2;
        EXPECT(load_from_file(csa2, file) && csa2.size() == csa.size());
        EXPECT(count(csa2, std::string("abra")) == 6 && extract(csa2, 12, 22) == "abracadabra");
        EXPECT(!load_from_file(csa2, "/nonexistent/dir/x.csa"));
        std::remove(file.c_str());
        std::stringstream s2, s3;
        wi.serialize(s2);
        wt_int<> wi2;
        wi2.load(s2);
        EXPECT(wi2.size() == wi.size() && wi2[5] == wi[5]);
        bit_vector bv(5000);
        for (uint64_t j = 0; j < 5000; j += 7)
            bv.set(j, true);
        sd_vector<> sd(bv), sd2;
        sd.serialize(s3);
        sd2.load(s3);
        sd_vector<>::rank_1_type r1(&sd), r2(&sd2);
        EXPECT(sd2.size() == 5000 && r1.rank(4321) == r2.rank(4321));
    }
    // ---- the support concept's serialize / load (rank_support.hpp:57-74, select_support.hpp:62-77): a vector and its
    //      supports written to one stream and read back in the same order, like util::init_support users do
    {
        bit_vector bv(70001);
        for (uint64_t j = 0; j < bv.size(); j += 3 + (j % 11))
            bv.set(j, true);
        rank_support_v<1> r1(&bv);
        rank_support_v<0> r0(&bv);
        select_support_mcl<1> s1(&bv);
        select_support_mcl<0> s0(&bv);
        std::stringstream ss;
        uint64_t total = bv.serialize(ss) + r1.serialize(ss) + r0.serialize(ss) + s1.serialize(ss) + s0.serialize(ss);
        EXPECT(total == ss.str().size());
        ss.write("tail", 4);
        bit_vector bv2;
        rank_support_v<1> q1;
        rank_support_v<0> q0;
        select_support_mcl<1> t1;
        select_support_mcl<0> t0;
        bv2.load(ss);
        q1.load(ss, &bv2);
        q0.load(ss, &bv2);
        t1.load(ss, &bv2);
        t0.load(ss, &bv2);
        char tail[5] = {0};
        ss.read(tail, 4);
        EXPECT(std::string(tail) == "tail"); // every load consumed exactly its own bytes
        EXPECT(bv2 == bv && q1 != r1 && q1 == rank_support_v<1>(&bv2));
        for (uint64_t j = 0; j <= bv.size(); j += 997)
            EXPECT(q1.rank(j) == r1.rank(j) && q0.rank(j) == r0.rank(j));
        uint64_t ones = r1.rank(bv.size());
        for (uint64_t k = 1; k <= ones; k += 1013)
            EXPECT(t1.select(k) == s1.select(k));
        EXPECT(t0.select(5) == s0.select(5));
        rank_support_v5<1> r5(&bv), q5;
        std::stringstream s5;
        uint64_t n5 = r5.serialize(s5);
        EXPECT(n5 == 8 + 8 * 2 * (((bv.size() + 63) >> 11) + 1) && n5 < r1.serialize(none_sink())); // rank_support_v5.hpp:73-79: a quarter of the table
        q5.load(s5, &bv2);
        EXPECT(q5.rank(31337) == r1.rank(31337) && r5(70001) == ones);
        rrr_vector<63> rrr(bv);
        rrr_vector<63>::rank_1_type rr(&rrr);
        std::stringstream none;
        EXPECT(rr.serialize(none) == 0 && none.str().empty());
    }
    std::printf("shim_test ok\n");
    return 0;
}
