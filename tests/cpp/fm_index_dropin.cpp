// fm_index_dropin.cpp — ONE source, two builds: against the reference's headers (-DUSE_REFERENCE -I<reference>/include)
// and against sdsl-lite_b200/include/sdsl_b200.hpp with nothing changed but the include and the namespace alias.
// It walks the interface of the hot path the way user code of the reference does (cf. its examples/fm-index.cpp:41-83:
// declare the index type, construct, count, locate, extract) plus the public members the search algorithms are
// written against (csa_wt.hpp:117-130) and prints everything; the two builds must print the same bytes
// (tests/test_dropin.py; the expected output is committed as tests/golden/fm_index_dropin.expected, made by the
// reference build).
#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#ifdef USE_REFERENCE
#include <sdsl/suffix_array_algorithm.hpp>
#include <sdsl/suffix_arrays.hpp>
#include <sdsl/wavelet_trees.hpp>
#else
#include "../../sdsl-lite_b200/include/sdsl_b200.hpp"
namespace sdsl = sdsl_b200;
#endif

using namespace std;

template <class t_csa>
void report(t_csa const & fm, vector<string> const & queries, string const & tag)
{
    static_assert(is_same<typename t_csa::index_category, sdsl::csa_tag>::value, "a CSA");
    static_assert(is_same<typename t_csa::wavelet_tree_type::index_category, sdsl::wt_tag>::value, "over a wavelet tree");
    cout << "== " << tag << ": size " << fm.size() << " sigma " << (unsigned)fm.sigma << " dens " << (unsigned)t_csa::sa_sample_dens << "/"
         << (unsigned)t_csa::isa_sample_dens << "\n";
    cout << "C:";
    for (size_t c = 0; c <= fm.sigma; ++c)
        cout << ' ' << fm.C[c];
    cout << "\ncomp2char:";
    for (size_t c = 0; c < fm.sigma; ++c)
        cout << ' ' << (unsigned)fm.comp2char[c] << "->" << (unsigned)fm.char2comp[fm.comp2char[c]];
    cout << "\n";
    for (string const & query : queries)
    {
        size_t m = query.size();
        size_t occs = sdsl::count(fm, query.begin(), query.end());
        cout << "[" << query << "] occurrences " << occs;
        auto iv = sdsl::lex_interval(fm, query.begin(), query.end());
        cout << " interval " << iv[0] << ' ' << iv[1] << (iv[1] + 1 - iv[0] == occs ? " ok" : " MISMATCH");
        if (occs > 0)
        {
            auto locations = sdsl::locate(fm, query.begin(), query.begin() + m);
            cout << " SA-order";
            for (size_t k = 0; k < min<size_t>(locations.size(), 6); ++k)
                cout << ' ' << locations[k];
            sort(locations.begin(), locations.end());
            size_t pos = locations[0], pre = min<size_t>(pos, 5), post = min<size_t>(fm.size() - 1 - pos - m, 5);
            auto s = sdsl::extract(fm, pos - pre, pos + m + post - 1);
            cout << " first " << pos << " context {" << s << "}";
        }
        cout << "\n";
    }
    // the members backward_search is written against, used directly
    typename t_csa::size_type l = 0, r = fm.size() - 1, l2 = 0, r2 = 0;
    string const probe = queries.empty() ? string("a") : queries[0];
    for (size_t k = probe.size(); k-- > 0;)
    {
        sdsl::backward_search(fm, l, r, (typename t_csa::char_type)probe[k], l2, r2);
        cout << "step '" << probe[k] << "' -> [" << l2 << ',' << r2 << "]\n";
        l = l2;
        r = r2;
        if (r + 1 - l == 0)
            break;
    }
    size_t const n = fm.size();
    for (size_t i : {size_t(0), n / 3, n / 2, n - 1})
    {
        auto c = fm.bwt[i];
        cout << "i " << i << " sa " << fm[i] << " bwt " << (unsigned)c << " rank " << fm.bwt.rank(i, c) << " lf " << fm.lf[i] << " wt " << (unsigned)fm.wavelet_tree[i]
             << " wt.rank " << fm.wavelet_tree.rank(i, c) << " inv " << fm.wavelet_tree.inverse_select(i).first << "\n";
    }
    cout << "wt size " << fm.wavelet_tree.size() << " sigma " << fm.wavelet_tree.sigma << " first symbols";
    auto it = fm.wavelet_tree.begin();
    for (size_t k = 0; k < min<size_t>(n, 8); ++k, ++it)
        cout << ' ' << (unsigned)*it;
    cout << "\n";
}

int main(int argc, char ** argv)
{
    if (argc < 3)
    {
        cerr << "usage: " << argv[0] << " text_file tmp_index_file < queries\n";
        return 1;
    }
    ifstream in(argv[1], ios::binary);
    string text((istreambuf_iterator<char>(in)), istreambuf_iterator<char>());
    vector<string> queries;
    for (string q; getline(cin, q);)
        queries.push_back(q);

    sdsl::csa_wt<sdsl::wt_huff<>> fm_index; // the reference's default byte FM-index
    sdsl::construct_im(fm_index, text, 1);
    report(fm_index, queries, "csa_wt<wt_huff<>>");
    cout << "size_in_bytes " << sdsl::size_in_bytes(fm_index) << "\n";

    // store, load into a second object, copy it: the answers survive (io.hpp:877-896, 992-1011)
    if (!sdsl::store_to_file(fm_index, argv[2]))
        return 2;
    sdsl::csa_wt<sdsl::wt_huff<>> loaded;
    if (!sdsl::load_from_file(loaded, argv[2]))
        return 3;
    sdsl::csa_wt<sdsl::wt_huff<>> copy(loaded);
    report(copy, queries, "stored, loaded and copied");

    sdsl::csa_wt<sdsl::wt_huff<sdsl::rrr_vector<63>>, 64, 128> fm_rrr; // other template arguments: compressed tree, sparser samples
    sdsl::construct_im(fm_rrr, text, 1);
    report(fm_rrr, queries, "csa_wt<wt_huff<rrr_vector<63>>, 64, 128>");

    sdsl::wt_huff<> wt;
    sdsl::construct_im(wt, text, 1);
    cout << "wt_huff: size " << wt.size() << " sigma " << wt.sigma << " rank(n/2,'a') " << wt.rank(wt.size() / 2, 'a') << " select(1,'a') "
         << (wt.rank(wt.size(), 'a') ? wt.select(1, 'a') : wt.size()) << "\n";

    sdsl::bit_vector bv(1000, 0);
    for (size_t i = 0; i < bv.size(); i += 7)
        bv[i] = 1;
    sdsl::bit_vector::rank_1_type rank1(&bv);
    sdsl::bit_vector::select_1_type select1(&bv);
    sdsl::sd_vector<> sd(bv);
    sdsl::sd_vector<>::rank_1_type sd_rank(&sd);
    sdsl::sd_vector<>::select_0_type sd_sel0(&sd);
    sdsl::rrr_vector<63> rrr(bv);
    sdsl::rrr_vector<63>::select_1_type rrr_sel(&rrr);
    cout << "bit vectors: " << rank1(500) << ' ' << select1(10) << ' ' << sd_rank(500) << ' ' << sd_sel0(100) << ' ' << rrr_sel(10) << "\n";
    return 0;
}
