// sais.h — host suffix-array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009), written for
// this engine's FM-index builder.  The reference delegates this step to a vendored divsufsort
// (construct_sa.hpp:103-137); any correct suffix sorter produces the same array, which is all the index
// needs (the BWT, the samples and therefore every query result are functions of the suffix array only).
//
// Requirements: s[n-1] is the unique smallest symbol (the 0 sentinel the CSA appends, construct.hpp:47-52).
// Index type IdxT must be signed (int32_t for n < 2^31, int64_t otherwise).
#pragma once
#include <cstdint>
#include <vector>

namespace sdslgpu
{

template <class CharT, class IdxT>
void sais(CharT const * s, IdxT * SA, IdxT n, IdxT K)
{
    if (n == 0)
        return;
    if (n == 1)
    {
        SA[0] = 0;
        return;
    }
    // suffix types: 1 = S-type, 0 = L-type
    std::vector<uint8_t> t((size_t)n);
    t[n - 1] = 1;
    t[n - 2] = 0;
    for (IdxT i = n - 3; i >= 0; --i)
        t[i] = (s[i] < s[i + 1] || (s[i] == s[i + 1] && t[i + 1])) ? 1 : 0;
    auto is_lms = [&](IdxT i) { return i > 0 && t[i] && !t[i - 1]; };

    std::vector<IdxT> bkt((size_t)K + 1);
    auto buckets = [&](bool end) {
        for (IdxT c = 0; c <= K; ++c)
            bkt[c] = 0;
        for (IdxT i = 0; i < n; ++i)
            ++bkt[(IdxT)s[i]];
        IdxT sum = 0;
        for (IdxT c = 0; c <= K; ++c)
        {
            sum += bkt[c];
            bkt[c] = end ? sum : sum - bkt[c];
        }
    };
    auto induce_l = [&]() {
        buckets(false);
        for (IdxT i = 0; i < n; ++i)
        {
            IdxT j = SA[i] - 1;
            if (j >= 0 && !t[j])
                SA[bkt[(IdxT)s[j]]++] = j;
        }
    };
    auto induce_s = [&]() {
        buckets(true);
        for (IdxT i = n - 1; i >= 0; --i)
        {
            IdxT j = SA[i] - 1;
            if (j >= 0 && t[j])
                SA[--bkt[(IdxT)s[j]]] = j;
        }
    };

    // stage 1: sort the LMS substrings
    buckets(true);
    for (IdxT i = 0; i < n; ++i)
        SA[i] = -1;
    for (IdxT i = 1; i < n; ++i)
        if (is_lms(i))
            SA[--bkt[(IdxT)s[i]]] = i;
    induce_l();
    induce_s();

    // compact the sorted LMS substrings and name them
    IdxT n1 = 0;
    for (IdxT i = 0; i < n; ++i)
        if (is_lms(SA[i]))
            SA[n1++] = SA[i];
    for (IdxT i = n1; i < n; ++i)
        SA[i] = -1;
    IdxT name = 0, prev = -1;
    for (IdxT i = 0; i < n1; ++i)
    {
        IdxT pos = SA[i];
        bool diff = false;
        for (IdxT d = 0; d < n; ++d)
        {
            if (prev == -1 || s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d])
            {
                diff = true;
                break;
            }
            else if (d > 0 && (is_lms(pos + d) || is_lms(prev + d)))
                break;
        }
        if (diff)
        {
            ++name;
            prev = pos;
        }
        SA[n1 + pos / 2] = name - 1;
    }
    for (IdxT i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0)
            SA[j--] = SA[i];

    // stage 2: order the LMS suffixes (recursively when names collide)
    IdxT * SA1 = SA;
    IdxT * s1 = SA + n - n1;
    if (name < n1)
        sais<IdxT, IdxT>(s1, SA1, n1, name - 1);
    else
        for (IdxT i = 0; i < n1; ++i)
            SA1[s1[i]] = i;

    // stage 3: induce the full order from the sorted LMS suffixes
    buckets(true);
    for (IdxT i = 1, j = 0; i < n; ++i)
        if (is_lms(i))
            s1[j++] = i;
    for (IdxT i = 0; i < n1; ++i)
        SA1[i] = s1[SA1[i]];
    for (IdxT i = n1; i < n; ++i)
        SA[i] = -1;
    for (IdxT i = n1 - 1; i >= 0; --i)
    {
        IdxT j = SA[i];
        SA[i] = -1;
        SA[--bkt[(IdxT)s[j]]] = j;
    }
    induce_l();
    induce_s();
}

} // namespace sdslgpu
