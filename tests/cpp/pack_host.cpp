// pack_host.cpp — CPU harness for sdsl-lite_b200/csrc/sdsl_pack.h: feeds the select_support_mcl writer with
// argument positions from a naive scan (the GPU library feeds it from the batched select kernel), so that
// tests/test_egress_pack.py can compare the bytes with the reference's serialize() on a box without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sdsl-lite_b200/csrc/sdsl_pack.h"

extern "C" uint64_t pack_select_mcl_naive(uint64_t const * words, uint64_t nbits, int b, uint8_t * out, uint64_t cap)
{
    std::vector<uint64_t> P;
    for (uint64_t i = 0; i < nbits; ++i)
        if ((int)((words[i >> 6] >> (i & 63)) & 1) == b)
            P.push_back(i);
    std::vector<uint8_t> blob;
    sdslgpu::pack::Sink sink{blob};
    sdslgpu::pack::SelectFn sel = [&](uint64_t const * keys, uint64_t n, uint64_t * pos) -> int {
        for (uint64_t k = 0; k < n; ++k)
        {
            if (keys[k] >= P.size())
                return -1;
            pos[k] = P[keys[k]];
        }
        return 0;
    };
    if (sdslgpu::pack::write_select_mcl(nbits, P.size(), sel, sink) != 0)
        return 0;
    if (blob.size() <= cap)
        std::memcpy(out, blob.data(), blob.size());
    return blob.size();
}
