// sais_host.cpp — CPU harness for sdsl-lite_b200/csrc/sais.h (the host suffix sorter behind the FM-index builder's
// fallback path): suffix array of text + 0 sentinel, 32- and 64-bit index types.
#include <cstdint>
#include <vector>

#include "../../sdsl-lite_b200/csrc/sais.h"

extern "C" void sais_host(uint8_t const * text, uint64_t len, int wide, uint64_t * sa_out)
{
    uint64_t n = len + 1;
    std::vector<uint8_t> t(text, text + len);
    t.push_back(0);
    if (wide)
    {
        std::vector<int64_t> sa(n);
        sdslgpu::sais<uint8_t, int64_t>(t.data(), sa.data(), (int64_t)n, 255);
        for (uint64_t i = 0; i < n; ++i)
            sa_out[i] = (uint64_t)sa[i];
    }
    else
    {
        std::vector<int32_t> sa(n);
        sdslgpu::sais<uint8_t, int32_t>(t.data(), sa.data(), (int32_t)n, 255);
        for (uint64_t i = 0; i < n; ++i)
            sa_out[i] = (uint64_t)sa[i];
    }
}
