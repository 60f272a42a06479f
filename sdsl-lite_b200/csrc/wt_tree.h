// wt_tree.h — the node table of a byte wavelet tree as every wt / fm kernel stages it into shared memory.
// Plain C++ (no CUDA): shared by the kernels (internal.h), the host shape builder (wt_shape.h) and the CPU tests.
#pragma once
#include <cstdint>

namespace sdslgpu
{

static constexpr uint16_t kWtUndef = 0xFFFF; // "no node" / "symbol absent" (wt_helper.hpp: undef)

// node table + per-symbol paths of a byte wavelet tree (wt_helper.hpp:219-225), as staged into shared
// memory by every wt / fm kernel.  ~14 KB.
struct alignas(16) WtTree
{
    static constexpr int kMaxNodes = 512; // 2*256-1 nodes + one sentinel slot (keeps every array 16-byte aligned)
    uint64_t bv_pos[kMaxNodes];
    uint64_t bv_pos_rank[kMaxNodes]; // leaves: the symbol
    uint16_t child[kMaxNodes][2];    // 0xFFFF = leaf
    uint16_t parent[kMaxNodes];
    uint16_t c_to_leaf[256]; // 0xFFFF = symbol absent
    uint64_t path[256];      // bits 0..55 path from the root (LSB first), bits 56..63 its length
    uint64_t occ[256];       // occurrences of each symbol (not in the reference's tree; bounds select)
    uint32_t nnodes;
    uint32_t pad_;
};

} // namespace sdslgpu
