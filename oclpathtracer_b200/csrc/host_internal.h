// host_internal.h -- declarations shared by the host-side translation units.
#pragma once

#include <cstdint>
#include <vector>

#include "ptb200.h"

namespace ptb {

// records the thread-local message behind ptb_last_error() and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

struct BuiltBvh {
    std::vector<ptb_bvh_node> nodes;   // binary tree; [0] is the root; first `smem_nodes` in BFS order
    std::vector<ptb_bvh_node4> nodes4; // the same tree collapsed to 4-wide nodes (only for scenes of <= 2048 triangles)
    int depth4 = 0, smem_nodes4 = 0;
    std::vector<ptb_bvh_leafbox> flat; // the leaf slots in leaf order (only for scenes of <= 32 leaves and <= 64 triangles)
    std::vector<ptb_bvh_tri> tris;     // BVH order, precomputed-edge layout
    std::vector<int32_t> tri_order;    // BVH position -> caller's triangle index
    int depth = 0;                     // longest root-to-leaf chain of internal nodes
    int smem_nodes = 0;                // nodes in the BFS-ordered prefix
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
};

// Deterministic binned-SAH build (bvh_build.cpp).  Returns PTB_OK or a PTB_E_* code.
int build_bvh(const ptb_triangle* tris, int n_tris, const ptb_bvh_params& params, BuiltBvh* out);

// Triangle in caller order, same 48-byte layout (brute-force path).
void make_edge_tris(const ptb_triangle* tris, int n_tris, std::vector<ptb_bvh_tri>* out);

}  // namespace ptb
