/* oracle_pt.h -- CPU ORACLE for the ray-cast + radiance-integration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (oclpathtracer_b200/,
 * include/) may include, link or call this.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * It restates, line by line and in plain C, the algorithm of the reference's
 * test/ClKernels/GenerateColors.cl (all 322 lines) and the host pieces of
 * test/RaytraceTest.cpp that feed it (structs :50-76, loadModel :87-198, frame
 * protocol :250-253, output transform :78-83,:277-287).  Each function cites the
 * lines it follows.
 *
 * PARITY STATUS: pinned against the reference's own output.  The reference holds
 * no golden vectors, known-answer tests or numeric fixtures for this path (its
 * RayCast test has zero assertions), so the pin is the
 * reference itself: oracle/_ref/adlTest64 (the unmodified test, built by
 * `make -C oracle ref`) run on a B200 through NVIDIA's OpenCL produced
 * tests/golden/reference_raycast_b200_opencl.npz (10000 frames); the CUDA path,
 * which is bit-identical to this oracle, reproduces it to rRMSE 9.95e-4 on the
 * 8-bit image (tests/test_gpu_parity.py, bound 1e-3 from the north star).  It
 * cannot be bit-exact: OpenCL leaves sin/cos/pow/normalize accuracy open.  Also
 * pinned:
 *   - the integer RNG (exactly specified by the source; KATs in
 *     tests/golden/rng_kat.json are derived from GenerateColors.cl:47-71),
 *   - the scene file (sha256 075b51a2...d18f62) and loader output,
 *   - structure layouts (64-byte records).
 * Where OpenCL C leaves arithmetic open (normalize/dot/cross association,
 * sin/cos/tan/pow accuracy, FMA contraction) this oracle CHOOSES and documents
 * (see "Numerics contract" in oracle_pt.c); those choices are then the
 * specification the CUDA path must reproduce bit for bit.
 */
#ifndef ORACLE_PT_H
#define ORACLE_PT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_float4 {
    float x, y, z, w;
} ora_float4;

/* GenerateColors.cl:12-19 / RaytraceTest.cpp:50-59 */
typedef struct ora_material {
    ora_float4 albedo;
    ora_float4 emissive;
    float roughness;
    int32_t type;
    char padding[24];
} ora_material;

/* GenerateColors.cl:21-28 / RaytraceTest.cpp:61-76 */
typedef struct ora_triangle {
    ora_float4 p1, p2, p3;
    int32_t id;
    char padding[12];
} ora_triangle;

/* BUILD-DEFINED acceleration structure: same byte layout as include/SharedHeader.h:ptb_bvh_node.
 * Built by the oracle's own builder (oracle_bvh.c: ora_bvh_build, the specification of the tree);
 * tests assert that the product's builders produce the same bytes.                              */
typedef struct ora_bvh_node { /* binary, 64 bytes (include/SharedHeader.h: ptb_bvh_node) */
    float c0[3]; /* child-0 box centre      */
    int32_t child0;
    float e0[3]; /* child-0 box half-extent */
    int32_t child1;
    float c1[3];
    int32_t pad0;
    float e1[3];
    int32_t pad1;
} ora_bvh_node;

typedef struct ora_bvh_node4 { /* 4-wide, 128 bytes (ptb_bvh_node4) */
    float c0[3];
    int32_t child0;
    float e0[3];
    int32_t child1;
    float c1[3];
    int32_t child2;
    float e1[3];
    int32_t child3;
    float c2[3];
    int32_t pad0;
    float e2[3];
    int32_t pad1;
    float c3[3];
    int32_t pad2;
    float e3[3];
    int32_t pad3;
} ora_bvh_node4;

/* FLAT form for scenes of <= 32 leaves and <= 64 triangles (include/SharedHeader.h: ptb_bvh_leafbox): the leaves of the
 * binary tree in leaf order, each with its padded box and the bit mask of its triangles' positions in the ordered array */
typedef struct ora_bvh_leafbox {
    float c[3];
    uint32_t mask_lo; /* triangles first .. first+count-1 of the ordered array, bits 0..31 */
    float e[3];
    uint32_t mask_hi; /* bits 32..63 */
} ora_bvh_leafbox;

typedef struct ora_bvh {
    const void* nodes; /* ora_bvh_node[n_nodes] (width 2), ora_bvh_node4[n_nodes] (width 4), ora_bvh_leafbox[n_nodes] (width 1) */
    int32_t n_nodes;
    const int32_t* tri_order; /* BVH position -> index into the caller's triangle array */
    int32_t n_tris;
    int32_t width;
    /* width 2 only: the QUANTISED encoding the product traverses (include/SharedHeader.h: ptb_bvh_nodeq), or NULL = walk the
     * fp32 boxes.  8 x uint32 per node: per child and axis one word (lo plane | hi plane << 16) on the grid below, then
     * the two child references. */
    const uint32_t* qnodes;
    float q_lo[3], q_step[3];
} ora_bvh;

/* BUILD-DEFINED quantisation of a binary tree ("Quantised binary nodes"): a pure function of the fp32
 * node array.  q_out: 8 * n_nodes words. */
int ora_bvh_quantize(const ora_bvh_node* nodes, int n_nodes, uint32_t* q_out, float q_lo[3], float q_step[3]);

enum { ORA_MODE_PRIMARY = 0, ORA_MODE_AO = 1, ORA_MODE_DIRECT = 2, ORA_MODE_PATH = 3 };
enum { ORA_ACCUM_REFERENCE = 0, ORA_ACCUM_LINEAR = 1 };

typedef struct ora_params {
    int32_t width, height;
    int32_t first_frame, n_frames;
    int32_t mode;
    int32_t accum;
    int32_t use_bvh; /* 0: brute force as the reference (ground truth for hit IDs) */
    int32_t max_depth;
    int32_t ao_samples;
    float ao_max_dist;
    int32_t light_quad;
    float light_p1[3], light_ea[3], light_eb[3]; /* area light parallelogram */
    int32_t shard_index, shard_count, shard_block;
    int32_t n_threads; /* 0 = all */
} ora_params;

/* per-pixel record for the LAST frame of a call (8 x 32-bit) */
typedef struct ora_pixel_stats {
    int32_t tri;          /* primary hit: index into the triangle array, -1 = miss */
    int32_t quad;         /* primary hit: Triangle.id, -1 = miss                  */
    uint32_t t_bits;      /* primary hit: bit pattern of t (0 on miss)             */
    uint32_t visits_primary;   /* internal BVH nodes fetched by the primary query  */
    uint32_t visits_secondary; /* ... by all other queries of the sample           */
    uint32_t count;       /* AO: unoccluded rays; DIRECT: 1 if lit; PATH: segments */
    uint32_t id_hash;     /* h = h*31 + (tri+2) over secondary queries, in order   */
    uint32_t tri_tests;   /* Moller-Trumbore tests started by the whole sample     */
} ora_pixel_stats;

typedef struct ora_counters {
    uint64_t rays_closest, rays_any, nodes, tri_tests, samples;
    uint64_t tri_u, tri_v, tri_t, tri_accept; /* tests that reached the u / v / t stage, accepted */
} ora_counters;

/* oracle_bvh.c: deterministic binned-SAH build (rules R1-R7 there).  width 2 -> ora_bvh_node[], width 4 -> ora_bvh_node4[]
 * (<= 2048 triangles), width 1 -> ora_bvh_leafbox[] (<= 32 leaves, <= 64 triangles).  *nodes and *tri_order are malloc'd: release with ora_free.  Returns 0 on success.   */
typedef struct ora_bvh_params {
    int32_t max_leaf;    /* 1..8, default 4 */
    float pad_rel;       /* boxes grow by pad_rel * scene diagonal, default 1e-4 */
    int32_t n_bins;      /* default 16 */
    int32_t smem_nodes;  /* breadth-first prefix, default 1024 */
    float traverse_cost; /* SAH cost of a node visit relative to a triangle test, default 1.2 */
} ora_bvh_params;
void ora_bvh_params_default(ora_bvh_params* p);
int ora_bvh_build(const ora_triangle* tris, int n_tris, const ora_bvh_params* params, int width, void** nodes,
                  int* n_nodes, int32_t** tri_order, int* depth, int* bfs_nodes);
void ora_free(void* p);

/* RaytraceTest.cpp:87-198 */
int ora_load_model(const char* path, ora_triangle* tris, int tri_cap, ora_material* mats, int mat_cap,
                   int* n_tris, int* n_mats);
/* BUILD-DEFINED C5 scene: every quad (2 consecutive triangles) -> k x k sub-quads */
int ora_tessellate(const ora_triangle* tris, int n_tris, int k, ora_triangle* out, int out_cap);
void ora_light_from_quad(const ora_triangle* tris, int n_tris, int quad, float p1[3], float ea[3], float eb[3]);

/* GenerateColors.cl:47-71 */
uint32_t ora_hash_uint32(uint32_t x);
float ora_random_float(uint32_t* seed);
void ora_rng_kat(uint32_t gid, uint32_t frame, int n, uint32_t* states, float* values);

/* numerics contract (see oracle_pt.c) */
void ora_sincos(float x, float* s, float* c);
float ora_tan(float x);
float ora_pow(float x, float y);
void ora_sincos_array(const float* x, int n, float* s, float* c);
void ora_pow_array(const float* x, int n, float y, float* out);

/* GenerateColors.cl:263-288 */
void ora_generate_ray(int gi, int gj, int width, int height, uint32_t* seed, float o[3], float d[3]);

/* Scene queries on arbitrary rays.  use_bvh=0: GenerateColors.cl:137-154 loop.
 * out_tri -1 on miss.  any_hit: first accepted triangle in visiting order.      */
void ora_trace(const ora_triangle* tris, int n_tris, const ora_bvh* bvh, int use_bvh, int any_hit, int n_rays,
               const float* o, const float* d, const float* tmax, int32_t* out_tri, float* out_t, float* out_u,
               float* out_v, uint32_t* out_visits, uint32_t* out_tests);

/* The hot path: GenerateColors.cl:302-322 per pixel x frame, plus the
 * BUILD-DEFINED modes.  fb: float4 per local pixel (in/out for
 * ORA_ACCUM_REFERENCE with first_frame > 0).  stats / counters may be NULL.    */
int ora_render(const ora_params* p, const ora_triangle* tris, int n_tris, const ora_material* mats, int n_mats,
               const ora_bvh* bvh, float* fb, ora_pixel_stats* stats, ora_counters* counters);

/* RaytraceTest.cpp:78-83,:280-285: c8 = min((int)(sqrtf(v)*255), 255) */
void ora_to_rgb8(const float* fb, int n_pixels, uint8_t* rgb);

int ora_max_threads(void);
int ora_local_pixel_count(const ora_params* p);

#ifdef __cplusplus
}
#endif
#endif
