/* oracle_bvh.c -- the oracle's OWN deterministic BVH builder.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_pt.h).  The reference has no acceleration
 * structure (test/ClKernels/GenerateColors.cl:137-154 loops over NUM_TRIANGLES), so the
 * tree is BUILD-DEFINED and this file is its specification, written from the rules in
 * DESIGN.md section 4 and sharing no code with the product's builder
 * (oclpathtracer_b200/csrc/bvh_build.cpp).  tests/ assert that both produce the same
 * bytes, so node-visit parity no longer leans on the product's own tree, and
 * bench.py --impl reference can run the BVH path without loading libptb200.so.
 *
 * The rules:
 *  R1 primitive = one caller triangle: box = exact fp32 min/max of its three vertices,
 *     centroid = 0.5f*lo + 0.5f*hi per axis.
 *  R2 a range of primitives becomes a LEAF when it has one primitive, or when it has
 *     <= max_leaf primitives, is not the root, and n * 1.0 <= traverse_cost + SAH(best
 *     split) / area(range) (all in double; area = dx*dy + dy*dz + dz*dx of the exact
 *     bounds).  Leaf primitives are stored in ascending caller index.
 *  R3 split = binned SAH: n_bins equal-width centroid bins per axis over the range's
 *     centroid bounds (bin = (int)((c - cmin) * (n_bins / (cmax - cmin))), clamped), axes
 *     x, y, z in that order, boundaries in ascending order, a candidate replaces the
 *     incumbent only when its cost is STRICTLY lower; cost = area(L)*|L| + area(R)*|R|.
 *     The range is partitioned in place by one forward sweep that swaps every left-side
 *     primitive to the front (so the left side keeps its relative order).  When no axis
 *     has a centroid extent the range is ordered by caller index and cut in half.
 *  R4 nodes are created in pre-order (node, left subtree, right subtree) and store their
 *     CHILDREN's boxes, grown by pad = pad_rel * scene diagonal, as centre/half-extent:
 *     c = 0.5f*lo + 0.5f*hi, h = 0.5f*hi - 0.5f*lo, e = h + 2.4e-7f*(|c| + h).
 *  R5 renumbering: the first smem_nodes nodes breadth-first from the root (children in
 *     slot order), then every subtree still queued depth-first (slot 0 first).
 *  R6 4-wide form (<= 2048 triangles): start from a binary node's two children and keep
 *     replacing the internal child with the largest box area (first such slot on ties) by
 *     its two children, in place, until four slots are used or only leaves remain;
 *     unused slots: ref 0x7fffffff, c = 0, e = -1e30.  Renumbered as R5.
 *  R7 flat form (<= 32 leaves and <= 64 triangles): the leaf slots of the binary tree in depth-first order (slot 0
 *     first), i.e. in the order of the triangle array; each record = the slot's padded box (R4) + the 64-bit mask of
 *     the positions its triangles occupy in the ordered array.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_pt.h"

typedef struct { float lo[3], hi[3]; } obox;
typedef struct { float lo[3], hi[3], c[3]; int32_t idx; } oprim;
typedef struct { int32_t child[2]; obox box[2]; } onode;

static void box_reset(obox* b) {
    for (int a = 0; a < 3; a++) { b->lo[a] = INFINITY; b->hi[a] = -INFINITY; }
}
static void box_grow(obox* b, const float* lo, const float* hi) {
    for (int a = 0; a < 3; a++) {
        if (lo[a] < b->lo[a]) b->lo[a] = lo[a];
        if (hi[a] > b->hi[a]) b->hi[a] = hi[a];
    }
}
static double box_area(const obox* b) { /* R2 */
    double dx = (double)b->hi[0] - b->lo[0], dy = (double)b->hi[1] - b->lo[1], dz = (double)b->hi[2] - b->lo[2];
    return dx * dy + dy * dz + dz * dx;
}
static int cmp_idx(const void* x, const void* y) {
    int32_t a = ((const oprim*)x)->idx, b = ((const oprim*)y)->idx;
    return (a > b) - (a < b);
}
static int bin_of(float c, float cmin, float scale, int n_bins) { /* R3 */
    int k = (int)((c - cmin) * scale);
    return k < 0 ? 0 : (k >= n_bins ? n_bins - 1 : k);
}
static void centre_extent(float lo, float hi, float* c, float* e) { /* R4 */
    *c = 0.5f * lo + 0.5f * hi;
    float h = 0.5f * hi - 0.5f * lo;
    *e = h + 2.4e-7f * (fabsf(*c) + h);
}

typedef struct {
    oprim* prims;
    onode* nodes; int n_nodes;
    int32_t* order; int n_order;
    int max_leaf, n_bins;
    double k_traverse;
} builder;

static int32_t emit_leaf(builder* B, int b, int e) { /* R2 */
    qsort(B->prims + b, (size_t)(e - b), sizeof(oprim), cmp_idx);
    int first = B->n_order;
    for (int i = b; i < e; i++) B->order[B->n_order++] = B->prims[i].idx;
    return ~(int32_t)(((uint32_t)first << 3) | (uint32_t)(e - b - 1));
}

/* explicit work stack instead of recursion (a chain of 1 | n-1 splits may be as deep as the triangle count):
 * tasks are popped last-in first-out and a node pushes its right range before its left one, which is exactly
 * the pre-order of R4 */
typedef struct { int b, e, parent, slot, force; } otask;

static int build_tree(builder* B, int n_tris) {
    otask* todo = (otask*)malloc(sizeof(otask) * (size_t)(n_tris + 2));
    obox* bin_box = (obox*)malloc(sizeof(obox) * (size_t)B->n_bins);
    int* bin_cnt = (int*)malloc(sizeof(int) * (size_t)B->n_bins);
    double* right_area = (double*)malloc(sizeof(double) * (size_t)B->n_bins);
    int* right_cnt = (int*)malloc(sizeof(int) * (size_t)B->n_bins);
    if (!todo || !bin_box || !bin_cnt || !right_area || !right_cnt) return -1;
    int nt = 0;
    todo[nt++] = (otask){0, n_tris, -1, 0, 1};
    while (nt) {
        const otask t = todo[--nt];
        const int n = t.e - t.b;
        oprim* P = B->prims;
        obox bb, cb;
        box_reset(&bb); box_reset(&cb);
        for (int i = t.b; i < t.e; i++) { box_grow(&bb, P[i].lo, P[i].hi); box_grow(&cb, P[i].c, P[i].c); }
        if (t.parent >= 0) B->nodes[t.parent].box[t.slot] = bb;
        int best_axis = -1, best_split = -1;
        double best_cost = INFINITY;
        if (n >= 2) {
            for (int axis = 0; axis < 3; axis++) {
                const float cmin = cb.lo[axis], cmax = cb.hi[axis];
                if (!(cmax > cmin)) continue;
                const float scale = (float)B->n_bins / (cmax - cmin);
                for (int k = 0; k < B->n_bins; k++) { box_reset(&bin_box[k]); bin_cnt[k] = 0; }
                for (int i = t.b; i < t.e; i++) {
                    int k = bin_of(P[i].c[axis], cmin, scale, B->n_bins);
                    box_grow(&bin_box[k], P[i].lo, P[i].hi);
                    bin_cnt[k]++;
                }
                obox acc; int cnt = 0;
                box_reset(&acc);
                for (int k = B->n_bins - 1; k >= 1; k--) {
                    if (bin_cnt[k]) box_grow(&acc, bin_box[k].lo, bin_box[k].hi);
                    cnt += bin_cnt[k];
                    right_area[k] = cnt ? box_area(&acc) : 0.0;
                    right_cnt[k] = cnt;
                }
                box_reset(&acc); cnt = 0;
                for (int k = 1; k < B->n_bins; k++) {
                    if (bin_cnt[k - 1]) box_grow(&acc, bin_box[k - 1].lo, bin_box[k - 1].hi);
                    cnt += bin_cnt[k - 1];
                    if (cnt == 0 || right_cnt[k] == 0) continue;
                    double cost = box_area(&acc) * cnt + right_area[k] * right_cnt[k];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = k; }
                }
            }
        }
        const double parent_area = box_area(&bb);
        double split_cost = INFINITY;
        if (best_axis >= 0) split_cost = parent_area > 0.0 ? B->k_traverse + 1.0 * best_cost / parent_area : B->k_traverse;
        const int can_leaf = n <= B->max_leaf && !t.force;
        if (n == 1 || (can_leaf && (double)n * 1.0 <= split_cost)) {
            int32_t ref = emit_leaf(B, t.b, t.e);
            if (t.parent >= 0) B->nodes[t.parent].child[t.slot] = ref;
            continue;
        }
        int mid;
        if (best_axis >= 0) {
            const float cmin = cb.lo[best_axis];
            const float scale = (float)B->n_bins / (cb.hi[best_axis] - cmin);
            mid = t.b;
            for (int i = t.b; i < t.e; i++)
                if (bin_of(P[i].c[best_axis], cmin, scale, B->n_bins) < best_split) {
                    oprim tmp = P[i]; P[i] = P[mid]; P[mid] = tmp;
                    mid++;
                }
        } else {
            qsort(P + t.b, (size_t)n, sizeof(oprim), cmp_idx);
            mid = t.b + n / 2;
        }
        const int me = B->n_nodes++;
        if (t.parent >= 0) B->nodes[t.parent].child[t.slot] = me;
        todo[nt++] = (otask){mid, t.e, me, 1, 0};
        todo[nt++] = (otask){t.b, mid, me, 0, 0};
    }
    free(todo); free(bin_box); free(bin_cnt); free(right_area); free(right_cnt);
    return 0;
}

/* R5 for any node width: refs[o*width + k] >= 0 and != EMPTY are internal children */
#define ORA_EMPTY 0x7fffffff
static int renumber(const int32_t* refs, int width, int n_nodes, int want_bfs, int32_t* new_of, int32_t* old_of, int* bfs_count) {
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_nodes + 1));
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_nodes + 1));
    if (!queue || !stack) return -1;
    for (int i = 0; i < n_nodes; i++) new_of[i] = -1;
    int qh = 0, qt = 0, n_out = 0;
    queue[qt++] = 0;
    while (qh < qt && n_out < want_bfs) {
        int32_t o = queue[qh++];
        new_of[o] = n_out; old_of[n_out++] = o;
        for (int k = 0; k < width; k++) {
            int32_t r = refs[(size_t)o * width + k];
            if (r >= 0 && r != ORA_EMPTY) queue[qt++] = r;
        }
    }
    *bfs_count = n_out;
    for (; qh < qt; qh++) {
        int sp = 0;
        stack[sp++] = queue[qh];
        while (sp) {
            int32_t o = stack[--sp];
            new_of[o] = n_out; old_of[n_out++] = o;
            for (int k = width - 1; k >= 0; k--) {
                int32_t r = refs[(size_t)o * width + k];
                if (r >= 0 && r != ORA_EMPTY) stack[sp++] = r;
            }
        }
    }
    free(queue); free(stack);
    return n_out == n_nodes ? 0 : -1;
}

static int chain_depth(const int32_t* refs, int width, int n_nodes) { /* longest chain of internal nodes */
    int32_t* st = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(n_nodes + 1));
    if (!st) return -1;
    int sp = 0, depth = 1;
    st[sp++] = 0; st[sp++] = 1;
    while (sp) {
        int dep = st[--sp]; int32_t ni = st[--sp];
        if (dep > depth) depth = dep;
        for (int k = 0; k < width; k++) {
            int32_t r = refs[(size_t)ni * width + k];
            if (r >= 0 && r != ORA_EMPTY) { st[sp++] = r; st[sp++] = dep + 1; }
        }
    }
    free(st);
    return depth;
}

void ora_bvh_params_default(ora_bvh_params* p) {
    memset(p, 0, sizeof *p);
    p->max_leaf = 4; p->pad_rel = 1e-4f; p->n_bins = 16; p->smem_nodes = 1024; p->traverse_cost = 1.2f;
}

void ora_free(void* p) { free(p); }

int ora_bvh_build(const ora_triangle* tris, int n_tris, const ora_bvh_params* params, int width, void** nodes_out,
                  int* n_nodes_out, int32_t** tri_order_out, int* depth_out, int* bfs_nodes_out) {
    if (!tris || n_tris < 1 || !nodes_out || !n_nodes_out || !tri_order_out || (width != 2 && width != 4 && width != 1)) return -1;
    if (width == 4 && n_tris > 2048) return -1;
    if (width == 1 && n_tris > 64) return -1;
    ora_bvh_params dp;
    if (!params) { ora_bvh_params_default(&dp); params = &dp; }
    builder B;
    memset(&B, 0, sizeof B);
    B.max_leaf = params->max_leaf < 1 ? 1 : (params->max_leaf > 8 ? 8 : params->max_leaf);
    B.n_bins = params->n_bins < 2 ? 2 : (params->n_bins > 256 ? 256 : params->n_bins);
    B.k_traverse = params->traverse_cost > 0.0f ? (double)params->traverse_cost : 1.2;
    B.prims = (oprim*)malloc(sizeof(oprim) * (size_t)n_tris);
    B.nodes = (onode*)calloc((size_t)n_tris + 1, sizeof(onode));
    B.order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_tris);
    if (!B.prims || !B.nodes || !B.order) return -1;
    obox scene;
    box_reset(&scene);
    for (int i = 0; i < n_tris; i++) { /* R1 */
        oprim* p = &B.prims[i];
        const float* v[3] = {&tris[i].p1.x, &tris[i].p2.x, &tris[i].p3.x};
        for (int a = 0; a < 3; a++) {
            float lo = v[0][a], hi = v[0][a];
            for (int k = 1; k < 3; k++) { if (v[k][a] < lo) lo = v[k][a]; if (v[k][a] > hi) hi = v[k][a]; }
            if (!(lo == lo) || !(hi == hi) || isinf(lo) || isinf(hi)) return -1;
            p->lo[a] = lo; p->hi[a] = hi;
            p->c[a] = 0.5f * lo + 0.5f * hi;
        }
        p->idx = i;
        box_grow(&scene, p->lo, p->hi);
    }
    if (n_tris == 1) { /* the root must be an internal node: both children are the same one-triangle leaf */
        int32_t leaf = emit_leaf(&B, 0, 1);
        B.nodes[0].child[0] = B.nodes[0].child[1] = leaf;
        B.nodes[0].box[0] = B.nodes[0].box[1] = scene;
        B.n_nodes = 1;
    } else if (build_tree(&B, n_tris)) {
        return -1;
    }
    const float dx = scene.hi[0] - scene.lo[0], dy = scene.hi[1] - scene.lo[1], dz = scene.hi[2] - scene.lo[2];
    const float diag = sqrtf(dx * dx + dy * dy + dz * dz);
    const float pad = (params->pad_rel > 0.0f ? params->pad_rel : 0.0f) * diag;

    int rc = -1;
    if (width == 1) { /* R7 */
        ora_bvh_leafbox* out = (ora_bvh_leafbox*)calloc(64, sizeof(ora_bvh_leafbox));
        int32_t* st = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(B.n_nodes + 1));
        if (!out || !st) return -1;
        int n_leaf = 0, sp = 0, too_many = 0;
        /* (node, slot) pairs still to visit, last in first out: slot 1 is pushed first so that slot 0 is handled first;
         * the one-triangle root repeats its leaf in both slots: keep one */
        if (n_tris > 1) { st[sp++] = 0; st[sp++] = 1; }
        st[sp++] = 0; st[sp++] = 0;
        while (sp) {
            const int slot = st[--sp]; const int32_t ni = st[--sp];
            const int32_t r = B.nodes[ni].child[slot];
            if (r >= 0) { st[sp++] = r; st[sp++] = 1; st[sp++] = r; st[sp++] = 0; continue; }
            if (n_leaf >= 32) { too_many = 1; break; }
            const uint32_t code = (uint32_t)(~r);
            const int first = (int)(code >> 3), count = (int)(code & 7u) + 1;
            const uint64_t m = ((count >= 64 ? 0ull : (1ull << count)) - 1ull) << first;
            ora_bvh_leafbox* d = &out[n_leaf++];
            for (int a = 0; a < 3; a++) centre_extent(B.nodes[ni].box[slot].lo[a] - pad, B.nodes[ni].box[slot].hi[a] + pad, &d->c[a], &d->e[a]);
            d->mask_lo = (uint32_t)m; d->mask_hi = (uint32_t)(m >> 32);
        }
        free(st);
        if (too_many) { free(out); free(B.prims); free(B.nodes); free(B.order); return -2; }
        if (depth_out) *depth_out = 1;
        if (bfs_nodes_out) *bfs_nodes_out = n_leaf;
        *nodes_out = out; *n_nodes_out = n_leaf;
        rc = 0;
    } else if (width == 2) {
        const int nn = B.n_nodes;
        int32_t* refs = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)nn);
        int32_t* new_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)nn);
        int32_t* old_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)nn);
        ora_bvh_node* out = (ora_bvh_node*)calloc((size_t)nn, sizeof(ora_bvh_node));
        if (!refs || !new_of || !old_of || !out) return -1;
        for (int i = 0; i < nn; i++) { refs[2 * i] = B.nodes[i].child[0]; refs[2 * i + 1] = B.nodes[i].child[1]; }
        int want = params->smem_nodes < 1 ? 1 : (params->smem_nodes > nn ? nn : params->smem_nodes), bfs = 0;
        if (renumber(refs, 2, nn, want, new_of, old_of, &bfs)) return -1;
        for (int ni = 0; ni < nn; ni++) {
            const onode* t = &B.nodes[old_of[ni]];
            ora_bvh_node* d = &out[ni];
            d->child0 = t->child[0] >= 0 ? new_of[t->child[0]] : t->child[0];
            d->child1 = t->child[1] >= 0 ? new_of[t->child[1]] : t->child[1];
            for (int a = 0; a < 3; a++) {
                centre_extent(t->box[0].lo[a] - pad, t->box[0].hi[a] + pad, &d->c0[a], &d->e0[a]);
                centre_extent(t->box[1].lo[a] - pad, t->box[1].hi[a] + pad, &d->c1[a], &d->e1[a]);
            }
        }
        for (int i = 0; i < nn; i++) { refs[2 * i] = out[i].child0; refs[2 * i + 1] = out[i].child1; }
        if (depth_out) *depth_out = chain_depth(refs, 2, nn);
        if (bfs_nodes_out) *bfs_nodes_out = bfs;
        *nodes_out = out; *n_nodes_out = nn;
        free(refs); free(new_of); free(old_of);
        rc = 0;
    } else { /* R6 */
        typedef struct { int32_t child[4]; obox box[4]; } owide;
        owide* wide = (owide*)calloc((size_t)B.n_nodes + 1, sizeof(owide));
        typedef struct { int32_t bin, parent; int slot; } witem;
        witem* todo = (witem*)malloc(sizeof(witem) * (size_t)(B.n_nodes + 1));
        if (!wide || !todo) return -1;
        int nw = 0, nt = 0;
        todo[nt++] = (witem){0, -1, 0};
        while (nt) {
            const witem it = todo[--nt];
            int32_t sref[4]; obox sbox[4]; int ns = 2;
            sref[0] = B.nodes[it.bin].child[0]; sbox[0] = B.nodes[it.bin].box[0];
            sref[1] = B.nodes[it.bin].child[1]; sbox[1] = B.nodes[it.bin].box[1];
            while (ns < 4) {
                int pick = -1; double best = -1.0;
                for (int k = 0; k < ns; k++)
                    if (sref[k] >= 0 && box_area(&sbox[k]) > best) { best = box_area(&sbox[k]); pick = k; }
                if (pick < 0) break;
                const onode* t = &B.nodes[sref[pick]];
                for (int k = ns; k > pick + 1; k--) { sref[k] = sref[k - 1]; sbox[k] = sbox[k - 1]; }
                sref[pick] = t->child[0]; sbox[pick] = t->box[0];
                sref[pick + 1] = t->child[1]; sbox[pick + 1] = t->box[1];
                ns++;
            }
            const int me = nw++;
            if (it.parent >= 0) wide[it.parent].child[it.slot] = me;
            for (int k = 0; k < 4; k++) {
                if (k < ns) { wide[me].child[k] = sref[k]; wide[me].box[k] = sbox[k]; }
                else wide[me].child[k] = ORA_EMPTY;
            }
            for (int k = ns - 1; k >= 0; k--)
                if (sref[k] >= 0) todo[nt++] = (witem){sref[k], me, k};
        }
        free(todo);
        int32_t* refs = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)nw);
        int32_t* new_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)nw);
        int32_t* old_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)nw);
        ora_bvh_node4* out = (ora_bvh_node4*)calloc((size_t)nw, sizeof(ora_bvh_node4));
        if (!refs || !new_of || !old_of || !out) return -1;
        for (int i = 0; i < nw; i++) for (int k = 0; k < 4; k++) refs[4 * i + k] = wide[i].child[k];
        int want = params->smem_nodes < 1 ? 1 : (params->smem_nodes > nw ? nw : params->smem_nodes), bfs = 0;
        if (renumber(refs, 4, nw, want, new_of, old_of, &bfs)) return -1;
        for (int ni = 0; ni < nw; ni++) {
            const owide* t = &wide[old_of[ni]];
            ora_bvh_node4* d = &out[ni];
            float* cs[4] = {d->c0, d->c1, d->c2, d->c3};
            float* es[4] = {d->e0, d->e1, d->e2, d->e3};
            int32_t* rf[4] = {&d->child0, &d->child1, &d->child2, &d->child3};
            for (int k = 0; k < 4; k++) {
                const int32_t r = t->child[k];
                *rf[k] = (r >= 0 && r != ORA_EMPTY) ? new_of[r] : r;
                for (int a = 0; a < 3; a++) {
                    if (r == ORA_EMPTY) { cs[k][a] = 0.0f; es[k][a] = -1e30f; }
                    else centre_extent(t->box[k].lo[a] - pad, t->box[k].hi[a] + pad, &cs[k][a], &es[k][a]);
                }
            }
        }
        for (int i = 0; i < nw; i++) {
            refs[4 * i] = out[i].child0; refs[4 * i + 1] = out[i].child1; refs[4 * i + 2] = out[i].child2; refs[4 * i + 3] = out[i].child3;
        }
        if (depth_out) *depth_out = chain_depth(refs, 4, nw);
        if (bfs_nodes_out) *bfs_nodes_out = bfs;
        *nodes_out = out; *n_nodes_out = nw;
        free(refs); free(new_of); free(old_of); free(wide);
        rc = 0;
    }
    *tri_order_out = B.order;
    free(B.prims); free(B.nodes);
    return rc;
}

/* Quantised binary nodes.  Rules, all in IEEE double unless stated:
 *  Q1 grid: per axis lo = min, hi = max over the ROOT's non-empty child boxes of (double)c -/+ (double)e; ext = hi - lo
 *     (1.0 when not positive); margin = ext / 1024; q_lo = (float)(lo - margin); q_step = (float)((ext + 2 * margin) / 65000).
 *  Q2 plane indices of a child box: ql = floor(((double)c - (double)e - (double)q_lo) / (double)q_step) - 1,
 *     qh = ceil(((double)c + (double)e - (double)q_lo) / (double)q_step) + 1, both clamped to [0, 65535]; an EMPTY child
 *     gets ql = 65535, qh = 0 (never hit).  The quantised box contains the fp32 box with a full grid step to spare on
 *     each side, which covers the half step the fp32 plane-distance formula of the traversal can be off by.
 *  Q3 record: words 0..2 = child 0 (x, y, z: ql | qh << 16), words 3..5 = child 1, words 6, 7 = the child references. */
int ora_bvh_quantize(const ora_bvh_node* nodes, int n_nodes, uint32_t* q_out, float q_lo[3], float q_step[3]) {
    if (!nodes || n_nodes < 1 || !q_out) return -1;
    for (int a = 0; a < 3; a++) {
        double lo = 0.0, hi = 0.0;
        int have = 0;
        const float* cs[2] = {nodes[0].c0, nodes[0].c1};
        const float* es[2] = {nodes[0].e0, nodes[0].e1};
        const int32_t refs[2] = {nodes[0].child0, nodes[0].child1};
        for (int k = 0; k < 2; k++) {
            if (refs[k] == ORA_EMPTY) continue;
            const double l = (double)cs[k][a] - (double)es[k][a], h = (double)cs[k][a] + (double)es[k][a];
            if (!have || l < lo) lo = l;
            if (!have || h > hi) hi = h;
            have = 1;
        }
        double ext = hi - lo;
        if (!(ext > 0.0)) ext = 1.0;
        const double margin = ext / 1024.0;
        q_lo[a] = (float)(lo - margin);
        q_step[a] = (float)((ext + 2.0 * margin) / 65000.0);
    }
    for (int i = 0; i < n_nodes; i++) {
        const float* cs[2] = {nodes[i].c0, nodes[i].c1};
        const float* es[2] = {nodes[i].e0, nodes[i].e1};
        const int32_t refs[2] = {nodes[i].child0, nodes[i].child1};
        for (int k = 0; k < 2; k++)
            for (int a = 0; a < 3; a++) {
                long ql = 65535, qh = 0;
                if (refs[k] != ORA_EMPTY) {
                    const double plo = (double)cs[k][a] - (double)es[k][a], phi = (double)cs[k][a] + (double)es[k][a];
                    const double fl = floor((plo - (double)q_lo[a]) / (double)q_step[a]) - 1.0;
                    const double fh = ceil((phi - (double)q_lo[a]) / (double)q_step[a]) + 1.0;
                    ql = fl < 0.0 ? 0 : fl > 65535.0 ? 65535 : (long)fl;
                    qh = fh < 0.0 ? 0 : fh > 65535.0 ? 65535 : (long)fh;
                }
                q_out[(size_t)i * 8 + (size_t)k * 3 + (size_t)a] = (uint32_t)ql | ((uint32_t)qh << 16);
            }
        q_out[(size_t)i * 8 + 6] = (uint32_t)nodes[i].child0;
        q_out[(size_t)i * 8 + 7] = (uint32_t)nodes[i].child1;
    }
    return 0;
}
