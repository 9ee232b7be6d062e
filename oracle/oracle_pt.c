/* oracle_pt.c -- CPU ORACLE (test infrastructure; see oracle_pt.h header note).
 *
 * Plain-C restatement of /root/reference/test/ClKernels/GenerateColors.cl and of
 * the host pieces of /root/reference/test/RaytraceTest.cpp that feed it.
 * PARITY: pinned against the unmodified reference's own 10000-frame output from a
 * B200 (tests/golden/reference_raycast_b200_opencl.npz, rRMSE 9.95e-4 <= 1e-3; see
 * oracle_pt.h), plus the integer-RNG known answers and the scene checksum.
 *
 * ---------------------------------------------------------------------------
 * Numerics contract (where OpenCL C leaves the arithmetic open, this file is the
 * definition; the CUDA path reproduces it bit for bit):
 *   N1. float = IEEE binary32, round-to-nearest-even, denormals kept, NO
 *       contraction of a*b+c in any expression the reference writes (build with
 *       -ffp-contract=off; nvcc --fmad=false).  Expressions are evaluated
 *       left-to-right exactly as GenerateColors.cl spells them.
 *   N2. dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z
 *       cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
 *                    (the host's own cross, RaytraceTest.cpp:19-28)
 *       normalize(v) = v * (1.0f / sqrtf(dot(v,v)))     (one IEEE div, one IEEE sqrt)
 *       max(x,y) = (x < y) ? y : x                       (OpenCL C 6.12.4, NaN kept in x)
 *       pow(x, 2.0f) in distributionGGX = x*x
 *   N3. sin/cos/tan/pow are BUILD-DEFINED polynomial kernels (ora_sincos,
 *       ora_pow below): explicit fmaf() steps only, so host and device agree
 *       exactly; measured max error vs libm double: sin/cos <= 1 ulp on [0, 2pi],
 *       pow correctly rounded to < 0.5000001 ulp (double intermediate).
 *   N4. BUILD-DEFINED BVH slab test uses explicit fmaf and fminf/fmaxf (below,
 *       bvh_slab: centre / half-extent boxes); everything pinned by the reference stays unfused.
 * ---------------------------------------------------------------------------
 */
#include "oracle_pt.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_DIFFUSE 1  /* GenerateColors.cl:3 */
#define ORA_SPECULAR 2 /* GenerateColors.cl:4 */
#define ORA_TWO_PI 6.28318530718f /* GenerateColors.cl:9  */
#define ORA_INV_PI 0.31830988618f /* GenerateColors.cl:10 */

typedef struct {
    float x, y, z;
} v3;

static inline v3 V3(float x, float y, float z) {
    v3 r = {x, y, z};
    return r;
}
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3s(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 neg3(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } /* N2 */
static inline v3 cross3(v3 a, v3 b) {                                                /* N2 */
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 normalize3(v3 v) { /* N2 */
    float inv = 1.0f / sqrtf(dot3(v, v));
    return V3(v.x * inv, v.y * inv, v.z * inv);
}
static inline float cl_max(float x, float y) { return (x < y) ? y : x; } /* N2 */

static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* ---- N3: deterministic transcendental kernels ------------------------------ */

/* sin and cos of x, |x| <= 64.  Cody-Waite reduction by pi/2 in four fmaf steps
 * (constants = 2 x Cephes DP1..3 plus the residual), then Cephes single-precision minimax
 * polynomials on [-pi/4, pi/4], Horner form in fmaf.                          */
void ora_sincos(float x, float* s_out, float* c_out) {
    float kf = floorf(x * 0.636619747f + 0.5f);
    int k = (int)kf;
    float r = fmaf(kf, -1.5703125f, x);
    r = fmaf(kf, -4.837512969970703125e-4f, r);
    r = fmaf(kf, -7.54978995489188e-8f, r);
    r = fmaf(kf, 1.7151245100058819e-15f, r); /* pi/2 - (C1+C2+C3) = -1.715e-15 */
    float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    float sn = fmaf(ps * z, r, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    float cs = fmaf(pc, z * z, fmaf(z, -0.5f, 1.0f));
    switch (k & 3) {
        case 0: *s_out = sn; *c_out = cs; break;
        case 1: *s_out = cs; *c_out = -sn; break;
        case 2: *s_out = -sn; *c_out = -cs; break;
        default: *s_out = -cs; *c_out = sn; break;
    }
}

float ora_tan(float x) {
    float s, c;
    ora_sincos(x, &s, &c);
    return s / c;
}

/* pow(x, y) for x >= 0 (the path only raises non-negative radiance to 2.2 and
 * 1/2.2).  Double-precision exp2(y*log2(x)) built from + - * / only (no FMA, no
 * libm) so it is bit-reproducible; result rounded once to float.
 * x = 0 -> 0, x = +inf -> +inf, NaN or x < 0 -> NaN.                            */
float ora_pow(float x, float y) {
    if (x != x) return x;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return 0.0f;
    if (x > 3.402823466e38f) return x; /* +inf */
    double xd = (double)x;             /* exact; float denormals are normal doubles */
    uint64_t b;
    memcpy(&b, &xd, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8); /* [1,2) */
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    p = p * s2;
    double ln_m = 2.0 * s + (2.0 * s) * p;
    double log2x = (double)e + ln_m * 1.4426950408889634;
    double t = (double)y * log2x;
    if (t > 130.0) return INFINITY;
    if (t < -160.0) return 0.0f;
    double n = floor(t + 0.5);
    double g = (t - n) * 0.6931471805599453;
    double q = 1.0 / 479001600.0; /* 1/12! */
    q = q * g + 1.0 / 39916800.0;
    q = q * g + 1.0 / 3628800.0;
    q = q * g + 1.0 / 362880.0;
    q = q * g + 1.0 / 40320.0;
    q = q * g + 1.0 / 5040.0;
    q = q * g + 1.0 / 720.0;
    q = q * g + 1.0 / 120.0;
    q = q * g + 1.0 / 24.0;
    q = q * g + 1.0 / 6.0;
    q = q * g + 0.5;
    q = q * g + 1.0;
    q = q * g + 1.0;
    uint64_t sb = (uint64_t)((int64_t)n + 1023) << 52;
    double scale;
    memcpy(&scale, &sb, 8);
    return (float)(q * scale);
}

void ora_sincos_array(const float* x, int n, float* s, float* c) {
    for (int i = 0; i < n; i++) ora_sincos(x[i], &s[i], &c[i]);
}
void ora_pow_array(const float* x, int n, float y, float* out) {
    for (int i = 0; i < n; i++) out[i] = ora_pow(x[i], y);
}

/* ---- RNG: GenerateColors.cl:47-71 ------------------------------------------ */

/* GenerateColors.cl:47-59 -- the Wang branch is '#if 0'; what runs is the LCG */
uint32_t ora_hash_uint32(uint32_t x) { return 1103515245u * x + 12345u; }

/* GenerateColors.cl:61-71 */
float ora_random_float(uint32_t* seed) {
    uint32_t s = *seed;
    s = (s ^ 61u) ^ (s >> 16);
    s = s + (s << 3);
    s = s ^ (s >> 4);
    s = s * 0x27d4eb2du;
    s = s ^ (s >> 15);
    s = 1103515245u * s + 12345u;
    *seed = s;
    return (float)s * 2.3283064365386963e-10f;
}

void ora_rng_kat(uint32_t gid, uint32_t frame, int n, uint32_t* states, float* values) {
    uint32_t seed = gid + ora_hash_uint32(frame); /* GenerateColors.cl:308 */
    for (int i = 0; i < n; i++) {
        values[i] = ora_random_float(&seed);
        states[i] = seed;
    }
}

/* ---- scene loading: RaytraceTest.cpp:87-198 -------------------------------- */

int ora_load_model(const char* path, ora_triangle* tris, int tri_cap, ora_material* mats, int mat_cap,
                   int* n_tris, int* n_mats) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size <= 0) {
        fclose(f);
        return -2;
    }
    char* buf = (char*)malloc((size_t)size);
    if (fread(buf, 1, (size_t)size, f) != (size_t)size) {
        fclose(f);
        free(buf);
        return -3;
    }
    fclose(f);
    const int32_t* c = (const int32_t*)buf; /* :117 */
    const int32_t* end = (const int32_t*)(buf + size);
    int n_mesh = *c++;
    int nt = 0, nm = 0, id = 0, rc = 0;
    for (int i = 0; i < n_mesh && rc == 0; i++) {
        if (c + 2 > end) { rc = -4; break; }
        int nf = *c++;                         /* :123 */
        float tag;
        memcpy(&tag, c++, 4);                  /* :124 */
        const int32_t* idx = c;
        c += 4 * (size_t)nf;                   /* :127-133 */
        if (c + 1 > end) { rc = -4; break; }
        int nv = *c++;                         /* :135 */
        const float* vtx = (const float*)c;
        c += 4 * (size_t)nv;                   /* :137-143 */
        if (c > end) { rc = -4; break; }

        ora_material m;
        memset(&m, 0, sizeof m);               /* roughness/padding are uninitialised in the reference: 0 here */
        m.type = ORA_DIFFUSE;                  /* :145 */
        if (tag != 0.5f) {                     /* :147-151 */
            m.emissive = (ora_float4){30.0f, 30.0f, 30.0f, 1.0f};
            m.albedo = (ora_float4){1.0f, 1.0f, 1.0f, 1.0f};
        } else {
            m.emissive = (ora_float4){0.0f, 0.0f, 0.0f, 1.0f}; /* :153 */
        }
        if (i == 0 || i == 1 || i == 2) m.albedo = (ora_float4){0.7f, 0.7f, 0.7f, 1.0f}; /* :165-166 */
        if (i == 3) m.albedo = (ora_float4){0.6f, 0.0f, 0.0f, 1.0f};                       /* :167-168 */
        if (i == 4) m.albedo = (ora_float4){0.0f, 0.6f, 0.0f, 1.0f};                       /* :169-170 */
        if (i == 5) {                                                                      /* :171-176 */
            m.albedo = (ora_float4){0.5f, 0.35f, 0.05f, 0.0f}; /* 3-element init: w = 0 */
            m.roughness = 0.008f;
            m.type = ORA_SPECULAR;
        }
        for (int j = 0; j < nf; j++) {         /* :179-194 */
            if (nt + 2 > tri_cap || nm + 1 > mat_cap) { rc = -5; break; }
            ora_float4 p[4];
            for (int k = 0; k < 4; k++) {
                int vi = idx[4 * j + k];
                if (vi < 0 || vi >= nv) { rc = -6; break; }
                p[k] = (ora_float4){vtx[4 * vi + 0], vtx[4 * vi + 1], vtx[4 * vi + 2], 0.0f};
            }
            if (rc) break;
            ora_triangle t1, t2;
            memset(&t1, 0, sizeof t1);
            memset(&t2, 0, sizeof t2);
            t1.p1 = p[0]; t1.p2 = p[1]; t1.p3 = p[2]; t1.id = id; /* :186 */
            t2.p1 = p[2]; t2.p2 = p[3]; t2.p3 = p[0]; t2.id = id; /* :187 */
            tris[nt++] = t1;
            tris[nt++] = t2;
            mats[nm++] = m;
            id++;
        }
    }
    free(buf);
    *n_tris = nt;
    *n_mats = nm;
    if (rc) return rc;
    return (nt / 2 == nm) ? 0 : -7; /* :197 */
}

/* BUILD-DEFINED (C5): every consecutive triangle pair (p1,p2,p3),(p3,p4,p1) is a
 * quad; split it into k x k sub-quads, each again (q1,q2,q3),(q3,q4,q1) with the
 * quad's id.  Vertex (i,j): a = p1 + (p2-p1)*s; b = p4 + (p3-p4)*s;
 * P = a + (b-a)*t with s = (float)i/(float)k, t = (float)j/(float)k.           */
static v3 tess_point(v3 p1, v3 p2, v3 p3, v3 p4, int i, int j, int k) {
    float s = (float)i / (float)k, t = (float)j / (float)k;
    v3 a = add3(p1, mul3s(sub3(p2, p1), s));
    v3 b = add3(p4, mul3s(sub3(p3, p4), s));
    return add3(a, mul3s(sub3(b, a), t));
}
static ora_float4 F4(v3 v) {
    ora_float4 r = {v.x, v.y, v.z, 0.0f};
    return r;
}
int ora_tessellate(const ora_triangle* tris, int n_tris, int k, ora_triangle* out, int out_cap) {
    int n = 0;
    for (int q = 0; q + 1 < n_tris; q += 2) {
        v3 p1 = V3(tris[q].p1.x, tris[q].p1.y, tris[q].p1.z);
        v3 p2 = V3(tris[q].p2.x, tris[q].p2.y, tris[q].p2.z);
        v3 p3 = V3(tris[q].p3.x, tris[q].p3.y, tris[q].p3.z);
        v3 p4 = V3(tris[q + 1].p2.x, tris[q + 1].p2.y, tris[q + 1].p2.z);
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) {
                if (n + 2 > out_cap) return -1;
                v3 q1 = tess_point(p1, p2, p3, p4, i, j, k);
                v3 q2 = tess_point(p1, p2, p3, p4, i + 1, j, k);
                v3 q3 = tess_point(p1, p2, p3, p4, i + 1, j + 1, k);
                v3 q4 = tess_point(p1, p2, p3, p4, i, j + 1, k);
                ora_triangle t1, t2;
                memset(&t1, 0, sizeof t1);
                memset(&t2, 0, sizeof t2);
                t1.p1 = F4(q1); t1.p2 = F4(q2); t1.p3 = F4(q3); t1.id = tris[q].id;
                t2.p1 = F4(q3); t2.p2 = F4(q4); t2.p3 = F4(q1); t2.id = tris[q].id;
                out[n++] = t1;
                out[n++] = t2;
            }
    }
    return n;
}

/* light parallelogram of quad `quad`: first triangle pair carrying that id */
void ora_light_from_quad(const ora_triangle* tris, int n_tris, int quad, float p1[3], float ea[3], float eb[3]) {
    for (int q = 0; q + 1 < n_tris; q += 2) {
        if (tris[q].id != quad) continue;
        p1[0] = tris[q].p1.x; p1[1] = tris[q].p1.y; p1[2] = tris[q].p1.z;
        ea[0] = tris[q].p2.x - tris[q].p1.x; ea[1] = tris[q].p2.y - tris[q].p1.y; ea[2] = tris[q].p2.z - tris[q].p1.z;
        eb[0] = tris[q + 1].p2.x - tris[q].p1.x; eb[1] = tris[q + 1].p2.y - tris[q].p1.y; eb[2] = tris[q + 1].p2.z - tris[q].p1.z;
        return;
    }
}

/* ---- scene queries ---------------------------------------------------------- */

typedef struct {
    uint64_t closest, any, nodes, tests, tu, tv, tt, acc;
} qctr;

typedef struct {
    float t, u, v;
    int tri;
} qhit;

static inline v3 P3(ora_float4 p) { return V3(p.x, p.y, p.z); }

/* GenerateColors.cl:89-125 up to and including the computation of t.
 * Returns 1 when every reject test of the reference passed and t > 0.0f
 * (the `t < tmax` half of :125 is applied by the caller).                      */
static inline int mt_core(v3 o, v3 d, const ora_triangle* tr, float* t_out, float* u_out, float* v_out, qctr* c) {
    c->tests++;
    v3 p1 = P3(tr->p1);
    v3 e1 = sub3(P3(tr->p2), p1);                 /* :92 */
    v3 e2 = sub3(P3(tr->p3), p1);                 /* :93 */
    v3 pvec = cross3(d, e2);                      /* :96 */
    float det = dot3(e1, pvec);                   /* :97 */
    if (det < 1e-8f || -det > 1e-8f) return 0;    /* :100 */
    c->tu++;
    float inv_det = 1.0f / det;                   /* :105 */
    v3 tvec = sub3(o, p1);                        /* :106 */
    float u = dot3(tvec, pvec) * inv_det;         /* :107 */
    if (u < 0.0f || u > 1.0f) return 0;           /* :109 */
    c->tv++;
    v3 qvec = cross3(tvec, e1);                   /* :114 */
    float v = dot3(d, qvec) * inv_det;            /* :115 */
    if (v < 0.0f || u + v > 1.0f) return 0;       /* :117 */
    c->tt++;
    float t = dot3(e2, qvec) * inv_det;           /* :122 */
    if (!(t > 0.0f)) return 0;                    /* :125 first half */
    *t_out = t;
    *u_out = u;
    *v_out = v;
    return 1;
}

/* GenerateColors.cl:137-154: ascending index, strict t < tmax, tmax shrinks */
static int closest_brute(const ora_triangle* tris, int n, v3 o, v3 d, qhit* h, qctr* c) {
    float tmax = 1e20f; /* :139 */
    int hit = 0;
    c->closest++;
    for (int i = 0; i < n; i++) {
        float t, u, v;
        if (mt_core(o, d, &tris[i], &t, &u, &v, c) && t < tmax) {
            c->acc++;
            tmax = t;
            h->t = t; h->u = u; h->v = v; h->tri = i;
            hit = 1;
        }
    }
    return hit;
}

static int any_brute(const ora_triangle* tris, int n, v3 o, v3 d, float tmax, int* blocker, qctr* c) {
    c->any++;
    for (int i = 0; i < n; i++) {
        float t, u, v;
        if (mt_core(o, d, &tris[i], &t, &u, &v, c) && t < tmax) {
            c->acc++;
            *blocker = i;
            return 1;
        }
    }
    *blocker = -1;
    return 0;
}

/* N4: BUILD-DEFINED slab test.  Returns entry distance in *tn. */
static inline float safe_rcp(float d) {
    if (fabsf(d) > 1e-20f) return 1.0f / d;
    return signbit(d) ? -1e20f : 1e20f;
}
/* box = [c - e, c + e]; per axis t_c = c*invd - o*invd, t_near = t_c - e*|invd|, t_far = t_c + e*|invd| */
static inline int bvh_slab(const float c[3], const float e[3], const float invd[3], const float ood[3],
                           float best_t, float* tn_out) {
    float tcx = fmaf(c[0], invd[0], -ood[0]), tcy = fmaf(c[1], invd[1], -ood[1]), tcz = fmaf(c[2], invd[2], -ood[2]);
    float ax = fabsf(invd[0]), ay = fabsf(invd[1]), az = fabsf(invd[2]);
    float nx = fmaf(-e[0], ax, tcx), ny = fmaf(-e[1], ay, tcy), nz = fmaf(-e[2], az, tcz);
    float fx = fmaf(e[0], ax, tcx), fy = fmaf(e[1], ay, tcy), fz = fmaf(e[2], az, tcz);
    float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, 0.0f));
    float tf = fminf(fminf(fx, fy), fminf(fz, best_t));
    *tn_out = tn;
    return tn <= tf;
}


/* N4q: BUILD-DEFINED slab test on a QUANTISED box (ora_bvh.qnodes; DESIGN.md section 4 "Quantised binary nodes").  Per ray and
 * axis A = q_step * invd and B = fma(-2^23, A, fma(q_lo, invd, -(o * invd))); the distance to grid plane q is ONE fma,
 * t = fma(2^23 + q, A, B), the factor being the float whose bits are 0x4B000000 | q (exact for q < 2^23).  The near plane
 * of an axis is the lo plane when invd >= 0, else the hi plane. */
typedef struct { float A[3], B[3]; int lo_near[3]; } qray;
static inline qray qray_init(const ora_bvh* bvh, v3 o, const float invd[3]) {
    qray q;
    const float oo[3] = {o.x, o.y, o.z};
    for (int a = 0; a < 3; a++) {
        q.A[a] = bvh->q_step[a] * invd[a];
        q.B[a] = fmaf(-8388608.0f, q.A[a], fmaf(bvh->q_lo[a], invd[a], -(oo[a] * invd[a])));
        q.lo_near[a] = invd[a] >= 0.0f;
    }
    return q;
}
static inline float qplane(uint32_t idx, float A, float B) {
    union { uint32_t u; float f; } m;
    m.u = 0x4B000000u | idx;
    return fmaf(m.f, A, B);
}
static inline int bvh_qslab(const uint32_t w[3], const qray* q, float best_t, float* tn_out) {
    float n[3], f[3];
    for (int a = 0; a < 3; a++) {
        const uint32_t lo = w[a] & 0xffffu, hi = w[a] >> 16;
        n[a] = qplane(q->lo_near[a] ? lo : hi, q->A[a], q->B[a]);
        f[a] = qplane(q->lo_near[a] ? hi : lo, q->A[a], q->B[a]);
    }
    float tn = fmaxf(fmaxf(n[0], n[1]), fmaxf(n[2], 0.0f));
    float tf = fminf(fminf(f[0], f[1]), fminf(f[2], best_t));
    *tn_out = tn;
    return tn <= tf;
}

/* BUILD-DEFINED traversal over the 4-wide tree (the specification of node-visit counts):
 *  - `cur` >= 0: fetch the node (visits++), slab-test its four child boxes against [0, best_t].
 *      closest-hit: descend into the NEAREST hit child = smallest key (bits of its entry distance tn with the
 *        two low mantissa bits replaced by the slot index); push the other hit children so that the lower
 *        slot pops first, each with its entry distance tn;
 *      any-hit: descend into the hit child with the lowest slot index, push the others so that the lower
 *        slot pops first.
 *  - `cur` < 0: leaf; test its triangles in stored order.  Closest: accept when t < best_t, or t == best_t
 *    and the triangle's index is lower than the current winner's (== the reference's "first index wins").
 *    Any: return at the first triangle with 0 < t < tmax.
 *  - pop: closest-hit entries whose stored entry distance exceeds best_t are discarded unvisited.         */
#define ORA_STACK 256

static int bvh_query4(const ora_triangle* tris, const ora_bvh* bvh, v3 o, v3 d, float tmax, int any_hit, qhit* h,
                     uint32_t* visits, qctr* c) {
    float invd[3] = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
    float ood[3] = {o.x * invd[0], o.y * invd[1], o.z * invd[2]};
    int32_t stack_ref[ORA_STACK];
    float stack_tn[ORA_STACK];
    int sp = 0;
    float best_t = tmax;
    int best_tri = -1;
    float best_u = 0, best_v = 0;
    int32_t cur = 0;
    if (any_hit) c->any++; else c->closest++;
    for (;;) {
        int descend = 0;
        if (cur >= 0) {
            const ora_bvh_node4* nd = &((const ora_bvh_node4*)bvh->nodes)[cur];
            (*visits)++;
            c->nodes++;
            const float* cs[4] = {nd->c0, nd->c1, nd->c2, nd->c3};
            const float* es[4] = {nd->e0, nd->e1, nd->e2, nd->e3};
            const int32_t refs[4] = {nd->child0, nd->child1, nd->child2, nd->child3};
            uint32_t key[4];
            float tnv[4];
            int hit[4];
            for (int k = 0; k < 4; k++) {
                hit[k] = bvh_slab(cs[k], es[k], invd, ood, best_t, &tnv[k]);
                key[k] = hit[k] ? ((f2u(tnv[k]) & ~3u) | (uint32_t)k) : 0xffffffffu;
            }
            if (hit[0] || hit[1] || hit[2] || hit[3]) {
                if (any_hit) {
                    int first = -1;
                    for (int k = 3; k >= 0; k--) {
                        if (!hit[k]) continue;
                        if (first >= 0) { stack_ref[sp] = refs[first]; stack_tn[sp] = 0.0f; sp++; }
                        first = k;
                    }
                    cur = refs[first];
                } else {
                    uint32_t kmin = 0xffffffffu;
                    for (int k = 0; k < 4; k++) if (key[k] < kmin) kmin = key[k];
                    for (int k = 3; k >= 0; k--)
                        if (hit[k] && key[k] != kmin) { stack_ref[sp] = refs[k]; stack_tn[sp] = tnv[k]; sp++; }
                    cur = refs[kmin & 3u];
                }
                descend = 1;
            }
        } else {
            uint32_t code = (uint32_t)(~cur);
            int first = (int)(code >> 3), count = (int)(code & 7u) + 1;
            for (int k = first; k < first + count; k++) {
                int idx = bvh->tri_order[k];
                float t, u, v;
                if (!mt_core(o, d, &tris[idx], &t, &u, &v, c)) continue;
                if (any_hit) {
                    if (t < best_t) {
                        c->acc++;
                        h->t = t; h->u = u; h->v = v; h->tri = idx;
                        return 1;
                    }
                } else if (t < best_t || (t == best_t && best_tri >= 0 && idx < best_tri)) {
                    c->acc++;
                    best_t = t; best_u = u; best_v = v; best_tri = idx;
                }
            }
        }
        if (descend) continue;
        /* pop */
        for (;;) {
            if (sp == 0) {
                if (best_tri >= 0 && !any_hit) {
                    h->t = best_t; h->u = best_u; h->v = best_v; h->tri = best_tri;
                    return 1;
                }
                h->tri = -1;
                return 0;
            }
            sp--;
            cur = stack_ref[sp];
            if (any_hit || stack_tn[sp] <= best_t) break;
        }
    }
}

/* BINARY tree (scenes traversed from L2/HBM): fetch the node (visits++), slab-test both child boxes against
 * [0, best_t]; both hit -> descend into the nearer (child 1 only if tn1 < tn0), push the other with its entry
 * distance; one hit -> descend.  Leaves, acceptance and pop as for the 4-wide tree.                        */
static int bvh_query2(const ora_triangle* tris, const ora_bvh* bvh, v3 o, v3 d, float tmax, int any_hit, qhit* h,
                     uint32_t* visits, qctr* c) {
    float invd[3] = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
    float ood[3] = {o.x * invd[0], o.y * invd[1], o.z * invd[2]};
    const qray qr = bvh->qnodes ? qray_init(bvh, o, invd) : (qray){{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    int32_t stack_ref[ORA_STACK];
    float stack_tn[ORA_STACK];
    int sp = 0;
    float best_t = tmax;
    int best_tri = -1;
    float best_u = 0, best_v = 0;
    int32_t cur = 0;
    if (any_hit) c->any++; else c->closest++;
    for (;;) {
        if (cur == 0x7fffffff) {
            /* empty child: nothing */
        } else if (cur >= 0) {
            const ora_bvh_node* nd = &((const ora_bvh_node*)bvh->nodes)[cur];
            (*visits)++;
            c->nodes++;
            float tn0, tn1;
            int h0, h1;
            if (bvh->qnodes) { /* the product walks the quantised encoding of this tree */
                const uint32_t* w = bvh->qnodes + (size_t)cur * 8;
                h0 = nd->child0 != 0x7fffffff && bvh_qslab(w, &qr, best_t, &tn0);
                h1 = nd->child1 != 0x7fffffff && bvh_qslab(w + 3, &qr, best_t, &tn1);
            } else {
                h0 = nd->child0 != 0x7fffffff && bvh_slab(nd->c0, nd->e0, invd, ood, best_t, &tn0);
                h1 = nd->child1 != 0x7fffffff && bvh_slab(nd->c1, nd->e1, invd, ood, best_t, &tn1);
            }
            if (h0 && h1) {
                if (tn1 < tn0) {
                    stack_ref[sp] = nd->child0; stack_tn[sp] = tn0; sp++;
                    cur = nd->child1;
                } else {
                    stack_ref[sp] = nd->child1; stack_tn[sp] = tn1; sp++;
                    cur = nd->child0;
                }
                continue;
            } else if (h0) {
                cur = nd->child0;
                continue;
            } else if (h1) {
                cur = nd->child1;
                continue;
            }
        } else {
            uint32_t code = (uint32_t)(~cur);
            int first = (int)(code >> 3), count = (int)(code & 7u) + 1;
            for (int k = first; k < first + count; k++) {
                int idx = bvh->tri_order[k];
                float t, u, v;
                if (!mt_core(o, d, &tris[idx], &t, &u, &v, c)) continue;
                if (any_hit) {
                    if (t < best_t) {
                        c->acc++;
                        h->t = t; h->u = u; h->v = v; h->tri = idx;
                        return 1;
                    }
                } else if (t < best_t || (t == best_t && best_tri >= 0 && idx < best_tri)) {
                    c->acc++;
                    best_t = t; best_u = u; best_v = v; best_tri = idx;
                }
            }
        }
        /* pop */
        for (;;) {
            if (sp == 0) {
                if (best_tri >= 0 && !any_hit) {
                    h->t = best_t; h->u = best_u; h->v = best_v; h->tri = best_tri;
                    return 1;
                }
                h->tri = -1;
                return 0;
            }
            sp--;
            cur = stack_ref[sp];
            if (stack_tn[sp] <= best_t) break;
        }
    }
}

/* FLAT form (scenes of <= 32 leaves and <= 64 triangles; BUILD-DEFINED, the specification of the product's flat query):
 *  phase A: slab-test EVERY leaf box against [0, tmax] (tmax = 1e20 for closest-hit) in record order; the leaves that
 *           pass contribute their triangle masks; visits = number of leaves that passed;
 *  phase B: test the selected triangles in ascending position of the ordered array.  Closest: accept when t < best_t, or
 *           t == best_t and the triangle's caller index is lower (== the reference's "first index wins"); no leaf is
 *           culled by best_t.  Any: return at the first triangle with 0 < t < tmax.                                   */
static int bvh_query_flat(const ora_triangle* tris, const ora_bvh* bvh, v3 o, v3 d, float tmax, int any_hit, qhit* h,
                          uint32_t* visits, qctr* c) {
    float invd[3] = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
    float ood[3] = {o.x * invd[0], o.y * invd[1], o.z * invd[2]};
    const ora_bvh_leafbox* lb = (const ora_bvh_leafbox*)bvh->nodes;
    uint64_t tm = 0;
    if (any_hit) c->any++; else c->closest++;
    for (int k = 0; k < bvh->n_nodes; k++) {
        float tn;
        if (bvh_slab(lb[k].c, lb[k].e, invd, ood, tmax, &tn)) {
            tm |= (uint64_t)lb[k].mask_lo | ((uint64_t)lb[k].mask_hi << 32);
            (*visits)++;
            c->nodes++;
        }
    }
    float best_t = tmax, best_u = 0, best_v = 0;
    int best_tri = -1;
    while (tm) {
        const int pos = __builtin_ctzll(tm);
        tm &= tm - 1;
        const int idx = bvh->tri_order[pos];
        float t, u, v;
        if (!mt_core(o, d, &tris[idx], &t, &u, &v, c)) continue;
        if (any_hit) {
            if (t < best_t) {
                c->acc++;
                h->t = t; h->u = u; h->v = v; h->tri = idx;
                return 1;
            }
        } else if (t < best_t || (t == best_t && best_tri >= 0 && idx < best_tri)) {
            c->acc++;
            best_t = t; best_u = u; best_v = v; best_tri = idx;
        }
    }
    if (best_tri >= 0 && !any_hit) {
        h->t = best_t; h->u = best_u; h->v = best_v; h->tri = best_tri;
        return 1;
    }
    h->tri = -1;
    return 0;
}

static int bvh_query(const ora_triangle* tris, const ora_bvh* bvh, v3 o, v3 d, float tmax, int any_hit, qhit* h,
                     uint32_t* visits, qctr* c) {
    if (bvh->width == 1) return bvh_query_flat(tris, bvh, o, d, tmax, any_hit, h, visits, c);
    if (bvh->width == 4) return bvh_query4(tris, bvh, o, d, tmax, any_hit, h, visits, c);
    return bvh_query2(tris, bvh, o, d, tmax, any_hit, h, visits, c);
}

/* query context */
typedef struct {
    const ora_triangle* tris;
    int n_tris;
    const ora_bvh* bvh; /* NULL -> brute force */
    qctr c;
} qctx;

static int q_closest(qctx* q, v3 o, v3 d, qhit* h, uint32_t* visits) {
    if (q->bvh) return bvh_query(q->tris, q->bvh, o, d, 1e20f, 0, h, visits, &q->c);
    return closest_brute(q->tris, q->n_tris, o, d, h, &q->c);
}
static int q_any(qctx* q, v3 o, v3 d, float tmax, int* blocker, uint32_t* visits) {
    if (q->bvh) {
        qhit h;
        int r = bvh_query(q->tris, q->bvh, o, d, tmax, 1, &h, visits, &q->c);
        *blocker = r ? h.tri : -1;
        return r;
    }
    return any_brute(q->tris, q->n_tris, o, d, tmax, blocker, &q->c);
}

void ora_trace(const ora_triangle* tris, int n_tris, const ora_bvh* bvh, int use_bvh, int any_hit, int n_rays,
               const float* o, const float* d, const float* tmax, int32_t* out_tri, float* out_t, float* out_u,
               float* out_v, uint32_t* out_visits, uint32_t* out_tests) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n_rays; i++) {
        qctx q;
        memset(&q, 0, sizeof q);
        q.tris = tris; q.n_tris = n_tris; q.bvh = use_bvh ? bvh : NULL;
        v3 ro = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
        v3 rd = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        uint32_t visits = 0;
        qhit h;
        memset(&h, 0, sizeof h);
        h.tri = -1;
        int hit;
        if (any_hit) {
            if (q.bvh) {
                hit = bvh_query(tris, bvh, ro, rd, tmax[i], 1, &h, &visits, &q.c);
            } else {
                /* brute any-hit needs t/u/v of the first accepted triangle too */
                hit = 0;
                q.c.any++;
                for (int k = 0; k < n_tris; k++) {
                    float t, u, v;
                    if (mt_core(ro, rd, &tris[k], &t, &u, &v, &q.c) && t < tmax[i]) {
                        h.t = t; h.u = u; h.v = v; h.tri = k;
                        hit = 1;
                        break;
                    }
                }
            }
        } else {
            hit = q_closest(&q, ro, rd, &h, &visits);
        }
        out_tri[i] = hit ? h.tri : -1;
        if (out_t) out_t[i] = hit ? h.t : 0.0f;
        if (out_u) out_u[i] = hit ? h.u : 0.0f;
        if (out_v) out_v[i] = hit ? h.v : 0.0f;
        if (out_visits) out_visits[i] = visits;
        if (out_tests) out_tests[i] = (uint32_t)q.c.tests;
    }
}

/* GenerateColors.cl:127-130: record fields of the accepted hit */
static void hit_point_normal(const ora_triangle* tr, v3 o, v3 d, const qhit* h, v3* p, v3* n) {
    v3 p1 = P3(tr->p1);
    v3 e1 = sub3(P3(tr->p2), p1);
    v3 e2 = sub3(P3(tr->p3), p1);
    v3 norm = cross3(e2, e1);                                     /* :123 */
    *p = add3(o, mul3s(d, h->t));                                 /* :128 */
    float w = 1.0f - h->u - h->v;
    v3 s = add3(add3(mul3s(norm, h->u), mul3s(norm, h->v)), mul3s(norm, w)); /* :130 */
    *n = normalize3(s);
}

/* ---- camera: GenerateColors.cl:263-288, getRay :73-87 ----------------------- */

typedef struct {
    v3 o, d;
} ray_t;

/* :73-87 -- invDir/sign are computed by the reference but never read */
static ray_t get_ray(v3 origin, v3 dir) {
    ray_t r;
    r.o = origin;
    r.d = normalize3(dir); /* :75 */
    return r;
}

static ray_t generate_ray(int xc, int yc, int width, int height, uint32_t* seed) {
    float inv_w = 1.0f / (float)width, inv_h = 1.0f / (float)height; /* :265 */
    float aspect = (float)width / (float)height;                     /* :266 */
    float fov = (float)((60.0f * 3.14159265358979323846) / 180.0f);  /* :267, M_PI is double in OpenCL C */
    float angle = ora_tan(0.5f * fov);                               /* :268 */
    const v3 eye = V3(0.0f, 2.75f, 4.0f);                            /* :270 */
    const v3 center = add3(eye, V3(0.0f, 0.0f, -1.0f));              /* :271 */
    const v3 up = V3(0.0f, 1.0f, 0.0f);                              /* :272 */
    const v3 view = normalize3(sub3(center, eye));                   /* :274 */
    const v3 hol = normalize3(cross3(view, up));                     /* :275 */
    const v3 upd = normalize3(cross3(hol, view));                    /* :276 */
    float x = (float)xc + ora_random_float(seed) - 0.5f;             /* :278 */
    float y = (float)yc + ora_random_float(seed) - 0.5f;             /* :279 */
    x = (2.0f * ((x + 0.5f) * inv_w) - 1) * angle * aspect;          /* :281 */
    y = -(1.0f - 2.0f * ((y + 0.5f) * inv_h)) * angle;               /* :282 */
    v3 dir = normalize3(add3(add3(mul3s(hol, x), mul3s(upd, -1.0f * y)), view)); /* :284 */
    v3 aimed = add3(eye, mul3s(dir, 4.0f));                          /* :285 */
    return get_ray(eye, normalize3(sub3(aimed, eye)));               /* :287 */
}

void ora_generate_ray(int gi, int gj, int width, int height, uint32_t* seed, float o[3], float d[3]) {
    ray_t r = generate_ray(gi, gj, width, height, seed);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z;
    d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z;
}

/* ---- BSDF sampling: GenerateColors.cl:156-221 ------------------------------- */

static v3 reflect3(v3 v, v3 n) { /* :156-159 */
    float k = 2.0f * dot3(v, n);
    return add3(neg3(v), mul3s(n, k));
}

static v3 frame_combine(v3 n, float phi, float sin_theta, float cos_theta) {
    float sp, cp;
    ora_sincos(phi, &sp, &cp);
    v3 axis = fabsf(n.x) > 0.001f ? V3(0.0f, 1.0f, 0.0f) : V3(1.0f, 0.0f, 0.0f); /* :167 / :187 */
    v3 t = normalize3(cross3(axis, n));                                           /* :168 / :188 */
    v3 s = cross3(n, t);                                                          /* :169 / :189 */
    v3 a = mul3s(mul3s(s, cp), sin_theta);
    v3 b = mul3s(mul3s(t, sp), sin_theta);
    v3 c = mul3s(n, cos_theta);
    return normalize3(add3(add3(a, b), c));                                       /* :171 / :191 */
}

static v3 sample_hemisphere_cosine(v3 n, uint32_t* seed) { /* :161-172 */
    float phi = ORA_TWO_PI * ora_random_float(seed);
    float s2 = ora_random_float(seed);
    float sin_theta = sqrtf(s2);
    return frame_combine(n, phi, sin_theta, sqrtf(1.0f - s2));
}

static float distribution_ggx(float cos_theta, float roughness) { /* :174-178 */
    float r2 = roughness * roughness;
    float x = cos_theta * cos_theta * (r2 - 1.0f) + 1.0f;
    return r2 * ORA_INV_PI / (x * x); /* pow(x, 2.0f) := x*x (N2) */
}

static v3 sample_ggx(v3 n, float roughness, float* cos_theta, uint32_t* seed) { /* :180-192 */
    float phi = ORA_TWO_PI * ora_random_float(seed);
    float xi = ora_random_float(seed);
    *cos_theta = sqrtf((1.0f - xi) / (xi * (roughness * roughness - 1.0f) + 1.0f));
    float sin_theta = sqrtf(cl_max(0.0f, 1.0f - (*cos_theta) * (*cos_theta)));
    return frame_combine(n, phi, sin_theta, *cos_theta);
}

/* :195-221.  Returns f in *f; wi, pdf by pointer. */
static ora_float4 brdf(v3 wo, v3* wi, float* pdf, v3 normal, const ora_material* mat, uint32_t* seed) {
    if (mat->type == ORA_DIFFUSE) {
        *wi = sample_hemisphere_cosine(normal, seed);
        *pdf = dot3(*wi, normal) * ORA_INV_PI;
        return (ora_float4){mat->albedo.x * ORA_INV_PI, mat->albedo.y * ORA_INV_PI, mat->albedo.z * ORA_INV_PI,
                            mat->albedo.w * ORA_INV_PI};
    } else if (mat->type == ORA_SPECULAR) {
        float cos_theta;
        v3 wh = sample_ggx(normal, mat->roughness, &cos_theta, seed);
        *wi = reflect3(wo, wh);
        if (dot3(*wi, normal) * dot3(wo, normal) < 0.0f) return (ora_float4){0, 0, 0, 0}; /* :211 */
        float D = distribution_ggx(cos_theta, mat->roughness);
        *pdf = D * cos_theta / (4.0f * dot3(wo, wh));                                      /* :215 */
        float k = D / (4.0f * dot3(*wi, normal) * dot3(wo, normal));                       /* :217 */
        return (ora_float4){k * mat->albedo.x * 2.0f, k * mat->albedo.y * 2.0f, k * mat->albedo.z * 2.0f,
                            k * mat->albedo.w * 2.0f};
    }
    return (ora_float4){0.0f, 0.0f, 0.0f, 1.0f}; /* :220 */
}

/* ---- per-sample integrators -------------------------------------------------- */

typedef struct {
    int32_t tri, quad;
    uint32_t t_bits, visits_primary, visits_secondary, count, id_hash;
} sstat;

static inline void hash_id(sstat* s, int tri) { s->id_hash = s->id_hash * 31u + (uint32_t)(tri + 2); }

/* GenerateColors.cl:223-261 with BOUNCES -> max_depth */
static ora_float4 trace_rays(qctx* q, const ora_material* mats, ray_t r, uint32_t* seed, int max_depth, sstat* st) {
    ora_float4 radiance = {0, 0, 0, 0};                    /* :225 */
    ora_float4 mask = {1.0f, 1.0f, 1.0f, 1.0f};            /* :226 */
    const ora_float4 bg = {0.45f, 0.45f, 0.45f, 1.0f};     /* :227 */
    for (int i = 0; i < max_depth; i++) {                  /* :229 */
        qhit h;
        uint32_t visits = 0;
        int hit = q_closest(q, r.o, r.d, &h, &visits);
        if (i == 0) {
            st->visits_primary = visits;
            st->tri = hit ? h.tri : -1;
            st->quad = hit ? q->tris[h.tri].id : -1;
            st->t_bits = hit ? f2u(h.t) : 0u;
        } else {
            st->visits_secondary += visits;
            hash_id(st, hit ? h.tri : -1);
        }
        st->count++;
        if (!hit) {                                        /* :233-237 */
            radiance.x += mask.x * cl_max(bg.x, 0.0f);
            radiance.y += mask.y * cl_max(bg.y, 0.0f);
            radiance.z += mask.z * cl_max(bg.z, 0.0f);
            radiance.w += mask.w * cl_max(bg.w, 0.0f);
            break;
        }
        const ora_triangle* tr = &q->tris[h.tri];
        v3 p, n;
        hit_point_normal(tr, r.o, r.d, &h, &p, &n);
        const ora_material* m = &mats[tr->id];             /* :239 */
        radiance.x += mask.x * m->emissive.x * 3.0f;       /* :241 */
        radiance.y += mask.y * m->emissive.y * 3.0f;
        radiance.z += mask.z * m->emissive.z * 3.0f;
        radiance.w += mask.w * m->emissive.w * 3.0f;
        n = dot3(n, r.d) < 0.0f ? n : mul3s(n, -1.0f);     /* :243 */
        v3 wi = V3(0, 0, 0);
        v3 wo = neg3(r.d);                                 /* :246 */
        float pdf = 0.0f;                                  /* :247 */
        ora_float4 color = brdf(wo, &wi, &pdf, n, m, seed);/* :249 */
        if (pdf <= 0.0f) break;                            /* :251 */
        float dw = dot3(wi, n);
        mask.x *= color.x * dw / pdf;                      /* :253-255 */
        mask.y *= color.y * dw / pdf;
        mask.z *= color.z * dw / pdf;
        mask.w *= color.w * dw / pdf;
        r = get_ray(add3(p, mul3s(wi, 0.01f)), wi);        /* :257 */
    }
    radiance.x = cl_max(radiance.x, 0.0f);                 /* :260 */
    radiance.y = cl_max(radiance.y, 0.0f);
    radiance.z = cl_max(radiance.z, 0.0f);
    radiance.w = cl_max(radiance.w, 0.0f);
    return radiance;
}

/* BUILD-DEFINED C1: primary ray only; colour = albedo of the hit quad, bg on miss */
static ora_float4 sample_primary(qctx* q, const ora_material* mats, ray_t r, sstat* st) {
    qhit h;
    uint32_t visits = 0;
    int hit = q_closest(q, r.o, r.d, &h, &visits);
    st->visits_primary = visits;
    st->tri = hit ? h.tri : -1;
    st->quad = hit ? q->tris[h.tri].id : -1;
    st->t_bits = hit ? f2u(h.t) : 0u;
    st->count = 1;
    if (!hit) return (ora_float4){0.45f, 0.45f, 0.45f, 1.0f};
    const ora_material* m = &mats[q->tris[h.tri].id];
    return (ora_float4){m->albedo.x, m->albedo.y, m->albedo.z, 1.0f};
}

/* BUILD-DEFINED C2: primary + ns cosine-hemisphere any-hit rays (same sampler,
 * same 0.01 origin offset and re-normalisation as the reference's bounce,
 * GenerateColors.cl:161-172,:257); value = unoccluded / ns; miss = 1.          */
static ora_float4 sample_ao(qctx* q, ray_t r, uint32_t* seed, int ns, float max_dist, sstat* st) {
    qhit h;
    uint32_t visits = 0;
    int hit = q_closest(q, r.o, r.d, &h, &visits);
    st->visits_primary = visits;
    st->tri = hit ? h.tri : -1;
    st->quad = hit ? q->tris[h.tri].id : -1;
    st->t_bits = hit ? f2u(h.t) : 0u;
    if (!hit) return (ora_float4){1.0f, 1.0f, 1.0f, 1.0f};
    v3 p, n;
    hit_point_normal(&q->tris[h.tri], r.o, r.d, &h, &p, &n);
    n = dot3(n, r.d) < 0.0f ? n : mul3s(n, -1.0f);
    uint32_t open = 0;
    for (int k = 0; k < ns; k++) {
        v3 wi = sample_hemisphere_cosine(n, seed);
        ray_t s = get_ray(add3(p, mul3s(wi, 0.01f)), wi);
        int blocker;
        uint32_t v2 = 0;
        int occ = q_any(q, s.o, s.d, max_dist, &blocker, &v2);
        st->visits_secondary += v2;
        hash_id(st, blocker);
        if (!occ) open++;
    }
    st->count = open;
    float v = (float)open / (float)ns;
    return (ora_float4){v, v, v, 1.0f};
}

/* BUILD-DEFINED C3: emission seen directly + one shadow ray to a uniformly
 * sampled point of the light parallelogram.                                   */
static ora_float4 sample_direct(qctx* q, const ora_material* mats, const ora_params* prm, ray_t r, uint32_t* seed,
                                sstat* st) {
    qhit h;
    uint32_t visits = 0;
    int hit = q_closest(q, r.o, r.d, &h, &visits);
    st->visits_primary = visits;
    st->tri = hit ? h.tri : -1;
    st->quad = hit ? q->tris[h.tri].id : -1;
    st->t_bits = hit ? f2u(h.t) : 0u;
    if (!hit) return (ora_float4){0.45f, 0.45f, 0.45f, 1.0f};
    const ora_triangle* tr = &q->tris[h.tri];
    const ora_material* m = &mats[tr->id];
    v3 p, n;
    hit_point_normal(tr, r.o, r.d, &h, &p, &n);
    ora_float4 c = {1.0f * m->emissive.x * 3.0f, 1.0f * m->emissive.y * 3.0f, 1.0f * m->emissive.z * 3.0f, 1.0f};
    n = dot3(n, r.d) < 0.0f ? n : mul3s(n, -1.0f);
    v3 wo = neg3(r.d);
    float xi1 = ora_random_float(seed);
    float xi2 = ora_random_float(seed);
    v3 lp = V3(prm->light_p1[0], prm->light_p1[1], prm->light_p1[2]);
    v3 ea = V3(prm->light_ea[0], prm->light_ea[1], prm->light_ea[2]);
    v3 eb = V3(prm->light_eb[0], prm->light_eb[1], prm->light_eb[2]);
    v3 P = add3(add3(lp, mul3s(ea, xi1)), mul3s(eb, xi2));
    v3 L = sub3(P, p);
    float dist2 = dot3(L, L);
    float dist = sqrtf(dist2);
    v3 wi = normalize3(L);
    v3 lc = cross3(ea, eb);
    float area = sqrtf(dot3(lc, lc));
    v3 nl = normalize3(lc);
    float cos_s = dot3(wi, n);
    float cos_l = -dot3(wi, nl);
    if (cos_s > 0.0f && cos_l > 0.0f) {
        ray_t s = get_ray(add3(p, mul3s(wi, 0.01f)), wi);
        int blocker;
        uint32_t v2 = 0;
        int occ = q_any(q, s.o, s.d, dist - 0.02f, &blocker, &v2);
        st->visits_secondary += v2;
        hash_id(st, blocker);
        if (!occ) {
            st->count = 1;
            const ora_material* lm = &mats[prm->light_quad];
            float fx, fy, fz;
            if (m->type == ORA_SPECULAR) {
                v3 wh = normalize3(add3(wo, wi));
                float D = distribution_ggx(dot3(n, wh), m->roughness);
                float k = D / (4.0f * dot3(wi, n) * dot3(wo, n));
                fx = k * m->albedo.x * 2.0f; fy = k * m->albedo.y * 2.0f; fz = k * m->albedo.z * 2.0f;
            } else {
                fx = m->albedo.x * ORA_INV_PI; fy = m->albedo.y * ORA_INV_PI; fz = m->albedo.z * ORA_INV_PI;
            }
            float G = cos_s * cos_l / dist2;
            c.x += fx * (lm->emissive.x * 3.0f) * G * area;
            c.y += fy * (lm->emissive.y * 3.0f) * G * area;
            c.z += fz * (lm->emissive.z * 3.0f) * G * area;
        }
    }
    c.x = cl_max(c.x, 0.0f);
    c.y = cl_max(c.y, 0.0f);
    c.z = cl_max(c.z, 0.0f);
    return c;
}

/* ---- accumulate: GenerateColors.cl:290-321 ----------------------------------- */

static ora_float4 gamma_correct(ora_float4 v) { /* :290-294 */
    const float g = 1.0f / 2.2f;
    return (ora_float4){ora_pow(v.x, g), ora_pow(v.y, g), ora_pow(v.z, g), 1.0f};
}
static ora_float4 read_from_gamma(ora_float4 v) { /* :296-300 */
    return (ora_float4){ora_pow(v.x, 2.2f), ora_pow(v.y, 2.2f), ora_pow(v.z, 2.2f), 1.0f};
}

static int owns(const ora_params* p, int gid) {
    if (p->shard_count <= 1) return 1;
    return (gid / p->shard_block) % p->shard_count == p->shard_index;
}
static int local_index(const ora_params* p, int gid) {
    if (p->shard_count <= 1) return gid;
    return (gid / (p->shard_block * p->shard_count)) * p->shard_block + gid % p->shard_block;
}
int ora_local_pixel_count(const ora_params* p) {
    int n = p->width * p->height;
    if (p->shard_count <= 1) return n;
    int c = 0;
    for (int b = p->shard_index; b * p->shard_block < n; b += p->shard_count) {
        int lo = b * p->shard_block, hi = lo + p->shard_block;
        c += (hi < n ? hi : n) - lo;
    }
    return c;
}

int ora_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ora_render(const ora_params* prm, const ora_triangle* tris, int n_tris, const ora_material* mats, int n_mats,
               const ora_bvh* bvh, float* fb, ora_pixel_stats* stats, ora_counters* counters) {
    (void)n_mats;
    const int W = prm->width, H = prm->height, npix = W * H;
    if (prm->use_bvh && !bvh) return -1;
    int nthreads = prm->n_threads > 0 ? prm->n_threads : ora_max_threads();
    qctr total;
    memset(&total, 0, sizeof total);
    uint64_t samples = 0;
#pragma omp parallel num_threads(nthreads)
    {
        qctx q;
        memset(&q, 0, sizeof q);
        q.tris = tris; q.n_tris = n_tris; q.bvh = prm->use_bvh ? bvh : NULL;
        uint64_t my_samples = 0;
#pragma omp for schedule(dynamic, 256)
        for (int gid = 0; gid < npix; gid++) {
            if (!owns(prm, gid)) continue;
            const int li = local_index(prm, gid);
            const int gi = gid % W, gj = gid / W;          /* GenerateColors.cl:305-306 */
            ora_float4 sum = {0, 0, 0, 0};
            sstat st;
            for (int f = prm->first_frame; f < prm->first_frame + prm->n_frames; f++) {
                uint32_t seed = (uint32_t)gid + ora_hash_uint32((uint32_t)f); /* :308 */
                ray_t r = generate_ray(gi, gj, W, H, &seed);                  /* :310 */
                memset(&st, 0, sizeof st);
                uint64_t tests0 = q.c.tests;
                ora_float4 c;
                switch (prm->mode) {
                    case ORA_MODE_PRIMARY: c = sample_primary(&q, mats, r, &st); break;
                    case ORA_MODE_AO: c = sample_ao(&q, r, &seed, prm->ao_samples, prm->ao_max_dist, &st); break;
                    case ORA_MODE_DIRECT: c = sample_direct(&q, mats, prm, r, &seed, &st); break;
                    default: c = trace_rays(&q, mats, r, &seed, prm->max_depth, &st); break; /* :312 */
                }
                my_samples++;
                if (stats && f == prm->first_frame + prm->n_frames - 1) {
                    ora_pixel_stats* ps = &stats[li];
                    ps->tri = st.tri; ps->quad = st.quad; ps->t_bits = st.t_bits;
                    ps->visits_primary = st.visits_primary; ps->visits_secondary = st.visits_secondary;
                    ps->count = st.count; ps->id_hash = st.id_hash;
                    ps->tri_tests = (uint32_t)(q.c.tests - tests0);
                }
                float* dst = &fb[4 * (size_t)li];
                if (prm->accum == ORA_ACCUM_REFERENCE) {
                    ora_float4 out;
                    if (f == 0) {                                            /* :314-317 */
                        out = gamma_correct(c);
                    } else {                                                 /* :318-321 */
                        ora_float4 prev = read_from_gamma((ora_float4){dst[0], dst[1], dst[2], dst[3]});
                        float zm1 = (float)(f - 1), z = (float)f;
                        ora_float4 mean = {(prev.x * zm1 + c.x) / z, (prev.y * zm1 + c.y) / z,
                                           (prev.z * zm1 + c.z) / z, (prev.w * zm1 + c.w) / z};
                        out = gamma_correct(mean);
                    }
                    dst[0] = out.x; dst[1] = out.y; dst[2] = out.z; dst[3] = out.w;
                } else {
                    sum.x += c.x; sum.y += c.y; sum.z += c.z;
                }
            }
            if (prm->accum == ORA_ACCUM_LINEAR) {
                float nf = (float)prm->n_frames;
                float* dst = &fb[4 * (size_t)li];
                dst[0] = sum.x / nf; dst[1] = sum.y / nf; dst[2] = sum.z / nf; dst[3] = 1.0f;
            }
        }
#pragma omp critical
        {
            total.closest += q.c.closest; total.any += q.c.any; total.nodes += q.c.nodes;
            total.tests += q.c.tests; total.tu += q.c.tu; total.tv += q.c.tv; total.tt += q.c.tt;
            total.acc += q.c.acc;
            samples += my_samples;
        }
    }
    if (counters) {
        counters->rays_closest = total.closest; counters->rays_any = total.any; counters->nodes = total.nodes;
        counters->tri_tests = total.tests; counters->samples = samples;
        counters->tri_u = total.tu; counters->tri_v = total.tv; counters->tri_t = total.tt;
        counters->tri_accept = total.acc;
    }
    return 0;
}

/* RaytraceTest.cpp:78-83 f2c and :280-285 */
void ora_to_rgb8(const float* fb, int n_pixels, uint8_t* rgb) {
    for (int i = 0; i < n_pixels; i++)
        for (int k = 0; k < 3; k++) {
            float a = sqrtf(fb[4 * i + k]); /* :283 */
            a *= 255;                       /* :80 */
            int b = (int)a;                 /* :81 */
            if (!(a == a)) b = 0;           /* (int)NaN is undefined in C; pin it to 0 */
            if (b > 255) b = 255;
            if (b < 0) b = 0;
            rgb[3 * i + k] = (uint8_t)b;
        }
}
