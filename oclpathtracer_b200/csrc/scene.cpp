// scene.cpp -- host-side scene ingestion and output transform (no CUDA).
//
// Mirrors the host pieces of the reference's test/RaytraceTest.cpp that sit on
// either side of the hot path: loadModel (:87-198) and the PPM writer
// (:78-83, :277-287).  Cited line numbers are relative to that file.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

#include "host_internal.h"

namespace {

struct Mesh {
    float tag = 0.f;                 // m_albedo (:124) -- only compared against 0.5
    std::vector<int32_t> quad_idx;   // 4 per quad (:127-133)
    std::vector<float> vtx;          // 4 per vertex (:137-143)
    ptb_material mat;
};

ptb_float4 vertex(const Mesh& m, int32_t i) {
    return ptb_float4{m.vtx[4 * i + 0], m.vtx[4 * i + 1], m.vtx[4 * i + 2], 0.0f};  // :181-184, w := 0
}

template <class T>
T* dup(const std::vector<T>& v) {
    T* p = static_cast<T*>(std::malloc(sizeof(T) * (v.empty() ? 1 : v.size())));
    if (p && !v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" int ptb_load_model(const char* path, ptb_triangle** tris_out, int* n_tris, ptb_material** mats_out,
                              int* n_mats) {
    if (!path || !tris_out || !n_tris || !mats_out || !n_mats) return ptb::fail(PTB_E_INVALID, "ptb_load_model: null argument");
    std::ifstream in(path, std::ios::binary);
    if (!in) return ptb::fail(PTB_E_IO, "ptb_load_model: cannot open %s", path);
    std::vector<char> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (bytes.size() < 4) return ptb::fail(PTB_E_IO, "ptb_load_model: %s is empty", path);

    // little-endian int32/float32 stream (:117-143)
    size_t pos = 0;
    auto need = [&](size_t n) { return pos + n <= bytes.size(); };
    auto rd_i = [&]() { int32_t v; std::memcpy(&v, &bytes[pos], 4); pos += 4; return v; };
    auto rd_f = [&]() { float v; std::memcpy(&v, &bytes[pos], 4); pos += 4; return v; };

    const int32_t n_mesh = rd_i();
    if (n_mesh < 0 || n_mesh > (1 << 20)) return ptb::fail(PTB_E_IO, "ptb_load_model: bad mesh count %d", n_mesh);
    std::vector<Mesh> meshes(n_mesh);
    for (int32_t mi = 0; mi < n_mesh; ++mi) {
        Mesh& m = meshes[mi];
        if (!need(8)) return ptb::fail(PTB_E_IO, "ptb_load_model: truncated mesh header %d", mi);
        const int32_t nq = rd_i();
        m.tag = rd_f();
        if (nq < 0 || !need(size_t(nq) * 16 + 4)) return ptb::fail(PTB_E_IO, "ptb_load_model: truncated indices");
        m.quad_idx.resize(size_t(nq) * 4);
        for (auto& v : m.quad_idx) v = rd_i();
        const int32_t nv = rd_i();
        if (nv < 0 || !need(size_t(nv) * 16)) return ptb::fail(PTB_E_IO, "ptb_load_model: truncated vertices");
        m.vtx.resize(size_t(nv) * 4);
        for (auto& v : m.vtx) v = rd_f();
        for (int32_t v : m.quad_idx)
            if (v < 0 || v >= nv) return ptb::fail(PTB_E_IO, "ptb_load_model: vertex index out of range");

        // material by tag, then by mesh position (:145-176).  The reference leaves
        // roughness (meshes 0-4) and the padding uninitialised; they are zero here.
        std::memset(&m.mat, 0, sizeof m.mat);
        m.mat.type = PTB_DIFFUSE;
        if (m.tag != 0.5f) {
            m.mat.emissive = ptb_float4{30.f, 30.f, 30.f, 1.f};
            m.mat.albedo = ptb_float4{1.f, 1.f, 1.f, 1.f};
        } else {
            m.mat.emissive = ptb_float4{0.f, 0.f, 0.f, 1.f};
        }
        switch (mi) {
            case 0: case 1: case 2: m.mat.albedo = ptb_float4{0.7f, 0.7f, 0.7f, 1.0f}; break;
            case 3: m.mat.albedo = ptb_float4{0.6f, 0.0f, 0.0f, 1.0f}; break;
            case 4: m.mat.albedo = ptb_float4{0.0f, 0.6f, 0.0f, 1.0f}; break;
            case 5:
                m.mat.albedo = ptb_float4{0.5f, 0.35f, 0.05f, 0.0f};  // three-value initialiser: w = 0
                m.mat.roughness = 0.008f;
                m.mat.type = PTB_SPECULAR;
                break;
            default: break;
        }
    }

    // quad j -> (p1,p2,p3) and (p3,p4,p1), both tagged with the running quad id; one material per quad (:161-195)
    std::vector<ptb_triangle> tris;
    std::vector<ptb_material> mats;
    int32_t quad_id = 0;
    for (const Mesh& m : meshes) {
        for (size_t q = 0; q * 4 < m.quad_idx.size(); ++q, ++quad_id) {
            const ptb_float4 a = vertex(m, m.quad_idx[4 * q + 0]), b = vertex(m, m.quad_idx[4 * q + 1]);
            const ptb_float4 c = vertex(m, m.quad_idx[4 * q + 2]), d = vertex(m, m.quad_idx[4 * q + 3]);
            ptb_triangle t;
            std::memset(&t, 0, sizeof t);
            t.id = quad_id;
            t.p1 = a; t.p2 = b; t.p3 = c;
            tris.push_back(t);
            t.p1 = c; t.p2 = d; t.p3 = a;
            tris.push_back(t);
            mats.push_back(m.mat);
        }
    }
    if (tris.size() / 2 != mats.size()) return ptb::fail(PTB_E_IO, "ptb_load_model: triangle/material mismatch");  // :197
    *tris_out = dup(tris);
    *mats_out = dup(mats);
    if (!*tris_out || !*mats_out) return ptb::fail(PTB_E_NOMEM, "ptb_load_model: out of memory");
    *n_tris = int(tris.size());
    *n_mats = int(mats.size());
    return PTB_OK;
}

extern "C" void ptb_free(void* p) { std::free(p); }

// BUILD-DEFINED (config C5): bilinear k x k split of every quad.  Grid point
// (i,j):  s = i/k, t = j/k;  a = p1 + (p2-p1)*s;  b = p4 + (p3-p4)*s;  P = a + (b-a)*t
// (fp32, evaluated in exactly this order; shared grid points are therefore
// bit-identical between neighbours and the mesh stays watertight).
extern "C" int ptb_tessellate(const ptb_triangle* tris, int n_tris, int k, ptb_triangle** out, int* n_out) {
    if (!tris || n_tris < 2 || (n_tris & 1) || k < 1 || !out || !n_out)
        return ptb::fail(PTB_E_INVALID, "ptb_tessellate: bad arguments");
    const size_t total = size_t(n_tris) * k * k;
    ptb_triangle* res = static_cast<ptb_triangle*>(std::calloc(total, sizeof(ptb_triangle)));
    if (!res) return ptb::fail(PTB_E_NOMEM, "ptb_tessellate: out of memory");
    std::vector<float> grid(size_t(k + 1) * (k + 1) * 3);
    size_t w = 0;
    for (int q = 0; q < n_tris; q += 2) {
        const float* P1 = &tris[q].p1.x;
        const float* P2 = &tris[q].p2.x;
        const float* P3 = &tris[q].p3.x;
        const float* P4 = &tris[q + 1].p2.x;
        for (int j = 0; j <= k; ++j)
            for (int i = 0; i <= k; ++i) {
                const float s = float(i) / float(k), t = float(j) / float(k);
                float* g = &grid[(size_t(j) * (k + 1) + i) * 3];
                for (int c = 0; c < 3; ++c) {
                    const float a = P1[c] + (P2[c] - P1[c]) * s;
                    const float b = P4[c] + (P3[c] - P4[c]) * s;
                    g[c] = a + (b - a) * t;
                }
            }
        auto pt = [&](int i, int j) {
            const float* g = &grid[(size_t(j) * (k + 1) + i) * 3];
            return ptb_float4{g[0], g[1], g[2], 0.0f};
        };
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) {
                ptb_triangle& t1 = res[w++];
                ptb_triangle& t2 = res[w++];
                t1.p1 = pt(i, j); t1.p2 = pt(i + 1, j); t1.p3 = pt(i + 1, j + 1); t1.id = tris[q].id;
                t2.p1 = t1.p3; t2.p2 = pt(i, j + 1); t2.p3 = t1.p1; t2.id = tris[q].id;
            }
    }
    *out = res;
    *n_out = int(total);
    return PTB_OK;
}

extern "C" int ptb_light_from_quad(const ptb_triangle* tris, int n_tris, int quad, float p1[3], float ea[3],
                                   float eb[3]) {
    if (!tris || !p1 || !ea || !eb) return ptb::fail(PTB_E_INVALID, "ptb_light_from_quad: null argument");
    for (int q = 0; q + 1 < n_tris; q += 2) {
        if (tris[q].id != quad) continue;
        const float* a = &tris[q].p1.x;
        const float* b = &tris[q].p2.x;
        const float* d = &tris[q + 1].p2.x;
        for (int c = 0; c < 3; ++c) {
            p1[c] = a[c];
            ea[c] = b[c] - a[c];
            eb[c] = d[c] - a[c];
        }
        return PTB_OK;
    }
    return ptb::fail(PTB_E_NOTFOUND, "ptb_light_from_quad: no triangle pair with id %d", quad);
}

// RaytraceTest.cpp:78-83 f2c applied to sqrtf(v) (:283)
extern "C" int ptb_to_rgb8(const float* rgba, int n_pixels, uint8_t* rgb) {
    if (!rgba || !rgb || n_pixels < 0) return ptb::fail(PTB_E_INVALID, "ptb_to_rgb8: bad arguments");
    for (int i = 0; i < n_pixels; ++i)
        for (int c = 0; c < 3; ++c) {
            const float a = std::sqrt(rgba[4 * size_t(i) + c]) * 255.0f;
            // (int)a is undefined for NaN and for values outside int's range: pinned to 0, the same explicit test as the
            // device transform (pt_kernels.cuh f2c_sqrt), so the two produce identical bytes by construction
            int b = 0;
            if (a == a && a < 2147483648.0f && a >= -2147483648.0f) b = int(a);
            rgb[3 * size_t(i) + c] = uint8_t(b > 255 ? 255 : (b < 0 ? 0 : b));
        }
    return PTB_OK;
}

// RaytraceTest.cpp:277-287: "P3\n%d %d\n%d\n" then "%d %d %d " per pixel
extern "C" int ptb_write_ppm(const char* path, const float* rgba, int width, int height) {
    if (!path || !rgba || width <= 0 || height <= 0) return ptb::fail(PTB_E_INVALID, "ptb_write_ppm: bad arguments");
    std::vector<uint8_t> rgb(size_t(width) * height * 3);
    ptb_to_rgb8(rgba, width * height, rgb.data());
    FILE* f = std::fopen(path, "w");
    if (!f) return ptb::fail(PTB_E_IO, "ptb_write_ppm: cannot open %s", path);
    std::fprintf(f, "P3\n%d %d\n%d\n", width, height, 255);
    for (size_t i = 0; i < rgb.size(); i += 3) std::fprintf(f, "%d %d %d ", rgb[i], rgb[i + 1], rgb[i + 2]);
    std::fclose(f);
    return PTB_OK;
}
