// raycast_main.cpp -- the reference's DeviceTest.RayCast flow (test/RaytraceTest.cpp:202-291)
// re-expressed against libptb200.so through the ADL-shaped shim: load cornellbox.bin, upload
// the 64-byte Triangle/Material records by map/unmap, launch GenerateColors once per frame
// with int4{W, H, frame, -}, wait, read the gamma-space framebuffer back, write the P3 PPM.
//
//   ptb_raycast [scene.bin] [out.ppm] [dimension=512] [frames=10000] [gpus=1 | PTB_GPUS]
// gpus > 1: the other GPUs of the box become helpers of device 0 (DeviceUtils::Config::m_nGpus -> ptb_device_add_helper);
// the loop below does not change and the image is bit-identical to one GPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Adl/Adl.h"
#include "SharedHeader.h"

using namespace adl;

typedef ptb_triangle Triangle;  // RaytraceTest.cpp:61-76
typedef ptb_material Material;  // RaytraceTest.cpp:50-59

int main(int argc, char** argv) {
    const char* scene = argc > 1 ? argv[1] : "../test/cornellbox.bin";  // :90
    const char* out_path = argc > 2 ? argv[2] : nullptr;
    const int dimension = argc > 3 ? std::atoi(argv[3]) : 512;          // :219
    const unsigned frames = argc > 4 ? unsigned(std::atoi(argv[4])) : 10000u;  // :250

    // fixture: DeviceTest::SetUp (test/TestBase.h:16-30)
    if (!adl::init(TYPE_CUDA)) {
        std::fprintf(stderr, "no CUDA device: %s\n", ptb_last_error());
        return 2;
    }
    DeviceUtils::Config cfg;
    if (argc > 5) cfg.m_nGpus = std::atoi(argv[5]);
    Device* m_d = DeviceUtils::allocate(TYPE_CUDA, cfg);
    if (!m_d) {
        std::fprintf(stderr, "device allocation failed: %s\n", ptb_last_error());
        return 2;
    }

    // loadModel (:87-198)
    Triangle* tri_rec = nullptr;
    Material* mat_rec = nullptr;
    int n_tri = 0, n_mat = 0;
    if (ptb_load_model(scene, &tri_rec, &n_tri, &mat_rec, &n_mat) != PTB_OK) {
        std::printf("Error loading model !!\n");  // :212
        std::fprintf(stderr, "%s\n", ptb_last_error());
        DeviceUtils::deallocate(m_d);
        return 1;
    }
    std::vector<Triangle> triangles(tri_rec, tri_rec + n_tri);
    std::vector<Material> materials(mat_rec, mat_rec + n_mat);
    ptb_free(tri_rec);
    ptb_free(mat_rec);

    int rc = 0;
    {
        // buffers (:216-223); the reference passes byte counts as element counts here -- kept
        Buffer<cl_float4> frameBuff(m_d, adlu64(dimension) * dimension);
        Buffer<Triangle> tBuffer(m_d, triangles.size() * sizeof(Triangle));
        Buffer<Material> materialBuffer(m_d, materials.size() * sizeof(Material));

        // upload through mapped pointers (:225-246)
        Triangle* tb = tBuffer.getHostPtr();
        Material* mb = materialBuffer.getHostPtr();
        DeviceUtils::waitForCompletion(m_d);
        for (size_t i = 0; i < triangles.size(); i++) tb[i] = triangles[i];
        for (size_t i = 0; i < materials.size(); i++) mb[i] = materials[i];
        tBuffer.returnHostPtr(tb);
        materialBuffer.returnHostPtr(mb);
        DeviceUtils::waitForCompletion(m_d);

        // one launch per sample (:248-268)
        const auto loop_t0 = std::chrono::steady_clock::now();
        unsigned frameCount = 0;
        while (frameCount != frames) {
            ptb_int4 res;
            res.x = dimension; res.y = dimension; res.z = int(frameCount++); res.w = 0;
            BufferInfo bInfo[] = {BufferInfo(&tBuffer), BufferInfo(&materialBuffer), BufferInfo(&frameBuff)};
            const Kernel* kernel = m_d->getKernel(SELECT_KERNELPATH1(m_d, "../test/", "GenerateColors"), "GenerateColors");
            if (!kernel) {
                std::fprintf(stderr, "getKernel failed: %s\n", ptb_last_error());
                rc = 3;
                break;
            }
            Launcher launcher(m_d, kernel);
            launcher.setBuffers(bInfo, sizeof(bInfo) / sizeof(BufferInfo));
            launcher.setConst(res);
            launcher.launch1D(dimension * dimension);
            DeviceUtils::waitForCompletion(m_d);
        }
        const double loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - loop_t0).count();

        // save the rendering (:270-290)
        if (rc == 0) {
            cl_float4* h = frameBuff.getHostPtr();
            DeviceUtils::waitForCompletion(m_d);
            char path[256];
            if (out_path) {
                std::snprintf(path, sizeof path, "%s", out_path);
            } else {
                char dev_name[128];
                m_d->getDeviceVersion(dev_name);
                for (char* c = dev_name; *c; ++c)
                    if (*c == ' ' || *c == '/') *c = '_';
                std::snprintf(path, sizeof path, "rayCastAo_%s.ppm", dev_name);  // getFilePath, TestBase.h:45-51
            }
            if (!h || ptb_write_ppm(path, &h[0].x, dimension, dimension) != PTB_OK) {
                std::fprintf(stderr, "writing %s failed: %s\n", path, ptb_last_error());
                rc = 4;
            } else {
                std::printf("%u frames of %dx%d on %d GPU(s) -> %s   (launch loop %.3f s = %.4f ms per frame; device creation, scene upload and the PPM write are outside)\n",
                            frames, dimension, dimension, 1 + m_d->m_nHelpers, path, loop_s, 1e3 * loop_s / (frames ? frames : 1));
            }
            DeviceUtils::waitForCompletion(m_d);
        }
    }  // buffers are released before the device, as DeviceUtils::deallocate expects
    DeviceUtils::deallocate(m_d);
    adl::quit(TYPE_CUDA);
    return rc;
}
