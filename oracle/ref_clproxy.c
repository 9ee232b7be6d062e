/* ref_clproxy.c -- OpenCL option interposer + stopwatch for running the UNMODIFIED reference on NVIDIA's OpenCL.
 *
 * TEST INFRASTRUCTURE (oracle/).  The reference hard-codes the AMD-only build option "-O0"
 * (Adl/CL/AdlKernelUtilsCL.cpp:260); NVIDIA's compiler rejects it ("Don't understand command line argument
 * -O0"), so the kernel never builds and the test renders a black frame.  This shared object is named
 * libOpenCL.so and placed first on LD_LIBRARY_PATH: the reference's clew loader dlopen()s it, every cl* symbol
 * except clBuildProgram resolves through its DT_NEEDED dependency (the real ICD loader), and clBuildProgram
 * replaces "-O0" by the standard spelling of the same request, "-cl-opt-disable".  Nothing else is touched.
 *
 * It also keeps three stopwatches, because the reference's wall time is NOT its renderer's speed: the test asks for its
 * kernel through Device::getKernel, whose default is cacheKernel = false (Adl/Adl.h:166-167), so KernelManager::query
 * deletes and REBUILDS the program on every one of the 10 000 frames (Adl/AdlKernel.cpp:132-140).  Reported at exit (stderr
 * and $PTB_REF_CLPROXY_REPORT): total wall time inside clBuildProgram, total wall time from each clEnqueueNDRangeKernel to the
 * return of the clFinish that follows it (the kernel-only time of the reference's own launch loop,
 * Adl/CL/AdlKernelUtilsCL.cpp:473 + Adl/CL/AdlCL.cpp:282-285), and the call counts.                                    */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <time.h>

typedef int (*build_fn)(void*, unsigned, const void*, const char*, void (*)(void*, void*), void*);
typedef int (*ndrange_fn)(void*, void*, unsigned, const size_t*, const size_t*, const size_t*, unsigned, const void*, void*);
typedef int (*finish_fn)(void*);

static double g_build_s = 0.0, g_kernel_s = 0.0, g_pending_t0 = 0.0;
static long g_builds = 0, g_launches = 0, g_finishes = 0;
static int g_pending = 0;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void report(void) {
    char line[512];
    snprintf(line, sizeof line,
             "{\"clBuildProgram_calls\": %ld, \"clBuildProgram_total_s\": %.3f, \"kernel_launches\": %ld, "
             "\"kernel_enqueue_to_finish_total_s\": %.3f, \"kernel_ms_per_launch\": %.4f, \"clFinish_calls\": %ld}",
             g_builds, g_build_s, g_launches, g_kernel_s, g_launches ? 1e3 * g_kernel_s / (double)g_launches : 0.0, g_finishes);
    fprintf(stderr, "ref_clproxy: %s\n", line);
    const char* path = getenv("PTB_REF_CLPROXY_REPORT");
    if (path) {
        FILE* f = fopen(path, "w");
        if (f) { fprintf(f, "%s\n", line); fclose(f); }
    }
}
__attribute__((constructor)) static void on_load(void) { atexit(report); }

int clEnqueueNDRangeKernel(void* queue, void* kernel, unsigned dim, const size_t* gwo, const size_t* gws, const size_t* lws,
                           unsigned n_wait, const void* wait, void* event) {
    static ndrange_fn real = 0;
    if (!real) real = (ndrange_fn)dlsym(RTLD_NEXT, "clEnqueueNDRangeKernel");
    if (!real) return -9999;
    if (!g_pending) { g_pending = 1; g_pending_t0 = now_s(); }
    g_launches++;
    return real(queue, kernel, dim, gwo, gws, lws, n_wait, wait, event);
}

int clFinish(void* queue) {
    static finish_fn real = 0;
    if (!real) real = (finish_fn)dlsym(RTLD_NEXT, "clFinish");
    if (!real) return -9999;
    const int rc = real(queue);
    g_finishes++;
    if (g_pending) { g_kernel_s += now_s() - g_pending_t0; g_pending = 0; }
    return rc;
}

int clBuildProgram(void* program, unsigned num_devices, const void* device_list, const char* options,
                   void (*notify)(void*, void*), void* user_data) {
    static build_fn real = 0;
    if (!real) real = (build_fn)dlsym(RTLD_NEXT, "clBuildProgram");
    if (!real) {
        fprintf(stderr, "ref_clproxy: no real clBuildProgram\n");
        return -9999;
    }
    char fixed[1024];
    fixed[0] = 0;
    if (options) {
        const char* p = options;
        size_t n = 0;
        while (*p && n + 32 < sizeof fixed) {
            if (strncmp(p, "-O0", 3) == 0 && (p[3] == 0 || p[3] == ' ')) {
                n += (size_t)snprintf(fixed + n, sizeof fixed - n, "-cl-opt-disable");
                p += 3;
            } else {
                fixed[n++] = *p++;
                fixed[n] = 0;
            }
        }
    }
    if (getenv("PTB_REF_CLPROXY_VERBOSE")) fprintf(stderr, "ref_clproxy: clBuildProgram options \"%s\" -> \"%s\"\n", options ? options : "", fixed);
    const double t0 = now_s();
    const int rc = real(program, num_devices, device_list, options ? fixed : 0, notify, user_data);
    g_build_s += now_s() - t0;
    g_builds++;
    return rc;
}
