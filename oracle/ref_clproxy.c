/* ref_clproxy.c -- OpenCL option interposer for running the UNMODIFIED reference on NVIDIA's OpenCL.
 *
 * TEST INFRASTRUCTURE (oracle/).  The reference hard-codes the AMD-only build option "-O0"
 * (Adl/CL/AdlKernelUtilsCL.cpp:260); NVIDIA's compiler rejects it ("Don't understand command line argument
 * -O0"), so the kernel never builds and the test renders a black frame.  This shared object is named
 * libOpenCL.so and placed first on LD_LIBRARY_PATH: the reference's clew loader dlopen()s it, every cl* symbol
 * except clBuildProgram resolves through its DT_NEEDED dependency (the real ICD loader), and clBuildProgram
 * replaces "-O0" by the standard spelling of the same request, "-cl-opt-disable".  Nothing else is touched.   */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int (*build_fn)(void*, unsigned, const void*, const char*, void (*)(void*, void*), void*);

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
    return real(program, num_devices, device_list, options ? fixed : 0, notify, user_data);
}
