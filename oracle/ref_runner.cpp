// ref_runner.cpp -- launcher for the UNMODIFIED reference test binary (oracle/_ref/adlTest64).
//
// TEST INFRASTRUCTURE (lives under oracle/).  The reference's DeviceTest.RayCast opens
// "../test/ClKernels/GenerateColors.cl" and "../test/cornellbox.bin" relative to its working directory.
// /root/reference does not exist on the GPU box, so oracle/Makefile embeds those two reference files into the
// binary as data (ld -r -b binary, from where they lie under /root/reference); this launcher writes them into a
// scratch tree, changes into <scratch>/build and then calls the reference's own main() (test/main.cpp, compiled
// with -Dmain=ref_main).  No reference code is modified or copied into the repository.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <unistd.h>

extern "C" {
extern const char _binary_GenerateColors_cl_start[], _binary_GenerateColors_cl_end[];
extern const char _binary_cornellbox_bin_start[], _binary_cornellbox_bin_end[];
}
int ref_main(int argc, char* argv[]);

static bool put(const std::string& path, const char* b, const char* e) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(b, 1, size_t(e - b), f) == size_t(e - b);
    std::fclose(f);
    return ok;
}

int main(int argc, char* argv[]) {
    const char* env = std::getenv("PTB_REF_WORKDIR");
    std::string root = env ? env : "/tmp/ptb_ref_run";
    mkdir(root.c_str(), 0775);
    mkdir((root + "/test").c_str(), 0775);
    mkdir((root + "/test/ClKernels").c_str(), 0775);
    mkdir((root + "/build").c_str(), 0775);
    if (!put(root + "/test/ClKernels/GenerateColors.cl", _binary_GenerateColors_cl_start, _binary_GenerateColors_cl_end) ||
        !put(root + "/test/cornellbox.bin", _binary_cornellbox_bin_start, _binary_cornellbox_bin_end) ||
        chdir((root + "/build").c_str()) != 0) {
        std::fprintf(stderr, "ref_runner: cannot prepare %s\n", root.c_str());
        return 2;
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = ref_main(argc, argv);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("REF_RUNNER wall_ms=%.1f workdir=%s/build rc=%d\n", ms, root.c_str(), rc);
    return rc;
}
