// Adl/Adl.h -- ADL-shaped C++ shim over the ptb200 C-ABI.
//
// Provides the subset of the reference's ADL API that its ray-cast flow uses
// (test/RaytraceTest.cpp:202-291), with the same names and argument meaning, so a
// maintainer can point that flow at libptb200.so by switching the include path:
//   adl::init / adl::quit                      (reference Adl/Adl.h:96-98)
//   adl::DeviceUtils::allocate / deallocate / waitForCompletion   (Adl/Adl.h:100-131)
//   adl::Device::getKernel / getDeviceVersion  (Adl/Adl.h:164-166)
//   adl::Buffer<T>                             (Adl/Adl.h:203-265)
//   adl::BufferInfo, adl::Launcher             (Adl/AdlKernel.h:94-202)
// Error convention: like release-mode ADL nothing throws; failures leave null
// handles / zero sizes and the message is available from ptb_last_error().
#pragma once

// standard headers the reference's Adl.h / AdlKernel.h make visible to their includers
#include <assert.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "ptb200.h"

// The reference's flow spells its records with OpenCL host types (clew.h, pulled in by the reference's Adl.h).
// Define the three it uses unless a CL header already did.
#if !defined(__OPENCL_CL_PLATFORM_H) && !defined(CLEW_HPP_INCLUDED) && !defined(ADL_SHIM_NO_CL_TYPES)
typedef float cl_float;
typedef int cl_int;
typedef unsigned int cl_uint;
union alignas(16) cl_float4 {
    cl_float s[4];
    struct { cl_float x, y, z, w; };
};
typedef cl_float4 cl_float3;  // a 3-vector occupies a 4-vector, as in OpenCL
#endif

namespace adl {

typedef unsigned long long adlu64;

// status codes and the assertion macro that ADL-side helpers (e.g. the reference's test/Array.h) spell out
// (names and values as Adl/AdlError.h:24-41; the release-mode macro evaluates and ignores, Adl/AdlError.h:51)
enum TahoeErrorCodes {
    TH_NO_ERROR = 0, TH_SUCCESS = 0, TH_FAILURE = 1, TH_ERROR_MEMORY = 2, TH_ERROR_IO = 3, TH_ERROR_PARAMETER = 4,
    TH_ERROR_INTERNAL = 5, TH_NOT_SUPPORTED = 11, TH_ERROR_NULLPTR = 14,
};
#ifndef ADLASSERT
#define ADLASSERT(x, code) do { if (x) {} } while (0)
#endif
#define ADL_SUCCESS 0
#define ADL_FAILURE 1

enum DeviceType { TYPE_CL = 0, TYPE_DX11 = 1, TYPE_HOST = 2, TYPE_METAL = 3, TYPE_VULKAN = 4, TYPE_CUDA = 5 };

inline bool init(DeviceType) {
    int n = 0;
    return ptb_device_count(&n) == PTB_OK && n > 0;
}
inline void quit(DeviceType) {}

struct Kernel {
    ptb_kernel* m_kernel = nullptr;
};

class Device {
  public:
    explicit Device(ptb_device* d) : m_type(TYPE_CUDA), m_dev(d) {}
    // name -> kernel, cached by the device for its lifetime; null when unknown
    const Kernel* getKernel(const char* fileName, const char* funcName, const char* /*option*/ = nullptr) const {
        for (int i = 0; i < m_nKernels; ++i)
            if (!std::strcmp(m_names[i], funcName)) return &m_kernels[i];
        ptb_kernel* k = nullptr;
        if (ptb_kernel_get(m_dev, fileName, funcName, &k) != PTB_OK || m_nKernels == kMaxKernels) return nullptr;
        std::strncpy(m_names[m_nKernels], funcName, sizeof m_names[0] - 1);
        m_kernels[m_nKernels].m_kernel = k;
        return &m_kernels[m_nKernels++];
    }
    void getDeviceVersion(char nameOut[128]) const { ptb_device_name(m_dev, nameOut); }
    void getDeviceName(char nameOut[128]) const { ptb_device_name(m_dev, nameOut); }
    void getBoardName(char nameOut[128]) const { ptb_device_name(m_dev, nameOut); }
    void getDeviceVendor(char nameOut[128]) const { std::strncpy(nameOut, "NVIDIA Corporation", 127); nameOut[127] = 0; }
    adlu64 getMaxAllocationSize() const {
        size_t f = 0, t = 0;
        ptb_device_memory(m_dev, &f, &t);
        return (adlu64)f;
    }
    DeviceType m_type;
    ptb_device* m_dev;
    ptb_device* m_helpers[PTB_MAX_PEERS] = {};  // other GPUs of the box that render part of this device's work (DeviceUtils::allocate)
    int m_nHelpers = 0;

  private:
    enum { kMaxKernels = 8 };
    mutable Kernel m_kernels[kMaxKernels];
    mutable char m_names[kMaxKernels][64] = {};
    mutable int m_nKernels = 0;
};

class DeviceUtils {
  public:
    struct Config {
        enum DeviceType { DEVICE_GPU, DEVICE_CPU };
        Config() : m_type(DEVICE_GPU), m_deviceIdx(0), m_nGpus(0) {}
        DeviceType m_type;
        int m_deviceIdx;
        // not in the reference (it picks ONE device): how many GPUs of the box serve this device; 0 = the environment
        // variable PTB_GPUS, else 1.  The flow does not change: the other GPUs become helpers (ptb_device_add_helper)
        // and ptb_launch1d deals its frame-ahead batches over them, bit-identical to one GPU.
        int m_nGpus;
    };
    static int getNDevices(DeviceType) {
        int n = 0;
        ptb_device_count(&n);
        return n;
    }
    // returns 0 when no device can be created (the reference returns a half-initialised object instead)
    static Device* allocate(DeviceType, Config cfg = Config()) {
        ptb_device* d = nullptr;
        if (ptb_device_create(cfg.m_deviceIdx, &d) != PTB_OK) return nullptr;
        Device* dev = new Device(d);
        int want = cfg.m_nGpus;
        if (want <= 0) {
            const char* e = getenv("PTB_GPUS");
            want = e ? atoi(e) : 1;
        }
        int have = 0;
        ptb_device_count(&have);
        for (int i = 1; i < want && i < have && dev->m_nHelpers < PTB_MAX_PEERS; ++i) {
            ptb_device* h = nullptr;
            if (ptb_device_create((cfg.m_deviceIdx + i) % have, &h) != PTB_OK) break;
            if (ptb_device_add_helper(d, h) != PTB_OK) { ptb_device_destroy(h); break; }
            dev->m_helpers[dev->m_nHelpers++] = h;
        }
        return dev;
    }
    static void deallocate(Device* device) {
        if (!device) return;
        ptb_device_destroy(device->m_dev);
        for (int i = 0; i < device->m_nHelpers; ++i) ptb_device_destroy(device->m_helpers[i]);
        delete device;
    }
    static void waitForCompletion(const Device* device) {
        if (device) ptb_device_sync(device->m_dev);
    }
};

struct BufferBase {
    enum BufferType { BUFFER, BUFFER_CONST, BUFFER_STAGING, BUFFER_APPEND, BUFFER_RAW, BUFFER_W_COUNTER, BUFFER_INDEX, BUFFER_VERTEX, BUFFER_ZERO_COPY };
};

template <typename T>
struct Buffer : public BufferBase {
    Buffer() : m_device(nullptr), m_size(0), m_buf(nullptr) {}
    Buffer(const Device* device, adlu64 nElems, BufferType = BUFFER) : m_device(nullptr), m_size(0), m_buf(nullptr) { allocate(device, nElems); }
    virtual ~Buffer() {
        if (m_buf) ptb_buffer_destroy(m_buf);
    }
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
    void allocate(const Device* device, adlu64 nElems, BufferType = BUFFER) {
        m_device = device;
        // on failure: null handle and zero size, no throw (Adl/CL/AdlCL.inl:190-197)
        m_size = (device && ptb_buffer_create(device->m_dev, size_t(nElems) * sizeof(T), &m_buf) == PTB_OK) ? nElems : 0;
    }
    void write(const T* hostSrcPtr, adlu64 nElems, adlu64 dstOffsetNElems = 0) {
        if (m_buf) ptb_buffer_write(m_buf, hostSrcPtr, size_t(nElems) * sizeof(T), size_t(dstOffsetNElems) * sizeof(T));
    }
    void read(T* hostDstPtr, adlu64 nElems, adlu64 srcOffsetNElems = 0) const {
        if (m_buf) ptb_buffer_read(m_buf, hostDstPtr, size_t(nElems) * sizeof(T), size_t(srcOffsetNElems) * sizeof(T));
    }
    void clear() {
        if (m_buf) ptb_buffer_clear(m_buf);
    }
    // map the whole buffer read/write; contents are valid after waitForCompletion
    T* getHostPtr(adlu64 /*size*/ = adlu64(-1), bool /*blocking*/ = false) const {
        void* p = nullptr;
        return (m_buf && ptb_buffer_map(m_buf, &p) == PTB_OK) ? static_cast<T*>(p) : nullptr;
    }
    void returnHostPtr(T* ptr) const {
        if (m_buf && ptr) ptb_buffer_unmap(m_buf, ptr);
    }
    adlu64 getSize() const { return m_size; }

    const Device* m_device;
    adlu64 m_size;
    ptb_buffer* m_buf;
};

struct BufferInfo {
    BufferInfo() : m_buffer(nullptr), m_isReadOnly(false) {}
    template <typename T>
    BufferInfo(const Buffer<T>* buff, bool isReadOnly = false) : m_buffer(buff ? buff->m_buf : nullptr), m_isReadOnly(isReadOnly) {}
    ptb_buffer* m_buffer;
    bool m_isReadOnly;
};

#define ADL_DEFAULT_LOCAL_SIZE_1D 64

class Launcher {
  public:
    enum { MAX_ARG_SIZE = 64, MAX_ARG_COUNT = 64 };
    Launcher(const Device* dd, const Kernel* kernel) : m_deviceData(dd), m_kernel(kernel), m_nBufs(0), m_constBytes(0) {}
    // positional arguments: buffers first, then by-value constants (auto-incrementing index)
    void setBuffers(BufferInfo* buffInfo, int n) {
        for (int i = 0; i < n && m_nBufs < MAX_ARG_COUNT; ++i) m_bufs[m_nBufs++] = buffInfo[i].m_buffer;
    }
    template <typename T>
    void setConst(const T& consts) {
        static_assert(sizeof(T) <= MAX_ARG_SIZE, "constant block too large");
        std::memcpy(m_const, &consts, sizeof(T));
        m_constBytes = sizeof(T);
    }
    // returns 0 (the reference returns elapsed ms only in profiling builds)
    float launch1D(int numThreads, int localSize = ADL_DEFAULT_LOCAL_SIZE_1D) {
        if (m_deviceData && m_kernel)
            ptb_launch1d(m_deviceData->m_dev, m_kernel->m_kernel, m_bufs, m_nBufs, m_const, m_constBytes, numThreads, localSize);
        return 0.f;
    }

    struct ExecInfo {  // reference Adl/AdlKernel.h:139-160
        int m_nWIs[3];
        int m_wgSize[3];
        int m_nDim;
        ExecInfo() {}
        ExecInfo(int nWIsX, int nWIsY = 1, int nWIsZ = 1, int wgSizeX = 64, int wgSizeY = 1, int wgSizeZ = 1, int nDim = 1) {
            m_nWIs[0] = nWIsX; m_nWIs[1] = nWIsY; m_nWIs[2] = nWIsZ;
            m_wgSize[0] = wgSizeX; m_wgSize[1] = wgSizeY; m_wgSize[2] = wgSizeZ;
            m_nDim = nDim;
        }
    };
    // dump the bound arguments + buffer contents in the reference's own file layout (AdlKernelUtilsCL.cpp:509-563)
    void serializeToFile(const char* filePath, const ExecInfo& info) {
        if (m_deviceData) ptb_launch_serialize(m_deviceData->m_dev, filePath, m_bufs, m_nBufs, m_const, m_constBytes, info.m_nWIs[0], info.m_wgSize[0]);
    }
    // re-create the buffers of a dump and bind everything; the caller then launches with the returned shape
    ExecInfo deserializeFromFile(const char* filePath, int buffCap, ptb_buffer** buffsOut, int* nBuffs) {
        ExecInfo info(0);
        int n = 0, ls = 64;
        m_nBufs = 0;
        if (m_deviceData && ptb_launch_deserialize(m_deviceData->m_dev, filePath, buffsOut, buffCap, nBuffs, m_const, &m_constBytes, &n, &ls) == PTB_OK) {
            for (int i = 0; i < *nBuffs && i < MAX_ARG_COUNT; ++i) m_bufs[m_nBufs++] = buffsOut[i];
            info = ExecInfo(n, 1, 1, ls);
        }
        return info;
    }

  private:
    const Device* m_deviceData;
    const Kernel* m_kernel;
    ptb_buffer* m_bufs[MAX_ARG_COUNT];
    int m_nBufs;
    unsigned char m_const[MAX_ARG_SIZE];
    size_t m_constBytes;
};

}  // namespace adl

#define SELECT_KERNELPATH1(device, path, name) (path "ClKernels/" name)
