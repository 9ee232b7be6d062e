// stub (replaced below)
#pragma once
#include "pt_kernels.cuh"
namespace ptd {
static int wavefront_render(cudaStream_t, void**, size_t*, unsigned long long*, int, const SceneDev&, const RenderArgs&, bool, bool, bool, int) { return -1; }
}
