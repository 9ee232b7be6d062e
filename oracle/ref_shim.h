/* ref_shim.h -- force-included when compiling the reference with g++: test/RaytraceTest.cpp:283 calls the
 * MSVC-only std::sqrtf.  One using-declaration; nothing else is changed. */
#include <math.h>
#ifdef __cplusplus
namespace std { using ::sqrtf; }
#endif
