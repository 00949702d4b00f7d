// lg_api.cu -- version, device check, thread-local error string of liblidargeom.so.
#include <stdarg.h>

#include "lg_common.cuh"

namespace lg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace lg

extern "C" int lg_version(void) { return LG_VERSION; }

extern "C" const char* lg_last_error_string(void) { return lg::g_err; }

extern "C" int lg_check_device(void) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        lg::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return LG_ERR_NO_DEVICE;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        lg::set_error("device %d has compute capability major %d; liblidargeom is built for sm_100a only", dev, major);
        return LG_ERR_NO_DEVICE;
    }
    return LG_OK;
}
