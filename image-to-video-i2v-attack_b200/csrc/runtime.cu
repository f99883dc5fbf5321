// Library-level plumbing behind the C ABI: version, error text, device check.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace i2v {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-free error so the next call reports its own
    return I2V_ECUDA;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            return 148;
    }
    return n;
}

}  // namespace i2v

extern "C" int i2v_version(void) { return I2V_VERSION; }

extern "C" const char* i2v_last_error(void) { return i2v::g_err; }

extern "C" int i2v_device_check(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return i2v::cuda_fail(e, "i2v_device_check");
    if (prop.major != 10) {
        i2v::set_error("i2v_b200 is built for sm_100a only; device %d is sm_%d%d (%s) and there is no fallback path",
                       device, prop.major, prop.minor, prop.name);
        return I2V_ECUDA;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return i2v::cuda_fail(e, "i2v_device_check");
    return I2V_OK;
}
