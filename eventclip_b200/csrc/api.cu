// api.cu -- error reporting and device checks of the C ABI (include/eventclip_b200.h).
#include <string.h>

#include "common.cuh"

namespace ec {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace ec

extern "C" const char *ec_last_error(void) { return ec::g_err; }

extern "C" int ec_version(void) { return 100; }

extern "C" int ec_device_check(void)
{
    int dev = 0, major = 0, minor = 0;
    EC_CUDA_CHECK(cudaGetDevice(&dev));
    EC_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    EC_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        ec::set_error("eventclip_b200 needs an sm_100 device (B200); found sm_%d%d. There is no fallback path.", major,
                      minor);
        return EC_ERR_UNSUPPORTED;
    }
    return EC_OK;
}
