// Library-wide helpers: error string, version, device count, pinned host memory.
#include <cstdarg>
#include "common.cuh"

namespace plvs {
int g_profiling = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace plvs

extern "C" {

const char* plvs_version(void) { return "plvs_b200 0.1 (sm_100a)"; }
const char* plvs_last_error(void) { return plvs::g_err; }

int plvs_set_profiling(int mask) { plvs::g_profiling = mask; return PLVS_OK; }

int plvs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int plvs_host_alloc(void** p, size_t bytes)
{
    if (!p) return PLVS_EINVAL;
    if (cudaHostAlloc(p, bytes, cudaHostAllocDefault) != cudaSuccess) { plvs::set_error("cudaHostAlloc(%zu) failed", bytes); return PLVS_ENOMEM; }
    return PLVS_OK;
}

int plvs_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) return PLVS_ENODEV;
    return PLVS_OK;
}

}  // extern "C"
