// Library-wide plumbing: version and the thread-local error string of the C ABI.
#include "common.cuh"

static thread_local char g_err[512] = "";

void vel_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

VEL_API int vel_version(void) { return 100; }  // 0.1.0

VEL_API const char* vel_last_error(void) { return g_err; }

void vel_keep_async_pool_cached()
{
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[dev] = true;
}
