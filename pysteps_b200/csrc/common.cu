// common.cu -- error plumbing and device queries of the C ABI.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace b200 {

long long launches();

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what,
              file, line);
    return (int)e;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

}  // namespace b200

extern "C" {

int b200_version(void) { return 100; }

long long b200_launch_count(void) { return b200::launches(); }

const char *b200_last_error(void) { return b200::g_err; }

int b200_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len) {
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (name && name_len > 0) {
        strncpy(name, p.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    // keep freed stream-ordered scratch cached in the pool instead of returning it to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return 0;
}

}  // extern "C"
