// sdb200 — library-level entry points: version, error string, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

namespace sdb {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// launches of this host thread that must NOT carry the programmatic attribute (see sdb_pdl_skip_next)
static thread_local int g_pdl_skip = 0;

bool pdl_enabled() {
    if (g_pdl_skip > 0) { --g_pdl_skip; return false; }
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SDB200_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: %s", what, cudaGetErrorString(e));
        return SDB_ERR_CUDA;
    }
    return SDB_OK;
}

}  // namespace sdb

extern "C" {

int sdb_version(void) { return 10000 * 0 + 100 * 1 + 0; }

const char* sdb_last_error_string(void) { return sdb::g_err; }

int sdb_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

unsigned long long sdb_launch_count(void) { return sdb::g_launches.load(std::memory_order_relaxed); }

int sdb_pdl_skip_next(int n) {
    const int prev = sdb::g_pdl_skip;
    sdb::g_pdl_skip = n > 0 ? n : 0;
    return prev;
}

}  // extern "C"
