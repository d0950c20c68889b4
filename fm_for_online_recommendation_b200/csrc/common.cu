// common.cu -- error plumbing and library identity for libfmb200.so
#include "fmb_common.cuh"
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

void fmb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

FMB_API const char* fmb_last_error(void) { return g_err; }
FMB_API int fmb_version(void) { return 100; }
// row pitch of the packed table in floats: rows are 64-byte aligned (HBM bursts are 64 B; a 48-byte row
// at a 48-byte pitch straddles two bursts 3 times out of 4 -- measured 35 MB of DRAM reads for 15 MB of rows)
FMB_API int fmb_rowp(int k) { return fmb_round_up(k + 1, 16); }
FMB_API int fmb_kp4(int k) { return fmb_round_up(k, 4); }

// number of visible CUDA devices (0 when there is no GPU / no driver): lets the Python side fail
// loudly instead of falling back to anything.
FMB_API int fmb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
