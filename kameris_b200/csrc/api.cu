// api.cu -- error channel, version and device info of the C ABI (include/kameris_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void kam_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int kam_sm_count() {
    static int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

extern "C" {

int kam_version(void) { return KAM_VERSION; }
const char *kam_last_error(void) { return g_err; }

int kam_device_info(int *sm_count, size_t *free_bytes, size_t *total_bytes) {
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    KAM_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        kam_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
        return KAM_E_UNSUPPORTED;
    }
    size_t f = 0, t = 0;
    KAM_CUDA(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return KAM_OK;
}

// A.1 record extraction (generation_cgr @0x10000d691-0x10000d83e): host-side, part of ingest.
static inline bool kam_isspace(unsigned char c) {
    return c == ' ' || (c >= '\t' && c <= '\r');
}

int64_t kam_fasta_record0(const char *buf, size_t n, char *out) {
    size_t len = 0;
    const char *p = buf, *end = buf + n;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *e = nl ? nl : end;
        const char *a = p, *b = e;
        while (a < b && kam_isspace((unsigned char)*a)) ++a;
        while (b > a && kam_isspace((unsigned char)b[-1])) --b;
        if (a == b || *a == '>') {
            if (len) return (int64_t)len;
        } else {
            memcpy(out + len, a, (size_t)(b - a));
            len += (size_t)(b - a);
        }
        p = e + 1;
    }
    return len ? (int64_t)len : -1;
}

}  // extern "C"
