// Error reporting and device queries shared by the C-ABI entry points.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace tvz {

char *err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("TVZ_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSMsB200;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace tvz

extern "C" {
const char *tvz_last_error(void) { return tvz::err_buf(); }
int tvz_abi_version(void) { return 1; }
}
