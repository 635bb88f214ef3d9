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

// Measurement helper: evict the L2 by rewriting and re-reading a buffer larger than it.  The kernel asks for
// the same shared-memory carve-out as the matcher's kernels, so a query timed right behind it starts with a
// cold L2 but does not also pay for an L1 / shared-memory reconfiguration that only the harness caused.
__global__ void __launch_bounds__(512, 2) flush_l2_kernel(uint4 *buf, size_t n16, unsigned long long *sink) {
    extern __shared__ unsigned char flush_smem[];
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t t0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (size_t i = t0; i < n16; i += stride) buf[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    unsigned long long acc = 0;
    for (size_t i = t0; i < n16; i += stride) {
        const uint4 v = buf[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (threadIdx.x == 0) flush_smem[0] = static_cast<unsigned char>(acc);
    if (acc == 0xdeadbeefull) *sink = acc + flush_smem[0];   // never true: keeps the reads alive
}

}  // namespace tvz

extern "C" {
/* Debug / measurement hook (not in the public header): rewrite and re-read `bytes` of device memory at d_buf. */
int tvz_debug_flush_l2(void *d_buf, int64_t bytes, void *stream) {
    TVZ_REQUIRE(d_buf && bytes >= 16, "bad arguments");
    static bool attr = false;
    constexpr int kSmem = 100 * 1024;
    if (!attr) {
        TVZ_CUDA(cudaFuncSetAttribute(tvz::flush_l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr = true;
    }
    tvz::flush_l2_kernel<<<2 * tvz::num_sms(), 512, kSmem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint4 *>(d_buf), static_cast<size_t>(bytes / 16), static_cast<unsigned long long *>(d_buf));
    TVZ_CUDA(cudaGetLastError());
    return TVZ_OK;
}
const char *tvz_last_error(void) { return tvz::err_buf(); }
int tvz_abi_version(void) { return 1; }
}
