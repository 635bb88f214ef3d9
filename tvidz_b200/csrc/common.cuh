// Shared helpers for the tvidz_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tvidz_b200.h"

namespace tvz {

constexpr int kNumSMsB200 = 148;

char *err_buf();                       // thread-local, 512 bytes
int set_error(int code, const char *fmt, ...);
int num_sms();                          // SM count of the current device (cached per device)

// Fused gather (multi-GPU matcher): where the compaction's last block stores this rank's hit
// record on every peer, and the flag it raises there afterwards.
constexpr int kMaxPeers = 8;
struct GatherTargets {
    int n_peers = 0;                 // 0 = no gather
    int n_dst = 0;                   // addresses in record[]: n_peers (one per peer), or 1 = a multicast address (NVSwitch replicates)
    unsigned epoch = 0;
    long long query_stride = 0;      // batched queries: ints between the records of consecutive queries
    int *record[kMaxPeers] = {};     // this rank's slot inside peer p's gather buffer (peer memory)
    unsigned *flag[kMaxPeers] = {};  // this rank's flag word on peer p (flag protocol: fragment compaction)
    // tagged protocol (match_tile_kernel): every 8-byte word of a record carries the query epoch, so a record is
    // complete when all its words show it -- no fence, no flag, no counter on the sender's side
    const int *my_slots = nullptr;   // this rank's own gather buffer: n_peers slots, slot_stride ints apart
    long long slot_stride = 0;
};

// compact.cu: single-pass ordered compaction of per-row results (fragment mode).
// `ticket` holds three u32 {next ticket = 0, query epoch = 1, finished blocks = 0}.
int compact_blocks(long long n_rows);
int compact_enqueue(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                    long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket,
                    const int *aux, int *aux_out, cudaStream_t st, const GatherTargets *gather = nullptr);
int compact_enqueue_keys(unsigned long long *keys, long long n_rows, int min_match, const int *vid, int *out,
                         long long *rows_out, long long cap, long long *n_hits_out, unsigned long long *state,
                         unsigned *ticket, int *delta_out, cudaStream_t st, const GatherTargets *gather = nullptr);
int gather_wait_enqueue(const unsigned *d_flags, int n_peers, unsigned epoch, cudaStream_t st);

// Fragment mode: a row's best candidate as ONE u64 so that concurrent warps can combine theirs with
// atomicMax.  Order = the spec's: higher score, then smaller |d|, then smaller d.  |d| < 2^30.
__host__ __device__ __forceinline__ unsigned long long frag_key(int score, int d) {
    const unsigned ad = static_cast<unsigned>(d < 0 ? -d : d);
    return (static_cast<unsigned long long>(static_cast<unsigned>(score)) << 32) |
           (0xffffffffu - (2u * ad + (d > 0 ? 1u : 0u)));
}
__host__ __device__ __forceinline__ int frag_key_score(unsigned long long k) { return static_cast<int>(k >> 32); }
__host__ __device__ __forceinline__ int frag_key_delta(unsigned long long k) {
    if (k == 0) return 0;  // no candidate: (score 0, offset 0)
    const unsigned x = 0xffffffffu - static_cast<unsigned>(k);
    const int ad = static_cast<int>(x >> 1);
    return (x & 1u) ? ad : -ad;
}

}  // namespace tvz
#include <exception>
#include <new>
namespace tvz {
// The C ABI never throws: host-side allocation failures become TVZ_ERR_NOMEM.
template <class F>
int guarded(F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return set_error(TVZ_ERR_NOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return set_error(TVZ_ERR_INVALID, "unexpected exception: %s", e.what());
    } catch (...) {
        return set_error(TVZ_ERR_INVALID, "unexpected exception");
    }
}

// Entry points run on the device their catalogue lives on, whatever device the calling thread has current
// (and leave the caller's choice as it was): bindings need no device context manager around a call.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

#define TVZ_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return tvz::set_error(TVZ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define TVZ_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) return tvz::set_error(TVZ_ERR_INVALID, __VA_ARGS__);                     \
    } while (0)

// ---- programmatic dependent launch (PDL) ----
// The kernels of one query (count -> compaction [-> next query's count]) are launched with the
// programmatic-stream-serialization attribute: a kernel's CTAs may become resident and run their
// prologue (first loads, shared-memory tables) while the predecessor's last CTAs drain, and block
// in pdl_wait() until the predecessor has completed and flushed -- which takes the grid launch
// latency off the critical path of a 50 us query.  Every such kernel calls pdl_wait() before it
// touches anything its predecessor writes (on all paths) and pdl_launch_dependents() right after,
// so a kernel two places down the stream never starts before the kernel two places up has finished.
bool pdl_enabled();  // common.cu: off with TVZ_NO_PDL=1
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1]{};
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- mbarrier / TMA bulk-copy PTX wrappers (shared::cta addresses as u32) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
// try_wait suspends for a hardware-chosen interval, so 2^26 polls are many seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls == (1u << 26)) {
#ifdef TVZ_DEBUG_WAIT
            printf("tvidz_b200: mbarrier wait timed out (tag %d, block %d, thread %d, bar 0x%x, parity %u)\n", tag,
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
#endif
            (void)tag;
            __trap();
        }
    }
}
// TMA 1-D bulk copy global -> shared::cta, completion reported on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}


}  // namespace tvz
