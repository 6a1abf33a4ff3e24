// selfcheck.cpp (cuemu) — TEST INFRASTRUCTURE: known-answer checks of the emulation itself, written in the
// form build.py rewrites kernels into (Launcher instead of <<< >>>, static_smem instead of __shared__).
//   selfcheck            all checks, prints "selfcheck ok"
//   selfcheck oob        writes past a device allocation: must abort with "out-of-bounds write"
//   selfcheck deadlock   a barrier only half the CTA reaches while the rest spins: must abort with "deadlock"
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#define CHECK(c)                                                       \
    do {                                                               \
        if (!(c)) {                                                    \
            printf("selfcheck FAILED at line %d: %s\n", __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

// block-wide inclusive scan through warp shuffles + shared memory, two barriers: the pattern of k2_scan
void k_scan(const uint32_t *in, uint32_t *out, uint32_t n) {
    typedef uint32_t wsum_t[32];
    wsum_t &wsum = *reinterpret_cast<wsum_t *>(cuemu::static_smem(1, sizeof(wsum_t)));
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t v = i < n ? in[i] : 0;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t k = 0; k < w; k++) base += wsum[k];
    __syncthreads();
    if (i < n) out[i] = base + v;
}

// ballot / all / reduce_min / match_any / shfl / shfl_xor with early-exited lanes in the last warp
void k_warp(uint32_t *out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;  // lanes beyond n exit: the others' full-mask primitives must not wait for them
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, (i & 1) != 0);
    const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, 1000u - i);
    const uint32_t g = __match_any_sync(0xFFFFFFFFu, i / 4);
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, i * 3, 0);
    const uint32_t x = __shfl_xor_sync(0xFFFFFFFFu, i, 1);
    const int all = __all_sync(0xFFFFFFFFu, i < n);
    __syncwarp();
    out[6 * i + 0] = b, out[6 * i + 1] = m, out[6 * i + 2] = g, out[6 * i + 3] = s, out[6 * i + 4] = x;
    out[6 * i + 5] = (uint32_t)all | (__activemask() << 1);
    (void)lane;
}

// shared memory must be private to a CTA and poisoned at CTA start; dynamic shared memory likewise
void k_smem(uint32_t *out) {
    uint32_t &flag = *reinterpret_cast<uint32_t *>(cuemu::static_smem(2, sizeof(uint32_t)));
    uint32_t *dyn = reinterpret_cast<uint32_t *>(cuemu::dyn_smem());
    const uint32_t seen = flag, seen_dyn = dyn[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) flag = blockIdx.x;
    dyn[threadIdx.x] = blockIdx.x * 1000 + threadIdx.x;
    __syncthreads();
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 0] = seen;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 1] = flag;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 2] = seen_dyn ^ dyn[blockDim.x - 1 - threadIdx.x];
}

// software grid barrier across co-resident CTAs (the pattern of k2_scan_fused)
void k_grid_barrier(uint32_t *counter, uint32_t *out) {
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        while (*(volatile uint32_t *)counter < gridDim.x) cuemu::yield_spin();
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = *counter;
}

void k_oob(uint32_t *p, uint32_t n) { p[n + threadIdx.x] = 1u; }

void k_deadlock(uint32_t *p) {
    if (threadIdx.x < 16) __syncthreads();
    else
        while (*(volatile uint32_t *)p == 0xEEEEEEEEu) cuemu::yield_spin();
}

int main(int argc, char **argv) {
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (argc > 1 && !strcmp(argv[1], "oob")) {
        uint32_t *p;
        cudaMalloc(&p, 64 * sizeof(uint32_t));
        cuemu::Launcher("k_oob", 1, 32, 0, st).run(k_oob, p, 64u);
        cudaStreamSynchronize(st);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "deadlock")) {
        uint32_t *p;
        cudaMalloc(&p, 4);
        cuemu::Launcher("k_deadlock", 1, 32, 0, st).run(k_deadlock, p);
        return 0;
    }
    {  // scan over 3 CTAs of 256
        const uint32_t n = 700;
        uint32_t *in, *out;
        cudaMalloc(&in, n * 4), cudaMalloc(&out, n * 4);
        std::vector<uint32_t> h(n), r(n);
        for (uint32_t i = 0; i < n; i++) h[i] = (i * 2654435761u) >> 28;
        cudaMemcpy(in, h.data(), n * 4, cudaMemcpyHostToDevice);
        cuemu::Launcher("k_scan", 3, 256, 0, st).run(k_scan, in, out, n);
        cudaMemcpy(r.data(), out, n * 4, cudaMemcpyDeviceToHost);
        for (uint32_t b = 0; b < 3; b++) {
            uint32_t acc = 0;
            for (uint32_t i = b * 256; i < std::min(n, (b + 1) * 256); i++) {
                acc += h[i];
                CHECK(r[i] == acc);
            }
        }
        cudaFree(in), cudaFree(out);
    }
    {  // warp primitives, 70 live threads in 3 warps of one CTA (last warp: 6 live lanes)
        const uint32_t n = 70;
        uint32_t *out;
        cudaMalloc(&out, n * 6 * 4);
        cuemu::Launcher("k_warp", 1, 96, 0, st).run(k_warp, out, n);
        std::vector<uint32_t> r(n * 6);
        cudaMemcpy(r.data(), out, n * 6 * 4, cudaMemcpyDeviceToHost);
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t w0 = i & ~31u, live = std::min(32u, n - w0);
            const uint32_t lmask = live == 32 ? 0xFFFFFFFFu : (1u << live) - 1u;
            CHECK(r[6 * i + 0] == (0xAAAAAAAAu & lmask));
            CHECK(r[6 * i + 1] == 1000u - (w0 + live - 1));
            uint32_t grp = 0xFu << ((i & 31u) & ~3u);
            CHECK(r[6 * i + 2] == (grp & lmask));
            CHECK(r[6 * i + 3] == w0 * 3);
            const uint32_t partner = i ^ 1u;
            CHECK(r[6 * i + 4] == (partner < n ? partner : i));
            CHECK((r[6 * i + 5] & 1u) == 1u);
        }
        cudaFree(out);
    }
    {  // shared memory private per CTA and poisoned
        uint32_t *out;
        cudaMalloc(&out, 4 * 64 * 3 * 4);
        cuemu::Launcher("k_smem", 4, 64, 64 * 4, st).run(k_smem, out);
        std::vector<uint32_t> r(4 * 64 * 3);
        cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
        for (uint32_t b = 0; b < 4; b++)
            for (uint32_t t = 0; t < 64; t++) {
                CHECK(r[(b * 64 + t) * 3 + 0] == 0xCDCDCDCDu);
                CHECK(r[(b * 64 + t) * 3 + 1] == b);
                CHECK(r[(b * 64 + t) * 3 + 2] == (0xCDCDCDCDu ^ (b * 1000 + 63 - t)));
            }
        cudaFree(out);
    }
    {  // grid barrier with all CTAs resident (8 x 256 = the fiber pool), inside a captured graph, replayed twice
        uint32_t *counter, *out;
        cudaMalloc(&counter, 4), cudaMalloc(&out, 8 * 256 * 4);
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        cudaMemsetAsync(counter, 0, 4, st);
        cuemu::Launcher("k_grid_barrier", 8, 256, 0, st).run(k_grid_barrier, counter, out);
        cudaStreamEndCapture(st, &graph);
        cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        std::vector<uint32_t> r(8 * 256);
        cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
        CHECK(r[0] == 0xEEEEEEEEu);  // capture does not execute
        for (int rep = 0; rep < 2; rep++) {
            cudaGraphLaunch(exec, st);
            cudaStreamSynchronize(st);
            cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
            for (uint32_t v : r) CHECK(v == 8u);
        }
        cudaGraphExecDestroy(exec);
        cudaFree(counter), cudaFree(out);
    }
    {  // launch limits: dynamic shared memory above 48 KB needs the opt-in; 1025 threads never launch
        uint32_t *out;
        cudaMalloc(&out, 4 * 64 * 3 * 4);
        cuemu::Launcher("k_smem", 1, 64, 100 * 1024, st).run(k_smem, out);
        CHECK(cudaGetLastError() == cudaErrorInvalidValue);
        CHECK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) == cudaSuccess);
        cuemu::Launcher("k_smem", 1, 64, 100 * 1024, st).run(k_smem, out);
        CHECK(cudaGetLastError() == cudaSuccess);
        cuemu::Launcher("k_smem", 1, 1025, 0, st).run(k_smem, out);
        CHECK(cudaGetLastError() == cudaErrorInvalidConfiguration);
        cudaFree(out);
    }
    printf("selfcheck ok\n");
    return 0;
}
