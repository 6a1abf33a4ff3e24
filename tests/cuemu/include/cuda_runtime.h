// cuda_runtime.h (cuemu) — TEST INFRASTRUCTURE, not product code.
//
// A lock-step CPU emulation of the small slice of CUDA that bendy2d_b200/csrc uses, so that the
// UNMODIFIED kernel sources (kernels.cuh) and host code (solver.cu) can be compiled with g++ and run
// through the same C ABI in a container that has no GPU.  Nothing in the product loads this: the
// emulated library is built by tests/cuemu/build.py into tests/cuemu/_build/ and only
// tests/test_emu_*.py point the loader at it.  It proves kernel LOGIC (indexing, barriers, warp
// exchanges, launch order, graph replay); it proves nothing about performance, memory-model races or
// the real hardware, which is what the `-m gpu` tests are for.
//
// Model: every CUDA thread is a fiber with its own stack; the CTAs of a launch run in batches; inside
// a batch fibers are resumed round-robin and block at __syncthreads() / *_sync() warp primitives until
// the other participants arrive.  Static __shared__ variables are rewritten by build.py into per-CTA
// storage that is POISONED (0xCD) at CTA start, dynamic shared memory likewise.  Device memory is host
// memory with canaries on both sides of every allocation.  Streams execute in issue order (a valid
// topological order of the event graph); stream capture records closures, graph launch replays them.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <tuple>
#include <utility>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types -------------------------------------------------------------------------------
struct alignas(8) float2 {
    float x, y;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
struct alignas(8) uint2 {
    uint32_t x, y;
};
struct uint3 {
    uint32_t x, y, z;
};
struct dim3 {
    uint32_t x, y, z;
    dim3(uint32_t x_ = 1, uint32_t y_ = 1, uint32_t z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

using std::isfinite;
using std::max;
using std::min;

// ---- runtime types ------------------------------------------------------------------------------
enum cudaError_t {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidConfiguration = 9,
    cudaErrorLaunchFailure = 719,
};
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };

namespace cuemu {
struct Stream;
struct Event;
struct Graph;
struct GraphExec;
}  // namespace cuemu
typedef cuemu::Stream *cudaStream_t;
typedef cuemu::Event *cudaEvent_t;
typedef cuemu::Graph *cudaGraph_t;
typedef cuemu::GraphExec *cudaGraphExec_t;

struct cudaLaunchAttribute {
    cudaLaunchAttributeID id;
    struct {
        int programmaticStreamSerializationAllowed;
    } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute *attrs;
    unsigned numAttrs;
};

// ---- scheduler interface (runtime.cpp) --------------------------------------------------------------
namespace cuemu {

struct Cta;
struct Fiber {
    void *sp;  // saved stack pointer while switched out
    void *stack;
    uint3 tid;
    uint32_t lane, warp;
    Cta *cta;
    bool done, at_barrier;
    uint32_t warp_op;
};
struct WarpSlot {
    uint32_t arrive, read;
    uint64_t vals[32];
};
struct WarpState {
    uint32_t live;
    WarpSlot slot[2];
};
struct Globals {
    Fiber *cur;
    uint64_t events;  // progress counter (deadlock detection)
};
extern Globals g;

void *dyn_smem();                            // this CTA's dynamic shared memory
void *static_smem(int id, size_t bytes);     // this CTA's copy of static __shared__ declaration `id`
void sync_block();                           // __syncthreads
void yield_spin();                           // inside a busy-wait on memory another CTA writes
WarpSlot &warp_arrive(uint32_t mask, uint64_t v, uint32_t *participants);
void warp_done(WarpSlot &s, uint32_t participants);
uint32_t warp_live();
void launch(const char *name, const void *fn_key, dim3 grid, dim3 block, size_t smem, cudaStream_t st,
            std::function<void()> body);
void set_max_dyn_smem(const void *fn_key, int bytes);

template <typename T>
inline uint64_t pack(T v) {
    static_assert(sizeof(T) <= 8, "warp exchange of at most 8 bytes");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T>
inline T unpack(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

// <<<grid, block, smem, stream>>> is rewritten by build.py into cuemu::Launcher(...).run(kernel, args...)
struct Launcher {
    dim3 grid, block;
    size_t smem;
    cudaStream_t st;
    const char *name;
    Launcher(const char *name_, dim3 g_, dim3 b_, size_t smem_ = 0, cudaStream_t st_ = nullptr)
        : grid(g_), block(b_), smem(smem_), st(st_), name(name_) {}
    template <typename... KArgs, typename... Args>
    void run(void (*kern)(KArgs...), Args &&...args) {
        // arguments are converted and copied by value, like a real launch
        auto tup = std::make_tuple(KArgs(args)...);
        launch(name, (const void *)kern, grid, block, smem, st, [kern, tup]() { std::apply(kern, tup); });
    }
};

}  // namespace cuemu

// plain globals (not macros: cudaLaunchConfig_t has members called gridDim / blockDim); the scheduler
// reloads threadIdx / blockIdx every time it switches to a fiber
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- device intrinsics ------------------------------------------------------------------------------
// Built with -ffp-contract=off -fno-fast-math on x86-64 (SSE2): every float operation below is one
// correctly rounded IEEE binary32 operation, exactly what the _rn intrinsics are.
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }  // one rounding (libm / FMA3)
// stand-in for MUFU.RCP (rcp.approx): the correctly rounded reciprocal.  kernels.cuh normalize2 is ptxas' own div.rn
// expansion with the reciprocal shared; like that expansion it is NOT independent of the seed (d = 0x1.fffffep+39,
// a = 1: a seed one ulp low ends in a tie that rounds the wrong way), so what the emulation checks is the guard
// logic around it; the quotient itself is checked on the device against IEEE division (tests/test_gpu_normalize.py)
static inline float cuemu_rcp_seed(float d) { return 1.0f / d; }
static inline int __float2int_rz(float a) {  // cvt.rzi.s32.f32: saturating, NaN -> 0
    if (!(a == a)) return 0;
    if (a >= 2147483648.0f) return 2147483647;
    if (a <= -2147483648.0f) return -2147483647 - 1;
    return (int)a;
}
static inline long long __float2ll_rn(float a) { return llrintf(a); }  // round-to-nearest-even (default mode)
static inline float __ll2float_rn(long long a) { return (float)a; }
static inline uint32_t __float_as_uint(float a) { return cuemu::unpack<uint32_t>(cuemu::pack(a)); }
static inline float __uint_as_float(uint32_t a) { return cuemu::unpack<float>(cuemu::pack(a)); }
static inline float __int_as_float(int a) { return cuemu::unpack<float>(cuemu::pack(a)); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <typename T>
static inline T __ldcg(const T *p) {
    return *p;
}
long long cuemu_clock64();
static inline long long clock64() { return cuemu_clock64(); }  // nanoseconds of the host's steady clock
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __syncthreads() { cuemu::sync_block(); }

// one OS thread runs every fiber, so a plain read-modify-write is atomic
template <typename T, typename U>
static inline T atomicAdd(T *p, U v) {
    T old = *p;
    *p = (T)(old + (T)v);
    return old;
}
template <typename T, typename U, typename V>
static inline T atomicCAS(T *p, U expected, V desired) {
    T old = *p;
    if (old == (T)expected) *p = (T)desired;
    return old;
}
template <typename T, typename U>
static inline T atomicMin(T *p, U v) {
    T old = *p;
    if ((T)v < old) *p = (T)v;
    return old;
}
template <typename T, typename U>
static inline T atomicMax(T *p, U v) {
    T old = *p;
    if ((T)v > old) *p = (T)v;
    return old;
}
template <typename T, typename U>
static inline T atomicOr(T *p, U v) {
    T old = *p;
    *p = (T)(old | (T)v);
    return old;
}

// warp primitives: every participating lane deposits its value, waits for the others, reads
template <typename T>
static inline T __shfl_sync(uint32_t mask, T v, int src) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, cuemu::pack(v), &part);
    const uint32_t l = (uint32_t)src & 31u;
    T r = (part >> l) & 1u ? cuemu::unpack<T>(s.vals[l]) : v;
    cuemu::warp_done(s, part);
    return r;
}
template <typename T>
static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, cuemu::pack(v), &part);
    const uint32_t lane = cuemu::g.cur->lane;
    T r = (lane >= d && ((part >> (lane - d)) & 1u)) ? cuemu::unpack<T>(s.vals[lane - d]) : v;
    cuemu::warp_done(s, part);
    return r;
}
template <typename T>
static inline T __shfl_xor_sync(uint32_t mask, T v, int d) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, cuemu::pack(v), &part);
    const uint32_t l = (cuemu::g.cur->lane ^ (uint32_t)d) & 31u;
    T r = (part >> l) & 1u ? cuemu::unpack<T>(s.vals[l]) : v;
    cuemu::warp_done(s, part);
    return r;
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, pred ? 1u : 0u, &part);
    uint32_t r = 0;
    for (uint32_t l = 0; l < 32; l++)
        if (((part >> l) & 1u) && s.vals[l]) r |= 1u << l;
    cuemu::warp_done(s, part);
    return r;
}
static inline int __all_sync(uint32_t mask, int pred) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, pred ? 1u : 0u, &part);
    int r = 1;
    for (uint32_t l = 0; l < 32; l++)
        if (((part >> l) & 1u) && !s.vals[l]) r = 0;
    cuemu::warp_done(s, part);
    return r;
}
static inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, v, &part);
    uint32_t r = 0xFFFFFFFFu;
    for (uint32_t l = 0; l < 32; l++)
        if ((part >> l) & 1u) r = std::min(r, (uint32_t)s.vals[l]);
    cuemu::warp_done(s, part);
    return r;
}
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, v, &part);
    uint32_t r = 0u;
    for (uint32_t l = 0; l < 32; l++)
        if ((part >> l) & 1u) r = std::max(r, (uint32_t)s.vals[l]);
    cuemu::warp_done(s, part);
    return r;
}
static inline uint32_t __match_any_sync(uint32_t mask, uint32_t v) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, v, &part);
    uint32_t r = 0;
    for (uint32_t l = 0; l < 32; l++)
        if (((part >> l) & 1u) && (uint32_t)s.vals[l] == v) r |= 1u << l;
    cuemu::warp_done(s, part);
    return r;
}
static inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu) {
    uint32_t part;
    cuemu::WarpSlot &s = cuemu::warp_arrive(mask, 0, &part);
    cuemu::warp_done(s, part);
}
// lanes of this warp that have not exited (divergence inside a warp is not modelled)
static inline uint32_t __activemask() { return cuemu::warp_live(); }

// ---- runtime API (runtime.cpp) ------------------------------------------------------------------------
const char *cudaGetErrorName(cudaError_t e);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest);
cudaError_t cuemuMalloc(void **p, size_t bytes);
template <typename T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
    return cuemuMalloc((void **)p, bytes);
}
cudaError_t cudaFree(void *p);
cudaError_t cuemuHostAlloc(void **p, size_t bytes);
template <typename T>
static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) {
    return cuemuHostAlloc((void **)p, bytes);
}
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode m);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);

template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F *fn, cudaFuncAttribute, int value) {
    cuemu::set_max_dyn_smem((const void *)fn, value);
    return cudaSuccess;
}
// the emulated device: 8 SMs, one resident CTA per SM for any kernel (only used to decide whether a
// kernel with a software grid barrier may be launched: its whole grid must fit one scheduler batch)
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F *, int, size_t) {
    *n = 1;
    return cudaSuccess;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kern)(KArgs...), Args &&...args) {
    cuemu::Launcher("cudaLaunchKernelEx", cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream)
        .run(kern, std::forward<Args>(args)...);
    return cudaPeekAtLastError();
}
