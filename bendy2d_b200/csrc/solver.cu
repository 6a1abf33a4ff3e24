// solver.cu — host side of libbendy2d_b200.so: scene store, link planning, launch sequencing,
// CUDA-graph capture of the substep, and the C ABI declared in include/bendy2d_b200.h.
//
// Mirrors the reference's `Solver` (src/solver.rs:20-116).  The substep order is the reference's
// (solver.rs:109-115): gravity -> links -> dynamic collisions -> bounds -> integrate, with gravity
// fused into the integrate kernel (acc is zero between substeps: particle.rs:24, solver.rs:110).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/bendy2d_b200.h"
#include "kernels.cuh"
#include "plan.h"

using namespace bendy;

namespace {

thread_local std::string g_last_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grows (never shrinks); contents are NOT preserved
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        cap = want;
        return cudaSuccess;
    }
};

struct PolyHost {
    uint32_t start = 0, nv = 0;  // into the polygon point arrays
    uint32_t link_start = 0, nl = 0;
    bool is_static = false;
    float2 center = {0.f, 0.f};
};

// NCCL is reached through dlopen so that libbendy2d_b200.so has no link-time dependency on it; in
// a torch process the already-loaded libnccl.so.2 (same SONAME) is reused.
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        // BENDY_NCCL_LIB names the NCCL build to use (a path or a soname); default: the one already in the process
        const char *user = getenv("BENDY_NCCL_LIB");
        for (const char *name : {user, "libnccl.so.2", "libnccl.so"}) {
            if (!name || !*name) continue;
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define NCCL_SYM(field, name)                                                    \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                 \
    if (!field) {                                                                \
        err = std::string("libnccl is missing ") + name;                         \
        lib = nullptr;                                                           \
        return false;                                                            \
    }
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(Send, "ncclSend")
        NCCL_SYM(Recv, "ncclRecv")
        NCCL_SYM(AllReduce, "ncclAllReduce")
        NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;

enum Phase { PHASE_ALL = 0, PHASE_A = 1, PHASE_B = 2, PHASE_C = 3 };

struct PendingEvent {
    int cls;
    cudaEvent_t e0, e1;
};

}  // namespace

struct bendy_solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};  // graph branches: circles, polygons, work statistics
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr}, ev_main = nullptr;
    std::string err;
    int sticky = BENDY_OK;

    // ---------------- host scene, USER order (solver.rs:24-28)
    std::vector<float2> p_pos, p_prev;  // free particles (acc is always 0 at add: solver.rs:52-54)
    std::vector<float> p_k;             // ext inverse-mass scale, empty = all 1
    std::vector<float2> c_pos, c_prev, c_acc;
    std::vector<float> c_rad, c_k;
    std::vector<uint32_t> pl_ab;  // particle links
    std::vector<float> pl_len;
    std::vector<GlobalLink> cl;   // circle links
    std::vector<PolyHost> polys;
    std::vector<float2> g_pos, g_prev, g_acc;  // polygon points, polygon-major
    std::vector<uint32_t> gl_ab;               // polygon-local link indices
    std::vector<float> gl_len;
    bool any_acc = false;  // some circle / polygon point was added with acc != 0

    float last_args[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // dt, gravity, bounds of the last update()
    bool have_last_args = false;
    uint16_t sub_steps = 1;
    float particle_radius = 0.f;
    float grid_cell = 0.f;
    bool polygon_contact = false;
    PlanParams plan_params;

    // ---------------- state location
    bool host_valid = true;     // host vectors hold the current pos/prev
    bool device_valid = false;  // device buffers hold the current pos/prev
    bool topo_dirty = true;     // plan / device tables must be rebuilt

    // ---------------- plan + device tables
    LinkPlan plan_p;
    uint32_t nP = 0, nC = 0, nG = 0, N = 0, Npad = 0;
    DevBuf<float2> d_pos, d_prev, d_accel, d_stage;
    DevBuf<float> d_k, d_crad;
    DevBuf<uint8_t> d_gstatic;
    DevBuf<uint32_t> d_rank;  // user particle -> internal
    DevBuf<uint32_t> d_part_start, d_part_cs;
    DevBuf<LocalLink> d_local;
    DevBuf<GlobalLink> d_global, d_clinks;
    bool accel_pending = false;
    bool has_k = false;
    // grid
    DevBuf<uint32_t> d_cell_count, d_cell_start, d_tile_sum, d_sorted_id, d_slot_of;
    bool poly_fused = true;     // BENDY_POLY_FUSED=0: the multi-launch polygon chain also for <= 1024 polygons
    uint32_t k3_threads = 128;  // BENDY_K3_THREADS
    bool halo_overlap = false;  // BENDY_HALO_OVERLAP
    bool small_scene = true;    // BENDY_SMALL_SCENE=0 forces the multi-kernel path for tiny scenes
    int pdl = 3;                // BENDY_PDL: 0 = off, 1 = programmatic dependent launch for links/scan/scatter,
                                // 2 = + narrowphase, 3 = + the circle / polygon branches (default); not yet used
                                // on the NCCL strip path (BENDY_PDL_NCCL=1)
    bool pdl_nccl = true;   // BENDY_PDL_NCCL=0 turns programmatic dependent launch off on the strip path
    DevBuf<float2> d_sorted_pos;
    // which end of the index range the narrowphase starts at (kernels.cuh, k_work_halves)
    DevBuf<uint32_t> d_chunk_work;                // clocks / 64 per warp of the narrowphase
    unsigned long long *h_work_halves = nullptr;  // pinned: their sums over the lower / upper half of the chunks, written
                                                  // by k_work_halves once per update, read (possibly a few updates
                                                  // stale) when the next one is enqueued
    int narrow_order_mode = 0;   // BENDY_NARROW_ORDER: 0 = auto (default), 1 = forward, 2 = reverse
    bool narrow_reverse = true;  // current choice; auto starts at the top: bodies created last tend to be the lowest
    bool record_work = false;    // set for the LAST substep of an update: only that one pays for the clocks
    DevBuf<uint32_t> d_circ_tile_count, d_circ_tile_ids;
    uint32_t n_scan_tiles = 0, n_circ_tiles = 0;
    DevBuf<unsigned long long> d_circ_acc;
    DevBuf<float2> d_circ_snap;
    uint32_t n_cells = 0;
    // polygons
    DevBuf<uint32_t> d_poly_start, d_poly_tiles;
    DevBuf<uint8_t> d_poly_static;
    DevBuf<float2> d_poly_center;
    DevBuf<uint32_t> d_poly_first_row, d_poly_link_start, d_poly_link_ab;
    DevBuf<float> d_poly_link_len;
    DevBuf<float4> d_poly_box;
    float poly_tile = 0.f;
    uint32_t n_poly_tiles = 0;
    DevBuf<int> d_flags;

    // ---------------- spatial strips (multi-GPU): ghost disc slots + halo exchange
    uint32_t nOwned = 0;      // free particles owned by this solver (== p_pos.size())
    uint32_t ghost_cap = 0;   // ghost slots per side; discs [nOwned, nOwned+cap) come from the left neighbour,
                              // [nOwned+cap, nOwned+2cap) from the right one
    bool halo_on = false;
    float halo_xl = -INFINITY, halo_xr = INFINITY, stray_xl = -INFINITY, stray_xr = INFINITY;
    float win_x0 = -INFINITY, win_x1 = INFINITY;  // x-range the broadphase grid covers (clipped to the bounds)
    DevBuf<float2> d_send[2];
    DevBuf<float> d_send_k[2];  // ext: inverse-mass scales of the packed discs
    DevBuf<uint32_t> d_send_cnt;
    ncclComm_t nccl_comm = nullptr;
    int comm_rank = -1, comm_world = 0;
    bendy_solver *peer[2] = {nullptr, nullptr};  // same-process transport (1-GPU emulation of strips)
    // links that cross a strip edge (cut bodies): host tables in USER particle indices, device tables in internal ones
    struct XlHost {
        uint32_t mine, slot;
        float len;
        uint32_t i_am_a;
    };
    std::vector<XlHost> xl;                  // sorted by colour
    std::vector<uint32_t> xl_colour_start;   // n_colours + 1
    std::vector<uint32_t> xl_send[2];        // my endpoints the left / right neighbour needs (user indices, agreed order)
    uint32_t xl_recv[2] = {0, 0};            // endpoints I receive from the left / right neighbour
    DevBuf<CrossLink> d_xl;
    DevBuf<uint32_t> d_xl_send_idx;          // [left block | right block] internal indices
    DevBuf<float2> d_xl_sendbuf, d_xl_ghost; // [left block | right block]
    cudaEvent_t ev_phase_a = nullptr, ev_xchg = nullptr;

    // ---------------- per-update params
    DevBuf<StepParams> d_prm;
    StepParams prm{};
    bool prm_valid = false;
    StepParams *h_prm_ring = nullptr;  // pinned
    uint32_t prm_ring_pos = 0;
    static constexpr uint32_t kPrmRing = 64;
    cudaEvent_t prm_ring_ev[kPrmRing] = {};

    // ---------------- launch machinery
    bool profiling = false;
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_substeps = 0;
    uint32_t kernels_per_substep = 0;
    uint64_t launches = 0;
    double k_ms[BENDY_K_CLASSES] = {};
    uint64_t k_launches[BENDY_K_CLASSES] = {};
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool capturing = false;
    uint32_t count_in_capture = 0;

    ~bendy_solver();
};

namespace {

// inside Ops member functions
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) return this->fail_cuda(_e, #call, __LINE__);                           \
    } while (0)

// per-call constants of one substep launch sequence
struct SubstepCtx {
    cudaStream_t st, qc, qg;  // main / circle-chain / polygon-chain streams (all == st when not capturing)
    const StepParams *prm;
    float2 *pos;
    const float *dk;
    bool K, discs, contact, halo, branch, acc, fuse_count;
    int pdl;  // programmatic dependent launch level for the particle chain (0 = plain launches)
    bool circ_joined;  // the circle chain's join event was already recorded (after the bins)
    uint32_t nPoly, n_in_parts;
    K3CountArgs ca;
    K1Args k1;
    K4Args k4;
};

struct Ops {  // helper with access to the solver; keeps bendy_solver a plain struct
    bendy_solver *s;
    int fail(int code, const std::string &msg) {
        s->err = msg;
        g_last_error = msg;
        return code;
    }
    int fail_cuda(cudaError_t e, const char *what, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s (%s) at solver.cu:%d: %s", cudaGetErrorName(e),
                 cudaGetErrorString(e), line, what);
        s->sticky = BENDY_ERR_CUDA;
        return fail(BENDY_ERR_CUDA, buf);
    }

    int bind() {
        cudaError_t e = cudaSetDevice(s->device);
        if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice", __LINE__);
        return BENDY_OK;
    }

    // ---- state movement -------------------------------------------------------------------
    int pull();         // device -> host vectors (user order)
    int rebuild();      // plan + upload everything
    int ensure_ready(); // before update / device reads
    int configure(float dt, float gx, float gy, float bx, float by, float bw, float bh);
    int grid_for(float bx, float by, float bw, float bh, StepParams *p, uint32_t *ncells);
    int enqueue_substeps(uint32_t count);
    int launch_substep(int phase = PHASE_ALL);
    int end_update(cudaStream_t q);
    bool work_stats_on() const;
    void choose_narrow_order();
    int halo_exchange_nccl(cudaStream_t q);
    SubstepCtx make_ctx();
    int launch_links_local(const SubstepCtx &c, cudaStream_t q, uint32_t p0, uint32_t p1, int halo_mode);
    int launch_links_global(const SubstepCtx &c, cudaStream_t q);
    int launch_links_cross(const SubstepCtx &c);
    int launch_polygon_chain(const SubstepCtx &c);
    int launch_circle_chain(SubstepCtx &c);
    int launch_count_unlinked(const SubstepCtx &c);
    int launch_halo_receive(const SubstepCtx &c, cudaStream_t q_clear);
    int launch_particle_links(const SubstepCtx &c, int phase, bool *ghosts_done);
    int launch_grid_build(const SubstepCtx &c);
    int launch_collide_integrate_discs(const SubstepCtx &c, int phase);
    int launch_circle_tail(const SubstepCtx &c);
    int launch_collide_integrate_plain(const SubstepCtx &c);
    int build_graph(uint32_t substeps);
    void drop_graph();
    int flush_events();
    int check_flags();

    template <typename F>
    int launch(int cls, F &&f);
};

void Ops::drop_graph() {
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    s->graph_exec = nullptr;
    s->graph_substeps = 0;
}

template <typename F>
int Ops::launch(int cls, F &&f) {
    if (s->capturing) {
        f();
        s->count_in_capture++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail_cuda(e, "kernel launch (capture)", __LINE__);
        return BENDY_OK;
    }
    if (s->profiling) {
        cudaEvent_t e0, e1;
        for (cudaEvent_t *pe : {&e0, &e1}) {
            if (!s->event_pool.empty()) {
                *pe = s->event_pool.back();
                s->event_pool.pop_back();
            } else {
                CK(cudaEventCreate(pe));
            }
        }
        CK(cudaEventRecord(e0, s->stream));
        f();
        CK(cudaEventRecord(e1, s->stream));
        s->pending.push_back(PendingEvent{cls, e0, e1});
    } else {
        f();
    }
    s->launches++;
    s->k_launches[cls]++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "kernel launch", __LINE__);
    if (s->pending.size() >= 8192) return flush_events();
    return BENDY_OK;
}

int Ops::flush_events() {
    if (s->pending.empty()) return BENDY_OK;
    CK(cudaStreamSynchronize(s->stream));
    for (PendingEvent &pe : s->pending) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pe.e0, pe.e1));
        s->k_ms[pe.cls] += ms;
        s->event_pool.push_back(pe.e0);
        s->event_pool.push_back(pe.e1);
    }
    s->pending.clear();
    return BENDY_OK;
}

int Ops::check_flags() {
    if (!s->d_flags.p) return BENDY_OK;
    int f = 0;
    CK(cudaMemcpyAsync(&f, s->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (f) {
        CK(cudaMemsetAsync(s->d_flags.p, 0, sizeof(int), s->stream));
        // Only the particle-polygon contact (ext) takes its candidates from the tile lists.  The reference's
        // polygon<->polygon pass is exact whatever the bins hold: the pair pre-scan answers an overflowing
        // tile or an unbinned polygon with "start at row 0" and the pass itself compares all boxes.
        const bool contact = s->polygon_contact && !s->polys.empty() && s->nOwned > 0;
        if (contact && (f & (FLAG_POLY_TILE_OVERFLOW | FLAG_POLY_SPAN_OVERFLOW)))
            return fail(BENDY_ERR_UNSUPPORTED,
                        "polygon broadphase overflow: more than 7 polygons share one tile or a polygon spans "
                        "more than 64 tiles; particle-polygon contacts were dropped for the overflowing tiles");
    }
    return BENDY_OK;
}

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// device -> host (USER order).  Used before the scene is edited after a run.
int Ops::pull() {
    if (s->host_valid) return BENDY_OK;
    if (!s->device_valid) return fail(BENDY_ERR_ARG, "internal: no valid state");
    if (int rc = bind()) return rc;
    CK(cudaStreamSynchronize(s->stream));
    std::vector<float2> tmp(s->N);
    for (int which = 0; which < 2; which++) {
        const float2 *src = which == 0 ? s->d_pos.p : s->d_prev.p;
        if (s->N) CK(cudaMemcpy(tmp.data(), src, (size_t)s->N * sizeof(float2), cudaMemcpyDeviceToHost));
        std::vector<float2> &P = which == 0 ? s->p_pos : s->p_prev;
        std::vector<float2> &Cc = which == 0 ? s->c_pos : s->c_prev;
        std::vector<float2> &G = which == 0 ? s->g_pos : s->g_prev;
        for (uint32_t i = 0; i < s->nOwned; i++) P[i] = tmp[s->plan_p.rank[i]];
        for (uint32_t i = 0; i < s->nC; i++) Cc[i] = tmp[s->nP + i];
        for (uint32_t i = 0; i < s->nG; i++) G[i] = tmp[s->nP + s->nC + i];
    }
    if (s->accel_pending && s->d_accel.p) {
        CK(cudaMemcpy(tmp.data(), s->d_accel.p, (size_t)s->N * sizeof(float2), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < s->nC; i++) s->c_acc[i] = tmp[s->nP + i];
        for (uint32_t i = 0; i < s->nG; i++) s->g_acc[i] = tmp[s->nP + s->nC + i];
    } else {
        // acc was consumed and cleared by the first integrate (particle.rs:24)
        std::fill(s->c_acc.begin(), s->c_acc.end(), make_float2(0.f, 0.f));
        for (size_t k = 0; k < s->polys.size(); k++)
            if (!s->polys[k].is_static)
                std::fill(s->g_acc.begin() + s->polys[k].start, s->g_acc.begin() + s->polys[k].start + s->polys[k].nv,
                          make_float2(0.f, 0.f));
        s->any_acc = false;
        for (const float2 &a : s->g_acc)
            if (a.x != 0.f || a.y != 0.f) s->any_acc = true;
    }
    // polygon centres: Polygon.center after update() is the mean of the pre-integrate positions
    // (polygon.rs:130 -> calc_center on what is now prev_pos); static polygons keep the value
    // computed by solve_links (polygon.rs:219), which the device stores in d_poly_center.
    if (!s->polys.empty()) {
        std::vector<float2> cen(s->polys.size());
        CK(cudaMemcpy(cen.data(), s->d_poly_center.p, cen.size() * sizeof(float2), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < s->polys.size(); k++) {
            PolyHost &P = s->polys[k];
            if (P.is_static) {
                P.center = cen[k];
            } else {
                float cx = 0.f, cy = 0.f;
                for (uint32_t v = 0; v < P.nv; v++) {
                    cx = cx + s->g_prev[P.start + v].x;
                    cy = cy + s->g_prev[P.start + v].y;
                }
                P.center = make_float2(cx / (float)P.nv, cy / (float)P.nv);
            }
        }
    }
    s->host_valid = true;
    return BENDY_OK;
}

template <typename T>
static cudaError_t upload(DevBuf<T> &d, const std::vector<T> &h, cudaStream_t st) {
    cudaError_t e = d.ensure(std::max<size_t>(h.size(), 1));
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

int Ops::rebuild() {
    if (int rc = bind()) return rc;
    if (int rc = pull()) return rc;
    // The reference accepts any link when it is added and panics when it SOLVES one whose indices do not satisfy
    // a < b < len (split_at_mut(b), split.0[a], split.1[0]: link.rs:19-21 / 37-39), i.e. inside the first update.
    // Same place here: the scene is about to be planned for the device.
    for (size_t k = 0; k < s->pl_len.size(); k++)
        if (!(s->pl_ab[2 * k] < s->pl_ab[2 * k + 1]) || !(s->pl_ab[2 * k + 1] < s->p_pos.size()))
            return fail(BENDY_ERR_LINK, "particle link " + std::to_string(k) + " (" + std::to_string(s->pl_ab[2 * k]) + ", " +
                                            std::to_string(s->pl_ab[2 * k + 1]) + ") needs a < b < " +
                                            std::to_string(s->p_pos.size()) + " particles: the reference panics here (link.rs:19-21)");
    for (size_t k = 0; k < s->cl.size(); k++)
        if (!(s->cl[k].a < s->cl[k].b) || !(s->cl[k].b < s->c_pos.size()))
            return fail(BENDY_ERR_LINK, "circle link " + std::to_string(k) + " (" + std::to_string(s->cl[k].a) + ", " +
                                            std::to_string(s->cl[k].b) + ") needs a < b < " + std::to_string(s->c_pos.size()) +
                                            " circles: the reference panics here (link.rs:37-39)");
    drop_graph();
    s->nOwned = (uint32_t)s->p_pos.size();
    s->nP = s->nOwned + (s->halo_on ? 2 * s->ghost_cap : 0);  // disc slots = owned + ghosts
    s->nC = (uint32_t)s->c_pos.size();
    // Polygons are fine in a strip: they never receive anything from a particle (solver.rs:178-187 only pairs
    // polygons with polygons; the particle-polygon extension only moves the particle), so every strip carries
    // an identical copy that evolves identically.  Circles are replicated as well: every strip runs the same
    // circle links and circle-circle pass, and the fixed-point corrections its OWN discs collected for each
    // Circle are summed over all strips (an integer all-reduce, hence order-free and bit-identical to the
    // unsharded sum) before the circles' tail applies them.  Inverse masses would need the neighbours' scales
    // next to the ghost positions: not yet.
    // Inverse masses (ext): the scale of every packed disc travels with its position into the neighbour's ghost
    // slots, so a ghost weighs in a contact exactly as it does on its owner.  Not combined with links across strip
    // edges yet (the remote endpoint's scale would have to travel with the endpoint positions).
    if (s->halo_on && (!s->p_k.empty() || !s->c_k.empty()) && !s->xl.empty())
        return fail(BENDY_ERR_UNSUPPORTED, "strips: inverse masses together with links across strip edges are not supported yet");
    s->nG = (uint32_t)s->g_pos.size();
    s->N = s->nP + s->nC + s->nG;
    s->Npad = (s->N + 1u) & ~1u;
    std::string perr;
    // strips: bodies close to a halo band are relaxed first so that the exchange can overlap the rest
    std::vector<uint8_t> prio;
    if (s->halo_on && s->ghost_cap && s->halo_overlap) {
        const float gl = std::isfinite(s->halo_xl) ? (s->halo_xl - s->stray_xl) / 3.0f : 0.f;  // = band / 2
        const float gr = std::isfinite(s->halo_xr) ? (s->stray_xr - s->halo_xr) / 3.0f : 0.f;
        prio.resize(s->nOwned);
        for (uint32_t i = 0; i < s->nOwned; i++) {
            const float x = s->p_pos[i].x;
            prio[i] = (x < s->halo_xl + gl || x > s->halo_xr - gr || !(x == x)) ? 1 : 0;
        }
    }
    if (!plan_links(s->nOwned, s->pl_ab.data(), s->pl_len.data(), s->pl_len.size(), s->plan_params, false, &s->plan_p,
                    &perr, prio.empty() ? nullptr : prio.data()))
        return fail(BENDY_ERR_UNSUPPORTED, perr);
    // ---- shared-memory opt-in: a CTA may use up to 227 KB, but anything above 48 KB (static + dynamic)
    // has to be requested per kernel
    {
        uint32_t maxp = 0;
        for (uint32_t p = 0; p < s->plan_p.n_parts(); p++)
            maxp = std::max(maxp, s->plan_p.part_start[p + 1] - s->plan_p.part_start[p]);
        const bool with_k = !s->p_k.empty() || !s->c_k.empty();
        const size_t k3_bytes = (size_t)maxp * (with_k ? 12 : 8);
        if (k3_bytes > 40 * 1024) {
            const int b = (int)k3_bytes;
            const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
            CK(cudaFuncSetAttribute(k3_links_local<false, false, 0>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<false, true, 0>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<true, false, 0>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<true, true, 0>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<false, true, 1>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<false, true, 2>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<true, true, 1>, at, b));
            CK(cudaFuncSetAttribute(k3_links_local<true, true, 2>, at, b));
        }
        // k_circles_exact: 45 KB of static tables + 12 B per circle (up to 4096 circles) of dynamic
        if (s->nC > 128 && s->nC <= 4096)
            CK(cudaFuncSetAttribute(k_circles_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12));
    }
    // ---- state upload (internal order)
    std::vector<float2> pos(s->Npad), prev(s->Npad);
    for (uint32_t i = 0; i < s->nOwned; i++) {
        pos[s->plan_p.rank[i]] = s->p_pos[i];
        prev[s->plan_p.rank[i]] = s->p_prev[i];
    }
    for (uint32_t i = s->nOwned; i < s->nP; i++) pos[i] = prev[i] = make_float2(NAN, NAN);  // empty ghost slots
    std::copy(s->c_pos.begin(), s->c_pos.end(), pos.begin() + s->nP);
    std::copy(s->c_prev.begin(), s->c_prev.end(), prev.begin() + s->nP);
    std::copy(s->g_pos.begin(), s->g_pos.end(), pos.begin() + s->nP + s->nC);
    std::copy(s->g_prev.begin(), s->g_prev.end(), prev.begin() + s->nP + s->nC);
    for (uint32_t i = s->N; i < s->Npad; i++) pos[i] = prev[i] = make_float2(0.f, 0.f);
    CK(upload(s->d_pos, pos, s->stream));
    CK(upload(s->d_prev, prev, s->stream));
    CK(s->d_stage.ensure(2 * std::max<size_t>(s->Npad, 1)));  // pos | prev halves of the host <-> device staging
    s->accel_pending = false;
    if (s->any_acc) {
        std::vector<float2> acc(s->Npad, make_float2(0.f, 0.f));
        std::copy(s->c_acc.begin(), s->c_acc.end(), acc.begin() + s->nP);
        std::copy(s->g_acc.begin(), s->g_acc.end(), acc.begin() + s->nP + s->nC);
        CK(upload(s->d_accel, acc, s->stream));
        s->accel_pending = true;
    }
    s->has_k = !s->p_k.empty() || !s->c_k.empty();
    if (s->has_k) {
        std::vector<float> k(s->Npad, 1.0f);
        if (!s->p_k.empty())
            for (uint32_t i = 0; i < s->nOwned; i++) k[s->plan_p.rank[i]] = s->p_k[i];
        if (!s->c_k.empty()) std::copy(s->c_k.begin(), s->c_k.end(), k.begin() + s->nP);
        CK(upload(s->d_k, k, s->stream));
    }
    CK(upload(s->d_crad, s->c_rad, s->stream));
    CK(upload(s->d_rank, s->plan_p.rank, s->stream));
    // ---- link tables
    CK(upload(s->d_part_start, s->plan_p.part_start, s->stream));
    CK(upload(s->d_part_cs, s->plan_p.part_colour_start, s->stream));
    CK(upload(s->d_local, s->plan_p.local_links, s->stream));
    CK(upload(s->d_global, s->plan_p.global_links, s->stream));
    {  // polygon-internal links stay in insertion order with polygon-local indices (polygon.rs:218-223)
        std::vector<uint32_t> ls(s->polys.size() + 1, 0);
        for (size_t k = 0; k < s->polys.size(); k++) ls[k] = s->polys[k].link_start, ls[k + 1] = s->polys[k].link_start + s->polys[k].nl;
        CK(upload(s->d_poly_link_start, ls, s->stream));
        CK(upload(s->d_poly_link_ab, s->gl_ab, s->stream));
        CK(upload(s->d_poly_link_len, s->gl_len, s->stream));
    }
    CK(upload(s->d_clinks, s->cl, s->stream));
    if (!s->xl.empty()) {  // links across strip edges: user -> internal indices
        std::vector<CrossLink> x(s->xl.size());
        for (size_t k = 0; k < x.size(); k++)
            x[k] = CrossLink{s->plan_p.rank[s->xl[k].mine], s->xl[k].slot, s->xl[k].len, s->xl[k].i_am_a};
        CK(upload(s->d_xl, x, s->stream));
        std::vector<uint32_t> idx;
        for (int side = 0; side < 2; side++)
            for (uint32_t u : s->xl_send[side]) idx.push_back(s->plan_p.rank[u]);
        CK(upload(s->d_xl_send_idx, idx, s->stream));
        CK(s->d_xl_sendbuf.ensure(std::max<size_t>(idx.size(), 1)));
        CK(s->d_xl_ghost.ensure(std::max<size_t>((size_t)s->xl_recv[0] + s->xl_recv[1], 1)));
    }
    // ---- polygons
    {
        std::vector<uint32_t> pstart(s->polys.size() + 1, 0);
        std::vector<uint8_t> pstatic(s->polys.size()), gstatic(s->nG);
        std::vector<float2> cen(s->polys.size());
        float ext = 0.f;
        for (size_t k = 0; k < s->polys.size(); k++) {
            const PolyHost &P = s->polys[k];
            pstart[k] = P.start;
            pstart[k + 1] = P.start + P.nv;
            pstatic[k] = P.is_static ? 1 : 0;
            cen[k] = P.center;
            float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
            for (uint32_t v = 0; v < P.nv; v++) {
                gstatic[P.start + v] = pstatic[k];
                float2 p = s->g_pos[P.start + v];
                x0 = std::fmin(x0, p.x), y0 = std::fmin(y0, p.y), x1 = std::fmax(x1, p.x), y1 = std::fmax(y1, p.y);
            }
            if (std::isfinite(x1 - x0)) ext = std::fmax(ext, x1 - x0);
            if (std::isfinite(y1 - y0)) ext = std::fmax(ext, y1 - y0);
        }
        s->poly_tile = ext > 0.f ? ext : 1.0f;
        CK(upload(s->d_poly_start, pstart, s->stream));
        CK(upload(s->d_poly_static, pstatic, s->stream));
        CK(upload(s->d_gstatic, gstatic, s->stream));
        CK(upload(s->d_poly_center, cen, s->stream));
        CK(s->d_poly_box.ensure(std::max<size_t>(s->polys.size(), 1)));
        CK(s->d_poly_first_row.ensure(1));
        CK(cudaMemsetAsync(s->d_poly_first_row.p, 0xFF, sizeof(uint32_t), s->stream));
    }
    if (!s->d_flags.p) {
        CK(s->d_flags.ensure(8));  // [0] error flags, [1..4] circle-pass fallbacks (statistics), [5] polygon tiles built
        CK(cudaMemsetAsync(s->d_flags.p, 0, 8 * sizeof(int), s->stream));
    }
    CK(cudaMemsetAsync(s->d_flags.p + 5, 0, sizeof(int), s->stream));  // the polygon tiles are rebuilt from scratch
    CK(s->d_prm.ensure(1));
    if (!s->h_prm_ring) {
        CK(cudaHostAlloc(&s->h_prm_ring, sizeof(StepParams) * bendy_solver::kPrmRing, cudaHostAllocDefault));
        for (uint32_t i = 0; i < bendy_solver::kPrmRing; i++)
            CK(cudaEventCreateWithFlags(&s->prm_ring_ev[i], cudaEventDisableTiming));
    }
    // grid buffers that depend only on the particle count
    if (s->nP) {
        CK(s->d_sorted_id.ensure(s->nP));
        CK(s->d_slot_of.ensure(s->nP));
        CK(s->d_sorted_pos.ensure(s->nP));
        const size_t n_work = (size_t)cdiv(s->nOwned ? s->nOwned : 1u, NARROW_THREADS) * NARROW_WARPS;
        CK(s->d_chunk_work.ensure(n_work));
        CK(cudaMemsetAsync(s->d_chunk_work.p, 0, n_work * sizeof(uint32_t), s->stream));
        if (!s->h_work_halves) CK(cudaHostAlloc(&s->h_work_halves, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
        s->h_work_halves[0] = s->h_work_halves[1] = 0ull;
    }
    if (s->halo_on && s->ghost_cap) {
        for (int side = 0; side < 2; side++) CK(s->d_send[side].ensure(s->ghost_cap));
        if (s->has_k)
            for (int side = 0; side < 2; side++) CK(s->d_send_k[side].ensure(s->ghost_cap));
        CK(s->d_send_cnt.ensure(8));
        CK(cudaMemsetAsync(s->d_send_cnt.p, 0, 8 * sizeof(uint32_t), s->stream));
        k_halo_clear<<<cdiv(s->ghost_cap, 256), 256, 0, s->stream>>>(s->d_send[0].p, s->d_send[1].p, s->d_send_cnt.p,
                                                                     s->ghost_cap);
        CK(cudaGetLastError());
    }
    if (s->nC) {
        CK(s->d_circ_snap.ensure(s->nC));
        CK(s->d_circ_acc.ensure(2 * (size_t)s->nC));
        CK(cudaMemsetAsync(s->d_circ_acc.p, 0, 2 * (size_t)s->nC * sizeof(unsigned long long), s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));  // host staging vectors die here
    s->prm_valid = false;
    s->n_cells = 0;
    s->topo_dirty = false;
    s->device_valid = true;
    return BENDY_OK;
}

int Ops::ensure_ready() {
    if (s->sticky != BENDY_OK) return fail(s->sticky, s->err);
    if (s->topo_dirty || !s->device_valid) return rebuild();
    return bind();
}

// broadphase grid for these bounds: origin = bounds.pos, h >= 2*r_p, cell count bounded
int Ops::grid_for(float bx, float by, float bw, float bh, StepParams *p, uint32_t *ncells) {
    double wx = std::isfinite(bw) && bw > 0.f ? bw : 1.0, wy = std::isfinite(bh) && bh > 0.f ? bh : 1.0;
    // a strip only needs cells where its own discs and ghosts can be: clip the grid to the window
    // (discs outside are clamped into the border cells, which stays correct)
    if (s->win_x0 > bx && s->win_x0 < bx + (float)wx) {
        wx -= (double)(s->win_x0 - bx);
        bx = s->win_x0;
    }
    if (s->win_x1 > bx && s->win_x1 < bx + (float)wx) wx = (double)(s->win_x1 - bx);
    float h = s->grid_cell;
    if (!(h > 0.f)) {
        // auto: 4.2*r_p, so that a disc's partners lie in a 2x2 block of cells with as few bystanders as possible;
        // larger only when that would be more than about four cells per particle (a sparse world: the per-cell
        // scan traffic would exceed the per-disc traffic).  Measured on strips: the end ranks' windows reach the
        // world's walls, and with the former two cells per particle their cells grew to 0.9 (4.7x the candidates
        // per disc, narrowphase 130 instead of 56 us for 2M discs).
        double np = std::max<double>(s->p_pos.size(), 1.0);
        h = std::max((float)std::sqrt(wx * wy / (4.0 * np)), 4.2f * s->particle_radius);
    }
    if (h < 2.0f * s->particle_radius) h = 2.0f * s->particle_radius;
    if (!(h > 0.f)) h = 1.0f;
    const double max_cells = 67108864.0;  // 2^26
    while (std::ceil(wx / h) * std::ceil(wy / h) > max_cells) h *= 1.25f;
    int nx = (int)std::ceil(wx / h), ny = (int)std::ceil(wy / h);
    nx = std::max(nx, 1), ny = std::max(ny, 1);
    p->gox = bx, p->goy = by, p->h = h, p->inv_h = 1.0f / h;
    p->nx = nx, p->ny = ny;
    p->quad = h >= 4.2f * s->particle_radius ? 1 : 0;
    p->tnx = (nx + (1 << BENDY_TILE_SHIFT) - 1) >> BENDY_TILE_SHIFT;
    p->tny = (ny + (1 << BENDY_TILE_SHIFT) - 1) >> BENDY_TILE_SHIFT;
    *ncells = (uint32_t)nx * (uint32_t)ny;
    return BENDY_OK;
}

int Ops::configure(float dt, float gx, float gy, float bx, float by, float bw, float bh) {
    StepParams p = s->prm;
    p.gx = gx, p.gy = gy, p.dt = dt;
    p.gdt2x = (gx * dt) * dt;  // particle.rs:23: acc * dt * dt, left-associative
    p.gdt2y = (gy * dt) * dt;
    p.lo_x = bx, p.lo_y = by;
    p.hi_x = bx + bw;  // particle.rs:32
    p.hi_y = by + bh;  // particle.rs:41
    p.rp = s->particle_radius;
    p.rs = p.rp + p.rp;  // circle.rs:36: r_a + r_b
    p.rs2 = p.rs * p.rs;
    p.rp2 = p.rp * p.rp;  // circle.rs:39-40
    {
        const volatile float two_rp2 = p.rp2 + p.rp2;  // (volatile: one rounding per operation, whatever the host flags)
        p.scale_u = 1.0f / two_rp2;                    // circle.rs:41
    }
    p.halo_xl = s->halo_on ? s->halo_xl : -INFINITY;
    p.halo_xr = s->halo_on ? s->halo_xr : INFINITY;
    p.stray_xl = s->halo_on ? s->stray_xl : -INFINITY;
    p.stray_xr = s->halo_on ? s->stray_xr : INFINITY;
    uint32_t ncells = 0;
    const bool discs = s->particle_radius > 0.f && s->nP > 0;
    if (discs) {
        grid_for(bx, by, bw, bh, &p, &ncells);
    } else {
        p.nx = p.ny = p.tnx = p.tny = 1, p.gox = bx, p.goy = by, p.h = 1.f, p.inv_h = 1.f, p.quad = 0;
    }
    const bool contact = s->polygon_contact && !s->polys.empty() && s->nP > 0;
    const bool poly_tiles = contact || s->polys.size() >= 2;  // also the polygon-polygon pre-scan bins
    uint32_t ptiles = 0;
    if (poly_tiles) {
        float t = s->poly_tile;
        double wx = std::isfinite(bw) && bw > 0.f ? bw : 1.0, wy = std::isfinite(bh) && bh > 0.f ? bh : 1.0;
        while (std::ceil(wx / t) * std::ceil(wy / t) > 4194304.0) t *= 1.25f;
        p.pox = bx, p.poy = by, p.psize = t, p.pinv = 1.0f / t;
        p.pnx = std::max(1, (int)std::ceil(wx / t)), p.pny = std::max(1, (int)std::ceil(wy / t));
        ptiles = (uint32_t)p.pnx * (uint32_t)p.pny;
    } else {
        p.pox = bx, p.poy = by, p.psize = 1.f, p.pinv = 1.f, p.pnx = p.pny = 1;
    }
    bool same = s->prm_valid && std::memcmp(&p, &s->prm, sizeof p) == 0;
    if (same) return BENDY_OK;
    // grid shape changes invalidate captured launch dimensions
    // (quad selects the narrowphase instance a captured graph holds)
    bool shape = !s->prm_valid || p.nx != s->prm.nx || p.ny != s->prm.ny || p.pnx != s->prm.pnx ||
                 p.pny != s->prm.pny || p.quad != s->prm.quad;
    if (shape) {
        drop_graph();
        if (discs) {
            s->n_cells = ncells;
            // n_cells + 1 counters (the extra cell collects non-finite points), padded to whole scan tiles
            s->n_scan_tiles = cdiv(ncells + 1, SCAN_TILE);
            const size_t padded = (size_t)s->n_scan_tiles * SCAN_TILE;
            CK(s->d_cell_count.ensure(padded));
            CK(cudaMemsetAsync(s->d_cell_count.p, 0, padded * sizeof(uint32_t), s->stream));
            CK(s->d_cell_start.ensure(padded));
            CK(s->d_tile_sum.ensure(s->n_scan_tiles));
            CK(cudaMemsetAsync(s->d_tile_sum.p, 0, (size_t)s->n_scan_tiles * sizeof(uint32_t), s->stream));
            if (s->nC) {
                s->n_circ_tiles = (uint32_t)p.tnx * (uint32_t)p.tny;
                CK(s->d_circ_tile_count.ensure(s->n_circ_tiles));
                CK(cudaMemsetAsync(s->d_circ_tile_count.p, 0, (size_t)s->n_circ_tiles * sizeof(uint32_t), s->stream));
                CK(s->d_circ_tile_ids.ensure((size_t)s->n_circ_tiles * BENDY_CIRC_CAP));
            }
        }
        if (poly_tiles) {
            s->n_poly_tiles = ptiles;
            CK(s->d_poly_tiles.ensure((size_t)ptiles * (BENDY_POLY_CAP + 1)));
        }
    }
    // the fused polygon chain only rebuilds its tile lists when an AABB moved: a new tile geometry invalidates them
    if (poly_tiles && s->d_flags.p &&
        (!s->prm_valid || p.pox != s->prm.pox || p.poy != s->prm.poy || p.pinv != s->prm.pinv || p.pnx != s->prm.pnx ||
         p.pny != s->prm.pny))
        CK(cudaMemsetAsync(s->d_flags.p + 5, 0, sizeof(int), s->stream));
    s->prm = p;
    s->prm_valid = true;
    // stage through a pinned ring so the async copy never reads a host value that changed later
    uint32_t slot = s->prm_ring_pos++ % bendy_solver::kPrmRing;
    if (s->prm_ring_pos > bendy_solver::kPrmRing) CK(cudaEventSynchronize(s->prm_ring_ev[slot]));
    s->h_prm_ring[slot] = p;
    CK(cudaMemcpyAsync(s->d_prm.p, &s->h_prm_ring[slot], sizeof p, cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->prm_ring_ev[slot], s->stream));
    return BENDY_OK;
}

// <<<grid, block, smem, q>>> with the programmatic-stream-serialisation attribute when `pdl` is set: the
// kernel may become resident while its stream predecessor drains and orders itself with pdl_wait()
template <typename... KArgs, typename... Args>
static void launch_k(bool pdl, void (*kern)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t q,
                     Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = q;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);  // errors surface through cudaGetLastError() in launch()
}

#define LAUNCH(cls, ...)                                         \
    do {                                                         \
        int _rc = launch((cls), [&]() { __VA_ARGS__; });         \
        if (_rc) return _rc;                                     \
    } while (0)

// one substep, reference order (solver.rs:109-115):
//   gravity (fused into integrate) -> links -> dynamic collisions -> bounds -> integrate.
// Free particles, circles and polygon points are disjoint worlds until the collision phase
// (solver.rs:143-153), so while capturing a graph the circle and polygon chains run as parallel
// branches beside the particle chain; eager (profiling) launches are simply serial.
int Ops::halo_exchange_nccl(cudaStream_t q) {
    if (!s->nccl_comm) return BENDY_OK;  // a single strip: nothing to exchange
    const size_t nflt = 2 * (size_t)s->ghost_cap;
    float2 *ghost = s->d_pos.p + s->nOwned;
    const int left = s->comm_rank - 1, right = s->comm_rank + 1;
    auto ck = [&](ncclResult_t r, const char *what) -> int {
        if (r == ncclSuccess) return BENDY_OK;
        s->sticky = BENDY_ERR_CUDA;
        return fail(BENDY_ERR_CUDA, std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
    };
    if (int rc = ck(g_nccl.GroupStart(), "ncclGroupStart")) return rc;
    float *ghost_k = s->has_k ? s->d_k.p + s->nOwned : nullptr;  // the ghosts' inverse-mass scales travel along
    if (left >= 0) {
        if (int rc = ck(g_nccl.Send(s->d_send[0].p, nflt, ncclFloat, left, s->nccl_comm, q), "ncclSend")) return rc;
        if (int rc = ck(g_nccl.Recv(ghost, nflt, ncclFloat, left, s->nccl_comm, q), "ncclRecv")) return rc;
        if (ghost_k) {
            if (int rc = ck(g_nccl.Send(s->d_send_k[0].p, s->ghost_cap, ncclFloat, left, s->nccl_comm, q), "ncclSend")) return rc;
            if (int rc = ck(g_nccl.Recv(ghost_k, s->ghost_cap, ncclFloat, left, s->nccl_comm, q), "ncclRecv")) return rc;
        }
    }
    if (right < s->comm_world) {
        if (int rc = ck(g_nccl.Send(s->d_send[1].p, nflt, ncclFloat, right, s->nccl_comm, q), "ncclSend")) return rc;
        if (int rc = ck(g_nccl.Recv(ghost + s->ghost_cap, nflt, ncclFloat, right, s->nccl_comm, q), "ncclRecv"))
            return rc;
        if (ghost_k) {
            if (int rc = ck(g_nccl.Send(s->d_send_k[1].p, s->ghost_cap, ncclFloat, right, s->nccl_comm, q), "ncclSend")) return rc;
            if (int rc = ck(g_nccl.Recv(ghost_k + s->ghost_cap, s->ghost_cap, ncclFloat, right, s->nccl_comm, q), "ncclRecv"))
                return rc;
        }
    }
    if (int rc = ck(g_nccl.GroupEnd(), "ncclGroupEnd")) return rc;
    if (s->capturing)
        s->count_in_capture++;
    else
        s->launches++, s->k_launches[BENDY_K_HALO]++;
    return BENDY_OK;
}

// ------------------------------------------------------------------------------------------------
// One substep, reference order (solver.rs:109-115):
//   gravity (fused into integrate) -> links -> dynamic collisions -> bounds -> integrate.
// Free particles, circles and polygon points are disjoint worlds until the collision phase
// (solver.rs:143-153).  While a graph is being captured the side streams were forked from the main
// stream by build_graph(): the circle chain runs on side[0], the polygon chain on side[1]; their
// tails (integrate) of substep k are queued behind the narrowphase of substep k and overlap the
// particle chain of substep k+1, which never touches circle or polygon state before ITS join.
// Eager (profiling) launches are simply serial on the main stream.
SubstepCtx Ops::make_ctx() {
    SubstepCtx c{};
    c.st = s->stream;
    c.branch = s->capturing;
    c.nPoly = (uint32_t)s->polys.size();
    c.qc = (c.branch && s->nC > 0) ? s->side[0] : c.st;
    c.qg = (c.branch && c.nPoly > 0) ? s->side[1] : c.st;
    c.prm = s->d_prm.p;
    c.pos = s->d_pos.p;
    c.K = s->has_k;
    c.dk = c.K ? s->d_k.p : nullptr;
    c.discs = s->particle_radius > 0.f && s->nP;
    c.contact = s->polygon_contact && c.nPoly && s->nP;
    c.halo = s->halo_on && c.discs && s->ghost_cap > 0;
    c.acc = s->accel_pending;
    c.pdl = (s->nccl_comm && !s->pdl_nccl) ? 0 : s->pdl;
    c.ca = K3CountArgs{c.prm,
                       s->n_cells,
                       s->d_cell_count.p,
                       s->d_tile_sum.p,
                       c.halo ? s->d_send[0].p : nullptr,
                       c.halo ? s->d_send[1].p : nullptr,
                       c.halo ? s->d_send_cnt.p : nullptr,
                       c.halo ? s->ghost_cap : 0u,
                       (c.halo && c.K) ? s->d_send_k[0].p : nullptr,
                       (c.halo && c.K) ? s->d_send_k[1].p : nullptr};
    const LinkPlan &P = s->plan_p;
    c.n_in_parts = P.n_parts() ? P.part_start.back() : 0u;
    // the histogram (and the halo packing) must see the positions after ALL links: global colours and links that
    // cross a strip edge run behind the partition kernel
    c.fuse_count = c.discs && P.n_global_colours() == 0 && c.n_in_parts > 0 && s->xl.empty();
    c.k1 = K1Args{c.pos, s->d_prev.p, c.acc ? s->d_accel.p : nullptr, c.dk, s->d_crad.p, s->d_gstatic.p, s->nP, s->nC, s->N};
    c.k4 = K4Args{c.pos + s->nP + s->nC, s->d_poly_start.p, s->d_poly_center.p, s->d_poly_box.p, s->d_poly_tiles.p,
                  s->d_poly_static.p};
    return c;
}

// shared-memory partitions [p0, p1) of the particle link plan; halo_mode: 0 none, 1 pack, 2 check only
int Ops::launch_links_local(const SubstepCtx &c, cudaStream_t q, uint32_t p0, uint32_t p1, int halo_mode) {
    const LinkPlan &P = s->plan_p;
    if (p1 <= p0 || P.local_links.empty()) return BENDY_OK;
    uint32_t maxp = 0;
    for (uint32_t p = 0; p < P.n_parts(); p++) maxp = std::max(maxp, P.part_start[p + 1] - P.part_start[p]);
    const size_t smem = (size_t)maxp * (c.K ? 12 : 8);
    const uint32_t T = s->k3_threads, np = p1 - p0, C = P.n_local_colours;
    const uint32_t *ps = s->d_part_start.p, *cs = s->d_part_cs.p;
    const LocalLink *ll = s->d_local.p;
    float2 *pos = c.pos;
    const float *dk = c.dk;
    const K3CountArgs &ca = c.ca;
    const bool pdl = c.pdl > 0;
#define K3L(HK, FC, HM) \
    LAUNCH(BENDY_K_LINKS_LOCAL, launch_k(pdl, k3_links_local<HK, FC, HM>, np, T, smem, q, pos, dk, 0u, ps, cs, ll, C, ca, p0))
    if (c.K && c.fuse_count && halo_mode == 1)
        K3L(true, true, 1);
    else if (c.K && c.fuse_count && halo_mode == 2)
        K3L(true, true, 2);
    else if (c.fuse_count && halo_mode == 1)
        K3L(false, true, 1);
    else if (c.fuse_count && halo_mode == 2)
        K3L(false, true, 2);
    else if (c.K && c.fuse_count)
        K3L(true, true, 0);
    else if (c.K)
        K3L(true, false, 0);
    else if (c.fuse_count)
        K3L(false, true, 0);
    else
        K3L(false, false, 0);
#undef K3L
    return BENDY_OK;
}

// links across partitions: one launch per global colour, after all local colours
int Ops::launch_links_global(const SubstepCtx &c, cudaStream_t q) {
    const LinkPlan &P = s->plan_p;
    for (uint32_t col = 0; col < P.n_global_colours(); col++) {
        const uint32_t l0 = P.gcolour_start[col], l1 = P.gcolour_start[col + 1];
        if (l0 == l1) continue;
        if (c.K)
            LAUNCH(BENDY_K_LINKS_GLOBAL,
                   k3_links_global<true><<<cdiv(l1 - l0, 256), 256, 0, q>>>(c.pos, c.dk, 0, s->d_global.p, l0, l1));
        else
            LAUNCH(BENDY_K_LINKS_GLOBAL,
                   k3_links_global<false><<<cdiv(l1 - l0, 256), 256, 0, q>>>(c.pos, c.dk, 0, s->d_global.p, l0, l1));
    }
    return BENDY_OK;
}

// polygon chain: centre (polygon.rs:219) -> own links (polygon.rs:220-222) -> AABB + bins, then the
// polygon half of solve_dynamic_collisions (solver.rs:178-187)
int Ops::launch_polygon_chain(const SubstepCtx &c) {
    if (!c.nPoly) return BENDY_OK;
    float2 *pts = c.pos + s->nP + s->nC;
    PolyArgs pa{pts,          s->d_poly_start.p, s->d_poly_static.p, c.nPoly, s->d_poly_center.p, s->d_poly_box.p,
                s->d_poly_tiles.p, s->d_flags.p,      s->d_poly_first_row.p};
    const bool bins = c.contact || c.nPoly >= 2;
    PolyLinkArgs la{s->d_poly_link_start.p, s->d_poly_link_ab.p, s->d_poly_link_len.p};
    if (c.nPoly <= 1024 && s->poly_fused) {
        // up to 1024 polygons: the whole chain in one launch of one CTA; the tiles are only rebuilt when an AABB moved
        const uint32_t threads = c.nPoly <= 128 ? 128 : (c.nPoly <= 256 ? 256 : (c.nPoly <= 512 ? 512 : 1024));
        LAUNCH(BENDY_K_POLY_PREP, launch_k(c.pdl > 2, k_polygons_fused, 1, threads, 0, c.qg, pts, pa, la, c.prm, s->n_poly_tiles,
                                           bins ? 1 : 0, s->d_flags.p + 5));
        return BENDY_OK;
    }
    if (bins)
        LAUNCH(BENDY_K_POLY_PREP, cudaMemsetAsync(s->d_poly_tiles.p, 0,
                                                  (size_t)s->n_poly_tiles * (BENDY_POLY_CAP + 1) * sizeof(uint32_t), c.qg));
    LAUNCH(BENDY_K_POLY_PREP, k_poly_prepare<<<cdiv(c.nPoly, 128), 128, 0, c.qg>>>(pts, pa, la, c.prm, bins ? 1 : 0));
    if (c.nPoly >= 2) {
        LAUNCH(BENDY_K_POLY_CONTACT, launch_k(c.pdl > 2, k4_poly_pair_prescan, cdiv(c.nPoly, 128), 128, 0, c.qg, pa, c.prm));
        LAUNCH(BENDY_K_POLY_CONTACT, launch_k(c.pdl > 2, k_polygons_exact, 1, 1024, 0, c.qg, pts, pa, c.prm, s->n_poly_tiles));
    }
    return BENDY_OK;
}

// circle chain: circle links (solver.rs:147-149) -> bins + entry snapshot -> circle-circle pass
// (solver.rs:168-177).  The disc contacts (ext) are tested against the circle centres at the ENTRY of
// the collision phase (the snapshot), so the narrowphase only waits for the bins (join event
// recorded here) and the exact pass overlaps it; the circles' tail follows both.
int Ops::launch_circle_chain(SubstepCtx &c) {
    float2 *cpos = c.pos + s->nP;
    if (!s->cl.empty())
        LAUNCH(BENDY_K_LINKS_CIRCLE, k3_circle_links<<<1, 32, 0, c.qc>>>(cpos, s->d_crad.p, s->d_clinks.p, (uint32_t)s->cl.size()));
    if (c.discs && s->nC) {
        LAUNCH(BENDY_K_CIRCLES, launch_k(c.pdl > 2, k2_circle_bin, cdiv(s->nC, 128), 128, 0, c.qc, cpos, s->d_crad.p, s->nC,
                                         c.prm, s->d_circ_tile_count.p, s->d_circ_tile_ids.p, s->d_circ_snap.p));
        if (c.branch && c.qc != c.st) {
            CK(cudaEventRecord(s->ev_join[0], c.qc));
            c.circ_joined = true;
        }
    }
    if (s->nC >= 2) {
        const size_t smem = s->nC <= 4096 ? (size_t)s->nC * 12 : 0;
        LAUNCH(BENDY_K_CIRCLE_PASS, launch_k(c.pdl > 2, k_circles_exact, 1, 1024, smem, c.qc, cpos, s->d_crad.p, s->nC,
                                             smem ? 1 : 0, s->d_flags.p + 1));
    }
    return BENDY_OK;
}

// histogram (+ halo packing) of the owned discs the link kernel did not cover
int Ops::launch_count_unlinked(const SubstepCtx &c) {
    if (!c.discs) return BENDY_OK;
    const uint32_t c0 = c.fuse_count ? c.n_in_parts : 0u;
    if (c0 >= s->nOwned) return BENDY_OK;
    if (c.halo)
        LAUNCH(BENDY_K_GRID_BUILD, k2_count<true><<<cdiv(s->nOwned - c0, 256), 256, 0, c.st>>>(c.pos, c.dk, c0, s->nOwned, c.ca));
    else
        LAUNCH(BENDY_K_GRID_BUILD, launch_k(c.pdl > 0, k2_count<false>, cdiv(s->nOwned - c0, 256), 256, 0, c.st, c.pos, c.dk, c0,
                                            s->nOwned, c.ca));
    return BENDY_OK;
}

// strips: my send buffers were consumed -> reset them; the received ghosts join the histogram
int Ops::launch_halo_receive(const SubstepCtx &c, cudaStream_t q_clear) {
    LAUNCH(BENDY_K_HALO, k_halo_clear<<<cdiv(s->ghost_cap, 256), 256, 0, q_clear>>>(s->d_send[0].p, s->d_send[1].p,
                                                                                    s->d_send_cnt.p, s->ghost_cap));
    if (q_clear != c.st) {
        CK(cudaEventRecord(s->ev_xchg, q_clear));
        CK(cudaStreamWaitEvent(c.st, s->ev_xchg, 0));
    }
    LAUNCH(BENDY_K_GRID_BUILD,
           k2_count<false><<<cdiv(s->nP - s->nOwned, 256), 256, 0, c.st>>>(c.pos, c.dk, s->nOwned, s->nP, c.ca));
    return BENDY_OK;
}

// particle links (solver.rs:144-146) with the fused histogram; in strip mode also the halo packing.
// *ghosts_done is set when the exchange AND the ghost histogram were already issued (overlap path).
int Ops::launch_particle_links(const SubstepCtx &c, int phase, bool *ghosts_done) {
    const LinkPlan &P = s->plan_p;
    const uint32_t n_parts = P.n_parts(), nb = P.n_priority_parts;
    *ghosts_done = false;
    // Strips, graph mode, opt-in (BENDY_HALO_OVERLAP=1): the partitions of the bodies near the halo
    // bands run first, their discs are packed, and the NCCL exchange proceeds on a side stream while
    // the interior partitions are relaxed.  Measured on 8 x B200 (C5, 2M discs per rank) the split
    // costs more than the exchange it hides (166 vs 153 us per substep), hence off by default.
    const bool overlap = s->halo_overlap && c.halo && phase == PHASE_ALL && c.branch && c.fuse_count && s->nccl_comm &&
                         nb > 0 && nb < n_parts && s->nC == 0;  // side[0] carries the circle chain when there are circles
    if (overlap) {
        cudaStream_t qx = s->side[0];  // free in strip mode (no circles)
        if (int rc = launch_links_local(c, c.st, 0, nb, 1)) return rc;
        if (int rc = launch_count_unlinked(c)) return rc;
        CK(cudaEventRecord(s->ev_main, c.st));
        CK(cudaStreamWaitEvent(qx, s->ev_main, 0));
        if (int rc = halo_exchange_nccl(qx)) return rc;
        if (int rc = launch_links_local(c, c.st, nb, n_parts, 2)) return rc;
        if (int rc = launch_halo_receive(c, qx)) return rc;
        *ghosts_done = true;
        return BENDY_OK;
    }
    if (int rc = launch_links_local(c, c.st, 0, n_parts, c.halo ? 1 : 0)) return rc;
    if (int rc = launch_links_global(c, c.st)) return rc;
    if (int rc = launch_links_cross(c)) return rc;
    return launch_count_unlinked(c);
}

// Links across strip edges, as trailing colours behind all local and global links of every strip.  Before each
// colour the ranks exchange the current positions of the endpoints the neighbour needs (ncclSend/Recv inside the
// captured graph; a few KB), then each rank relaxes its links of that colour and keeps its own endpoints' halves.
// The sequential order this equals: [strip 0's schedule][strip 1's] ... [cross colour 0][cross colour 1] ...
int Ops::launch_links_cross(const SubstepCtx &c) {
    if (s->xl.empty()) return BENDY_OK;
    if (!s->nccl_comm)
        return fail(BENDY_ERR_UNSUPPORTED, "links across strip edges need the NCCL transport (bendy_halo_comm_nccl); the "
                                           "same-process strip group does not carry them");
    auto ck = [&](ncclResult_t r, const char *what) -> int {
        if (r == ncclSuccess) return BENDY_OK;
        s->sticky = BENDY_ERR_CUDA;
        return fail(BENDY_ERR_CUDA, std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
    };
    const uint32_t ns[2] = {(uint32_t)s->xl_send[0].size(), (uint32_t)s->xl_send[1].size()};
    const uint32_t n_send = ns[0] + ns[1];
    const int nb[2] = {s->comm_rank - 1, s->comm_rank + 1};
    const uint32_t n_colours = (uint32_t)s->xl_colour_start.size() - 1;
    for (uint32_t col = 0; col < n_colours; col++) {
        const uint32_t l0 = s->xl_colour_start[col], l1 = s->xl_colour_start[col + 1];
        if (n_send)
            LAUNCH(BENDY_K_LINKS_GLOBAL,
                   k_xl_pack<<<cdiv(n_send, 256), 256, 0, c.st>>>(c.pos, s->d_xl_send_idx.p, n_send, s->d_xl_sendbuf.p));
        if (int rc = ck(g_nccl.GroupStart(), "ncclGroupStart")) return rc;
        for (int side = 0; side < 2; side++) {
            float2 *sb = s->d_xl_sendbuf.p + (side ? ns[0] : 0u), *gb = s->d_xl_ghost.p + (side ? s->xl_recv[0] : 0u);
            if (ns[side])
                if (int rc = ck(g_nccl.Send(sb, 2 * (size_t)ns[side], ncclFloat, nb[side], s->nccl_comm, c.st), "ncclSend")) return rc;
            if (s->xl_recv[side])
                if (int rc = ck(g_nccl.Recv(gb, 2 * (size_t)s->xl_recv[side], ncclFloat, nb[side], s->nccl_comm, c.st), "ncclRecv"))
                    return rc;
        }
        if (int rc = ck(g_nccl.GroupEnd(), "ncclGroupEnd")) return rc;
        if (s->capturing)
            s->count_in_capture++;
        else
            s->launches++, s->k_launches[BENDY_K_HALO]++;
        if (l1 > l0)
            LAUNCH(BENDY_K_LINKS_GLOBAL,
                   k_xl_links<<<cdiv(l1 - l0, 256), 256, 0, c.st>>>(c.pos, s->d_xl_ghost.p, s->d_xl.p, l0, l1));
    }
    return BENDY_OK;
}

// K2 grid build: exclusive scan of the cell histogram (one pass: the scan-tile totals came with the histogram),
// then the counting-sort scatter
int Ops::launch_grid_build(const SubstepCtx &c) {
    if (!c.discs) return BENDY_OK;
    cudaStream_t st = c.st;
    LAUNCH(BENDY_K_GRID_BUILD, launch_k(c.pdl > 0, k2_scan, s->n_scan_tiles, SCAN_THREADS, 0, st, s->d_cell_count.p,
                                        s->d_tile_sum.p, s->d_cell_start.p));
    const uint32_t blocks = cdiv(std::max(s->nP, s->n_scan_tiles), 256);
    if (c.K)
        LAUNCH(BENDY_K_GRID_BUILD, launch_k(c.pdl > 0, k2_scatter<true>, blocks, 256, 0, st, c.pos, s->nP, c.prm, s->n_cells,
                                            s->d_cell_start.p, s->d_tile_sum.p, s->n_scan_tiles, s->d_sorted_pos.p,
                                            s->d_slot_of.p, s->d_sorted_id.p));
    else
        LAUNCH(BENDY_K_GRID_BUILD, launch_k(c.pdl > 0, k2_scatter<false>, blocks, 256, 0, st, c.pos, s->nP, c.prm, s->n_cells,
                                            s->d_cell_start.p, s->d_tile_sum.p, s->n_scan_tiles, s->d_sorted_pos.p,
                                            s->d_slot_of.p, s->d_sorted_id.p));
    return BENDY_OK;
}

// disc grid on: narrowphase + polygon contact + bounds + integrate for the free particles in one
// launch; then the tails (circles: apply + bounds + integrate; polygon points: bounds + integrate),
// each behind the narrowphase on its own branch
int Ops::launch_collide_integrate_discs(const SubstepCtx &c, int phase) {
    cudaStream_t st = c.st;
    if (phase == PHASE_C) return launch_circle_tail(c);  // same-process strips: the corrections were summed by the group
    K2Args a{c.pos,    s->d_prev.p, c.dk,  s->d_slot_of.p, s->d_sorted_id.p,       s->d_sorted_pos.p,    s->d_cell_start.p, s->n_cells,
             s->nP,    s->nOwned,   s->nC, s->d_crad.p,    s->d_circ_tile_count.p, s->d_circ_tile_ids.p, s->d_circ_acc.p,
             s->d_circ_snap.p, s->record_work ? s->d_chunk_work.p : nullptr, s->narrow_reverse ? 1u : 0u};
    const uint32_t blocks = cdiv(s->nOwned, NARROW_THREADS);
    const bool quad = s->prm.quad != 0;  // two candidate rows at most: the instance with the one-step index map
#define NARROW(HK, HP)                                                                                                   \
    do {                                                                                                                 \
        if (quad)                                                                                                        \
            LAUNCH(BENDY_K_NARROWPHASE, launch_k(c.pdl > 1, k2_narrow_contact_integrate<HK, HP, true>, blocks,           \
                                                 NARROW_THREADS, 0, st, a, c.k4, c.prm));                                \
        else                                                                                                             \
            LAUNCH(BENDY_K_NARROWPHASE, launch_k(c.pdl > 1, k2_narrow_contact_integrate<HK, HP, false>, blocks,          \
                                                 NARROW_THREADS, 0, st, a, c.k4, c.prm));                                \
    } while (0)
    if (c.K && c.contact)
        NARROW(true, true);
    else if (c.K)
        NARROW(true, false);
    else if (c.contact)
        NARROW(false, true);
    else
        NARROW(false, false);
#undef NARROW
    if (c.branch && (c.qc != st || c.qg != st)) CK(cudaEventRecord(s->ev_main, st));
    if (s->nC && !(phase == PHASE_B && c.halo)) {
        if (c.qc != st) CK(cudaStreamWaitEvent(c.qc, s->ev_main, 0));
        if (c.halo && s->nccl_comm) {
            // strips: the Circles are replicated; sum the corrections every strip's own discs collected for them
            ncclResult_t r = g_nccl.AllReduce(s->d_circ_acc.p, s->d_circ_acc.p, 2 * (size_t)s->nC, ncclUint64, ncclSum,
                                              s->nccl_comm, c.qc);
            if (r != ncclSuccess) {
                s->sticky = BENDY_ERR_CUDA;
                return fail(BENDY_ERR_CUDA, std::string("NCCL error in ncclAllReduce: ") + g_nccl.GetErrorString(r));
            }
            if (s->capturing)
                s->count_in_capture++;
            else
                s->launches++, s->k_launches[BENDY_K_HALO]++;
        }
        if (int rc = launch_circle_tail(c)) return rc;
    }
    if (s->nG) {
        if (c.qg != st) CK(cudaStreamWaitEvent(c.qg, s->ev_main, 0));
        const uint32_t first = s->nP + s->nC, n = s->nG, blocks_g = cdiv(n, 256);
        if (c.acc && c.K)
            LAUNCH(BENDY_K_INTEGRATE, launch_k(c.pdl > 2, k1_integrate_range<true, true>, blocks_g, 256, 0, c.qg, c.k1, first, n, c.prm));
        else if (c.acc)
            LAUNCH(BENDY_K_INTEGRATE, launch_k(c.pdl > 2, k1_integrate_range<true, false>, blocks_g, 256, 0, c.qg, c.k1, first, n, c.prm));
        else if (c.K)
            LAUNCH(BENDY_K_INTEGRATE, launch_k(c.pdl > 2, k1_integrate_range<false, true>, blocks_g, 256, 0, c.qg, c.k1, first, n, c.prm));
        else
            LAUNCH(BENDY_K_INTEGRATE, launch_k(c.pdl > 2, k1_integrate_range<false, false>, blocks_g, 256, 0, c.qg, c.k1, first, n, c.prm));
    }
    return BENDY_OK;
}

// tail of the substep for the Circles: apply the particles' corrections, bounds, integrate
int Ops::launch_circle_tail(const SubstepCtx &c) {
    {
        const uint32_t blocks_c = cdiv(s->nC, 128);
#define CTAIL(A, KK)                                                                                            \
    LAUNCH(BENDY_K_CIRCLES, launch_k(c.pdl > 2, k_circle_tail<A, KK, true>, blocks_c, 128, 0, c.qc, c.k1,           \
                                     s->d_circ_acc.p, s->d_circ_tile_count.p, s->n_circ_tiles, c.prm))
        if (c.acc && c.K)
            CTAIL(true, true);
        else if (c.acc)
            CTAIL(true, false);
        else if (c.K)
            CTAIL(false, true);
        else
            CTAIL(false, false);
#undef CTAIL
    }
    return BENDY_OK;
}

// disc grid off (reference semantics for free particles): optional particle-polygon contact, then
// solve_boundary_collisions + update_positions (solver.rs:113-114) for ALL points in one launch
int Ops::launch_collide_integrate_plain(const SubstepCtx &c) {
    cudaStream_t st = c.st;
    if (c.contact) {
        if (c.K)
            LAUNCH(BENDY_K_POLY_CONTACT, k4_poly_contact<true><<<cdiv(s->nP, 128), 128, 0, st>>>(c.pos, c.dk, s->nP, c.k4, c.prm));
        else
            LAUNCH(BENDY_K_POLY_CONTACT, k4_poly_contact<false><<<cdiv(s->nP, 128), 128, 0, st>>>(c.pos, c.dk, s->nP, c.k4, c.prm));
    }
    const uint32_t blocks = cdiv(s->Npad / 2, 256);
    if (blocks) {
        if (c.acc && c.K)
            LAUNCH(BENDY_K_INTEGRATE, k1_integrate<true, true><<<blocks, 256, 0, st>>>(c.k1, c.prm));
        else if (c.acc)
            LAUNCH(BENDY_K_INTEGRATE, k1_integrate<true, false><<<blocks, 256, 0, st>>>(c.k1, c.prm));
        else if (c.K)
            LAUNCH(BENDY_K_INTEGRATE, k1_integrate<false, true><<<blocks, 256, 0, st>>>(c.k1, c.prm));
        else
            LAUNCH(BENDY_K_INTEGRATE, k1_integrate<false, false><<<blocks, 256, 0, st>>>(c.k1, c.prm));
    }
    if (c.branch) {  // the next substep's circle / polygon chains must see the integrated state
        CK(cudaEventRecord(s->ev_main, st));
        if (c.qc != st) CK(cudaStreamWaitEvent(c.qc, s->ev_main, 0));
        if (c.qg != st) CK(cudaStreamWaitEvent(c.qg, s->ev_main, 0));
    }
    return BENDY_OK;
}

// phase: PHASE_ALL = the whole substep (NCCL exchange in-stream); PHASE_A = up to the halo packing;
// PHASE_B = from the ghost histogram on (bendy_update_group moves the halo between A and B).
int Ops::launch_substep(int phase) {
    SubstepCtx c = make_ctx();
    bool ghosts_done = false;
    if (phase == PHASE_C) return (c.discs && c.halo && s->nC) ? launch_collide_integrate_discs(c, PHASE_C) : BENDY_OK;
    if (phase != PHASE_B) {
        if (int rc = launch_polygon_chain(c)) return rc;
        if (int rc = launch_circle_chain(c)) return rc;
        if (int rc = launch_particle_links(c, phase, &ghosts_done)) return rc;
        if (phase == PHASE_A) {
            CK(cudaEventRecord(s->ev_phase_a, c.st));
            return BENDY_OK;
        }
        if (c.halo && !ghosts_done)
            if (int rc = halo_exchange_nccl(c.st)) return rc;
    }
    if (c.halo && !ghosts_done)
        if (int rc = launch_halo_receive(c, c.st)) return rc;
    if (int rc = launch_grid_build(c)) return rc;
    // ---- join: the collision phase needs all three worlds
    if (c.branch) {
        if (c.qc != c.st) {
            if (!c.circ_joined) CK(cudaEventRecord(s->ev_join[0], c.qc));
            CK(cudaStreamWaitEvent(c.st, s->ev_join[0], 0));
        }
        if (c.qg != c.st) {
            CK(cudaEventRecord(s->ev_join[1], c.qg));
            CK(cudaStreamWaitEvent(c.st, s->ev_join[1], 0));
        }
    }
    return c.discs ? launch_collide_integrate_discs(c, phase) : launch_collide_integrate_plain(c);
}

// once per update(): the work the narrowphase's warps recorded, summed per half of the index range, goes to the
// host.  In a captured graph this is a branch of its own that runs beside the update's first substeps (the figures
// it carries are those of the update before, give or take the substep already running - they steer a heuristic,
// nothing else).
bool Ops::work_stats_on() const { return s->particle_radius > 0.f && s->nP && s->d_chunk_work.p && s->h_work_halves; }
int Ops::end_update(cudaStream_t q) {
    if (!work_stats_on()) return BENDY_OK;
    // (pinned host memory is addressable from the device: the two sums are written straight into it)
    LAUNCH(BENDY_K_GRID_BUILD, k_work_halves<<<1, 1024, 0, q>>>(s->d_chunk_work.p, cdiv(s->nOwned, NARROW_THREADS), s->h_work_halves));
    return BENDY_OK;
}

// before an update is enqueued: start the narrowphase at the end of the index range that cost more (by a quarter)
// in the most recent update whose figures have arrived; a captured graph has the choice baked in
void Ops::choose_narrow_order() {
    bool want = s->narrow_reverse;
    if (s->narrow_order_mode == 1)
        want = false;
    else if (s->narrow_order_mode == 2)
        want = true;
    else if (s->h_work_halves) {
        const unsigned long long lo = s->h_work_halves[0], hi = s->h_work_halves[1];
        if (hi > lo + lo / 4)
            want = true;
        else if (lo > hi + hi / 4)
            want = false;
    }
    if (want != s->narrow_reverse) {
        s->narrow_reverse = want;
        drop_graph();
    }
}

int Ops::build_graph(uint32_t substeps) {
    drop_graph();
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    s->capturing = true;
    s->count_in_capture = 0;
    int rc = BENDY_OK;
    // fork the circle / polygon branches off the main stream once; join them at the end
    const bool use_c = s->nC > 0, use_g = !s->polys.empty();
    cudaError_t ce = cudaSuccess;
    if (use_c || use_g) ce = cudaEventRecord(s->ev_fork, s->stream);
    if (ce == cudaSuccess && use_c) ce = cudaStreamWaitEvent(s->side[0], s->ev_fork, 0);
    if (ce == cudaSuccess && use_g) ce = cudaStreamWaitEvent(s->side[1], s->ev_fork, 0);
    const bool use_w = work_stats_on();
    if (use_w && ce == cudaSuccess) {
        if (!(use_c || use_g)) ce = cudaEventRecord(s->ev_fork, s->stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s->side[2], s->ev_fork, 0);
        if (ce == cudaSuccess) rc = end_update(s->side[2]);
    }
    for (uint32_t k = 0; k < substeps && rc == BENDY_OK && ce == cudaSuccess; k++) {
        s->record_work = k + 1 == substeps;
        rc = launch_substep();
    }
    s->record_work = false;
    if (ce == cudaSuccess && use_w) ce = cudaEventRecord(s->ev_join[2], s->side[2]);
    if (ce == cudaSuccess && use_w) ce = cudaStreamWaitEvent(s->stream, s->ev_join[2], 0);
    if (ce == cudaSuccess && use_c) ce = cudaEventRecord(s->ev_join[0], s->side[0]);
    if (ce == cudaSuccess && use_c) ce = cudaStreamWaitEvent(s->stream, s->ev_join[0], 0);
    if (ce == cudaSuccess && use_g) ce = cudaEventRecord(s->ev_join[1], s->side[1]);
    if (ce == cudaSuccess && use_g) ce = cudaStreamWaitEvent(s->stream, s->ev_join[1], 0);
    s->capturing = false;
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (ce != cudaSuccess && rc == BENDY_OK) {
        if (graph) cudaGraphDestroy(graph);
        return fail_cuda(ce, "graph branch fork/join", __LINE__);
    }
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaStreamEndCapture", __LINE__);
    e = cudaGraphInstantiate(&s->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGraphInstantiate", __LINE__);
    s->graph_substeps = substeps;
    s->kernels_per_substep = substeps ? s->count_in_capture / substeps : 0;
    return BENDY_OK;
}

// `count` update() calls worth of substeps
int Ops::enqueue_substeps(uint32_t updates) {
    const uint32_t S = s->sub_steps;
    uint32_t done = 0;
    choose_narrow_order();
    // the first substep after circles/polygons were added with a pending acc runs eagerly with the
    // accel variant of K1 (particle.rs:23-24); afterwards acc == 0 and gravity is fused.
    if (s->accel_pending) {
        if (int rc = launch_substep()) return rc;
        s->accel_pending = false;
        drop_graph();
        for (uint32_t k = 1; k < S; k++) {
            s->record_work = k + 1 == S;
            int rc = launch_substep();
            s->record_work = false;
            if (rc) return rc;
        }
        if (int rc = end_update(s->stream)) return rc;
        done = 1;
    }
    if (done == updates) return BENDY_OK;
    // small scenes: one launch per update() with all its substeps (see k_small_scene)
    const bool small = s->small_scene && !s->profiling && !s->halo_on && !s->has_k && s->polys.empty() && s->nC <= 1 &&
                       s->cl.empty() && !(s->particle_radius > 0.f) && s->N > 0 && s->N <= 2900 &&
                       s->plan_p.n_parts() <= 1 && s->plan_p.global_links.empty() &&
                       (s->plan_p.n_parts() == 0 || s->plan_p.part_start[0] == 0);
    if (small) {
        K1Args k1{s->d_pos.p, s->d_prev.p, nullptr, nullptr, s->d_crad.p, s->d_gstatic.p, s->nP, s->nC, s->N};
        const uint32_t C = s->plan_p.n_parts() ? s->plan_p.n_local_colours : 0u;
        for (uint32_t u = done; u < updates; u++)
            LAUNCH(BENDY_K_FUSED, k_small_scene<<<1, 256, (size_t)s->N * 16, s->stream>>>(
                                      k1, s->d_part_cs.p, s->d_local.p, C, s->d_prm.p, S));
        return BENDY_OK;
    }
    if (s->profiling) {
        for (uint32_t u = done; u < updates; u++) {
            for (uint32_t k = 0; k < S; k++) {
                s->record_work = k + 1 == S;
                int rc = launch_substep();
                s->record_work = false;
                if (rc) return rc;
            }
            if (int rc = end_update(s->stream)) return rc;
        }
        return BENDY_OK;
    }
    if (!s->graph_exec || s->graph_substeps != S)
        if (int rc = build_graph(S)) return rc;
    for (uint32_t u = done; u < updates; u++) {
        CK(cudaGraphLaunch(s->graph_exec, s->stream));
        s->launches += (uint64_t)s->kernels_per_substep * S;
    }
    return BENDY_OK;
}

}  // namespace

bendy_solver::~bendy_solver() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (PendingEvent &pe : pending) {
        cudaEventDestroy(pe.e0);
        cudaEventDestroy(pe.e1);
    }
    for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    if (h_work_halves) cudaFreeHost(h_work_halves);
    if (h_prm_ring) {
        for (uint32_t i = 0; i < kPrmRing; i++)
            if (prm_ring_ev[i]) cudaEventDestroy(prm_ring_ev[i]);
        cudaFreeHost(h_prm_ring);
    }
    if (nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(nccl_comm);
    for (int sd = 0; sd < 2; sd++)
        if (peer[sd]) peer[sd]->peer[1 - sd] = nullptr;
    for (cudaEvent_t e : {ev_fork, ev_join[0], ev_join[1], ev_join[2], ev_main, ev_phase_a, ev_xchg})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t q : side)
        if (q) cudaStreamDestroy(q);
    if (stream) cudaStreamDestroy(stream);
}

// =================================================================================================
// C ABI
// =================================================================================================
#undef CK
// inside the extern "C" functions (an `ops` local is in scope)
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) return ops.fail_cuda(_e, #call, __LINE__);                             \
    } while (0)
#define OPS Ops ops{s}
#define NEED(s)                                     \
    if (!(s)) {                                     \
        g_last_error = "null solver handle";        \
        return BENDY_ERR_ARG;                       \
    }

extern "C" {

int bendy_abi_version(void) { return BENDY_ABI_VERSION; }

const char *bendy_last_error(const bendy_solver *s) { return s ? s->err.c_str() : g_last_error.c_str(); }

bendy_solver *bendy_create(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_last_error = std::string("no CUDA device (") + cudaGetErrorString(e) +
                       "): libbendy2d_b200 has no CPU path";
        return nullptr;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) {
        g_last_error = "device index out of range";
        return nullptr;
    }
    std::unique_ptr<bendy_solver> s(new bendy_solver());
    s->device = device;
    if (const char *v = getenv("BENDY_K3_THREADS")) {  // tuning knobs (not part of the ABI)
        int t = atoi(v);
        if (t >= 32 && t <= 1024 && t % 32 == 0) s->k3_threads = (uint32_t)t;
    }
    if (const char *v = getenv("BENDY_HALO_OVERLAP")) s->halo_overlap = atoi(v) != 0;
    if (const char *v = getenv("BENDY_NARROW_ORDER")) {
        s->narrow_order_mode = !strcmp(v, "forward") ? 1 : (!strcmp(v, "reverse") ? 2 : 0);
        if (s->narrow_order_mode) s->narrow_reverse = s->narrow_order_mode == 2;
    }
    if (const char *v = getenv("BENDY_SMALL_SCENE")) s->small_scene = atoi(v) != 0;
    if (const char *v = getenv("BENDY_POLY_FUSED")) s->poly_fused = atoi(v) != 0;
    if (const char *v = getenv("BENDY_PDL")) s->pdl = atoi(v);
    if (const char *v = getenv("BENDY_PDL_NCCL")) s->pdl_nccl = atoi(v) != 0;
    // BENDY_SIDE_PRIORITY=1: the circle / polygon branches get the highest stream priority, so their few
    // CTAs are dispatched ahead of the queued CTAs of the particle chain (kernel nodes keep the priority
    // of the stream they were captured on)
    int prio_side = 0;
    if ((e = cudaSetDevice(device)) == cudaSuccess) {
        const char *v = getenv("BENDY_SIDE_PRIORITY");
        int least = 0, greatest = 0;
        if (v && atoi(v) != 0 && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) prio_side = greatest;
    }
    if (e != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&s->side[0], cudaStreamNonBlocking, prio_side)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&s->side[1], cudaStreamNonBlocking, prio_side)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&s->side[2], cudaStreamNonBlocking, prio_side)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_join[2], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_join[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_join[1], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_main, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_phase_a, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&s->ev_xchg, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreate(&s->t0)) != cudaSuccess || (e = cudaEventCreate(&s->t1)) != cudaSuccess) {
        g_last_error = std::string("CUDA init failed: ") + cudaGetErrorString(e);
        return nullptr;
    }
    return s.release();
}

void bendy_destroy(bendy_solver *s) { delete s; }

bendy_solver *bendy_clone(bendy_solver *s) {
    if (!s) return nullptr;
    OPS;
    if (ops.bind() || ops.pull()) return nullptr;
    bendy_solver *c = bendy_create(s->device);
    if (!c) return nullptr;
    c->p_pos = s->p_pos, c->p_prev = s->p_prev, c->p_k = s->p_k;
    c->c_pos = s->c_pos, c->c_prev = s->c_prev, c->c_acc = s->c_acc, c->c_rad = s->c_rad, c->c_k = s->c_k;
    c->pl_ab = s->pl_ab, c->pl_len = s->pl_len, c->cl = s->cl;
    c->polys = s->polys, c->g_pos = s->g_pos, c->g_prev = s->g_prev, c->g_acc = s->g_acc;
    c->gl_ab = s->gl_ab, c->gl_len = s->gl_len, c->any_acc = s->any_acc;
    c->sub_steps = s->sub_steps, c->particle_radius = s->particle_radius, c->grid_cell = s->grid_cell;
    c->polygon_contact = s->polygon_contact, c->plan_params = s->plan_params;
    memcpy(c->last_args, s->last_args, sizeof c->last_args), c->have_last_args = s->have_last_args;
    // strip configuration travels too; the transport (NCCL communicator / local peers) does not
    c->halo_on = s->halo_on, c->ghost_cap = s->ghost_cap, c->halo_xl = s->halo_xl, c->halo_xr = s->halo_xr;
    c->stray_xl = s->stray_xl, c->stray_xr = s->stray_xr, c->win_x0 = s->win_x0, c->win_x1 = s->win_x1;
    return c;
}

// ---- snapshot: the whole scene as one flat little-endian file of 4-byte words -------------------
// (the reference's only checkpoint is #[derive(Clone)], solver.rs:19; this is the same state on disk,
// so a parity failure can be replayed elsewhere: bendy2d_b200/snapshot.py reads it with numpy)
//   magic "B2DSNAP1" | u32 version, sub_steps | f32 particle_radius, grid_cell | u32 polygon_contact,
//   pack_points, max_points, has_last_update | f32 last dt, gx, gy, bx, by, bw, bh, 0 |
//   u64 nP, nC, nPoly, nG, nPL, nCL, nGL, has_particle_k, has_circle_k |
//   p_pos[2nP] p_prev[2nP] (p_k[nP]) | c_pos c_prev c_acc[2nC each] c_rad[nC] (c_k[nC]) |
//   pl_ab[2nPL] pl_len[nPL] | cl_ab[2nCL] cl_len[nCL] |
//   per polygon: u32 start, nv, link_start, nl, is_static, f32 cx, cy | g_pos g_prev g_acc[2nG each] |
//   gl_ab[2nGL] gl_len[nGL]
extern "C++" {
namespace {
const char kSnapMagic[8] = {'B', '2', 'D', 'S', 'N', 'A', 'P', '1'};
struct SnapHeader {
    char magic[8];
    uint32_t version, sub_steps;
    float particle_radius, grid_cell;
    uint32_t polygon_contact, pack_points, max_points, has_last;
    float last[8];
    uint64_t n[9];  // nP, nC, nPoly, nG, nPL, nCL, nGL, has_pk, has_ck
};
static_assert(sizeof(SnapHeader) == 8 + 8 + 8 + 16 + 32 + 72, "snapshot header layout");
struct SnapPoly {
    uint32_t start, nv, link_start, nl, is_static;
    float cx, cy;
};
template <typename T>
bool put(FILE *f, const std::vector<T> &v) {
    return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
}
template <typename T>
bool get(FILE *f, std::vector<T> &v, uint64_t n) {
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
}  // namespace
}  // extern "C++"

int bendy_save_snapshot(bendy_solver *s, const char *path) {
    NEED(s);
    OPS;
    if (!path) return ops.fail(BENDY_ERR_ARG, "bendy_save_snapshot: null path");
    if (s->sticky != BENDY_OK) return ops.fail(s->sticky, s->err);
    if (!s->host_valid) {
        if (int rc = ops.bind()) return rc;
        if (int rc = ops.pull()) return rc;
    }
    SnapHeader h{};
    memcpy(h.magic, kSnapMagic, 8);
    h.version = 1, h.sub_steps = s->sub_steps;
    h.particle_radius = s->particle_radius, h.grid_cell = s->grid_cell;
    h.polygon_contact = (s->polygon_contact ? 1u : 0u) | (s->plan_params.reference_order ? 2u : 0u);  // bit 1: link schedule
    h.pack_points = s->plan_params.pack_points;
    h.max_points = s->plan_params.max_points, h.has_last = s->have_last_args;
    for (int i = 0; i < 7; i++) h.last[i] = s->last_args[i];
    const uint64_t n[9] = {s->p_pos.size(), s->c_pos.size(), s->polys.size(), s->g_pos.size(), s->pl_len.size(),
                           s->cl.size(),    s->gl_len.size(), s->p_k.empty() ? 0u : 1u, s->c_k.empty() ? 0u : 1u};
    for (int i = 0; i < 9; i++) h.n[i] = n[i];
    std::vector<uint32_t> cl_ab;
    std::vector<float> cl_len;
    for (const GlobalLink &l : s->cl) cl_ab.push_back(l.a), cl_ab.push_back(l.b), cl_len.push_back(l.len);
    std::vector<SnapPoly> polys;
    for (const PolyHost &P : s->polys)
        polys.push_back(SnapPoly{P.start, P.nv, P.link_start, P.nl, P.is_static ? 1u : 0u, P.center.x, P.center.y});
    FILE *f = fopen(path, "wb");
    if (!f) return ops.fail(BENDY_ERR_ARG, std::string("bendy_save_snapshot: cannot open ") + path);
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && put(f, s->p_pos) && put(f, s->p_prev) && put(f, s->p_k) &&
              put(f, s->c_pos) && put(f, s->c_prev) && put(f, s->c_acc) && put(f, s->c_rad) && put(f, s->c_k) &&
              put(f, s->pl_ab) && put(f, s->pl_len) && put(f, cl_ab) && put(f, cl_len) && put(f, polys) &&
              put(f, s->g_pos) && put(f, s->g_prev) && put(f, s->g_acc) && put(f, s->gl_ab) && put(f, s->gl_len);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return ops.fail(BENDY_ERR_ARG, std::string("bendy_save_snapshot: short write to ") + path);
    return BENDY_OK;
}

static bendy_solver *load_snapshot_impl(const char *path, int device);

bendy_solver *bendy_load_snapshot(const char *path, int device) {
    try {  // nothing may unwind through the C boundary (a corrupt file must not take the host process down)
        return load_snapshot_impl(path, device);
    } catch (const std::exception &e) {
        g_last_error = std::string("bendy_load_snapshot: ") + e.what();
        return nullptr;
    }
}

static bendy_solver *load_snapshot_impl(const char *path, int device) {
    if (!path) {
        g_last_error = "bendy_load_snapshot: null path";
        return nullptr;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        g_last_error = std::string("bendy_load_snapshot: cannot open ") + path;
        return nullptr;
    }
    // the file is parsed and validated completely before a device is touched
    std::unique_ptr<bendy_solver> t(new bendy_solver());
    SnapHeader h{};
    std::vector<uint32_t> cl_ab;
    std::vector<float> cl_len;
    std::vector<SnapPoly> polys;
    const char *why = nullptr;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kSnapMagic, 8) != 0)
        why = "not a bendy2d snapshot (bad magic)";
    else if (h.version != 1)
        why = "unsupported snapshot version";
    else if (h.sub_steps == 0 || h.sub_steps > 0xFFFFu)
        why = "corrupt header (sub_steps)";
    else if (!(h.particle_radius >= 0.f) || !(h.grid_cell >= 0.f) || !std::isfinite(h.particle_radius) || !std::isfinite(h.grid_cell))
        why = "corrupt header (particle radius / grid cell: the setters only take finite values >= 0)";
    else {
        for (int i = 0; i < 7; i++)
            if (h.n[i] > 0x7FFFFFF0ull) why = "corrupt header (counts)";
    }
    if (!why) {
        // the payload the counts announce must be exactly what is left of the file: checked BEFORE anything is
        // allocated (a corrupt count must not ask for gigabytes)
        const uint64_t nP = h.n[0], nC = h.n[1], nPoly = h.n[2], nG = h.n[3], nPL = h.n[4], nCL = h.n[5], nGL = h.n[6];
        const uint64_t want = nP * 16 + (h.n[7] ? nP * 4 : 0) + nC * 28 + (h.n[8] ? nC * 4 : 0) + nPL * 12 + nCL * 12 +
                              nPoly * sizeof(SnapPoly) + nG * 24 + nGL * 12;
        const long here = ftell(f);
        long end = -1;
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) end = ftell(f);
        if (here < 0 || end < 0 || fseek(f, here, SEEK_SET) != 0)
            why = "cannot determine the file size";
        else if ((uint64_t)(end - here) < want)
            why = "truncated snapshot";
        else if ((uint64_t)(end - here) > want)
            why = "trailing bytes after the snapshot";
    }
    if (!why) {
        const uint64_t nP = h.n[0], nC = h.n[1], nPoly = h.n[2], nG = h.n[3], nPL = h.n[4], nCL = h.n[5], nGL = h.n[6];
        bool ok = get(f, t->p_pos, nP) && get(f, t->p_prev, nP) && get(f, t->p_k, h.n[7] ? nP : 0) &&
                  get(f, t->c_pos, nC) && get(f, t->c_prev, nC) && get(f, t->c_acc, nC) && get(f, t->c_rad, nC) &&
                  get(f, t->c_k, h.n[8] ? nC : 0) && get(f, t->pl_ab, 2 * nPL) && get(f, t->pl_len, nPL) &&
                  get(f, cl_ab, 2 * nCL) && get(f, cl_len, nCL) && get(f, polys, nPoly) && get(f, t->g_pos, nG) &&
                  get(f, t->g_prev, nG) && get(f, t->g_acc, nG) && get(f, t->gl_ab, 2 * nGL) && get(f, t->gl_len, nGL);
        if (!ok)
            why = "truncated snapshot";
        else if (fgetc(f) != EOF)
            why = "trailing bytes after the snapshot";
        // a snapshot is the state of a scene that can be stepped: a link the reference could not solve (link.rs:19-21)
        // makes the file invalid (the add_* calls accept such a link and the next update refuses it)
        for (uint64_t k = 0; !why && k < nPL; k++)
            if (!(t->pl_ab[2 * k] < t->pl_ab[2 * k + 1]) || !(t->pl_ab[2 * k + 1] < nP)) why = "particle link out of range";
        for (uint64_t k = 0; !why && k < nCL; k++)
            if (!(cl_ab[2 * k] < cl_ab[2 * k + 1]) || !(cl_ab[2 * k + 1] < nC)) why = "circle link out of range";
        uint64_t pts = 0, lks = 0;
        for (uint64_t k = 0; !why && k < nPoly; k++) {
            const SnapPoly &P = polys[k];
            if (P.start != pts || P.link_start != lks || P.nv == 0 || pts + P.nv > nG || lks + P.nl > nGL)
                why = "polygon table inconsistent";
            for (uint32_t l = 0; !why && l < P.nl; l++) {
                const uint32_t a = t->gl_ab[2 * (lks + l)], b = t->gl_ab[2 * (lks + l) + 1];
                if (!(a < b) || !(b < P.nv)) why = "polygon link out of range";
            }
            pts += P.nv, lks += P.nl;
        }
        if (!why && (pts != nG || lks != nGL)) why = "polygon table does not cover the polygon points";
    }
    fclose(f);
    if (why) {
        g_last_error = std::string("bendy_load_snapshot: ") + why + " (" + path + ")";
        return nullptr;
    }
    bendy_solver *s = bendy_create(device);
    if (!s) return nullptr;
    s->p_pos.swap(t->p_pos), s->p_prev.swap(t->p_prev), s->p_k.swap(t->p_k);
    s->c_pos.swap(t->c_pos), s->c_prev.swap(t->c_prev), s->c_acc.swap(t->c_acc), s->c_rad.swap(t->c_rad);
    s->c_k.swap(t->c_k), s->pl_ab.swap(t->pl_ab), s->pl_len.swap(t->pl_len);
    for (size_t k = 0; k < cl_len.size(); k++) s->cl.push_back(GlobalLink{cl_ab[2 * k], cl_ab[2 * k + 1], cl_len[k]});
    for (const SnapPoly &P : polys) {
        PolyHost H;
        H.start = P.start, H.nv = P.nv, H.link_start = P.link_start, H.nl = P.nl, H.is_static = P.is_static != 0;
        H.center = make_float2(P.cx, P.cy);
        s->polys.push_back(H);
    }
    s->g_pos.swap(t->g_pos), s->g_prev.swap(t->g_prev), s->g_acc.swap(t->g_acc);
    s->gl_ab.swap(t->gl_ab), s->gl_len.swap(t->gl_len);
    for (const float2 &a : s->c_acc)
        if (a.x != 0.f || a.y != 0.f) s->any_acc = true;
    for (const PolyHost &P : s->polys)
        for (uint32_t v = 0; v < P.nv && !P.is_static; v++)
            if (s->g_acc[P.start + v].x != 0.f || s->g_acc[P.start + v].y != 0.f) s->any_acc = true;
    s->sub_steps = (uint16_t)h.sub_steps;
    s->particle_radius = h.particle_radius, s->grid_cell = h.grid_cell;
    s->polygon_contact = (h.polygon_contact & 1u) != 0;
    s->plan_params.reference_order = (h.polygon_contact & 2u) != 0;
    s->plan_params.pack_points = h.pack_points, s->plan_params.max_points = h.max_points;
    s->have_last_args = h.has_last != 0;
    for (int i = 0; i < 7; i++) s->last_args[i] = h.last[i];
    return s;
}

int bendy_get_last_update_args(const bendy_solver *s, float *dt_g_bounds7, int *valid) {
    NEED(s);
    if (!dt_g_bounds7 || !valid) {
        g_last_error = "bendy_get_last_update_args: null output";
        return BENDY_ERR_ARG;
    }
    for (int i = 0; i < 7; i++) dt_g_bounds7[i] = s->last_args[i];
    *valid = s->have_last_args ? 1 : 0;
    return BENDY_OK;
}

// ---- scene construction --------------------------------------------------------------------
static int edit_begin(bendy_solver *s) {
    OPS;
    if (s->sticky != BENDY_OK) return ops.fail(s->sticky, s->err);
    if (!s->host_valid) {
        if (int rc = ops.pull()) return rc;
    }
    return BENDY_OK;
}
static void edit_end(bendy_solver *s) {
    s->topo_dirty = true;
    s->device_valid = false;
}

int bendy_add_particles(bendy_solver *s, const float *pos_xy, size_t n) {
    NEED(s);
    OPS;
    if (n && !pos_xy) return ops.fail(BENDY_ERR_ARG, "bendy_add_particles: null positions");
    if (s->p_pos.size() + n > 0x7FFFFFF0u) return ops.fail(BENDY_ERR_ARG, "too many particles");
    if (int rc = edit_begin(s)) return rc;
    for (size_t i = 0; i < n; i++) {
        float2 p = make_float2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        s->p_pos.push_back(p);
        s->p_prev.push_back(p);  // particle.rs:15
    }
    if (!s->p_k.empty()) s->p_k.resize(s->p_pos.size(), 1.0f);
    edit_end(s);
    return BENDY_OK;
}

int bendy_add_circles(bendy_solver *s, const float *pos_xy, const float *prev_xy, const float *acc_xy,
                      const float *radius, size_t n) {
    NEED(s);
    OPS;
    if (n && (!pos_xy || !radius)) return ops.fail(BENDY_ERR_ARG, "bendy_add_circles: null positions or radii");
    if (int rc = edit_begin(s)) return rc;
    for (size_t i = 0; i < n; i++) {
        float2 p = make_float2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        s->c_pos.push_back(p);
        s->c_prev.push_back(prev_xy ? make_float2(prev_xy[2 * i], prev_xy[2 * i + 1]) : p);
        float2 a = acc_xy ? make_float2(acc_xy[2 * i], acc_xy[2 * i + 1]) : make_float2(0.f, 0.f);
        if (a.x != 0.f || a.y != 0.f) s->any_acc = true;
        s->c_acc.push_back(a);
        s->c_rad.push_back(radius[i]);
    }
    if (!s->c_k.empty()) s->c_k.resize(s->c_pos.size(), 1.0f);
    edit_end(s);
    return BENDY_OK;
}

int bendy_add_polygon(bendy_solver *s, const float *pos_xy, const float *prev_xy, const float *acc_xy, size_t nv,
                      const uint32_t *link_ab, const float *link_len, size_t nl, int is_static, float cx, float cy) {
    NEED(s);
    OPS;
    if (nv == 0 || !pos_xy) return ops.fail(BENDY_ERR_ARG, "bendy_add_polygon: no points");
    if (nl && (!link_ab || !link_len)) return ops.fail(BENDY_ERR_ARG, "bendy_add_polygon: null links");
    for (size_t k = 0; k < nl; k++)  // link.rs:19-21 on the polygon's own particle vector
        if (!(link_ab[2 * k] < link_ab[2 * k + 1]) || !(link_ab[2 * k + 1] < nv))
            return ops.fail(BENDY_ERR_LINK, "bendy_add_polygon: link needs a < b < point count (the reference panics)");
    if (int rc = edit_begin(s)) return rc;
    PolyHost P;
    P.start = (uint32_t)s->g_pos.size();
    P.nv = (uint32_t)nv;
    P.link_start = (uint32_t)s->gl_len.size();
    P.nl = (uint32_t)nl;
    P.is_static = is_static != 0;
    P.center = make_float2(cx, cy);
    for (size_t i = 0; i < nv; i++) {
        float2 p = make_float2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        s->g_pos.push_back(p);
        s->g_prev.push_back(prev_xy ? make_float2(prev_xy[2 * i], prev_xy[2 * i + 1]) : p);
        float2 a = acc_xy ? make_float2(acc_xy[2 * i], acc_xy[2 * i + 1]) : make_float2(0.f, 0.f);
        if ((a.x != 0.f || a.y != 0.f) && !P.is_static) s->any_acc = true;
        s->g_acc.push_back(a);
    }
    for (size_t k = 0; k < nl; k++) {
        s->gl_ab.push_back(link_ab[2 * k]);
        s->gl_ab.push_back(link_ab[2 * k + 1]);
        s->gl_len.push_back(link_len[k]);
    }
    s->polys.push_back(P);
    edit_end(s);
    return BENDY_OK;
}

int bendy_add_particle_links(bendy_solver *s, const uint32_t *ab, const float *len, size_t n) {
    NEED(s);
    OPS;
    if (n && (!ab || !len)) return ops.fail(BENDY_ERR_ARG, "bendy_add_particle_links: null arrays");
    // like the reference, any pair is accepted here (solver.rs:62-64 pushes the link; a program may add its links
    // before the particles they name): the indices are checked where link.rs:19-21 would panic, inside update
    if (int rc = edit_begin(s)) return rc;
    s->pl_ab.insert(s->pl_ab.end(), ab, ab + 2 * n);
    s->pl_len.insert(s->pl_len.end(), len, len + n);
    edit_end(s);
    return BENDY_OK;
}

int bendy_add_circle_links(bendy_solver *s, const uint32_t *ab, const float *len, size_t n) {
    NEED(s);
    OPS;
    if (n && (!ab || !len)) return ops.fail(BENDY_ERR_ARG, "bendy_add_circle_links: null arrays");
    if (int rc = edit_begin(s)) return rc;  // indices are checked inside update, like the reference (link.rs:37-39)
    for (size_t k = 0; k < n; k++) s->cl.push_back(GlobalLink{ab[2 * k], ab[2 * k + 1], len[k]});
    edit_end(s);
    return BENDY_OK;
}

// ---- hot path ---------------------------------------------------------------------------------
int bendy_update_n(bendy_solver *s, uint32_t n, float dt, float gx, float gy, float bx, float by, float bw,
                   float bh) {
    NEED(s);
    OPS;
    if (n == 0) return BENDY_OK;
    if (int rc = ops.ensure_ready()) return rc;
    // solver.rs:107-108
    float mult = 1.0f / (float)s->sub_steps;
    float delta = dt * mult;
    if (int rc = ops.configure(delta, gx, gy, bx, by, bw, bh)) return rc;
    if (int rc = ops.enqueue_substeps(n)) return rc;
    s->host_valid = false;
    const float args[7] = {dt, gx, gy, bx, by, bw, bh};
    memcpy(s->last_args, args, sizeof args);
    s->have_last_args = true;
    return BENDY_OK;
}

int bendy_update(bendy_solver *s, float dt, float gx, float gy, float bx, float by, float bw, float bh) {
    return bendy_update_n(s, 1, dt, gx, gy, bx, by, bw, bh);
}

int bendy_synchronize(bendy_solver *s) {
    NEED(s);
    OPS;
    if (s->sticky != BENDY_OK) return ops.fail(s->sticky, s->err);
    if (!s->stream) return BENDY_OK;
    if (int rc = ops.bind()) return rc;
    if (int rc = ops.flush_events()) return rc;
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return ops.fail_cuda(e, "cudaStreamSynchronize", __LINE__);
    return ops.check_flags();
}

// ---- getters ------------------------------------------------------------------------------------
size_t bendy_particle_len(const bendy_solver *s) { return s ? s->p_pos.size() : 0; }
size_t bendy_circle_len(const bendy_solver *s) { return s ? s->c_pos.size() : 0; }
size_t bendy_polygon_len(const bendy_solver *s) { return s ? s->polys.size() : 0; }
size_t bendy_particle_link_len(const bendy_solver *s) { return s ? s->pl_len.size() : 0; }
size_t bendy_circle_link_len(const bendy_solver *s) { return s ? s->cl.size() : 0; }
size_t bendy_polygon_point_len(const bendy_solver *s, size_t k) { return s && k < s->polys.size() ? s->polys[k].nv : 0; }
size_t bendy_polygon_link_len(const bendy_solver *s, size_t k) { return s && k < s->polys.size() ? s->polys[k].nl : 0; }

static void copy_out(const std::vector<float2> &v, size_t first, size_t n, float *out) {
    if (out && n) std::memcpy(out, v.data() + first, n * sizeof(float2));
}

int bendy_read_particles(bendy_solver *s, size_t first, size_t n, float *pos_xy, float *prev_xy) {
    NEED(s);
    OPS;
    if (first + n > s->p_pos.size()) return ops.fail(BENDY_ERR_ARG, "bendy_read_particles: range out of bounds");
    if (n == 0) return BENDY_OK;
    if (s->host_valid) {
        copy_out(s->p_pos, first, n, pos_xy);
        copy_out(s->p_prev, first, n, prev_xy);
        return BENDY_OK;
    }
    // device is authoritative: gather USER order on the device, copy straight into the caller's buffers
    if (int rc = ops.ensure_ready()) return rc;
    if (int rc = ops.flush_events()) return rc;
    // one gather launch for both arrays into the two halves of the staging buffer, two copies, ONE synchronise
    // (check_flags reads its word with the same synchronise)
    if (pos_xy || prev_xy) {
        float2 *st0 = s->d_stage.p, *st1 = s->d_stage.p + s->Npad;
        k_gather2<<<cdiv((uint32_t)n, 256), 256, 0, s->stream>>>(pos_xy ? s->d_pos.p : nullptr, prev_xy ? s->d_prev.p : nullptr,
                                                                  s->d_rank.p + first, (uint32_t)n, st0, st1);
        s->launches++;
        CK(cudaGetLastError());
        if (pos_xy) CK(cudaMemcpyAsync(pos_xy, st0, n * sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
        if (prev_xy) CK(cudaMemcpyAsync(prev_xy, st1, n * sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
    }
    return ops.check_flags();
}

int bendy_write_particles(bendy_solver *s, size_t first, size_t n, const float *pos_xy, const float *prev_xy) {
    NEED(s);
    OPS;
    if (first + n > s->p_pos.size()) return ops.fail(BENDY_ERR_ARG, "bendy_write_particles: range out of bounds");
    if (n == 0) return BENDY_OK;
    if (s->topo_dirty || !s->device_valid) {  // nothing on the device yet: edit the host scene
        if (int rc = edit_begin(s)) return rc;
        if (pos_xy) std::memcpy(s->p_pos.data() + first, pos_xy, n * sizeof(float2));
        if (prev_xy) std::memcpy(s->p_prev.data() + first, prev_xy, n * sizeof(float2));
        s->device_valid = false;
        return BENDY_OK;
    }
    if (int rc = ops.ensure_ready()) return rc;
    if (pos_xy || prev_xy) {
        float2 *st0 = s->d_stage.p, *st1 = s->d_stage.p + s->Npad;
        if (pos_xy) CK(cudaMemcpyAsync(st0, pos_xy, n * sizeof(float2), cudaMemcpyHostToDevice, s->stream));
        if (prev_xy) CK(cudaMemcpyAsync(st1, prev_xy, n * sizeof(float2), cudaMemcpyHostToDevice, s->stream));
        k_scatter2<<<cdiv((uint32_t)n, 256), 256, 0, s->stream>>>(pos_xy ? st0 : nullptr, prev_xy ? st1 : nullptr,
                                                                   s->d_rank.p + first, (uint32_t)n, s->d_pos.p, s->d_prev.p);
        s->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s->stream));  // the caller's buffers may be reused after return
    }
    s->host_valid = false;
    return BENDY_OK;
}

int bendy_read_circles(bendy_solver *s, size_t first, size_t n, float *pos_xy, float *prev_xy, float *radius) {
    NEED(s);
    OPS;
    if (first + n > s->c_pos.size()) return ops.fail(BENDY_ERR_ARG, "bendy_read_circles: range out of bounds");
    if (n == 0) return BENDY_OK;
    if (radius) std::memcpy(radius, s->c_rad.data() + first, n * sizeof(float));
    if (s->host_valid) {
        copy_out(s->c_pos, first, n, pos_xy);
        copy_out(s->c_prev, first, n, prev_xy);
        return BENDY_OK;
    }
    if (int rc = ops.ensure_ready()) return rc;
    if (int rc = ops.flush_events()) return rc;
    if (pos_xy)
        CK(cudaMemcpyAsync(pos_xy, s->d_pos.p + s->nP + first, n * sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
    if (prev_xy)
        CK(cudaMemcpyAsync(prev_xy, s->d_prev.p + s->nP + first, n * sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return ops.check_flags();
}

int bendy_read_polygon(bendy_solver *s, size_t k, float *pos_xy, float *prev_xy, float *center_xy, int *is_static) {
    NEED(s);
    OPS;
    if (k >= s->polys.size()) return ops.fail(BENDY_ERR_ARG, "bendy_read_polygon: index out of bounds");
    if (!s->host_valid) {
        if (int rc = ops.ensure_ready()) return rc;
        if (int rc = ops.flush_events()) return rc;
        if (int rc = ops.pull()) return rc;
    }
    const PolyHost &P = s->polys[k];
    copy_out(s->g_pos, P.start, P.nv, pos_xy);
    copy_out(s->g_prev, P.start, P.nv, prev_xy);
    if (center_xy) center_xy[0] = P.center.x, center_xy[1] = P.center.y;
    if (is_static) *is_static = P.is_static ? 1 : 0;
    return BENDY_OK;
}

int bendy_read_particle_links(const bendy_solver *s, size_t first, size_t n, uint32_t *ab, float *len) {
    if (!s || first + n > s->pl_len.size()) return BENDY_ERR_ARG;
    if (ab && n) std::memcpy(ab, s->pl_ab.data() + 2 * first, 2 * n * sizeof(uint32_t));
    if (len && n) std::memcpy(len, s->pl_len.data() + first, n * sizeof(float));
    return BENDY_OK;
}
int bendy_read_circle_links(const bendy_solver *s, size_t first, size_t n, uint32_t *ab, float *len) {
    if (!s || first + n > s->cl.size()) return BENDY_ERR_ARG;
    for (size_t k = 0; k < n; k++) {
        if (ab) ab[2 * k] = s->cl[first + k].a, ab[2 * k + 1] = s->cl[first + k].b;
        if (len) len[k] = s->cl[first + k].len;
    }
    return BENDY_OK;
}
int bendy_read_polygon_links(const bendy_solver *s, size_t k, uint32_t *ab, float *len) {
    if (!s || k >= s->polys.size()) return BENDY_ERR_ARG;
    const PolyHost &P = s->polys[k];
    if (ab && P.nl) std::memcpy(ab, s->gl_ab.data() + 2 * (size_t)P.link_start, 2 * (size_t)P.nl * sizeof(uint32_t));
    if (len && P.nl) std::memcpy(len, s->gl_len.data() + P.link_start, (size_t)P.nl * sizeof(float));
    return BENDY_OK;
}

// ---- additive API -----------------------------------------------------------------------------
int bendy_set_sub_steps(bendy_solver *s, uint16_t n) {
    NEED(s);
    OPS;
    if (n == 0) return ops.fail(BENDY_ERR_ARG, "sub_steps must be >= 1");
    s->sub_steps = n;
    return BENDY_OK;
}
int bendy_set_particle_radius(bendy_solver *s, float r) {
    NEED(s);
    OPS;
    if (!(r >= 0.f)) return ops.fail(BENDY_ERR_ARG, "particle radius must be >= 0");
    if (r != s->particle_radius) {
        s->particle_radius = r;
        s->prm_valid = false;
        ops.drop_graph();
    }
    return BENDY_OK;
}
int bendy_set_grid_cell(bendy_solver *s, float h) {
    NEED(s);
    OPS;
    if (!(h >= 0.f)) return ops.fail(BENDY_ERR_ARG, "grid cell must be >= 0");
    if (h != s->grid_cell) {
        s->grid_cell = h;
        s->prm_valid = false;
        ops.drop_graph();
    }
    return BENDY_OK;
}
int bendy_set_polygon_contact(bendy_solver *s, int on) {
    NEED(s);
    OPS;
    if ((on != 0) != s->polygon_contact) {
        s->polygon_contact = on != 0;
        s->prm_valid = false;
        ops.drop_graph();
    }
    return BENDY_OK;
}
int bendy_set_particle_inv_mass(bendy_solver *s, size_t first, size_t n, const float *k) {
    NEED(s);
    OPS;
    if (first + n > s->p_pos.size() || (n && !k)) return ops.fail(BENDY_ERR_ARG, "bendy_set_particle_inv_mass: bad range");
    for (size_t i = 0; i < n; i++)
        if (!(k[i] >= 0.f)) return ops.fail(BENDY_ERR_ARG, "inverse-mass scale must be >= 0");
    if (int rc = edit_begin(s)) return rc;
    if (s->p_k.empty()) s->p_k.assign(s->p_pos.size(), 1.0f);
    std::copy(k, k + n, s->p_k.begin() + first);
    edit_end(s);
    return BENDY_OK;
}
int bendy_set_circle_inv_mass(bendy_solver *s, size_t first, size_t n, const float *k) {
    NEED(s);
    OPS;
    if (first + n > s->c_pos.size() || (n && !k)) return ops.fail(BENDY_ERR_ARG, "bendy_set_circle_inv_mass: bad range");
    for (size_t i = 0; i < n; i++)
        if (!(k[i] >= 0.f)) return ops.fail(BENDY_ERR_ARG, "inverse-mass scale must be >= 0");
    if (int rc = edit_begin(s)) return rc;
    if (s->c_k.empty()) s->c_k.assign(s->c_pos.size(), 1.0f);
    std::copy(k, k + n, s->c_k.begin() + first);
    edit_end(s);
    return BENDY_OK;
}
int bendy_set_plan_params(bendy_solver *s, uint32_t pack_points, uint32_t max_points) {
    NEED(s);
    OPS;
    // a partition lives in one CTA's shared memory (12 B per point with inverse masses, 227 KB per CTA)
    // and its link records address points with 16 bits
    if (max_points == 1 || max_points > 16384 || pack_points > 16384)
        return ops.fail(BENDY_ERR_ARG, "bendy_set_plan_params: pack_points <= 16384 and 2 <= max_points <= 16384");
    if (pack_points) s->plan_params.pack_points = pack_points;
    if (max_points) s->plan_params.max_points = max_points;
    s->topo_dirty = true;
    return BENDY_OK;
}

int bendy_set_link_schedule(bendy_solver *s, int mode) {
    NEED(s);
    OPS;
    if (mode != BENDY_LINKS_COLOURED && mode != BENDY_LINKS_REFERENCE_ORDER)
        return ops.fail(BENDY_ERR_ARG, "bendy_set_link_schedule: mode is BENDY_LINKS_COLOURED or BENDY_LINKS_REFERENCE_ORDER");
    const bool ref = mode == BENDY_LINKS_REFERENCE_ORDER;
    if (ref != s->plan_params.reference_order) {
        s->plan_params.reference_order = ref;
        s->topo_dirty = true;
    }
    return BENDY_OK;
}

// ---- schedule export ----------------------------------------------------------------------------
static void fill_info(const LinkPlan &P, const LinkPlan *G, bendy_schedule_info *out) {
    std::memset(out, 0, sizeof *out);
    out->n_partitions = P.n_parts();
    out->n_local_colours = P.n_local_colours;
    out->n_global_colours = P.n_global_colours();
    out->n_local_links = (uint32_t)P.local_links.size();
    out->n_global_links = (uint32_t)P.global_links.size();
    if (G) out->n_poly_partitions = G->n_parts();
    out->n_priority_partitions = P.n_priority_parts;
}

int bendy_get_schedule_info(bendy_solver *s, bendy_schedule_info *out) {
    NEED(s);
    OPS;
    if (!out) return ops.fail(BENDY_ERR_ARG, "null out");
    if (int rc = ops.ensure_ready()) return rc;
    fill_info(s->plan_p, nullptr, out);
    out->kernels_per_substep = s->kernels_per_substep;
    return BENDY_OK;
}
int bendy_get_link_order(bendy_solver *s, uint32_t *perm, size_t n) {
    NEED(s);
    OPS;
    if (n != s->pl_len.size() || (n && !perm)) return ops.fail(BENDY_ERR_ARG, "bendy_get_link_order: n must equal the link count");
    if (int rc = ops.ensure_ready()) return rc;
    std::vector<uint32_t> p = s->plan_p.perm();
    std::copy(p.begin(), p.end(), perm);
    return BENDY_OK;
}
int bendy_get_point_rank(bendy_solver *s, uint32_t *rank, size_t n) {
    NEED(s);
    OPS;
    if (n != s->p_pos.size() || (n && !rank)) return ops.fail(BENDY_ERR_ARG, "bendy_get_point_rank: n must equal the particle count");
    if (int rc = ops.ensure_ready()) return rc;
    std::copy(s->plan_p.rank.begin(), s->plan_p.rank.end(), rank);
    return BENDY_OK;
}
int bendy_get_grid(bendy_solver *s, float bx, float by, float bw, float bh, float *ox, float *oy, float *inv_h,
                   int *nx, int *ny) {
    NEED(s);
    OPS;
    StepParams p{};
    uint32_t nc = 0;
    ops.grid_for(bx, by, bw, bh, &p, &nc);
    if (ox) *ox = p.gox;
    if (oy) *oy = p.goy;
    if (inv_h) *inv_h = p.inv_h;
    if (nx) *nx = p.nx;
    if (ny) *ny = p.ny;
    return BENDY_OK;
}

// ---- measurement ----------------------------------------------------------------------------------
int bendy_set_profiling(bendy_solver *s, int profile) {
    NEED(s);
    OPS;
    if (int rc = ops.flush_events()) return rc;
    s->profiling = profile != 0;
    return BENDY_OK;
}
int bendy_get_kernel_times(bendy_solver *s, double *ms, uint64_t *launches, int n_classes, int reset) {
    NEED(s);
    OPS;
    if (int rc = ops.bind()) return rc;
    if (int rc = ops.flush_events()) return rc;
    for (int k = 0; k < n_classes && k < BENDY_K_CLASSES; k++) {
        if (ms) ms[k] = s->k_ms[k];
        if (launches) launches[k] = s->k_launches[k];
    }
    if (reset) {
        std::memset(s->k_ms, 0, sizeof s->k_ms);
        std::memset(s->k_launches, 0, sizeof s->k_launches);
    }
    return BENDY_OK;
}
uint64_t bendy_launch_count(const bendy_solver *s) { return s ? s->launches : 0; }

int bendy_get_stats(bendy_solver *s, uint64_t *out, int n) {
    NEED(s);
    OPS;
    if (!out || n < 1) return ops.fail(BENDY_ERR_ARG, "bendy_get_stats: bad arguments");
    for (int k = 0; k < n; k++) out[k] = 0;
    if (n > 4) out[4] = s->n_scan_tiles;
    if (n > 5) out[5] = s->n_cells;
    if (n > 6) out[6] = s->narrow_reverse ? 1 : 0;  // the narrowphase currently starts at the top of the index range
    if (!s->d_flags.p) return BENDY_OK;
    if (int rc = ops.bind()) return rc;
    int h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(h, s->d_flags.p, sizeof h, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n && k < 4; k++) out[k] = (uint64_t)h[1 + k];  // total, path bound, list overflow, scale
    return BENDY_OK;
}

int bendy_timer_start(bendy_solver *s) {
    NEED(s);
    OPS;
    if (int rc = ops.bind()) return rc;
    CK(cudaEventRecord(s->t0, s->stream));
    return BENDY_OK;
}
int bendy_timer_stop(bendy_solver *s, float *ms) {
    NEED(s);
    OPS;
    if (int rc = ops.bind()) return rc;
    CK(cudaEventRecord(s->t1, s->stream));
    CK(cudaEventSynchronize(s->t1));
    float v = 0.f;
    CK(cudaEventElapsedTime(&v, s->t0, s->t1));
    if (ms) *ms = v;
    return BENDY_OK;
}
void *bendy_get_stream(const bendy_solver *s) { return s ? (void *)s->stream : nullptr; }
int bendy_get_device(const bendy_solver *s) { return s ? s->device : -1; }

int bendy_get_device_buffers(bendy_solver *s, void **pos, void **prev, size_t *n_points) {
    NEED(s);
    OPS;
    if (int rc = ops.ensure_ready()) return rc;
    if (pos) *pos = s->d_pos.p;
    if (prev) *prev = s->d_prev.p;
    if (n_points) *n_points = s->N;
    return BENDY_OK;
}

// ---- spatial strips: halo exchange ----------------------------------------------------------------
int bendy_set_grid_window(bendy_solver *s, float x0, float x1) {
    NEED(s);
    OPS;
    if (!(x0 < x1)) return ops.fail(BENDY_ERR_ARG, "bendy_set_grid_window: need x0 < x1");
    s->win_x0 = x0, s->win_x1 = x1;
    s->prm_valid = false;
    ops.drop_graph();
    return BENDY_OK;
}

int bendy_halo_configure(bendy_solver *s, uint32_t ghost_cap, float x_left, float x_right, float stray_left,
                         float stray_right) {
    NEED(s);
    OPS;
    if (ghost_cap > 0x3FFFFFFFu) return ops.fail(BENDY_ERR_ARG, "ghost capacity too large");
    if (int rc = edit_begin(s)) return rc;
    s->halo_on = ghost_cap > 0;
    s->ghost_cap = ghost_cap;
    s->halo_xl = x_left, s->halo_xr = x_right;
    s->stray_xl = stray_left, s->stray_xr = stray_right;
    s->prm_valid = false;
    edit_end(s);
    return BENDY_OK;
}

int bendy_nccl_unique_id(void *out128) {
    if (!out128) return BENDY_ERR_ARG;
    if (!g_nccl.load()) {
        g_last_error = g_nccl.err;
        return BENDY_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) {
        g_last_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
        return BENDY_ERR_CUDA;
    }
    std::memcpy(out128, &id, sizeof id);
    return BENDY_OK;
}

int bendy_halo_comm_nccl(bendy_solver *s, const void *unique_id128, int rank, int world) {
    NEED(s);
    OPS;
    if (!unique_id128 || rank < 0 || rank >= world) return ops.fail(BENDY_ERR_ARG, "bendy_halo_comm_nccl: bad arguments");
    if (!g_nccl.load()) return ops.fail(BENDY_ERR_UNSUPPORTED, g_nccl.err);
    if (int rc = ops.bind()) return rc;
    if (s->nccl_comm) {
        g_nccl.CommDestroy(s->nccl_comm);
        s->nccl_comm = nullptr;
    }
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&s->nccl_comm, world, id, rank);
    if (r != ncclSuccess) return ops.fail(BENDY_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    s->comm_rank = rank, s->comm_world = world;
    ops.drop_graph();
    return BENDY_OK;
}

int bendy_strip_set_cross_links(bendy_solver *s, size_t n, const uint32_t *mine, const uint32_t *slot, const uint8_t *i_am_a,
                                const float *len, uint32_t n_colours, const uint32_t *colour_start, size_t n_send_left,
                                const uint32_t *send_left, size_t n_send_right, const uint32_t *send_right,
                                size_t n_recv_left, size_t n_recv_right) {
    NEED(s);
    OPS;
    if (n && (!mine || !slot || !i_am_a || !len || !colour_start || n_colours == 0))
        return ops.fail(BENDY_ERR_ARG, "bendy_strip_set_cross_links: null table");
    if ((n_send_left && !send_left) || (n_send_right && !send_right))
        return ops.fail(BENDY_ERR_ARG, "bendy_strip_set_cross_links: null send list");
    if (n_recv_left + n_recv_right > 0x7FFFFFFFu || n > 0x7FFFFFFFu) return ops.fail(BENDY_ERR_ARG, "too many cross links");
    if (n && (colour_start[0] != 0 || colour_start[n_colours] != n))
        return ops.fail(BENDY_ERR_ARG, "bendy_strip_set_cross_links: colour_start must run from 0 to n");
    const size_t np = s->p_pos.size();
    for (size_t k = 0; k < n; k++)
        if (mine[k] >= np || slot[k] >= n_recv_left + n_recv_right)
            return ops.fail(BENDY_ERR_LINK, "bendy_strip_set_cross_links: endpoint or ghost slot out of range");
    for (uint32_t c = 0; c < n_colours && n; c++)
        if (colour_start[c] > colour_start[c + 1]) return ops.fail(BENDY_ERR_ARG, "colour_start must be non-decreasing");
    for (size_t k = 0; k < n_send_left; k++)
        if (send_left[k] >= np) return ops.fail(BENDY_ERR_LINK, "bendy_strip_set_cross_links: send index out of range");
    for (size_t k = 0; k < n_send_right; k++)
        if (send_right[k] >= np) return ops.fail(BENDY_ERR_LINK, "bendy_strip_set_cross_links: send index out of range");
    if (int rc = edit_begin(s)) return rc;
    s->xl.resize(n);
    for (size_t k = 0; k < n; k++) s->xl[k] = bendy_solver::XlHost{mine[k], slot[k], len[k], i_am_a[k] ? 1u : 0u};
    s->xl_colour_start.assign(n ? colour_start : nullptr, n ? colour_start + n_colours + 1 : nullptr);
    s->xl_send[0].assign(send_left, send_left + n_send_left);
    s->xl_send[1].assign(send_right, send_right + n_send_right);
    s->xl_recv[0] = (uint32_t)n_recv_left, s->xl_recv[1] = (uint32_t)n_recv_right;
    edit_end(s);
    return BENDY_OK;
}

int bendy_halo_connect_local(bendy_solver *left, bendy_solver *right) {
    if (!left || !right || left == right) return BENDY_ERR_ARG;
    if (left->ghost_cap != right->ghost_cap || !left->halo_on || !right->halo_on) {
        g_last_error = left->err = "bendy_halo_connect_local: both strips need the same ghost capacity";
        return BENDY_ERR_ARG;
    }
    left->peer[1] = right;
    right->peer[0] = left;
    return BENDY_OK;
}

int bendy_halo_stats(bendy_solver *s, uint32_t *sent_left, uint32_t *sent_right, uint32_t *overflow,
                     uint32_t *strayed) {
    NEED(s);
    OPS;
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (s->d_send_cnt.p) {
        if (int rc = ops.bind()) return rc;
        CK(cudaStreamSynchronize(s->stream));
        CK(cudaMemcpy(h, s->d_send_cnt.p, sizeof h, cudaMemcpyDeviceToHost));
    }
    if (sent_left) *sent_left = h[4];
    if (sent_right) *sent_right = h[5];
    if (overflow) *overflow = h[2];
    if (strayed) *strayed = h[3];
    return BENDY_OK;
}

// Lock-step stepping of several strips that live in ONE process (same-process transport): for every
// substep all strips run phase A, the halos are copied device-to-device, all strips run phase B.
// This is the 1-GPU emulation of the multi-GPU path; the arithmetic is that of bendy_update.
int bendy_update_group(bendy_solver **group, int n, uint32_t n_updates, float dt, float gx, float gy, float bx,
                       float by, float bw, float bh) {
    if (!group || n <= 0) return BENDY_ERR_ARG;
    for (int k = 0; k < n; k++) {
        bendy_solver *s = group[k];
        NEED(s);
        OPS;
        if (int rc = ops.ensure_ready()) return rc;
        if (int rc = ops.flush_events()) return rc;
        float delta = dt * (1.0f / (float)s->sub_steps);
        if (int rc = ops.configure(delta, gx, gy, bx, by, bw, bh)) return rc;
        if (s->sub_steps != group[0]->sub_steps) return ops.fail(BENDY_ERR_ARG, "group members need equal sub_steps");
        const float args[7] = {dt, gx, gy, bx, by, bw, bh};
        memcpy(s->last_args, args, sizeof args);
        s->have_last_args = true;
    }
    const uint32_t total = n_updates * group[0]->sub_steps;
    for (uint32_t step = 0; step < total; step++) {
        for (int k = 0; k < n; k++) {
            bendy_solver *s = group[k];
            OPS;
            if (int rc = ops.bind()) return rc;
            if (int rc = ops.launch_substep(PHASE_A)) return rc;
            if (!(s->halo_on && s->particle_radius > 0.f && s->nP)) CK(cudaEventRecord(s->ev_phase_a, s->stream));
        }
        for (int k = 0; k < n; k++) {
            bendy_solver *s = group[k];
            OPS;
            if (int rc = ops.bind()) return rc;
            const size_t bytes = (size_t)s->ghost_cap * sizeof(float2);
            for (int side = 0; side < 2; side++) {
                bendy_solver *p = s->peer[side];
                if (!p || !bytes) continue;
                CK(cudaStreamWaitEvent(s->stream, p->ev_phase_a, 0));
                // my left ghosts = the left peer's right-going send buffer, and vice versa
                CK(cudaMemcpyAsync(s->d_pos.p + s->nOwned + (size_t)side * s->ghost_cap, p->d_send[1 - side].p, bytes,
                                   cudaMemcpyDeviceToDevice, s->stream));
                if (s->has_k != p->has_k) return ops.fail(BENDY_ERR_ARG, "group members must all have (or all lack) inverse masses");
                if (s->has_k)  // the ghosts' inverse-mass scales travel along
                    CK(cudaMemcpyAsync(s->d_k.p + s->nOwned + (size_t)side * s->ghost_cap, p->d_send_k[1 - side].p,
                                       (size_t)s->ghost_cap * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
            }
            CK(cudaEventRecord(s->ev_xchg, s->stream));
        }
        for (int k = 0; k < n; k++) {
            bendy_solver *s = group[k];
            OPS;
            if (int rc = ops.bind()) return rc;
            for (int side = 0; side < 2; side++)
                if (s->peer[side]) CK(cudaStreamWaitEvent(s->stream, s->peer[side]->ev_xchg, 0));
            if (int rc = ops.launch_substep(PHASE_B)) return rc;
        }
        // replicated Circles: every strip's own discs collected fixed-point corrections for them; sum them over
        // the group (integers: any order gives the same bits), hand every strip the total, then run the tails
        if (group[0]->nC && group[0]->halo_on && group[0]->particle_radius > 0.f && group[0]->nP) {
            bendy_solver *s0 = group[0];
            const uint32_t n_acc = 2 * s0->nC;
            for (int k = 0; k < n; k++) {
                if (group[k]->nC != s0->nC) {
                    g_last_error = s0->err = "group members need the same (replicated) circles";
                    return BENDY_ERR_ARG;
                }
                bendy_solver *s = group[k];
                OPS;
                CK(cudaStreamSynchronize(s->stream));
            }
            {
                bendy_solver *s = s0;
                OPS;
                if (int rc = ops.bind()) return rc;
                for (int k = 1; k < n; k++) {
                    k_acc_add<<<cdiv(n_acc, 128), 128, 0, s0->stream>>>(s0->d_circ_acc.p, group[k]->d_circ_acc.p, n_acc);
                    CK(cudaGetLastError());
                }
                for (int k = 1; k < n; k++)
                    CK(cudaMemcpyAsync(group[k]->d_circ_acc.p, s0->d_circ_acc.p, n_acc * sizeof(unsigned long long),
                                       cudaMemcpyDeviceToDevice, s0->stream));
                CK(cudaStreamSynchronize(s0->stream));
            }
            for (int k = 0; k < n; k++) {
                bendy_solver *s = group[k];
                OPS;
                if (int rc = ops.bind()) return rc;
                if (int rc = ops.launch_substep(PHASE_C)) return rc;
            }
        }
        for (int k = 0; k < n; k++) {
            bendy_solver *s = group[k];
            s->accel_pending = false;
            s->host_valid = false;
        }
    }
    return BENDY_OK;
}

// ---- host-only planning ----------------------------------------------------------------------------
int bendy_debug_normalize(int device, const float *dx, const float *dy, const float *norm, size_t n, float *nx, float *ny) {
    if (n == 0) return BENDY_OK;
    if (!dx || !dy || !norm || !nx || !ny || n > 0x7FFFFFF0u) {
        g_last_error = "bendy_debug_normalize: null array or n too large";
        return BENDY_ERR_ARG;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        g_last_error = "bendy_debug_normalize: no such device";
        return BENDY_ERR_NO_DEVICE;
    }
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 5 * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d, dx, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, dy, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + 2 * n, norm, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_debug_normalize<<<((uint32_t)n + 255u) / 256u, 256>>>(d, d + n, d + 2 * n, (uint32_t)n, d + 3 * n, d + 4 * n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(nx, d + 3 * n, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ny, d + 4 * n, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (e != cudaSuccess) {
        g_last_error = std::string("bendy_debug_normalize: ") + cudaGetErrorString(e);
        return BENDY_ERR_CUDA;
    }
    return BENDY_OK;
}

int bendy_plan_links(size_t n_points, const uint32_t *ab, size_t n_links, uint32_t pack_points, uint32_t max_points,
                     uint32_t *rank, uint32_t *perm, uint32_t *link_colour, uint32_t *link_partition,
                     bendy_schedule_info *info) {
    return bendy_plan_links_scheduled(n_points, ab, n_links, pack_points, max_points, BENDY_LINKS_COLOURED, rank, perm,
                                      link_colour, link_partition, info);
}

int bendy_plan_links_scheduled(size_t n_points, const uint32_t *ab, size_t n_links, uint32_t pack_points,
                               uint32_t max_points, int link_schedule, uint32_t *rank, uint32_t *perm,
                               uint32_t *link_colour, uint32_t *link_partition, bendy_schedule_info *info) {
    if (n_links && !ab) return BENDY_ERR_ARG;
    if (link_schedule != BENDY_LINKS_COLOURED && link_schedule != BENDY_LINKS_REFERENCE_ORDER) return BENDY_ERR_ARG;
    for (size_t k = 0; k < n_links; k++)
        if (!(ab[2 * k] < ab[2 * k + 1]) || !(ab[2 * k + 1] < n_points)) {
            g_last_error = "bendy_plan_links: link needs a < b < n_points";
            return BENDY_ERR_LINK;
        }
    if (max_points == 1 || max_points > 16384 || pack_points > 16384) {
        g_last_error = "bendy_plan_links: pack_points <= 16384 and 2 <= max_points <= 16384";
        return BENDY_ERR_ARG;
    }
    PlanParams pp;
    if (pack_points) pp.pack_points = pack_points;
    if (max_points) pp.max_points = max_points;
    pp.reference_order = link_schedule == BENDY_LINKS_REFERENCE_ORDER;
    std::vector<float> len(n_links, 1.0f);
    LinkPlan P;
    std::string err;
    if (!plan_links(n_points, ab, len.data(), n_links, pp, false, &P, &err)) {
        g_last_error = err;
        return BENDY_ERR_UNSUPPORTED;
    }
    if (rank) std::copy(P.rank.begin(), P.rank.end(), rank);
    if (perm) {
        std::vector<uint32_t> p = P.perm();
        std::copy(p.begin(), p.end(), perm);
    }
    if (link_colour || link_partition) {
        const size_t stride = (size_t)P.n_local_colours + 1;
        for (uint32_t p = 0; p < P.n_parts(); p++)
            for (uint32_t c = 0; c < P.n_local_colours; c++)
                for (uint32_t l = P.part_colour_start[p * stride + c]; l < P.part_colour_start[p * stride + c + 1]; l++) {
                    if (link_colour) link_colour[P.local_user[l]] = c;
                    if (link_partition) link_partition[P.local_user[l]] = p;
                }
        for (uint32_t c = 0; c < P.n_global_colours(); c++)
            for (uint32_t l = P.gcolour_start[c]; l < P.gcolour_start[c + 1]; l++) {
                if (link_colour) link_colour[P.global_user[l]] = P.n_local_colours + c;
                if (link_partition) link_partition[P.global_user[l]] = 0xFFFFFFFFu;
            }
    }
    if (info) fill_info(P, nullptr, info);
    return BENDY_OK;
}

}  // extern "C"

#ifdef BENDY_TIMESTAMPS
// measurement builds only (profiles/substep_timeline.py): reset / read the per-SM kernel time stamps
extern "C" int bendy_debug_ts_reset() {
    static unsigned long long h[4][3][256];
    for (auto &k : h)
        for (int w = 0; w < 3; w++)
            for (int i = 0; i < 256; i++) k[w][i] = w == 2 ? 0ull : ~0ull;
    return cudaMemcpyToSymbol(bendy::g_ts, h, sizeof h) == cudaSuccess ? 0 : -3;
}
extern "C" int bendy_debug_ts_read(unsigned long long *out) {
    return cudaMemcpyFromSymbol(out, bendy::g_ts, sizeof(unsigned long long) * 4 * 3 * 256) == cudaSuccess ? 0 : -3;
}
#endif
