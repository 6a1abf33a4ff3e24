/*
 * bendy2d_b200.h — C ABI of libbendy2d_b200.so: the B200 (sm_100a) solver substep behind
 * bendy2d's public `Solver` API.
 *
 * The reference (Murrent/bendy2d, Rust) has no FFI; its boundary for this path is the public
 * API of `Solver` in src/solver.rs.  Each entry point below names the reference item it replaces
 * (file:line relative to the reference repo).  A Rust `bendy2d-sys` crate binds exactly these
 * symbols (see INTEGRATION.md and rust/).
 *
 * Conventions
 *  - plain C types only; every pointer is caller-owned HOST memory, copied before return;
 *  - xy arrays are interleaved f32 pairs (x0,y0,x1,y1,...) — the layout of nalgebra Vector2<f32>;
 *  - int return: BENDY_OK or a negative error; bendy_last_error() gives the text;
 *  - a handle is bound to ONE CUDA device and ONE stream; thread-compatible, not thread-safe
 *    (the reference takes &mut self for every mutation: solver.rs:52-67,106);
 *  - bendy_update() only enqueues work; reads synchronise.  There is NO CPU fallback: without a
 *    CUDA device bendy_create() fails.
 */
#ifndef BENDY2D_B200_H
#define BENDY2D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BENDY_ABI_VERSION 1

typedef struct bendy_solver bendy_solver;

enum {
    BENDY_OK = 0,
    BENDY_ERR_ARG = -1,         /* null pointer / out-of-range argument */
    BENDY_ERR_LINK = -2,        /* link with a >= b or b >= len: the reference panics (link.rs:19-21) */
    BENDY_ERR_CUDA = -3,        /* CUDA runtime error (sticky; text in bendy_last_error) */
    BENDY_ERR_UNSUPPORTED = -4, /* scene outside what the device path implements */
    BENDY_ERR_NO_DEVICE = -5    /* no CUDA device: there is no CPU path */
};

/* ---------------------------------------------------------------- lifetime */
/* Solver::new (solver.rs:34-50).  device < 0 = the calling thread's current device.
 * Returns NULL on failure (bendy_last_error(NULL) has the reason). */
bendy_solver *bendy_create(int device);
void bendy_destroy(bendy_solver *s);
/* #[derive(Clone)] on Solver (solver.rs:19): deep copy of host tables and device buffers. */
bendy_solver *bendy_clone(bendy_solver *s);
const char *bendy_last_error(const bendy_solver *s);
int bendy_abi_version(void);
/* The same state as bendy_clone, on disk (the reference has no file format, Clone at
 * solver.rs:19 is its only checkpoint): every particle / circle / polygon point with pos, prev and pending
 * acc, all links in insertion order, polygon tables, inverse-mass scales, sub_steps, particle radius, grid
 * cell, polygon-contact switch, plan parameters and the arguments of the last update.  A flat file of
 * little-endian 4-byte words (layout in bendy2d_b200/csrc/solver.cu and bendy2d_b200/snapshot.py); strip /
 * halo configuration is not part of it.  A loaded solver continues bit-identically to the saved one.
 * bendy_load_snapshot validates the whole file (sizes, link index rules of link.rs:19-21) before it
 * touches a device and returns NULL on failure (bendy_last_error(NULL) has the reason). */
int bendy_save_snapshot(bendy_solver *s, const char *path);
bendy_solver *bendy_load_snapshot(const char *path, int device);
/* dt, gravity x/y, bounds x/y/w/h of the most recent update (pub fields gravity / bounds of the reference's
 * Solver, solver.rs:21-22, which this ABI takes per call); *valid = 0 when no update has run yet */
int bendy_get_last_update_args(const bendy_solver *s, float *dt_g_bounds7, int *valid);

/* ---------------------------------------------------------------- scene construction */
/* Solver::add_particle (solver.rs:52-54) x n: Particle::new => prev=pos, acc=0 (particle.rs:12-18) */
int bendy_add_particles(bendy_solver *s, const float *pos_xy, size_t n);
/* Solver::add_circle (solver.rs:55-57) x n. prev_xy NULL => prev=pos; acc_xy NULL => 0 */
int bendy_add_circles(bendy_solver *s, const float *pos_xy, const float *prev_xy, const float *acc_xy,
                      const float *radius, size_t n);
/* Solver::add_polygon (solver.rs:58-60): points, polygon-local links (a<b), is_static, cached centre
 * (polygon.rs:8-14). */
int bendy_add_polygon(bendy_solver *s, const float *pos_xy, const float *prev_xy, const float *acc_xy, size_t nv,
                      const uint32_t *link_ab, const float *link_len, size_t nl, int is_static, float cx, float cy);
/* Solver::add_particle_link (solver.rs:62-64) x n; ab = (a0,b0,a1,b1,...).  Like the reference, any pair is accepted
 * here (links may be added before the particles they name); a link that does not satisfy a < b < len when the scene
 * is stepped makes bendy_update return BENDY_ERR_LINK where the reference panics (link.rs:19-21). */
int bendy_add_particle_links(bendy_solver *s, const uint32_t *ab, const float *len, size_t n);
/* Solver::add_circle_link (solver.rs:65-67) x n; indices are checked inside update as well (link.rs:37-39) */
int bendy_add_circle_links(bendy_solver *s, const uint32_t *ab, const float *len, size_t n);

/* ---------------------------------------------------------------- the hot path */
/* Solver::update(dt) (solver.rs:106-116).  The pub fields gravity and bounds (solver.rs:21-22) are
 * passed by value on every call because the user may mutate them between calls.  bounds_active
 * is never read by the reference (solver.rs:155-165) so it is not passed.  Asynchronous. */
int bendy_update(bendy_solver *s, float dt, float gx, float gy, float bx, float by, float bw, float bh);
/* n back-to-back update(dt) calls in one enqueue (same arithmetic as calling bendy_update n times) */
int bendy_update_n(bendy_solver *s, uint32_t n, float dt, float gx, float gy, float bx, float by, float bw,
                   float bh);
int bendy_synchronize(bendy_solver *s);

/* ---------------------------------------------------------------- getters */
size_t bendy_particle_len(const bendy_solver *s);      /* get_particle_len solver.rs:69 */
size_t bendy_circle_len(const bendy_solver *s);        /* get_circles_len  solver.rs:72 */
size_t bendy_polygon_len(const bendy_solver *s);       /* get_polygons_len solver.rs:75 */
size_t bendy_particle_link_len(const bendy_solver *s); /* get_particle_links().len() solver.rs:79 */
size_t bendy_circle_link_len(const bendy_solver *s);   /* get_circle_links().len()   solver.rs:82 */
size_t bendy_polygon_point_len(const bendy_solver *s, size_t polygon);
size_t bendy_polygon_link_len(const bendy_solver *s, size_t polygon);
/* get_particles / get_particle (solver.rs:86,96): copies [first, first+n) in USER order; any pointer may be NULL */
int bendy_read_particles(bendy_solver *s, size_t first, size_t n, float *pos_xy, float *prev_xy);
/* get_circles / get_circle (solver.rs:89,99) */
int bendy_read_circles(bendy_solver *s, size_t first, size_t n, float *pos_xy, float *prev_xy, float *radius);
/* get_polygons / get_polygon (solver.rs:92,102): points, cached centre, static flag */
int bendy_read_polygon(bendy_solver *s, size_t polygon, float *pos_xy, float *prev_xy, float *center_xy,
                       int *is_static);
int bendy_read_particle_links(const bendy_solver *s, size_t first, size_t n, uint32_t *ab, float *len);
int bendy_read_circle_links(const bendy_solver *s, size_t first, size_t n, uint32_t *ab, float *len);
int bendy_read_polygon_links(const bendy_solver *s, size_t polygon, uint32_t *ab, float *len);

/* ---------------------------------------------------------------- additive API (not in the reference) */
/* sub_steps is a private field fixed at 1 in the reference (solver.rs:29,47); default 1 */
int bendy_set_sub_steps(bendy_solver *s, uint16_t n);
/* overwrite pos/prev of free particles [first, first+n) from host buffers (snapshot restore / streaming) */
int bendy_write_particles(bendy_solver *s, size_t first, size_t n, const float *pos_xy, const float *prev_xy);
/* ext: r > 0 makes every free particle a disc of radius r that collides with particles and Circles
 * through the uniform-grid narrowphase; 0 (default) = reference semantics (no particle contact) */
int bendy_set_particle_radius(bendy_solver *s, float r);
/* ext: grid cell edge; 0 = auto (2*r rounded up so the cell count stays bounded) */
int bendy_set_grid_cell(bendy_solver *s, float h);
/* ext: free particles vs static convex polygons (closest-edge contact); default off */
int bendy_set_polygon_contact(bendy_solver *s, int on);
/* ext: inverse-mass scale per particle / circle; default 1 (= reference arithmetic); 0 pins the point */
int bendy_set_particle_inv_mass(bendy_solver *s, size_t first, size_t n, const float *k);
int bendy_set_circle_inv_mass(bendy_solver *s, size_t first, size_t n, const float *k);
/* planner knobs: points per shared-memory partition (pack target, hard cap); 0 keeps the default (512, 4096).
 * Limits: pack_points <= 16384, 2 <= max_points <= 16384 (one partition = one CTA's shared memory). */
int bendy_set_plan_params(bendy_solver *s, uint32_t pack_points, uint32_t max_points);
/* Order in which the particle links are relaxed (solver.rs:143-146 walks them sequentially in insertion order).
 * BENDY_LINKS_COLOURED (default): greedy graph colouring, few colours; the results equal the reference fed the
 *   links in the order bendy_get_link_order exports (the contract: "fed links in the same colour order").
 * BENDY_LINKS_REFERENCE_ORDER: dependency-level colours - every link runs after all EARLIER links that share a
 *   point with it, so the results equal the reference's own insertion-order walk bit for bit, whatever order
 *   the user added the links in; costs more colours (one CTA barrier or one launch each). */
#define BENDY_LINKS_COLOURED 0
#define BENDY_LINKS_REFERENCE_ORDER 1
int bendy_set_link_schedule(bendy_solver *s, int mode);

/* ---------------------------------------------------------------- schedule export (parity replay) */
typedef struct bendy_schedule_info {
    uint32_t n_partitions;      /* shared-memory link partitions (free particles) */
    uint32_t n_local_colours;   /* max colours inside one partition */
    uint32_t n_global_colours;  /* colours of cross-partition links and of links that found no local colour (one launch each) */
    uint32_t n_local_links;
    uint32_t n_global_links;
    uint32_t n_poly_partitions; /* partitions holding polygon-internal links */
    uint32_t kernels_per_substep;
    uint32_t n_priority_partitions; /* strips: leading partitions relaxed before the halo exchange starts */
} bendy_schedule_info;
/* Builds the schedule if links changed, then reports it. */
int bendy_get_schedule_info(bendy_solver *s, bendy_schedule_info *out);
/* perm[k] = user index of the k-th particle link in the sequential order that is arithmetically
 * identical to the device's parallel schedule (feed it to the reference/oracle). n = link count */
int bendy_get_link_order(bendy_solver *s, uint32_t *perm, size_t n);
/* rank[i] = internal index of user particle i (fixes the in-cell accumulation order of the grid) */
int bendy_get_point_rank(bendy_solver *s, uint32_t *rank, size_t n);
/* the broadphase grid bendy_update would use for these bounds */
int bendy_get_grid(bendy_solver *s, float bx, float by, float bw, float bh, float *ox, float *oy, float *inv_h,
                   int *nx, int *ny);

/* ---------------------------------------------------------------- measurement */
enum {
    BENDY_K_INTEGRATE = 0, /* K1 bounds + gravity + Verlet integrate */
    BENDY_K_LINKS_LOCAL,   /* K3 shared-memory partition relaxation */
    BENDY_K_LINKS_GLOBAL,  /* K3 cross-partition colours */
    BENDY_K_LINKS_CIRCLE,  /* CircleLink relaxation */
    BENDY_K_GRID_BUILD,    /* K2 hash/count/scan/scatter/canonicalise */
    BENDY_K_NARROWPHASE,   /* K2 3x3 narrowphase */
    BENDY_K_CIRCLES,       /* circle-circle lexicographic pass + circle binning/apply */
    BENDY_K_POLY_PREP,     /* polygon centre / AABB / binning */
    BENDY_K_POLY_CONTACT,  /* K4 particle-polygon closest edge (+ polygon-polygon) */
    BENDY_K_FUSED,         /* small scenes: all substeps of an update in one single-CTA launch */
    BENDY_K_HALO,          /* strips: halo exchange (NCCL send/recv or peer copy) + send-buffer reset */
    BENDY_K_CIRCLE_PASS,   /* exact circle-circle pass (solver.rs:168-177) */
    BENDY_K_CLASSES
};
/* profile != 0: launch kernels one by one with cudaEvent pairs (slower; for per-kernel timing).
 * profile == 0 (default): CUDA-graph replay of the substep. */
int bendy_set_profiling(bendy_solver *s, int profile);
/* accumulated device ms and launch counts per kernel class since the last reset */
int bendy_get_kernel_times(bendy_solver *s, double *ms, uint64_t *launches, int n_classes, int reset);
/* diagnostic counters: out[0] = substeps in which the parallel exact circle pass had to fall back to
 * the plain sequential pass (its path-length bound did not hold), out[1..3] = by cause (path bound,
 * near-list overflow, coordinate scale), out[4] = scan tiles (2048 cells each) of the broadphase histogram,
 * out[5] = grid cells */
int bendy_get_stats(bendy_solver *s, uint64_t *out, int n);
/* kernels launched (graph nodes included) since creation: the gpu_launches figure of bench.py */
uint64_t bendy_launch_count(const bendy_solver *s);
/* cudaEvent timer on the solver's stream */
int bendy_timer_start(bendy_solver *s);
int bendy_timer_stop(bendy_solver *s, float *ms); /* synchronises */
/* the cudaStream_t and device the handle runs on (for callers that interleave their own work) */
void *bendy_get_stream(const bendy_solver *s);
int bendy_get_device(const bendy_solver *s);

/* ---------------------------------------------------------------- multi-GPU strips (halo exchange)
 * One process per GPU; each solver owns the bodies of one vertical strip of the world.  Every
 * substep, after the link pass, owned discs lying left of x_left / right of x_right (the strip
 * edges moved inwards by the halo band) are packed on the device and sent to the left / right
 * neighbour, where they occupy read-only ghost slots of the broadphase grid.  No reference
 * counterpart: the reference is single-threaded (solver.rs:109-188). */
/* ghost_cap = ghost disc slots per side (0 turns strips off); owned discs with x < x_left go to the
 * left neighbour, x > x_right to the right one (-inf/+inf when there is no neighbour on that side);
 * an owned disc outside [stray_left, stray_right] raises the `strayed` flag of bendy_halo_stats: it
 * is so deep in a neighbour's strip that ownership has to be rebalanced by the host layer */
int bendy_halo_configure(bendy_solver *s, uint32_t ghost_cap, float x_left, float x_right, float stray_left,
                         float stray_right);
/* restricts the broadphase grid to x in [x0, x1] (clipped to the bounds): a strip needs cells only
 * for its own discs and ghosts.  Results do not depend on it (the grid only prunes pairs). */
/* Links whose two ends are owned by neighbouring strips (a body cut by a strip edge; the reference relaxes any link
 * between any two particles: link.rs:18-27, loop solver.rs:144-146).  Each rank registers its share: link k joins
 * MY particle mine[k] (user index) and the neighbour's particle that arrives in slot[k] of my link-ghost buffer
 * (slots [0, n_recv_left) come from the left neighbour, the rest from the right one); i_am_a[k] tells which end I
 * am (link.rs:22 computes a - b).  The links are sorted by colour (colour_start[n_colours + 1]; a colour holds
 * vertex-disjoint links, and both ranks of a link use the same colour).  send_left / send_right: my particles whose
 * positions the neighbour needs, in the order of ITS slots.  The cross colours run after all the strip's own
 * links, every colour behind an exchange of the endpoint positions (NCCL transport only), which equals the
 * sequential order [strip 0's schedule][strip 1's] ... [cross colour 0][cross colour 1] ... */
int bendy_strip_set_cross_links(bendy_solver *s, size_t n, const uint32_t *mine, const uint32_t *slot, const uint8_t *i_am_a,
                                const float *len, uint32_t n_colours, const uint32_t *colour_start, size_t n_send_left,
                                const uint32_t *send_left, size_t n_send_right, const uint32_t *send_right,
                                size_t n_recv_left, size_t n_recv_right);
int bendy_set_grid_window(bendy_solver *s, float x0, float x1);
/* rank 0 creates the NCCL id (128 bytes); the host layer (torch.distributed) broadcasts it */
int bendy_nccl_unique_id(void *out128);
/* joins the strip communicator: neighbours are rank-1 and rank+1; send/recv run on the solver's stream
 * inside the captured substep graph.  Circles and polygons are replicated on every strip; the fixed-point
 * corrections a strip's own discs collect for the Circles are all-reduced (ncclUint64 sum) over the
 * communicator before they are applied, so every copy stays identical to the unsharded run. */
int bendy_halo_comm_nccl(bendy_solver *s, const void *unique_id128, int rank, int world);
/* same-process transport between two neighbouring strips (1-GPU emulation, tests) */
int bendy_halo_connect_local(bendy_solver *left, bendy_solver *right);
/* lock-step update of several same-process strips (phase A | halo copy | phase B | sum of the Circles'
 * corrections over the group | the Circles' tail, per substep) */
int bendy_update_group(bendy_solver **group, int n, uint32_t n_updates, float dt, float gx, float gy, float bx,
                       float by, float bw, float bh);
/* discs packed for each neighbour in the last substep; overflow != 0 means ghost_cap was too small;
 * strayed != 0 means ownership is stale (see bendy_halo_configure) */
int bendy_halo_stats(bendy_solver *s, uint32_t *sent_left, uint32_t *sent_right, uint32_t *overflow,
                     uint32_t *strayed);

/* Device pointers of the internal SoA (pos, prev as float2 arrays in INTERNAL order) so that a host
 * layer (torch.distributed / NCCL or CUDA IPC) can move halo particles without a host round trip. */
int bendy_get_device_buffers(bendy_solver *s, void **pos, void **prev, size_t *n_points);

/* ---------------------------------------------------------------- host-only planning (no GPU needed) */
/* The link planner used by the solver, callable on host arrays (unit tests, offline tools).
 * Outputs (any may be NULL): rank[n_points], perm[n_links]; info filled like bendy_get_schedule_info. */
int bendy_plan_links(size_t n_points, const uint32_t *ab, size_t n_links, uint32_t pack_points,
                     uint32_t max_points, uint32_t *rank, uint32_t *perm, uint32_t *link_colour,
                     uint32_t *link_partition, bendy_schedule_info *info);
/* the same with the link schedule of bendy_set_link_schedule (BENDY_LINKS_COLOURED / BENDY_LINKS_REFERENCE_ORDER) */
int bendy_plan_links_scheduled(size_t n_points, const uint32_t *ab, size_t n_links, uint32_t pack_points,
                               uint32_t max_points, int link_schedule, uint32_t *rank, uint32_t *perm,
                               uint32_t *link_colour, uint32_t *link_partition, bendy_schedule_info *info);

/* Test hook, no solver needed: the normalize() of link.rs:24 / circle.rs:37 as the kernels compute it (both
 * components over one refined reciprocal, kernels.cuh normalize2) for n host triples (dx, dy, norm) on `device`
 * (-1 = current).  The parity tests compare it with IEEE binary32 division bit for bit over the operand ranges,
 * the guard's edges, zeros of both signs, denormals, infinities and NaN. */
int bendy_debug_normalize(int device, const float *dx, const float *dy, const float *norm, size_t n, float *nx,
                          float *ny);

#ifdef __cplusplus
}
#endif
#endif /* BENDY2D_B200_H */
