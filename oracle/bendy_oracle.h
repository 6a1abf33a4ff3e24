/*
 * bendy_oracle.h — CPU ORACLE for the bendy2d solver substep.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a single-threaded C++ restatement of the reference's Rust solver
 * (/root/reference/src/{particle,link,circle,polygon,common,solver}.rs) used as the parity
 * checker and as the CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libbendy2d_b200.so) never
 * links, loads or calls anything in oracle/.
 *
 * PARITY PIN STATUS: the reference ships no tests, fixtures or golden vectors and cannot be
 * compiled here (no Rust toolchain, nalgebra 0.32.x not vendored).  The oracle is pinned by
 * (i) the cited source lines, (ii) hand-derived known-answer tests (tests/test_oracle_kat.py,
 * SURVEY.md §4), (iii) a second restatement in numpy float32 written separately from the Rust
 * lines (tests/np_restatement.py), compared bit for bit on random inputs and whole updates
 * (tests/test_oracle_vs_numpy.py).  None of these is the reference itself: PARITY UNPINNED.
 * Every function that has NO reference counterpart (the "ext" functions:
 * disc Jacobi contact, particle-vs-polygon closest-edge contact, inv_mass) is
 * "parity unpinned by the reference" and says so at its definition.
 *
 * All arithmetic is IEEE binary32, evaluated in the operator order of the cited lines;
 * build with -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#ifndef BENDY_ORACLE_H
#define BENDY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bo_world bo_world;

#define BO_OK 0
#define BO_ERR_PANIC (-1)   /* the reference would panic (link.rs:19-21 index checks) */
#define BO_ERR_ARG (-2)

/* ---- world lifetime: Solver::new defaults, solver.rs:34-50 ---- */
bo_world *bo_create(void);
void bo_destroy(bo_world *w);
bo_world *bo_clone(const bo_world *w); /* #[derive(Clone)] solver.rs:19 */

/* pub fields gravity / bounds, solver.rs:21-22 */
void bo_set_gravity(bo_world *w, float gx, float gy);
void bo_set_bounds(bo_world *w, float bx, float by, float sx, float sy);

/* ---- scene construction, solver.rs:52-67 ---- */
void bo_add_particle(bo_world *w, float x, float y); /* Particle::new: prev=pos, acc=0 */
void bo_add_particles(bo_world *w, const float *pos_xy, size_t n);                       /* bulk form */
void bo_add_particle_links(bo_world *w, const uint32_t *ab, const float *len, size_t n); /* bulk form */
void bo_add_circle(bo_world *w, float px, float py, float qx, float qy, float ax, float ay, float radius);
int bo_add_polygon(bo_world *w, const float *pos_xy, const float *prev_xy, const float *acc_xy, size_t nv,
                   const uint32_t *link_ab, const float *link_len, size_t nl, int is_static, float cx,
                   float cy);
int bo_add_polygon_new(bo_world *w, const float *pts_xy, size_t nv, int is_static);        /* polygon.rs:84-123 */
int bo_add_polygon_circle(bo_world *w, float radius, float px, float py, size_t n, int is_static); /* polygon.rs:17-82 */
void bo_add_particle_link(bo_world *w, size_t a, size_t b, float len); /* no validation, like the reference */
void bo_add_circle_link(bo_world *w, size_t a, size_t b, float len);

/* ---- the hot path: Solver::update, solver.rs:106-116 ---- */
int bo_update(bo_world *w, float dt);

/* ---- read-back ---- */
size_t bo_particle_len(const bo_world *w);
size_t bo_circle_len(const bo_world *w);
size_t bo_polygon_len(const bo_world *w);
size_t bo_particle_link_len(const bo_world *w);
void bo_read_particles(const bo_world *w, float *pos_xy, float *prev_xy);
void bo_write_particles(bo_world *w, const float *pos_xy, const float *prev_xy);
void bo_read_circles(const bo_world *w, float *pos_xy, float *prev_xy, float *radius);
size_t bo_polygon_point_len(const bo_world *w, size_t poly);
size_t bo_polygon_link_len(const bo_world *w, size_t poly);
void bo_read_polygon(const bo_world *w, size_t poly, float *pos_xy, float *prev_xy, float *center_xy);
void bo_read_polygon_links(const bo_world *w, size_t poly, uint32_t *ab, float *len);

/* ---- schedule-order replay (SURVEY §8.3 iii) ----
 * perm[k] = index (insertion order) of the k-th particle link to solve.  NULL/0 restores
 * insertion order.  Within a colour all links are vertex-disjoint, so replaying the GPU's
 * colour-bucket order sequentially is arithmetically identical to the parallel execution. */
int bo_set_link_order(bo_world *w, const uint32_t *perm, size_t n);
/* additive: sub_steps setter (the reference field is private and fixed at 1, solver.rs:29,47) */
void bo_set_sub_steps(bo_world *w, uint16_t n);

/* ---- ext: features with NO reference semantics (parity unpinned by the reference) ---- */
/* radius > 0 turns on disc contact for free particles (particle-particle, particle-Circle),
 * per-pair rule = Circle::solve_circle (circle.rs:32-45), Jacobi update discipline. */
void bo_ext_set_particle_radius(bo_world *w, float r);
/* broadphase grid used ONLY to fix the accumulation order (origin, 1/h, dims) */
void bo_ext_set_grid(bo_world *w, float ox, float oy, float inv_h, int nx, int ny);
/* rank[i] orders free particles inside one cell (the GPU's internal index). NULL = identity */
void bo_ext_set_point_rank(bo_world *w, const uint32_t *rank, size_t n);
void bo_ext_set_particle_inv_mass(bo_world *w, size_t first, size_t n, const float *k);
void bo_ext_set_circle_inv_mass(bo_world *w, size_t first, size_t n, const float *k);
/* particle vs static convex polygon closest-edge contact on/off */
void bo_ext_set_polygon_contact(bo_world *w, int on);

/* ---- single primitives, exported for known-answer and randomised tests ---- */
void bo_prim_particle_update(float *pos, float *prev, float *acc, float dt);                 /* particle.rs:20-25 */
void bo_prim_particle_bounds(float *pos, float *prev, float bx, float by, float sx, float sy); /* particle.rs:27-46 */
void bo_prim_circle_bounds(float *pos, float *prev, float r, float bx, float by, float sx, float sy); /* circle.rs:11-30 */
void bo_prim_link_solve(float *a, float *b, float len);                                       /* link.rs:18-27 */
void bo_prim_circle_link_solve(float *a, float *b, float ra, float rb, float len);            /* link.rs:36-48 */
int bo_prim_circle_solve(float *p1, float *p2, float r1, float r2);                           /* circle.rs:32-45 */
int bo_prim_line_intersection(const float *p1, const float *p2, const float *p3, const float *p4, float *out); /* common.rs:4-26 */
/* polygon.rs:164-216; out = new_a, new_b, new_q (6 floats); returns 1 on Some */
int bo_prim_resolve_line_intersection(const float *a, const float *b, const float *q, const float *other_center,
                                      const float *self_center, float *out);
/* polygon.rs:147-162 on two explicit polygons (positions updated in place) */
void bo_prim_solve_polygon_single(float *self_xy, size_t nself, const float *self_center, float *other_xy,
                                  size_t nother, const float *other_center);

#ifdef __cplusplus
}
#endif
#endif
