// bendy_oracle.cpp — CPU ORACLE (test infrastructure only; see bendy_oracle.h header comment).
//
// Line-by-line restatement of the reference's solver.  Every function cites the reference
// file:line it follows (paths relative to /root/reference/).  nalgebra 0.32.x semantics assumed
// for Vector2<f32> (source not on this box — residual risk):
//   dot(a,b)      = a.x*b.x + a.y*b.y        (two rounded products, one rounded add, no FMA)
//   norm_squared  = dot(v,v);  magnitude = norm = sqrt(norm_squared)
//   normalize     = v / norm                  (component-wise DIVISION)
//   v*s, s*v, v/s = component-wise, left-associative as written in the source
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).

#include "bendy_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct V2 {
    float x, y;
};
static inline V2 v2(float x, float y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
static inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
static inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float norm_squared(V2 a) { return dot(a, a); }
static inline float magnitude(V2 a) { return std::sqrt(norm_squared(a)); }
static inline V2 normalize(V2 a) { return a / magnitude(a); }

struct Bounds {  // solver.rs:13-17
    V2 pos, size;
};

struct Particle {  // particle.rs:5-9
    V2 pos, prev_pos, acc;
};

// particle.rs:12-18
static inline Particle particle_new(V2 pos) { return Particle{pos, pos, v2(0.0f, 0.0f)}; }

// particle.rs:20-25
static inline void particle_update(Particle &p, float dt) {
    V2 vel = p.pos - p.prev_pos;
    p.prev_pos = p.pos;
    p.pos = p.pos + vel + p.acc * dt * dt;
    p.acc = v2(0.0f, 0.0f);
}

// particle.rs:27-46
static inline void particle_solve_bounds(Particle &p, const Bounds &b) {
    if (p.pos.x < b.pos.x) {
        float vel_x = p.prev_pos.x - p.pos.x;
        p.prev_pos.x = b.pos.x - vel_x;
        p.pos.x = b.pos.x;
    } else if (p.pos.x > b.pos.x + b.size.x) {
        float vel_x = p.prev_pos.x - p.pos.x;
        p.prev_pos.x = b.pos.x + b.size.x - vel_x;
        p.pos.x = b.pos.x + b.size.x;
    }
    if (p.pos.y < b.pos.y) {
        float vel_y = p.prev_pos.y - p.pos.y;
        p.prev_pos.y = b.pos.y - vel_y;
        p.pos.y = b.pos.y;
    } else if (p.pos.y > b.pos.y + b.size.y) {
        float vel_y = p.prev_pos.y - p.pos.y;
        p.prev_pos.y = b.pos.y + b.size.y - vel_y;
        p.pos.y = b.pos.y + b.size.y;
    }
}

struct Link {  // link.rs:5-10
    size_t a, b;
    float target_distance;
};

// link.rs:18-27 — ParticleLink::solve.  Returns false where the reference would panic
// (split_at_mut(b) needs b <= len; split.0[a] needs a < b; split.1[0] needs b < len).
static inline bool particle_link_solve(const Link &l, std::vector<Particle> &ps) {
    if (!(l.b < ps.size()) || !(l.a < l.b)) return false;
    Particle &pa = ps[l.a];
    Particle &pb = ps[l.b];
    V2 dist_vec = pa.pos - pb.pos;
    float dist = magnitude(dist_vec);
    V2 normal = normalize(dist_vec);
    V2 ca = normal * (dist - l.target_distance) * 0.5f;
    V2 cb = normal * (dist - l.target_distance) * 0.5f;
    pa.pos = pa.pos - ca;
    pb.pos = pb.pos + cb;
    return true;
}

struct Circle {  // circle.rs:5-8
    Particle point;
    float radius;
};

// circle.rs:11-30
static inline void circle_solve_bounds(Circle &c, const Bounds &b) {
    Particle &p = c.point;
    float r = c.radius;
    if (p.pos.x < b.pos.x + r) {
        float vel_x = p.prev_pos.x - p.pos.x;
        p.prev_pos.x = b.pos.x + r - vel_x;
        p.pos.x = b.pos.x + r;
    } else if (p.pos.x > b.pos.x + b.size.x - r) {
        float vel_x = p.prev_pos.x - p.pos.x;
        p.prev_pos.x = b.pos.x + b.size.x - r - vel_x;
        p.pos.x = b.pos.x + b.size.x - r;
    }
    if (p.pos.y < b.pos.y + r) {
        float vel_y = p.prev_pos.y - p.pos.y;
        p.prev_pos.y = b.pos.y + r - vel_y;
        p.pos.y = b.pos.y + r;
    } else if (p.pos.y > b.pos.y + b.size.y - r) {
        float vel_y = p.prev_pos.y - p.pos.y;
        p.prev_pos.y = b.pos.y + b.size.y - r - vel_y;
        p.pos.y = b.pos.y + b.size.y - r;
    }
}

// circle.rs:32-45 — Circle::solve_circle (strict '<'; coincident centres give NaN, no guard)
static inline bool circle_solve_circle(Circle &self, Circle &other) {
    V2 dist = self.point.pos - other.point.pos;
    float dist_sqr = norm_squared(dist);
    float radius_sum = self.radius + other.radius;
    if (dist_sqr < radius_sum * radius_sum) {
        V2 normal = normalize(dist);
        float overlap = radius_sum - std::sqrt(dist_sqr);
        float self_rad_sqr = self.radius * self.radius;
        float circle_rad_sqr = other.radius * other.radius;
        float scale = 1.0f / (self_rad_sqr + circle_rad_sqr);
        self.point.pos = self.point.pos + normal * scale * overlap * circle_rad_sqr;
        other.point.pos = other.point.pos - normal * scale * overlap * self_rad_sqr;
        return true;
    }
    return false;
}

// link.rs:36-48 — CircleLink::solve
static inline bool circle_link_solve(const Link &l, std::vector<Circle> &cs) {
    if (!(l.b < cs.size()) || !(l.a < l.b)) return false;
    Circle &ca = cs[l.a];
    Circle &cb = cs[l.b];
    V2 dist_vec = ca.point.pos - cb.point.pos;
    float dist = magnitude(dist_vec);
    V2 normal = normalize(dist_vec);
    float c_a_rad_sqr = ca.radius * ca.radius;
    float c_b_rad_sqr = cb.radius * cb.radius;
    float scale = 1.0f / (c_a_rad_sqr + c_b_rad_sqr);
    V2 da = normal * (dist - l.target_distance) * scale * c_b_rad_sqr;
    V2 db = normal * (dist - l.target_distance) * scale * c_a_rad_sqr;
    ca.point.pos = ca.point.pos - da;
    cb.point.pos = cb.point.pos + db;
    return true;
}

// common.rs:4-26 — segment/segment intersection; the denominator is computed twice (:15-16);
// parallel segments divide by zero -> inf/NaN -> every comparison false -> None.
static inline bool line_intersection(V2 p1, V2 p2, V2 p3, V2 p4, V2 *out) {
    float s1_x = p2.x - p1.x;
    float s1_y = p2.y - p1.y;
    float s2_x = p4.x - p3.x;
    float s2_y = p4.y - p3.y;
    float s = (-s1_y * (p1.x - p3.x) + s1_x * (p1.y - p3.y)) / (-s2_x * s1_y + s1_x * s2_y);
    float t = (s2_x * (p1.y - p3.y) - s2_y * (p1.x - p3.x)) / (-s2_x * s1_y + s1_x * s2_y);
    if (s >= 0.0f && s <= 1.0f && t >= 0.0f && t <= 1.0f) {
        float i_x = p1.x + (t * s1_x);
        float i_y = p1.y + (t * s1_y);
        *out = v2(i_x, i_y);
        return true;
    }
    return false;
}

struct Polygon {  // polygon.rs:8-14
    std::vector<Particle> particles;
    std::vector<Link> particle_links;
    bool is_static;
    V2 center;
    float scale;
};

// polygon.rs:231-237 — sequential f32 sum in index order, then divide by n as f32
static inline void polygon_calc_center(Polygon &p) {
    p.center = v2(0.0f, 0.0f);
    for (const Particle &pt : p.particles) p.center = p.center + pt.pos;
    p.center = p.center / (float)p.particles.size();
}

// polygon.rs:164-216 — Polygon::resolve_line_intersection
static inline bool resolve_line_intersection(V2 self_center, V2 a, V2 b, V2 q, V2 other_center, V2 *new_a,
                                             V2 *new_b, V2 *new_q) {
    V2 intersection;
    if (!line_intersection(a, b, q, other_center, &intersection)) return false;  // :171-173
    V2 normal_line = normalize(b - a);                                           // :175
    V2 center_proj =
        (dot(normal_line, self_center - intersection) / dot(normal_line, normal_line)) * normal_line;  // :177-179
    V2 normal_in = normalize(self_center - (intersection + center_proj));                             // :181
    float dist_to_a = magnitude(intersection - a);                                                    // :183
    float dist_to_b = magnitude(intersection - b);                                                    // :184
    float dist_a_to_b = dist_to_a + dist_to_b;                                                        // :186
    float influence_a = dist_to_b / dist_a_to_b;                                                      // :188
    float influence_b = dist_to_a / dist_a_to_b;                                                      // :189
    V2 diff_int_to_point = intersection - q;                                                          // :191
    V2 intersection_on_normal =
        (dot(normal_in, diff_int_to_point) / dot(normal_in, normal_in)) * normal_in;  // :193-194
    V2 displace_third = intersection_on_normal / 3.0f;                               // :196
    V2 displace_line = displace_third * 2.0f;                                        // :198
    V2 displacement_a = influence_a * displace_line;                                 // :200
    V2 displacement_b = influence_b * displace_line;                                 // :201
    *new_a = a - displacement_a;                                                     // :203
    *new_b = b - displacement_b;                                                     // :204
    V2 np;
    if (line_intersection(a, b, q, q - normal_in * 10000.0f, &np)) {  // :206-209
        *new_q = np;
        return true;
    }
    return false;  // :213
}

// polygon.rs:147-162 — outer loop edges of self (end points copied ONCE per edge), inner loop
// points of other; hits overwrite self[i], self[i+1] and the point (last writer wins).
static inline void solve_polygon_single(Polygon &self, Polygon &other) {
    size_t n = self.particles.size();
    for (size_t i = 0; i < n; i++) {
        Particle point_a = self.particles[i];
        size_t b_id = (i + 1) % n;
        Particle point_b = self.particles[b_id];
        for (Particle &others_point : other.particles) {
            V2 na, nb, nq;
            if (resolve_line_intersection(self.center, point_a.pos, point_b.pos, others_point.pos, other.center,
                                          &na, &nb, &nq)) {
                self.particles[i].pos = na;
                self.particles[b_id].pos = nb;
                others_point.pos = nq;
            }
        }
    }
}

// polygon.rs:142-145
static inline void solve_polygon(Polygon &a, Polygon &b) {
    solve_polygon_single(a, b);
    solve_polygon_single(b, a);
}

// polygon.rs:218-223 — calc_center() FIRST, then the polygon's own links (local indices)
static inline bool polygon_solve_links(Polygon &p) {
    polygon_calc_center(p);
    for (const Link &l : p.particle_links)
        if (!particle_link_solve(l, p.particles)) return false;
    return true;
}

// polygon.rs:125-134
static inline void polygon_update(Polygon &p, float dt) {
    if (p.is_static) return;
    polygon_calc_center(p);
    for (Particle &pt : p.particles) particle_update(pt, dt);
}

static const float FIX_SCALE = 1099511627776.0f;        // 2^40
static const float FIX_INV = 1.0f / 1099511627776.0f;   // 2^-40 (exact)
static const float FIX_LIMIT = 1048576.0f;              // |c| must be < 2^20 else contributes 0

// ext: order-independent fixed-point accumulation of a Circle's Jacobi correction
static inline int64_t to_fix(float c) {
    if (!(std::fabs(c) < FIX_LIMIT)) return 0;  // NaN / inf / huge -> 0 (same rule in the CUDA kernel)
    return (int64_t)std::llrintf(c * FIX_SCALE);
}
static inline float from_fix(int64_t a) { return (float)a * FIX_INV; }

}  // namespace

struct bo_world {
    // solver.rs:20-31
    V2 gravity;
    Bounds bounds;
    bool bounds_active;  // never read by the reference (solver.rs:155-165 has no check)
    std::vector<Particle> particles;
    std::vector<Link> particle_links;
    std::vector<Circle> circles;
    std::vector<Link> circle_links;
    std::vector<Polygon> polygons;
    uint16_t sub_steps;
    float sub_steps_multiplier;
    // replay
    std::vector<uint32_t> link_order;  // empty = insertion order
    // ext
    float particle_radius = 0.0f;
    float grid_ox = 0, grid_oy = 0, grid_inv_h = 1;
    int grid_nx = 1, grid_ny = 1;
    bool grid_set = false;
    std::vector<uint32_t> point_rank;
    std::vector<float> particle_k;  // empty = all 1
    std::vector<float> circle_k;    // empty = all 1
    bool polygon_contact = false;
};

namespace {

// solver.rs:130-141
static void apply_gravity(bo_world &w) {
    V2 g = w.gravity;
    for (Particle &p : w.particles) p.acc = p.acc + g;
    for (Circle &c : w.circles) c.point.acc = c.point.acc + g;
    for (Polygon &poly : w.polygons)
        for (Particle &p : poly.particles) p.acc = p.acc + g;  // polygon.rs:225-229, static included
}

static inline float pk(const bo_world &w, size_t i) { return w.particle_k.empty() ? 1.0f : w.particle_k[i]; }
static inline float ck(const bo_world &w, size_t i) { return w.circle_k.empty() ? 1.0f : w.circle_k[i]; }

// solver.rs:143-153
static bool apply_links(bo_world &w) {
    const bool uniform = w.particle_k.empty();
    size_t nl = w.particle_links.size();
    for (size_t k = 0; k < nl; k++) {
        const Link &l = w.particle_links[w.link_order.empty() ? k : w.link_order[k]];
        if (uniform) {
            if (!particle_link_solve(l, w.particles)) return false;
        } else {
            // ext (parity unpinned): inverse-mass weighted split; equals link.rs:25-26 bit-for-bit
            // when both weights are 1 (1/(1+1) == 0.5 exactly).
            if (!(l.b < w.particles.size()) || !(l.a < l.b)) return false;
            float ka = w.particle_k[l.a], kb = w.particle_k[l.b];
            if (ka == 0.0f && kb == 0.0f) continue;
            Particle &pa = w.particles[l.a];
            Particle &pb = w.particles[l.b];
            V2 dist_vec = pa.pos - pb.pos;
            float dist = magnitude(dist_vec);
            V2 normal = normalize(dist_vec);
            float wa = ka / (ka + kb), wb = kb / (ka + kb);
            pa.pos = pa.pos - normal * (dist - l.target_distance) * wa;
            pb.pos = pb.pos + normal * (dist - l.target_distance) * wb;
        }
    }
    for (const Link &l : w.circle_links)
        if (!circle_link_solve(l, w.circles)) return false;
    for (Polygon &p : w.polygons)
        if (!polygon_solve_links(p)) return false;
    return true;
}

static inline int cell_coord(float x, float o, float inv_h, int n) {
    float f = (x - o) * inv_h;
    if (!(f >= 0.0f)) return 0;  // negative and NaN
    if (f >= (float)n) return n - 1;
    return (int)f;
}

// ext — PARITY UNPINNED BY THE REFERENCE (free particles collide with nothing there,
// solver.rs:167-188).  Spec: every free particle is a disc of radius r_p; the
// per-pair rule is Circle::solve_circle (circle.rs:32-45) seen from the disc being updated;
// update discipline is Jacobi: all overlap tests and normals use the positions at phase entry.
// Each per-pair correction (an f32 computed exactly like circle.rs:42) is converted to 2^-40 fixed
// point and summed in int64, so the sum does not depend on the visiting order; the disc then moves
// by pos + (float)sum.  A correction >= 2^-17 converts exactly, so for a contact graph that is a
// matching this equals the reference's sequential pass bit-for-bit.  The grid only prunes pairs.
static void ext_disc_contacts(bo_world &w, const std::vector<V2> &QC) {
    const size_t n = w.particles.size();
    const size_t nc = w.circles.size();
    if (n == 0) return;
    const float rp = w.particle_radius;
    std::vector<V2> Q(n);
    for (size_t i = 0; i < n; i++) Q[i] = w.particles[i].pos;

    const int nx = w.grid_nx, ny = w.grid_ny;
    std::vector<uint32_t> cell(n);
    const uint32_t NOCELL = 0xFFFFFFFFu;  // non-finite positions overlap nothing (all compares false)
    for (size_t i = 0; i < n; i++) {
        if (!std::isfinite(Q[i].x) || !std::isfinite(Q[i].y)) {
            cell[i] = NOCELL;
            continue;
        }
        int cx = cell_coord(Q[i].x, w.grid_ox, w.grid_inv_h, nx);
        int cy = cell_coord(Q[i].y, w.grid_oy, w.grid_inv_h, ny);
        cell[i] = (uint32_t)cy * (uint32_t)nx + (uint32_t)cx;
    }
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (cell[a] != cell[b]) return cell[a] < cell[b];
        return a < b;
    });
    std::vector<uint32_t> sorted_cell(n);
    for (size_t s = 0; s < n; s++) sorted_cell[s] = cell[order[s]];
    auto cell_begin = [&](uint32_t c) {
        return (size_t)(std::lower_bound(sorted_cell.begin(), sorted_cell.end(), c) - sorted_cell.begin());
    };

    std::vector<int64_t> accx(nc, 0), accy(nc, 0);
    const float rp_sq = rp * rp;
    std::vector<V2> out(n);
    for (size_t i = 0; i < n; i++) {
        out[i] = Q[i];
        if (cell[i] == NOCELL) continue;
        const float ki = pk(w, i);
        int64_t sx = 0, sy = 0;
        bool moved = false;
        int cx = (int)(cell[i] % (uint32_t)nx), cy = (int)(cell[i] / (uint32_t)nx);
        for (int dy = -1; dy <= 1; dy++) {
            int yy = cy + dy;
            if (yy < 0 || yy >= ny) continue;
            int x0 = std::max(cx - 1, 0), x1 = std::min(cx + 1, nx - 1);
            size_t s0 = cell_begin((uint32_t)yy * nx + x0);
            size_t s1 = cell_begin((uint32_t)yy * nx + x1 + 1);
            for (size_t s = s0; s < s1; s++) {
                uint32_t j = order[s];
                if (j == i) continue;
                V2 dist = Q[i] - Q[j];
                float dist_sqr = norm_squared(dist);
                float radius_sum = rp + rp;
                if (dist_sqr < radius_sum * radius_sum) {
                    if (ki == 0.0f) continue;
                    const float kj = pk(w, j);
                    V2 normal = normalize(dist);
                    float overlap = radius_sum - std::sqrt(dist_sqr);
                    float wi = ki * rp_sq, wj = kj * rp_sq;  // k_i*r_j^2, k_j*r_i^2
                    float scale = 1.0f / (wj + wi);
                    V2 c = normal * scale * overlap * wi;
                    sx += to_fix(c.x), sy += to_fix(c.y);
                    moved = true;
                }
            }
        }
        for (size_t c = 0; c < nc; c++) {
            V2 dist = Q[i] - QC[c];
            float dist_sqr = norm_squared(dist);
            float R = w.circles[c].radius;
            float radius_sum = rp + R;
            if (dist_sqr < radius_sum * radius_sum) {
                const float kc = ck(w, c);
                if (ki == 0.0f && kc == 0.0f) continue;
                V2 normal = normalize(dist);
                float overlap = radius_sum - std::sqrt(dist_sqr);
                float wi = ki * (R * R), wc = kc * rp_sq;
                float scale = 1.0f / (wc + wi);
                V2 x = normal * scale * overlap;
                V2 ci = x * wi;
                sx += to_fix(ci.x), sy += to_fix(ci.y);
                moved = true;
                V2 cc = x * wc;
                accx[c] += to_fix(-cc.x);
                accy[c] += to_fix(-cc.y);
            }
        }
        if (moved) out[i] = v2(Q[i].x + from_fix(sx), Q[i].y + from_fix(sy));
    }
    for (size_t i = 0; i < n; i++) w.particles[i].pos = out[i];
    for (size_t c = 0; c < nc; c++) {
        if (accx[c] == 0 && accy[c] == 0) continue;
        // the Circle receives its share AFTER the reference's circle-circle pass has run
        V2 cur = w.circles[c].point.pos;
        w.circles[c].point.pos = v2(cur.x + from_fix(accx[c]), cur.y + from_fix(accy[c]));
    }
}

// ext — PARITY UNPINNED BY THE REFERENCE (a free particle never meets a polygon there).
// Spec: static convex polygons are immovable obstacles for free particles.
// For a particle q whose position lies in the polygon's AABB (points and centre): per edge e=(a,b) build the inward
// normal exactly as polygon.rs:175-181 does with the edge start as the on-line point; q is inside
// iff every signed inward distance is > 0; the closest edge (smallest distance, lowest index on
// ties) receives q via the reference's projection polygon.rs:206-209
// (line_intersection(edge, (q, q - n_in*10000))).
// Several obstacles (they may overlap): the CANDIDATES of a particle are the static polygons whose AABB
// contains its position at the ENTRY of this step (after the disc contacts), taken in ascending polygon
// index; each candidate is then tested (AABB again, inside test, projection) against the particle's
// CURRENT position, i.e. after the projections of the earlier candidates.  A polygon the particle was not
// inside the AABB of at entry is not revisited in this substep, even if a projection moves the particle
// into it (it becomes a candidate in the next substep).  The rule is independent of any binning.
static void ext_polygon_contacts(bo_world &w) {
    const size_t n = w.particles.size();
    struct Box {
        float x0, y0, x1, y1;
    };
    std::vector<Box> boxes(w.polygons.size());
    for (size_t k = 0; k < w.polygons.size(); k++) {
        const Polygon &P = w.polygons[k];
        Box b{INFINITY, INFINITY, -INFINITY, -INFINITY};
        for (const Particle &pt : P.particles) {
            b.x0 = std::fmin(b.x0, pt.pos.x);
            b.y0 = std::fmin(b.y0, pt.pos.y);
            b.x1 = std::fmax(b.x1, pt.pos.x);
            b.y1 = std::fmax(b.y1, pt.pos.y);
        }
        // the AABB covers the polygon's points AND its centre (the one solve_links computed, polygon.rs:219):
        // the inward normals are built from that centre, and after a polygon-polygon contact has squashed a
        // static polygon the centre can lie outside the hull of the points
        b.x0 = std::fmin(b.x0, P.center.x), b.y0 = std::fmin(b.y0, P.center.y);
        b.x1 = std::fmax(b.x1, P.center.x), b.y1 = std::fmax(b.y1, P.center.y);
        boxes[k] = b;
    }
    // coarse bins over the polygon boxes so the oracle stays O(n) at test sizes
    float wx0 = INFINITY, wy0 = INFINITY, wx1 = -INFINITY, wy1 = -INFINITY, maxext = 0.0f;
    for (const Box &b : boxes) {
        wx0 = std::fmin(wx0, b.x0), wy0 = std::fmin(wy0, b.y0);
        wx1 = std::fmax(wx1, b.x1), wy1 = std::fmax(wy1, b.y1);
        maxext = std::fmax(maxext, std::fmax(b.x1 - b.x0, b.y1 - b.y0));
    }
    if (!(maxext > 0.0f)) return;
    const float bin = maxext;
    const int bx = std::min(4096, std::max(1, (int)std::ceil((wx1 - wx0) / bin)));
    const int by = std::min(4096, std::max(1, (int)std::ceil((wy1 - wy0) / bin)));
    std::vector<std::vector<uint32_t>> bins((size_t)bx * by);
    auto bcoord = [&](float v, float o, int nb) {
        int c = (int)std::floor((v - o) / bin);
        return std::min(nb - 1, std::max(0, c));
    };
    for (size_t k = 0; k < boxes.size(); k++) {
        if (!w.polygons[k].is_static) continue;
        for (int yy = bcoord(boxes[k].y0, wy0, by); yy <= bcoord(boxes[k].y1, wy0, by); yy++)
            for (int xx = bcoord(boxes[k].x0, wx0, bx); xx <= bcoord(boxes[k].x1, wx0, bx); xx++)
                bins[(size_t)yy * bx + xx].push_back((uint32_t)k);
    }
    for (size_t i = 0; i < n; i++) {
        if (pk(w, i) == 0.0f) continue;
        V2 q = w.particles[i].pos;
        const V2 q0 = q;  // position at entry: decides the candidate set
        if (!(q.x >= wx0 && q.x <= wx1 && q.y >= wy0 && q.y <= wy1)) continue;
        const std::vector<uint32_t> &cand = bins[(size_t)bcoord(q.y, wy0, by) * bx + bcoord(q.x, wx0, bx)];
        for (uint32_t k : cand) {  // ascending polygon index by construction
            const Box &b = boxes[k];
            if (!(q0.x >= b.x0 && q0.x <= b.x1 && q0.y >= b.y0 && q0.y <= b.y1)) continue;  // not a candidate
            if (!(q.x >= b.x0 && q.x <= b.x1 && q.y >= b.y0 && q.y <= b.y1)) continue;      // q may have moved
            const Polygon &P = w.polygons[k];
            const size_t E = P.particles.size();
            bool inside = true;
            float best = INFINITY;
            size_t best_e = 0;
            V2 best_nin = v2(0, 0);
            for (size_t e = 0; e < E; e++) {
                V2 a = P.particles[e].pos, bb = P.particles[(e + 1) % E].pos;
                V2 normal_line = normalize(bb - a);
                V2 center_proj = (dot(normal_line, P.center - a) / dot(normal_line, normal_line)) * normal_line;
                V2 normal_in = normalize(P.center - (a + center_proj));
                float sd = dot(normal_in, q - a);
                if (!(sd > 0.0f)) {
                    inside = false;
                    break;
                }
                if (sd < best) best = sd, best_e = e, best_nin = normal_in;
            }
            if (!inside) continue;
            V2 a = P.particles[best_e].pos, bb = P.particles[(best_e + 1) % E].pos;
            V2 np;
            if (line_intersection(a, bb, q, q - best_nin * 10000.0f, &np)) {
                q = np;
                w.particles[i].pos = q;
            }
        }
    }
}

// solver.rs:167-188
static void solve_dynamic_collisions(bo_world &w) {
    // ext: the disc contacts below are tested against the circle centres at phase ENTRY (a Jacobi
    // snapshot, like the particle positions), so that on the device the circle-circle pass can
    // overlap the narrowphase
    std::vector<V2> circle_entry;
    if (w.particle_radius > 0.0f) {
        circle_entry.resize(w.circles.size());
        for (size_t c = 0; c < w.circles.size(); c++) circle_entry[c] = w.circles[c].point.pos;
    }
    size_t length = w.circles.size();
    for (size_t i = 0; i < length; i++)
        for (size_t j = i + 1; j < length; j++) circle_solve_circle(w.circles[i], w.circles[j]);
    length = w.polygons.size();
    for (size_t i = 0; i < length; i++)
        for (size_t j = i + 1; j < length; j++) solve_polygon(w.polygons[i], w.polygons[j]);
    // ext phases (off by default => reference semantics)
    if (w.particle_radius > 0.0f) ext_disc_contacts(w, circle_entry);
    if (w.polygon_contact) ext_polygon_contacts(w);
}

// solver.rs:155-165
static void solve_boundary_collisions(bo_world &w) {
    for (size_t i = 0; i < w.particles.size(); i++) {
        if (pk(w, i) == 0.0f) continue;  // ext: pinned
        particle_solve_bounds(w.particles[i], w.bounds);
    }
    for (size_t i = 0; i < w.circles.size(); i++) {
        if (ck(w, i) == 0.0f) continue;
        circle_solve_bounds(w.circles[i], w.bounds);
    }
    for (Polygon &poly : w.polygons)
        for (Particle &p : poly.particles) particle_solve_bounds(p, w.bounds);  // polygon.rs:136-140
}

// solver.rs:118-128
static void update_positions(bo_world &w, float dt) {
    for (size_t i = 0; i < w.particles.size(); i++) {
        if (pk(w, i) == 0.0f) {  // ext: pinned point keeps pos/prev; acc cleared like update()
            w.particles[i].acc = v2(0, 0);
            continue;
        }
        particle_update(w.particles[i], dt);
    }
    for (size_t i = 0; i < w.circles.size(); i++) {
        if (ck(w, i) == 0.0f) {
            w.circles[i].point.acc = v2(0, 0);
            continue;
        }
        particle_update(w.circles[i].point, dt);
    }
    for (Polygon &poly : w.polygons) polygon_update(poly, dt);
}

static bool links_valid(const bo_world &w) {
    for (const Link &l : w.particle_links)
        if (!(l.b < w.particles.size()) || !(l.a < l.b)) return false;
    for (const Link &l : w.circle_links)
        if (!(l.b < w.circles.size()) || !(l.a < l.b)) return false;
    for (const Polygon &p : w.polygons)
        for (const Link &l : p.particle_links)
            if (!(l.b < p.particles.size()) || !(l.a < l.b)) return false;
    return true;
}

}  // namespace

extern "C" {

bo_world *bo_create(void) {
    bo_world *w = new bo_world();
    w->gravity = v2(0.0f, 98.2f);                                // solver.rs:36
    w->bounds = Bounds{v2(0.0f, 0.0f), v2(100.0f, 100.0f)};      // solver.rs:37-40
    w->bounds_active = true;
    w->sub_steps = 1;                                            // solver.rs:47
    w->sub_steps_multiplier = 0.0f;
    return w;
}
void bo_destroy(bo_world *w) { delete w; }
bo_world *bo_clone(const bo_world *w) { return new bo_world(*w); }

void bo_set_gravity(bo_world *w, float gx, float gy) { w->gravity = v2(gx, gy); }
void bo_set_bounds(bo_world *w, float bx, float by, float sx, float sy) {
    w->bounds = Bounds{v2(bx, by), v2(sx, sy)};
}

void bo_add_particle(bo_world *w, float x, float y) { w->particles.push_back(particle_new(v2(x, y))); }
void bo_add_particles(bo_world *w, const float *pos_xy, size_t n) {
    w->particles.reserve(w->particles.size() + n);
    for (size_t i = 0; i < n; i++) w->particles.push_back(particle_new(v2(pos_xy[2 * i], pos_xy[2 * i + 1])));
}
void bo_add_particle_links(bo_world *w, const uint32_t *ab, const float *len, size_t n) {
    w->particle_links.reserve(w->particle_links.size() + n);
    for (size_t k = 0; k < n; k++) w->particle_links.push_back(Link{ab[2 * k], ab[2 * k + 1], len[k]});
}
void bo_add_circle(bo_world *w, float px, float py, float qx, float qy, float ax, float ay, float radius) {
    w->circles.push_back(Circle{Particle{v2(px, py), v2(qx, qy), v2(ax, ay)}, radius});
}
int bo_add_polygon(bo_world *w, const float *pos_xy, const float *prev_xy, const float *acc_xy, size_t nv,
                   const uint32_t *link_ab, const float *link_len, size_t nl, int is_static, float cx, float cy) {
    if (nv == 0 || !pos_xy) return BO_ERR_ARG;
    Polygon p;
    for (size_t i = 0; i < nv; i++) {
        Particle pt;
        pt.pos = v2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        pt.prev_pos = prev_xy ? v2(prev_xy[2 * i], prev_xy[2 * i + 1]) : pt.pos;
        pt.acc = acc_xy ? v2(acc_xy[2 * i], acc_xy[2 * i + 1]) : v2(0, 0);
        p.particles.push_back(pt);
    }
    for (size_t k = 0; k < nl; k++) p.particle_links.push_back(Link{link_ab[2 * k], link_ab[2 * k + 1], link_len[k]});
    p.is_static = is_static != 0;
    p.center = v2(cx, cy);
    p.scale = 1.0f;
    w->polygons.push_back(std::move(p));
    return BO_OK;
}

// polygon.rs:84-123 — Polygon::new: perimeter links i <-> (i+1)%n with a<b, length = initial distance
int bo_add_polygon_new(bo_world *w, const float *pts_xy, size_t nv, int is_static) {
    if (nv == 0 || !pts_xy) return BO_ERR_ARG;
    Polygon p;
    V2 center = v2(0.0f, 0.0f);
    for (size_t i = 0; i < nv; i++) {
        Particle pt = particle_new(v2(pts_xy[2 * i], pts_xy[2 * i + 1]));
        p.particles.push_back(pt);
        center = center + pt.pos;
    }
    center = center / (float)nv;
    for (size_t i = 0; i < nv; i++) {
        size_t a_id = i, b_id = (i + 1) % nv;
        if (a_id > b_id) std::swap(a_id, b_id);
        V2 dist_vec = p.particles[a_id].pos - p.particles[b_id].pos;
        p.particle_links.push_back(Link{a_id, b_id, magnitude(dist_vec)});
    }
    p.is_static = is_static != 0;
    p.center = center;
    p.scale = 1.0f;
    w->polygons.push_back(std::move(p));
    return BO_OK;
}

// polygon.rs:17-82 — Polygon::circle: f32 angle accumulation, chords i<->i+2n/3 and i<->i+n/3.
// cosf/sinf stand in for Rust's f32::cos/sin (libm-dependent in the last ulp).
int bo_add_polygon_circle(bo_world *w, float radius, float px, float py, size_t n, int is_static) {
    if (n == 0) return BO_ERR_ARG;
    Polygon p;
    V2 center = v2(0.0f, 0.0f);
    float angle = 0.0f;
    for (size_t i = 0; i < n; i++) {
        float x = radius * std::cos(angle);
        float y = radius * std::sin(angle);
        Particle pt = particle_new(v2(px, py) + v2(x, y));
        p.particles.push_back(pt);
        center = center + pt.pos;
        angle += 2.0f * 3.14159265358979323846f / (float)n;
    }
    center = center / (float)n;
    for (size_t i = 0; i < n; i++) {
        for (int pass = 0; pass < 2; pass++) {
            size_t a_id = i;
            size_t b_id = pass == 0 ? (i + 2 * n / 3) % n : (i + n / 3) % n;
            if (a_id > b_id) std::swap(a_id, b_id);
            V2 dist_vec = p.particles[a_id].pos - p.particles[b_id].pos;
            p.particle_links.push_back(Link{a_id, b_id, magnitude(dist_vec)});
        }
    }
    p.is_static = is_static != 0;
    p.center = center;
    p.scale = 1.0f;
    w->polygons.push_back(std::move(p));
    return BO_OK;
}

void bo_add_particle_link(bo_world *w, size_t a, size_t b, float len) { w->particle_links.push_back(Link{a, b, len}); }
void bo_add_circle_link(bo_world *w, size_t a, size_t b, float len) { w->circle_links.push_back(Link{a, b, len}); }

// solver.rs:106-116
int bo_update(bo_world *w, float dt) {
    if (!links_valid(*w)) return BO_ERR_PANIC;  // link.rs:19-21 would panic inside apply_links
    if (!w->link_order.empty() && w->link_order.size() != w->particle_links.size()) return BO_ERR_ARG;
    w->sub_steps_multiplier = 1.0f / (float)w->sub_steps;
    float delta = dt * w->sub_steps_multiplier;
    for (uint16_t s = 0; s < w->sub_steps; s++) {
        apply_gravity(*w);
        if (!apply_links(*w)) return BO_ERR_PANIC;
        solve_dynamic_collisions(*w);
        solve_boundary_collisions(*w);
        update_positions(*w, delta);
    }
    return BO_OK;
}

size_t bo_particle_len(const bo_world *w) { return w->particles.size(); }
size_t bo_circle_len(const bo_world *w) { return w->circles.size(); }
size_t bo_polygon_len(const bo_world *w) { return w->polygons.size(); }
size_t bo_particle_link_len(const bo_world *w) { return w->particle_links.size(); }

void bo_read_particles(const bo_world *w, float *pos_xy, float *prev_xy) {
    for (size_t i = 0; i < w->particles.size(); i++) {
        if (pos_xy) pos_xy[2 * i] = w->particles[i].pos.x, pos_xy[2 * i + 1] = w->particles[i].pos.y;
        if (prev_xy) prev_xy[2 * i] = w->particles[i].prev_pos.x, prev_xy[2 * i + 1] = w->particles[i].prev_pos.y;
    }
}
void bo_write_particles(bo_world *w, const float *pos_xy, const float *prev_xy) {
    for (size_t i = 0; i < w->particles.size(); i++) {
        if (pos_xy) w->particles[i].pos = v2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        if (prev_xy) w->particles[i].prev_pos = v2(prev_xy[2 * i], prev_xy[2 * i + 1]);
    }
}
void bo_read_circles(const bo_world *w, float *pos_xy, float *prev_xy, float *radius) {
    for (size_t i = 0; i < w->circles.size(); i++) {
        const Circle &c = w->circles[i];
        if (pos_xy) pos_xy[2 * i] = c.point.pos.x, pos_xy[2 * i + 1] = c.point.pos.y;
        if (prev_xy) prev_xy[2 * i] = c.point.prev_pos.x, prev_xy[2 * i + 1] = c.point.prev_pos.y;
        if (radius) radius[i] = c.radius;
    }
}
size_t bo_polygon_point_len(const bo_world *w, size_t poly) {
    return poly < w->polygons.size() ? w->polygons[poly].particles.size() : 0;
}
size_t bo_polygon_link_len(const bo_world *w, size_t poly) {
    return poly < w->polygons.size() ? w->polygons[poly].particle_links.size() : 0;
}
void bo_read_polygon(const bo_world *w, size_t poly, float *pos_xy, float *prev_xy, float *center_xy) {
    if (poly >= w->polygons.size()) return;
    const Polygon &p = w->polygons[poly];
    for (size_t i = 0; i < p.particles.size(); i++) {
        if (pos_xy) pos_xy[2 * i] = p.particles[i].pos.x, pos_xy[2 * i + 1] = p.particles[i].pos.y;
        if (prev_xy) prev_xy[2 * i] = p.particles[i].prev_pos.x, prev_xy[2 * i + 1] = p.particles[i].prev_pos.y;
    }
    if (center_xy) center_xy[0] = p.center.x, center_xy[1] = p.center.y;
}
void bo_read_polygon_links(const bo_world *w, size_t poly, uint32_t *ab, float *len) {
    if (poly >= w->polygons.size()) return;
    const Polygon &p = w->polygons[poly];
    for (size_t k = 0; k < p.particle_links.size(); k++) {
        ab[2 * k] = (uint32_t)p.particle_links[k].a;
        ab[2 * k + 1] = (uint32_t)p.particle_links[k].b;
        len[k] = p.particle_links[k].target_distance;
    }
}

int bo_set_link_order(bo_world *w, const uint32_t *perm, size_t n) {
    if (!perm || n == 0) {
        w->link_order.clear();
        return BO_OK;
    }
    if (n != w->particle_links.size()) return BO_ERR_ARG;
    std::vector<uint8_t> seen(n, 0);
    for (size_t k = 0; k < n; k++) {
        if (perm[k] >= n || seen[perm[k]]) return BO_ERR_ARG;
        seen[perm[k]] = 1;
    }
    w->link_order.assign(perm, perm + n);
    return BO_OK;
}
void bo_set_sub_steps(bo_world *w, uint16_t n) { w->sub_steps = n ? n : 1; }

void bo_ext_set_particle_radius(bo_world *w, float r) { w->particle_radius = r; }
void bo_ext_set_grid(bo_world *w, float ox, float oy, float inv_h, int nx, int ny) {
    w->grid_ox = ox, w->grid_oy = oy, w->grid_inv_h = inv_h;
    w->grid_nx = nx > 0 ? nx : 1, w->grid_ny = ny > 0 ? ny : 1;
    w->grid_set = true;
}
void bo_ext_set_point_rank(bo_world *w, const uint32_t *rank, size_t n) {
    if (!rank || n == 0)
        w->point_rank.clear();
    else
        w->point_rank.assign(rank, rank + n);
}
void bo_ext_set_particle_inv_mass(bo_world *w, size_t first, size_t n, const float *k) {
    if (w->particle_k.size() < w->particles.size()) w->particle_k.resize(w->particles.size(), 1.0f);
    for (size_t i = 0; i < n && first + i < w->particle_k.size(); i++) w->particle_k[first + i] = k[i];
}
void bo_ext_set_circle_inv_mass(bo_world *w, size_t first, size_t n, const float *k) {
    if (w->circle_k.size() < w->circles.size()) w->circle_k.resize(w->circles.size(), 1.0f);
    for (size_t i = 0; i < n && first + i < w->circle_k.size(); i++) w->circle_k[first + i] = k[i];
}
void bo_ext_set_polygon_contact(bo_world *w, int on) { w->polygon_contact = on != 0; }

/* ---- primitives ---- */
void bo_prim_particle_update(float *pos, float *prev, float *acc, float dt) {
    Particle p{v2(pos[0], pos[1]), v2(prev[0], prev[1]), v2(acc[0], acc[1])};
    particle_update(p, dt);
    pos[0] = p.pos.x, pos[1] = p.pos.y, prev[0] = p.prev_pos.x, prev[1] = p.prev_pos.y;
    acc[0] = p.acc.x, acc[1] = p.acc.y;
}
void bo_prim_particle_bounds(float *pos, float *prev, float bx, float by, float sx, float sy) {
    Particle p{v2(pos[0], pos[1]), v2(prev[0], prev[1]), v2(0, 0)};
    particle_solve_bounds(p, Bounds{v2(bx, by), v2(sx, sy)});
    pos[0] = p.pos.x, pos[1] = p.pos.y, prev[0] = p.prev_pos.x, prev[1] = p.prev_pos.y;
}
void bo_prim_circle_bounds(float *pos, float *prev, float r, float bx, float by, float sx, float sy) {
    Circle c{Particle{v2(pos[0], pos[1]), v2(prev[0], prev[1]), v2(0, 0)}, r};
    circle_solve_bounds(c, Bounds{v2(bx, by), v2(sx, sy)});
    pos[0] = c.point.pos.x, pos[1] = c.point.pos.y, prev[0] = c.point.prev_pos.x, prev[1] = c.point.prev_pos.y;
}
void bo_prim_link_solve(float *a, float *b, float len) {
    std::vector<Particle> ps{particle_new(v2(a[0], a[1])), particle_new(v2(b[0], b[1]))};
    particle_link_solve(Link{0, 1, len}, ps);
    a[0] = ps[0].pos.x, a[1] = ps[0].pos.y, b[0] = ps[1].pos.x, b[1] = ps[1].pos.y;
}
void bo_prim_circle_link_solve(float *a, float *b, float ra, float rb, float len) {
    std::vector<Circle> cs{Circle{particle_new(v2(a[0], a[1])), ra}, Circle{particle_new(v2(b[0], b[1])), rb}};
    circle_link_solve(Link{0, 1, len}, cs);
    a[0] = cs[0].point.pos.x, a[1] = cs[0].point.pos.y, b[0] = cs[1].point.pos.x, b[1] = cs[1].point.pos.y;
}
int bo_prim_circle_solve(float *p1, float *p2, float r1, float r2) {
    Circle c1{particle_new(v2(p1[0], p1[1])), r1}, c2{particle_new(v2(p2[0], p2[1])), r2};
    bool hit = circle_solve_circle(c1, c2);
    p1[0] = c1.point.pos.x, p1[1] = c1.point.pos.y, p2[0] = c2.point.pos.x, p2[1] = c2.point.pos.y;
    return hit ? 1 : 0;
}
int bo_prim_line_intersection(const float *p1, const float *p2, const float *p3, const float *p4, float *out) {
    V2 o;
    if (line_intersection(v2(p1[0], p1[1]), v2(p2[0], p2[1]), v2(p3[0], p3[1]), v2(p4[0], p4[1]), &o)) {
        out[0] = o.x, out[1] = o.y;
        return 1;
    }
    return 0;
}
int bo_prim_resolve_line_intersection(const float *a, const float *b, const float *q, const float *other_center,
                                      const float *self_center, float *out) {
    V2 na, nb, nq;
    if (resolve_line_intersection(v2(self_center[0], self_center[1]), v2(a[0], a[1]), v2(b[0], b[1]),
                                  v2(q[0], q[1]), v2(other_center[0], other_center[1]), &na, &nb, &nq)) {
        out[0] = na.x, out[1] = na.y, out[2] = nb.x, out[3] = nb.y, out[4] = nq.x, out[5] = nq.y;
        return 1;
    }
    return 0;
}
void bo_prim_solve_polygon_single(float *self_xy, size_t nself, const float *self_center, float *other_xy,
                                  size_t nother, const float *other_center) {
    Polygon A, B;
    for (size_t i = 0; i < nself; i++) A.particles.push_back(particle_new(v2(self_xy[2 * i], self_xy[2 * i + 1])));
    for (size_t i = 0; i < nother; i++) B.particles.push_back(particle_new(v2(other_xy[2 * i], other_xy[2 * i + 1])));
    A.center = v2(self_center[0], self_center[1]);
    B.center = v2(other_center[0], other_center[1]);
    A.is_static = B.is_static = false;
    solve_polygon_single(A, B);
    for (size_t i = 0; i < nself; i++) self_xy[2 * i] = A.particles[i].pos.x, self_xy[2 * i + 1] = A.particles[i].pos.y;
    for (size_t i = 0; i < nother; i++) other_xy[2 * i] = B.particles[i].pos.x, other_xy[2 * i + 1] = B.particles[i].pos.y;
}

}  // extern "C"
