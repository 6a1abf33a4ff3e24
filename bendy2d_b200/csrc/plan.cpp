// plan.cpp — see plan.h.  Pure host C++ (no CUDA) so that it can be unit-tested without a GPU.
#include "plan.h"

#include <algorithm>
#include <numeric>

namespace bendy {

namespace {

struct ColourMask {  // 256 colours per vertex
    uint64_t w[4] = {0, 0, 0, 0};
};

static inline int lowest_free(const ColourMask &x, const ColourMask &y) {
    for (int k = 0; k < 4; k++) {
        uint64_t used = x.w[k] | y.w[k];
        if (~used) return k * 64 + __builtin_ctzll(~used);
    }
    return -1;
}
static inline void mark(ColourMask &m, int c) { m.w[c >> 6] |= (1ull << (c & 63)); }

static uint32_t uf_find(std::vector<uint32_t> &p, uint32_t x) {
    while (p[x] != x) {
        p[x] = p[p[x]];
        x = p[x];
    }
    return x;
}

}  // namespace

std::vector<uint32_t> LinkPlan::perm() const {
    std::vector<uint32_t> p;
    p.reserve(local_user.size() + global_user.size());
    p.insert(p.end(), local_user.begin(), local_user.end());
    p.insert(p.end(), global_user.begin(), global_user.end());
    return p;
}

bool plan_links(size_t n_points, const uint32_t *ab, const float *len, size_t n_links, const PlanParams &pp_in,
                bool keep_order, LinkPlan *out, std::string *err, const uint8_t *priority) {
    PlanParams pp = pp_in;
    if (pp.max_points == 0) pp.max_points = PlanParams().max_points;
    if (pp.pack_points == 0) pp.pack_points = PlanParams().pack_points;
    pp.max_points = std::min<uint32_t>(pp.max_points, 16384);  // 16-bit local indices, 128 KB of smem
    pp.pack_points = std::min(pp.pack_points, pp.max_points);

    LinkPlan &P = *out;
    P = LinkPlan();
    P.n_points = n_points;
    const uint32_t n = (uint32_t)n_points;
    P.rank.resize(n);
    P.order.resize(n);

    // ---- 1. renumber: linked components first (by first appearance), unlinked points last
    size_t n_linked = 0;
    size_t n_priority_points = 0;
    if (keep_order || n_links == 0) {
        std::iota(P.rank.begin(), P.rank.end(), 0u);
        std::iota(P.order.begin(), P.order.end(), 0u);
        if (n_links) {
            uint32_t hi = 0;
            for (size_t k = 0; k < n_links; k++) hi = std::max(hi, std::max(ab[2 * k], ab[2 * k + 1]));
            n_linked = (size_t)hi + 1;
        }
    } else {
        std::vector<uint32_t> parent(n);
        std::iota(parent.begin(), parent.end(), 0u);
        std::vector<uint8_t> linked(n, 0);
        for (size_t k = 0; k < n_links; k++) {
            uint32_t a = ab[2 * k], b = ab[2 * k + 1];
            linked[a] = linked[b] = 1;
            uint32_t ra = uf_find(parent, a), rb = uf_find(parent, b);
            if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);  // root = smallest index
        }
        // component label in order of first appearance == order of the root index
        std::vector<uint32_t> comp_size(n, 0);
        for (uint32_t i = 0; i < n; i++)
            if (linked[i]) comp_size[uf_find(parent, i)]++;
        // components holding a priority point (strips: bodies near a strip edge) are numbered first so
        // that their partitions form the leading range [0, n_priority_parts)
        std::vector<uint8_t> comp_prio(priority ? n : 0, 0);
        if (priority)
            for (uint32_t i = 0; i < n; i++)
                if (linked[i] && priority[i]) comp_prio[uf_find(parent, i)] = 1;
        std::vector<uint32_t> comp_off(n, 0);
        uint32_t run = 0;
        for (int pass = priority ? 0 : 1; pass < 2; pass++) {
            for (uint32_t r = 0; r < n; r++) {
                if (!comp_size[r]) continue;
                const bool prio = priority && comp_prio[r];
                if (priority && prio != (pass == 0)) continue;
                comp_off[r] = run;
                run += comp_size[r];
            }
            if (pass == 0) n_priority_points = run;
        }
        n_linked = run;
        uint32_t free_run = run;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t dst = linked[i] ? comp_off[uf_find(parent, i)]++ : free_run++;
            P.rank[i] = dst;
            P.order[dst] = i;
        }
    }
    if (n_links == 0) return true;

    // ---- 2. cut the linked range into partitions at "clean" boundaries where possible
    // crossing[i] = number of links spanning the boundary between internal points i-1 and i
    std::vector<int32_t> crossing(n_linked + 1, 0);
    for (size_t k = 0; k < n_links; k++) {
        uint32_t a = P.rank[ab[2 * k]], b = P.rank[ab[2 * k + 1]];
        if (a > b) std::swap(a, b);
        crossing[a + 1] += 1;
        if ((size_t)b + 1 <= n_linked) crossing[b + 1] -= 1;
    }
    for (size_t i = 1; i <= n_linked; i++) crossing[i] += crossing[i - 1];
    crossing[n_linked] = 0;
    P.part_start.push_back(0);
    size_t pos = 0;
    while (pos < n_linked) {
        if (pos == n_priority_points) P.n_priority_parts = P.n_parts();
        size_t lim = std::min(n_linked, pos + (size_t)pp.max_points);
        if (pos < n_priority_points) lim = std::min(lim, n_priority_points);  // never pack across the group edge
        size_t best = 0;
        for (size_t e = pos + 1; e <= lim; e++) {
            if (crossing[e] != 0) continue;
            if (e - pos <= pp.pack_points) {
                best = e;
            } else {
                if (!best) best = e;
                break;
            }
        }
        if (!best) best = lim;  // no clean boundary: dirty cut, spanning links become global
        P.part_start.push_back((uint32_t)best);
        pos = best;
    }
    if (n_priority_points >= n_linked) P.n_priority_parts = P.n_parts();
    const uint32_t n_parts = P.n_parts();
    std::vector<uint32_t> part_of(n_linked);
    for (uint32_t p = 0; p < n_parts; p++)
        for (uint32_t i = P.part_start[p]; i < P.part_start[p + 1]; i++) part_of[i] = p;

    // ---- 3. colour: local links per partition, global links over the whole graph (user order)
    // A link that finds no free local colour (its partition's 255-entry table is full: a point with more
    // than ~128 links) is scheduled with the cross-partition links instead, and a cross-partition link
    // that finds none of the 256 mask colours free gets an EXTRA colour: every vertex remembers the first
    // extra colour it has not used yet, and the link takes the larger of its two ends' (so the links at a
    // vertex get strictly increasing colours, which is all a colouring needs).  The reference relaxes any
    // graph (solver.rs:144-146); so does this schedule - a hub of degree d costs about d launches.
    std::vector<ColourMask> lmask(pp.reference_order ? 0 : n_linked), gmask;
    std::vector<uint32_t> gnext;  // per vertex: first extra global colour (>= 256) not used at it yet
    std::vector<uint32_t> colour(n_links);
    std::vector<uint8_t> is_global(n_links, 0);
    uint32_t C = 0, G = 0;
    size_t n_global = 0;
    (void)err;
    if (pp.reference_order) {
        // Dependency levels in insertion order (see PlanParams::reference_order).  Partitions are independent
        // of each other only while no link crosses them: then the levels are taken per partition and run as the
        // partition kernel's colours.  A crossing link, or a partition that needs more levels than the kernel's
        // colour table holds, makes every link a "global" one: one launch per level over the whole graph.
        std::vector<uint32_t> next(n_linked, 0u);
        bool local_ok = true;
        for (size_t k = 0; k < n_links && local_ok; k++) {
            const uint32_t a = P.rank[ab[2 * k]], b = P.rank[ab[2 * k + 1]];
            if (part_of[a] != part_of[b]) local_ok = false;
            const uint32_t lv = std::max(next[a], next[b]);
            if (lv >= kMaxLocalColours) local_ok = false;
            colour[k] = lv;
            next[a] = next[b] = lv + 1;
            C = std::max(C, lv + 1);
        }
        if (!local_ok) {
            std::fill(next.begin(), next.end(), 0u);
            C = 0;
            for (size_t k = 0; k < n_links; k++) {
                const uint32_t a = P.rank[ab[2 * k]], b = P.rank[ab[2 * k + 1]];
                const uint32_t lv = std::max(next[a], next[b]);
                colour[k] = lv;
                next[a] = next[b] = lv + 1;
                is_global[k] = 1;
                G = std::max(G, lv + 1);
            }
            n_global = n_links;
        }
    }
    for (size_t k = 0; k < n_links && !pp.reference_order; k++) {
        uint32_t a = P.rank[ab[2 * k]], b = P.rank[ab[2 * k + 1]];
        int c = -1;
        if (part_of[a] == part_of[b]) {
            c = lowest_free(lmask[a], lmask[b]);
            if (c >= (int)kMaxLocalColours) c = -1;
        }
        if (c >= 0) {
            mark(lmask[a], c), mark(lmask[b], c);
            colour[k] = (uint32_t)c;
            C = std::max(C, (uint32_t)c + 1);
        } else {
            if (gmask.empty()) gmask.resize(n_linked);
            uint32_t gc;
            c = lowest_free(gmask[a], gmask[b]);
            if (c >= 0) {
                mark(gmask[a], c), mark(gmask[b], c);
                gc = (uint32_t)c;
            } else {
                if (gnext.empty()) gnext.assign(n_linked, 256u);
                gc = std::max(gnext[a], gnext[b]);
                gnext[a] = gnext[b] = gc + 1;
            }
            colour[k] = gc;
            is_global[k] = 1;
            G = std::max(G, gc + 1);
            n_global++;
        }
    }
    P.n_local_colours = C;

    // ---- 4. bucket local links (partition-major, colour-major, stable in user order)
    const size_t stride = (size_t)C + 1;
    P.part_colour_start.assign((size_t)n_parts * stride, 0);
    {
        std::vector<uint32_t> count((size_t)n_parts * stride, 0);
        for (size_t k = 0; k < n_links; k++) {
            if (is_global[k]) continue;
            uint32_t p = part_of[P.rank[ab[2 * k]]];
            count[(size_t)p * stride + colour[k]]++;
        }
        uint32_t run = 0;
        for (uint32_t p = 0; p < n_parts; p++) {
            for (uint32_t c = 0; c < C; c++) {
                P.part_colour_start[(size_t)p * stride + c] = run;
                run += count[(size_t)p * stride + c];
            }
            P.part_colour_start[(size_t)p * stride + C] = run;
        }
        P.local_links.resize(run);
        P.local_user.resize(run);
        std::vector<uint32_t> cursor(P.part_colour_start);
        for (size_t k = 0; k < n_links; k++) {
            if (is_global[k]) continue;
            uint32_t a = P.rank[ab[2 * k]], b = P.rank[ab[2 * k + 1]];
            uint32_t p = part_of[a];
            uint32_t slot = cursor[(size_t)p * stride + colour[k]]++;
            uint32_t base = P.part_start[p];
            // keep the user's (a,b) orientation: link.rs:22 computes a - b
            P.local_links[slot] = LocalLink{(uint16_t)(a - base), (uint16_t)(b - base), len[k]};
            P.local_user[slot] = (uint32_t)k;
        }
    }
    // ---- 5. bucket global links by colour
    if (n_global) {
        P.gcolour_start.assign((size_t)G + 1, 0);
        for (size_t k = 0; k < n_links; k++)
            if (is_global[k]) P.gcolour_start[colour[k] + 1]++;
        for (uint32_t c = 0; c < G; c++) P.gcolour_start[c + 1] += P.gcolour_start[c];
        P.global_links.resize(n_global);
        P.global_user.resize(n_global);
        std::vector<uint32_t> cursor(P.gcolour_start.begin(), P.gcolour_start.end() - 1);
        for (size_t k = 0; k < n_links; k++) {
            if (!is_global[k]) continue;
            uint32_t slot = cursor[colour[k]]++;
            P.global_links[slot] = GlobalLink{P.rank[ab[2 * k]], P.rank[ab[2 * k + 1]], len[k]};
            P.global_user[slot] = (uint32_t)k;
        }
    }
    return true;
}

}  // namespace bendy
