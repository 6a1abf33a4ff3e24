// plan.h — host-side link schedule: partitioning into shared-memory blocks + graph colouring.
//
// Replaces the reference's sequential Gauss-Seidel walk over `particle_links` in insertion order
// (solver.rs:143-146, link.rs:18-27) by an arithmetically identical parallel schedule:
//   * points are renumbered so that each partition is one contiguous index range small enough to
//     live in one CTA's shared memory; links with both ends in one partition are "local";
//   * inside a partition local links are greedily edge-coloured: links of one colour share no
//     vertex, so they can be relaxed concurrently; colours run in order with a CTA barrier;
//   * links across partitions are "global", coloured over the whole graph, one launch per colour.
// The sequential order that reproduces the schedule bit for bit is
//   [partition 0: colour 0.., colour 1.., ...][partition 1 ...] ... [global colour 0][global colour 1]...
// and is exported as `perm` (user link indices) for the oracle replay.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace bendy {

struct alignas(8) LocalLink {  // 8 B record streamed by the partition kernel (one 64-bit load)
    uint16_t a, b;  // partition-local point indices
    float len;
};
struct GlobalLink {  // 12 B record of a cross-partition link (internal point indices)
    uint32_t a, b;
    float len;
};

// the partition kernel keeps the colour offsets of its partition in a 256-entry shared-memory table
constexpr uint32_t kMaxLocalColours = 255;

struct PlanParams {
    uint32_t pack_points = 512;  // keep adding whole components to a partition up to this many points
    uint32_t max_points = 4096;  // hard cap (shared memory: 8 B per point)
    // false: greedy edge colouring (few colours; results equal the reference fed the links in the exported
    //        colour order, the contract of north_star);
    // true:  REFERENCE ORDER - colours are dependency levels in insertion order: level = max(next[a], next[b]),
    //        next[a] = next[b] = level + 1.  Every link then runs after all EARLIER links that share a vertex
    //        with it and links of one level are vertex-disjoint, so the level-parallel schedule is arithmetically
    //        identical to the reference's sequential walk in insertion order (solver.rs:144-146) - at the price
    //        of more colours (a 20x20 lattice: ~60 levels instead of 8 colours).
    bool reference_order = false;
};

struct LinkPlan {
    size_t n_points = 0;
    std::vector<uint32_t> rank;              // user point -> internal index
    std::vector<uint32_t> order;             // internal index -> user point
    std::vector<uint32_t> part_start;        // n_parts+1 internal ranges (linked points only)
    uint32_t n_local_colours = 0;            // C = max colours in a partition
    std::vector<uint32_t> part_colour_start; // n_parts*(C+1) offsets into local_links
    std::vector<LocalLink> local_links;
    std::vector<uint32_t> local_user;        // user link index of each local record
    std::vector<uint32_t> gcolour_start;     // G+1 offsets into global_links
    std::vector<GlobalLink> global_links;
    std::vector<uint32_t> global_user;
    uint32_t n_priority_parts = 0;           // leading partitions made of components with a priority point
    uint32_t n_parts() const { return part_start.empty() ? 0u : (uint32_t)part_start.size() - 1; }
    uint32_t n_global_colours() const { return gcolour_start.empty() ? 0u : (uint32_t)gcolour_start.size() - 1; }
    // sequential-equivalent order of user link indices
    std::vector<uint32_t> perm() const;
};

// keep_order = true: points keep their indices (rank = identity); partitions are consecutive index
// ranges cut at max_points (used for polygon points, whose order is part of the polygon's shape).
// Returns false and sets err on failure (vertex degree too high for the colour masks).
bool plan_links(size_t n_points, const uint32_t *ab, const float *len, size_t n_links, const PlanParams &pp,
                bool keep_order, LinkPlan *out, std::string *err, const uint8_t *priority = nullptr);

}  // namespace bendy
