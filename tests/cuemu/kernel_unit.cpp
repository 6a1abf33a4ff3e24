// kernel_unit.cpp (cuemu) — TEST INFRASTRUCTURE: kernel-level check that needs no solver around it.
// Includes the REWRITTEN product kernels (tests/cuemu/_build/.../kernels.cuh, made by build.py).
// The counting sort of the broadphase with the ONE-PASS scan: k2_count builds the histogram AND the scan-tile
// totals (per-CTA shared-memory table, runs of equal tiles; discs spread over more scan tiles than the table
// holds go to the global totals directly), k2_scan places every cell from those totals without any inter-CTA
// waiting, k2_scatter sorts and re-zeroes the totals.  Checked cell by cell for a grid of 4 and of 40 scan tiles
// (the emulated device holds 8 CTAs at a time: a grid larger than the device must work too - there is no barrier).
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "kernels.cuh"

using namespace bendy;

#define CHECK(c)                                                         \
    do {                                                                 \
        if (!(c)) {                                                      \
            printf("kernel_unit FAILED at line %d: %s\n", __LINE__, #c); \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main() {
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int tiles : {4, 40}) {
        const uint32_t n_cells = (uint32_t)tiles * SCAN_TILE, n = 5000;
        StepParams hp{};
        hp.gox = 0.f, hp.goy = 0.f, hp.h = 1.f, hp.inv_h = 1.f, hp.nx = 2048, hp.ny = tiles;
        hp.halo_xl = -INFINITY, hp.halo_xr = INFINITY, hp.stray_xl = -INFINITY, hp.stray_xr = INFINITY;
        std::vector<float2> hpos(n);
        std::vector<uint32_t> hcount(n_cells, 0), htile(tiles, 0);
        for (uint32_t i = 0; i < n; i++) {
            // consecutive discs jump between cell rows = scan tiles: a CTA's 256 discs cover up to 40 tiles (table: 32)
            hpos[i] = make_float2((float)((i * 7919u) % 2048u) + 0.5f, (float)((i * 31u) % (uint32_t)tiles) + 0.5f);
            const uint32_t c = (uint32_t)hpos[i].y * 2048u + (uint32_t)hpos[i].x;
            hcount[c]++, htile[c >> SCAN_TILE_SHIFT]++;
        }
        hpos[17] = make_float2(NAN, 3.0f);  // a non-finite disc is in no cell
        {
            const uint32_t c = (uint32_t)(((17u * 31u) % (uint32_t)tiles)) * 2048u + (uint32_t)((17u * 7919u) % 2048u);
            hcount[c]--, htile[c >> SCAN_TILE_SHIFT]--;
        }
        uint32_t *count, *tile_sum, *cell_start, *slot_of, *sorted_id;
        float2 *pos, *sorted_pos;
        StepParams *prm;
        cudaMalloc(&count, n_cells * 4), cudaMalloc(&tile_sum, tiles * 4), cudaMalloc(&cell_start, n_cells * 4);
        cudaMalloc(&slot_of, n * 4), cudaMalloc(&sorted_id, n * 4), cudaMalloc(&pos, n * 8);
        cudaMalloc(&sorted_pos, n * 8), cudaMalloc(&prm, sizeof hp);
        cudaMemcpy(pos, hpos.data(), n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(prm, &hp, sizeof hp, cudaMemcpyHostToDevice);
        cudaMemsetAsync(count, 0, n_cells * 4, st), cudaMemsetAsync(tile_sum, 0, tiles * 4, st);
        K3CountArgs ca{};
        ca.prm = prm, ca.n_cells = n_cells, ca.cell_count = count, ca.tile_sum = tile_sum;
        for (int round = 0; round < 2; round++) {  // twice: the second substep finds the counters re-zeroed
            cuemu::Launcher("k2_count", (n + 255) / 256, 256, 0, st).run(k2_count<false>, (const float2 *)pos, (const float *)nullptr, 0u, n, ca);
            cudaStreamSynchronize(st);
            std::vector<uint32_t> gt(tiles);
            cudaMemcpy(gt.data(), tile_sum, tiles * 4, cudaMemcpyDeviceToHost);
            for (int t = 0; t < tiles; t++) CHECK(gt[t] == htile[t]);
            cuemu::Launcher("k2_scan", tiles, SCAN_THREADS, 0, st).run(k2_scan, count, (const uint32_t *)tile_sum, cell_start);
            cuemu::Launcher("k2_scatter", (n + 255) / 256, 256, 0, st)
                .run(k2_scatter<true>, (const float2 *)pos, n, (const StepParams *)prm, n_cells, cell_start, tile_sum, (uint32_t)tiles, sorted_pos, slot_of, sorted_id);
            cudaStreamSynchronize(st);
            std::vector<uint32_t> hslot(n), hid(n), hend(n_cells), hcnt(n_cells);
            std::vector<float2> hsorted(n);
            cudaMemcpy(hslot.data(), slot_of, n * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hid.data(), sorted_id, n * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hend.data(), cell_start, n_cells * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hcnt.data(), count, n_cells * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hsorted.data(), sorted_pos, n * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(gt.data(), tile_sum, tiles * 4, cudaMemcpyDeviceToHost);
            uint32_t run = 0;
            for (uint32_t c = 0; c < n_cells; c++) {  // after the scatter cell_start[c] is the END of cell c
                run += hcount[c];
                CHECK(hend[c] == run);
                CHECK(hcnt[c] == 0);  // histogram re-zeroed for the next substep
            }
            for (int t = 0; t < tiles; t++) CHECK(gt[t] == 0);  // and so are the scan-tile totals
            std::vector<int> seen(n, 0);
            for (uint32_t i = 0; i < n; i++) {
                if (i == 17) {
                    CHECK(hslot[i] == NO_CELL);
                    continue;
                }
                const uint32_t c = (uint32_t)hpos[i].y * 2048u + (uint32_t)hpos[i].x;
                CHECK(hslot[i] < hend[c] && hslot[i] >= hend[c] - hcount[c]);
                CHECK(hid[hslot[i]] == i && hsorted[hslot[i]].x == hpos[i].x && hsorted[hslot[i]].y == hpos[i].y);
                seen[hslot[i]]++;
            }
            for (uint32_t i = 0; i + 1 < n; i++) CHECK(seen[i] == 1);
        }
        for (void *p : {(void *)count, (void *)tile_sum, (void *)cell_start, (void *)slot_of, (void *)sorted_id, (void *)pos,
                        (void *)sorted_pos, (void *)prm})
            cudaFree(p);
    }
    printf("kernel_unit ok\n");
    return 0;
}
