// kernel_unit.cpp (cuemu) — TEST INFRASTRUCTURE: kernel-level checks that need no solver around them.
// Includes the REWRITTEN product kernels (tests/cuemu/_build/.../kernels.cuh, made by build.py).
//   1. a software grid barrier whose grid is not resident must time out, raise FLAG_GRID_BARRIER_TIMEOUT
//      and let every CTA leave - not hang (k2_scan_scatter_fused, k2_scan_fused_mt);
//   2. the same kernels on a resident grid produce the exclusive scan / a valid counting sort.
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
    for (int tiles : {4, 40}) {  // 4 CTAs fit the 8-CTA device; 40 do not
        const uint32_t n_cells = (uint32_t)tiles * SCAN_TILE, n = 5000;
        StepParams hp{};
        hp.gox = 0.f, hp.goy = 0.f, hp.h = 1.f, hp.inv_h = 1.f, hp.nx = 2048, hp.ny = tiles;
        std::vector<float2> hpos(n);
        std::vector<uint32_t> hcount(n_cells, 0);
        for (uint32_t i = 0; i < n; i++) {
            hpos[i] = make_float2((float)((i * 7919u) % 2048u) + 0.5f, (float)((i * 31u) % (uint32_t)tiles) + 0.5f);
            hcount[(uint32_t)hpos[i].y * 2048u + (uint32_t)hpos[i].x]++;
        }
        uint32_t *count, *tile_sum, *cell_start, *bar, *slot_of, *sorted_id;
        float2 *pos, *sorted_pos;
        StepParams *prm;
        int *flags;
        cudaMalloc(&count, n_cells * 4), cudaMalloc(&tile_sum, tiles * 4), cudaMalloc(&cell_start, n_cells * 4);
        cudaMalloc(&bar, 16), cudaMalloc(&slot_of, n * 4), cudaMalloc(&sorted_id, n * 4), cudaMalloc(&pos, n * 8);
        cudaMalloc(&sorted_pos, n * 8), cudaMalloc(&prm, sizeof hp), cudaMalloc(&flags, 32);
        cudaMemcpy(count, hcount.data(), n_cells * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(pos, hpos.data(), n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(prm, &hp, sizeof hp, cudaMemcpyHostToDevice);
        cudaMemsetAsync(bar, 0, 16, st), cudaMemsetAsync(flags, 0, 32, st), cudaMemsetAsync(tile_sum, 0, tiles * 4, st);
        cuemu::Launcher("k2_scan_scatter_fused", tiles, SCAN_THREADS, 0, st)
            .run(k2_scan_scatter_fused<true>, count, tile_sum, cell_start, bar, flags, pos, n, prm, n_cells, sorted_pos, slot_of, sorted_id);
        cudaStreamSynchronize(st);
        int hflags = 0;
        cudaMemcpy(&hflags, flags, 4, cudaMemcpyDeviceToHost);
        if (tiles > 8) {
            CHECK(hflags & FLAG_GRID_BARRIER_TIMEOUT);  // and we got here: nobody hung
        } else {
            CHECK(hflags == 0);
            std::vector<uint32_t> hslot(n), hid(n), hend(n_cells), hcnt(n_cells);
            std::vector<float2> hsorted(n);
            cudaMemcpy(hslot.data(), slot_of, n * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hid.data(), sorted_id, n * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hend.data(), cell_start, n_cells * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hcnt.data(), count, n_cells * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hsorted.data(), sorted_pos, n * 8, cudaMemcpyDeviceToHost);
            uint32_t run = 0;
            for (uint32_t c = 0; c < n_cells; c++) {  // after the scatter cell_start[c] is the END of cell c
                run += hcount[c];
                CHECK(hend[c] == run);
                CHECK(hcnt[c] == 0);  // histogram re-zeroed for the next substep
            }
            std::vector<int> seen(n, 0);
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t c = (uint32_t)hpos[i].y * 2048u + (uint32_t)hpos[i].x;
                CHECK(hslot[i] < hend[c] && hslot[i] >= hend[c] - hcount[c]);
                CHECK(hid[hslot[i]] == i && hsorted[hslot[i]].x == hpos[i].x && hsorted[hslot[i]].y == hpos[i].y);
                seen[hslot[i]]++;
            }
            for (uint32_t i = 0; i < n; i++) CHECK(seen[i] == 1);
        }
        // the multi-tile scan on the same histogram (2 tiles per CTA): 2 CTAs fit, 20 do not
        cudaMemcpy(count, hcount.data(), n_cells * 4, cudaMemcpyHostToDevice);
        cudaMemsetAsync(bar, 0, 16, st), cudaMemsetAsync(flags, 0, 32, st);
        cuemu::Launcher("k2_scan_fused_mt", tiles / 2, SCAN_THREADS, 0, st).run(k2_scan_fused_mt<2>, count, tile_sum, cell_start, bar, flags);
        cudaStreamSynchronize(st);
        cudaMemcpy(&hflags, flags, 4, cudaMemcpyDeviceToHost);
        if (tiles / 2 > 8) {
            CHECK(hflags & FLAG_GRID_BARRIER_TIMEOUT);
        } else {
            CHECK(hflags == 0);
            std::vector<uint32_t> hstart(n_cells);
            cudaMemcpy(hstart.data(), cell_start, n_cells * 4, cudaMemcpyDeviceToHost);
            uint32_t run = 0;
            for (uint32_t c = 0; c < n_cells; c++) {
                CHECK(hstart[c] == run);
                run += hcount[c];
            }
        }
        for (void *p : {(void *)count, (void *)tile_sum, (void *)cell_start, (void *)bar, (void *)slot_of, (void *)sorted_id,
                        (void *)pos, (void *)sorted_pos, (void *)prm, (void *)flags})
            cudaFree(p);
    }
    printf("kernel_unit ok\n");
    return 0;
}
