# C3 graph substep for a few broadphase cell sizes (r_p = 0.1): auto = 4.2 r_p with the 2x2-cell candidate block,
# below that 3x3 blocks of smaller cells (fewer candidates per disc, more cells to scan)
mkdir -p gpurun_out
for h in "" 0.2 0.21 0.256 0.3; do
  QUICK_C3_CELL=$h timeout 200 python profiles/quick_c3.py "cell=${h:-auto}" 2>&1 | tail -1
done | tee gpurun_out/r2_cell_sweep.txt
