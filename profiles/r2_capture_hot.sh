# per-source-line capture of the two hot kernels (narrowphase, link kernel) on C3, free fall and piled up
mkdir -p gpurun_out
K='regex:k2_narrow|k3_links_local'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 50 -c 2 -f -o gpurun_out/r2_hot_c3_early \
    python profiles/ncu_c3_chain.py 3 1 > gpurun_out/r2_hot1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -s 434 -c 2 -f -o gpurun_out/r2_hot_c3_late \
    python profiles/ncu_c3_chain.py 27 1 > gpurun_out/r2_hot2.log 2>&1
tail -2 gpurun_out/r2_hot*.log; ls -la gpurun_out/*.ncu-rep
