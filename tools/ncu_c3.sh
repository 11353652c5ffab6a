timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_raster_bwd2|k_visible_faces|k_corr_fwd|k_corr_bwd" -c 4 -o gpurun_out/r02_C3_step -f python bench.py --config C3 --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-other-kernels --no-reference-gpu > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_C3_step.ncu-rep > gpurun_out/r02_raster_C3_ncu_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_C3_step.ncu-rep k_visible_faces 2>/dev/null | head -30 >> gpurun_out/r02_raster_C3_ncu_summary.txt
grep "kernel:\|time_duration" gpurun_out/r02_raster_C3_ncu_summary.txt | cut -c1-120
rm -f gpurun_out/r02_C3_step.ncu-rep
