# what the driver runs at round end (same flags), plus one ncu --set full capture of the C3 step's kernels
set -x
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | head -c 400; echo
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_driver_flags.json; head -c 300 gpurun_out/r02_bench_driver_flags.json; echo
timeout 300 python bench.py --config C3 --no-cpu-baseline --no-other-kernels --no-reference-gpu 2>/dev/null | tail -1 > gpurun_out/r02_bench_C3.json; head -c 260 gpurun_out/r02_bench_C3.json; echo
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_raster_bwd2|k_visible_faces|k_corr_fwd|k_corr_bwd|k_raster_fwd3" -c 5 -o gpurun_out/r02_C3_step -f python bench.py --config C3 --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-other-kernels --no-reference-gpu > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_C3_step.ncu-rep > gpurun_out/r02_raster_C3_ncu_summary.txt 2>&1; grep -c kernel: gpurun_out/r02_raster_C3_ncu_summary.txt
rm -f gpurun_out/r02_C3_step.ncu-rep
