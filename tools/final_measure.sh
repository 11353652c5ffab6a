# round-end measurement pass on one B200 (gpurun): GPU suite, smoke, bench lines of the three BASELINE configs, launch lists
set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; head -c 250 gpurun_out/r02_bench_C2.json; echo
timeout 400 python bench.py --config C3 --no-cpu-baseline > gpurun_out/r02_bench_C3.json 2> gpurun_out/r02_bench_C3.err; head -c 250 gpurun_out/r02_bench_C3.json; echo
timeout 400 python bench.py --config C5 --no-cpu-baseline > gpurun_out/r02_bench_C5.json 2> gpurun_out/r02_bench_C5.err; head -c 250 gpurun_out/r02_bench_C5.json; echo
for c in C2 C3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${c}_step.csv python bench.py --config $c --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-other-kernels --no-reference-gpu > /dev/null 2>&1
  python tools/launch_list.py gpurun_out/r02_launches_${c}_step.csv > gpurun_out/r02_launches_${c}_step.txt; grep -A14 "share by kernel" gpurun_out/r02_launches_${c}_step.txt
done
timeout 300 python tools/train_step_bench.py 2>&1 | tail -4 | tee gpurun_out/r02_train_step.txt
