set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; tail -c 600 gpurun_out/r02_bench_C2.json
timeout 400 python bench.py --config C3 --no-cpu-baseline > gpurun_out/r02_bench_C3.json 2> gpurun_out/r02_bench_C3.err; head -c 300 gpurun_out/r02_bench_C3.json
timeout 400 python bench.py --config C5 --no-cpu-baseline > gpurun_out/r02_bench_C5.json 2> gpurun_out/r02_bench_C5.err; head -c 300 gpurun_out/r02_bench_C5.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_C3_step.csv python bench.py --config C3 --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-other-kernels --no-reference-gpu > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r02_launches_C3_step.csv > gpurun_out/r02_launches_C3_step.txt; tail -25 gpurun_out/r02_launches_C3_step.txt
