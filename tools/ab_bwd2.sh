# same-box A/B: warp-level texel-gradient pre-reduction off / on (UMR_TEXGRAD_PRE) at C2, C3-like and C5-like shapes
for pre in 0 1; do
  echo "== UMR_TEXGRAD_PRE=$pre"
  UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py --iters 30 2>&1 | grep "kernel time" | cut -c1-150
  UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py --iters 10 --B 8 --is 1024 --subdiv 3 2>&1 | grep "kernel time" | cut -c1-150
  UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py --iters 10 --B 8 --is 1024 --subdiv 4 2>&1 | grep "kernel time" | cut -c1-150
  UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py --iters 20 --B 16 --is 256 --subdiv 3 --R 4 2>&1 | grep "kernel time" | cut -c1-150
  UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py --iters 20 --B 16 --is 512 --subdiv 3 2>&1 | grep "kernel time" | cut -c1-150
done
