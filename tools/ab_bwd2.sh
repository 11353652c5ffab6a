# same-box A/B: warp-level texel pre-reduction forced off / on (UMR_TEXGRAD_PRE) with the vector-RED build
for pre in 0 1; do
  echo "== UMR_TEXGRAD_PRE=$pre"
  for args in "--iters 30" "--iters 10 --B 8 --is 1024 --subdiv 3" "--iters 10 --B 8 --is 1024 --subdiv 4" "--iters 20 --B 16 --is 512 --subdiv 3" "--iters 10 --B 4 --is 1024 --subdiv 2"; do
    UMR_TEXGRAD_PRE=$pre timeout 120 python tools/quick_bench.py $args 2>&1 | grep "kernel time" | sed 's/kernel time //; s/backward//g; s/(detached geometry)//; s/(constant textures)//'
  done
done
