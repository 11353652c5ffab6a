# same-box A/B of build variants (python -m umr_b200.build --variant NAME -DFLAG ...): kernel times from the library's events
for n in "" _f5; do
  echo "== variant ${n:-default}"
  for args in "--iters 30" "--iters 10 --B 8 --is 1024 --subdiv 4"; do
    UMR_B200_LIB=/root/repo/umr_b200/libumr_b200$n.so timeout 120 python tools/quick_bench.py $args 2>&1 | grep "kernel time" | cut -c1-75
  done
done
