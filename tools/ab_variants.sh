# same-box A/B of build variants (python -m umr_b200.build --variant NAME -DFLAG ...): kernel times from the library's events
for n in "" _VARIANT; do   # replace _VARIANT by the --variant name(s) built
  echo "== variant ${n:-default}"
  UMR_B200_LIB=/root/repo/umr_b200/libumr_b200$n.so timeout 120 python tools/quick_bench.py --iters 40 2>&1 | grep "kernel time" | cut -c1-60
  UMR_B200_LIB=/root/repo/umr_b200/libumr_b200$n.so timeout 120 python tools/visibility_bench.py 2>&1 | head -1
done
