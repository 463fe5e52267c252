# Developer A/B: quick_bench of variant libraries (python -m jacks_b200.build --variant=<v> -D...) at two line counts.
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-base}; do
  for L in ${LINES:-400}; do
    if [ "$v" = base ]; then lib=jacks_b200/libjacks_b200.so; else lib=jacks_b200/libjacks_b200_$v.so; fi
    echo "== variant '$v' L=$L"
    JACKS_B200_LIB=$PWD/$lib QB_SHAPES=default QB_LINES=$L timeout 300 python tools/quick_bench.py 2>&1 | grep -v "^fp64"
  done
done
