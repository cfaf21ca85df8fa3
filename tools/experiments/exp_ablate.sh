# where does the headline's wall time go?  one class of work compiled out at a time
for d in "" "-DMDPP_EXP_NO_SLOW" "-DMDPP_EXP_NO_FILL" "-DMDPP_EXP_NO_CHAIN_PHILOX" "-DMDPP_EXP_NO_STATS" "-DMDPP_EXP_SKIP=3" "-DMDPP_EXP_SKIP=7" \
  "-DMDPP_EXP_NO_FILL -DMDPP_EXP_NO_CHAIN_PHILOX" "-DMDPP_EXP_NO_FILL -DMDPP_EXP_NO_CHAIN_PHILOX -DMDPP_EXP_NO_STATS -DMDPP_EXP_SKIP=7"; do
  echo "== $d"
  MDPP_JIT_EXTRA="$d" python tools/time_one.py fp64 2>&1 | grep frac
done
echo "== fast"; python tools/time_one.py fast 2>&1 | grep frac
echo "== fast no chain philox"; MDPP_JIT_EXTRA="-DMDPP_EXP_NO_CHAIN_PHILOX" python tools/time_one.py fast 2>&1 | grep frac
