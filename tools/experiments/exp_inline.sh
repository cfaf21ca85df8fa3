for d in "-DMDPP_EXP_NO_STAGE" "-DMDPP_EXP_NO_STAGE -DMDPP_EXP_NO_SLOW"; do
  echo "== $d"
  MDPP_JIT_EXTRA="$d" python tools/time_one.py fp64 2>&1 | grep frac
  MDPP_JIT_CHUNK=4 MDPP_JIT_EXTRA="$d" python tools/time_one.py fp64 2>&1 | grep frac
done
