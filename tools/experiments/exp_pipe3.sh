for o in 0 1 2; do for s in "" "-DMDPP_EXP_NO_SLOW"; do
  echo "== order $o $s"
  MDPP_JIT_EXTRA="-DMDPP_PIPE_ORDER=$o $s" python tools/time_one.py fp64 2>&1 | grep frac
done; done
