for s in "-DMDPP_EXP_SKIP=4" "-DMDPP_EXP_SKIP=4 -DMDPP_EXP_NO_SLOW" "-DMDPP_EXP_SKIP=3" ; do
  echo "== pipe $s"
  MDPP_JIT_EXTRA="$s" python tools/time_one.py fp64 2>&1 | grep frac
done
echo "== nopipe skip4"; MDPP_ZIG_PIPE=0 MDPP_JIT_EXTRA="-DMDPP_EXP_SKIP=4" python tools/time_one.py fp64 2>&1 | grep frac
