echo "== pipe rare_slow"; MDPP_JIT_EXTRA="-DMDPP_EXP_RARE_SLOW" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe rare_slow 1 phase"; MDPP_JIT_EXTRA="-DMDPP_EXP_RARE_SLOW -DMDPP_PIPE_PHASES=1" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe no_slow"; MDPP_JIT_EXTRA="-DMDPP_EXP_NO_SLOW" python tools/time_one.py fp64 2>&1 | grep frac
