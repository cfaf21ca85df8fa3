MDPP_ZIG_PIPE=0 python tools/time_one.py fp64 2>&1 | grep frac
python tools/time_one.py fp64 2>&1 | grep frac
MDPP_ZIG_PIPE=0 python tools/time_one.py fp64 2>&1 | grep frac
python tools/time_one.py fp64 2>&1 | grep frac
MDPP_JIT_EXTRA="-DMDPP_EXP_NO_SLOW" python tools/time_one.py fp64 2>&1 | grep frac
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
