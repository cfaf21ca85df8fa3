python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_EXP_NO_SLOW" python tools/time_one.py fp64
