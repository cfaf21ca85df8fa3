python tools/time_one.py fp64
MDPP_JIT_MINBLOCKS=7 python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_NOINLINE" python tools/time_one.py fp64
MDPP_JIT_MINBLOCKS=7 MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_NOINLINE" python tools/time_one.py fp64
MDPP_JIT_MINBLOCKS=7 python tools/time_one.py fast
MDPP_JIT_MINBLOCKS=6 python tools/time_one.py fp64
