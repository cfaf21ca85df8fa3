python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_UNROLL=8" python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_UNROLL=2" python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_UNROLL=16" python tools/time_one.py fp64
python tools/time_one.py fast
