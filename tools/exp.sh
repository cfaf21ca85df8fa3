python tools/time_hetero.py fp64
MDPP_JIT_CHUNK=4 python tools/time_hetero.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_UNROLL=2" python tools/time_hetero.py fp64
MDPP_JIT_CHUNK=4 MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_UNROLL=2" python tools/time_hetero.py fp64
MDPP_JIT_CHUNK=4 python tools/time_hetero.py fast
python tools/time_hetero.py fast
