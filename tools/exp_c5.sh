# C5 experiments: dynamic last-window fill, window size, steps per launch
python tools/time_hetero.py fp64 100
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_DYN" python tools/time_hetero.py fp64 100
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_DYN" MDPP_ZIG_WINDOW=32 python tools/time_hetero.py fp64 100
MDPP_ZIG_WINDOW=32 python tools/time_hetero.py fp64 100
python tools/time_hetero.py fp64 96
python tools/time_hetero.py fp64 128
python tools/time_hetero.py fp64 1000
MDPP_ZIG_WINDOW=32 python tools/time_hetero.py fp64 1000
python tools/time_hetero.py fast 1000
python tools/time_one.py fp64
MDPP_JIT_EXTRA="-DMDPP_ZIG_FILL_DYN" python tools/time_one.py fp64
