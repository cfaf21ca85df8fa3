python tools/time_one.py fp64
python tools/time_hetero.py fp64
MDPP_ZIG_WINDOW=32 python tools/time_hetero.py fp64
python tools/time_one.py fp64 1048576 100 5
MDPP_ZIG_WINDOW=16 python tools/time_one.py fp64 1048576 100 5
