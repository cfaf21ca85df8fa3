python tools/time_one.py fp64
python tools/time_one.py fp64 1048576 100 5
python tools/time_one.py fast 1048576 100 5
