python -m pytest tests/test_cuda_discrete.py tests/test_cuda_hetero.py -q -x -k "pipelined or long_launch or bench_signature" 2>&1 | tail -5
python tools/time_one.py fp64 2>&1 | grep frac
MDPP_ZIG_PIPE=0 python tools/time_one.py fp64 2>&1 | grep frac
python tools/time_hetero.py fp64 1000 2>&1 | grep frac
MDPP_ZIG_PIPE=0 python tools/time_hetero.py fp64 1000 2>&1 | grep frac
MDPP_ZIG_PIPE=1 python tools/time_hetero.py fp64 100 2>&1 | grep frac
