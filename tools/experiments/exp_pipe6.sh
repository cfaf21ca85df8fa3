python -m pytest tests/test_cuda_discrete.py tests/test_cuda_hetero.py -q -x -k "pipelined or long_launch" 2>&1 | tail -3
echo "== pipe, 4 phases"; python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe, 1 phase"; MDPP_JIT_EXTRA="-DMDPP_PIPE_PHASES=1" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe, 2 phases"; MDPP_JIT_EXTRA="-DMDPP_PIPE_PHASES=2" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe off"; MDPP_ZIG_PIPE=0 python tools/time_one.py fp64 2>&1 | grep frac
echo "== C5 pipe"; python tools/time_hetero.py fp64 1000 2>&1 | grep frac
