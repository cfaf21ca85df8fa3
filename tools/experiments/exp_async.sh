echo "== base"; python tools/time_one.py fp64 2>&1 | grep frac
for n in 2 3; do echo "== cp.async buffers $n"; MDPP_JIT_EXTRA="-DMDPP_ACT_ASYNC=$n" python tools/time_one.py fp64 2>&1 | grep frac; done
echo "== fast base"; python tools/time_one.py fast 2>&1 | grep frac
echo "== fast cp.async 2"; MDPP_JIT_EXTRA="-DMDPP_ACT_ASYNC=2" python tools/time_one.py fast 2>&1 | grep frac
MDPP_JIT_EXTRA="-DMDPP_ACT_ASYNC=2" python -m pytest tests/test_cuda_discrete.py -q -x -k "long_launch" 2>&1 | tail -2
