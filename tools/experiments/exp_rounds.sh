echo "== base"; python tools/time_one.py fp64 2>&1 | grep frac
echo "== 7 rounds (timing only)"; MDPP_JIT_EXTRA="-DMDPP_EXP_PHILOX_ROUNDS=7" python tools/time_one.py fp64 2>&1 | grep frac
echo "== fast, 7 rounds"; MDPP_JIT_EXTRA="-DMDPP_EXP_PHILOX_ROUNDS=7" python tools/time_one.py fast 2>&1 | grep frac
