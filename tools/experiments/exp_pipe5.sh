echo "== pipe rej_bits"; MDPP_JIT_EXTRA="-DMDPP_ZIG_REJ_BITS" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe rej_bits chunk4"; MDPP_JIT_CHUNK=4 MDPP_JIT_EXTRA="-DMDPP_ZIG_REJ_BITS" python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe chunk4"; MDPP_JIT_CHUNK=4 python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe window16"; MDPP_ZIG_WINDOW=16 python tools/time_one.py fp64 2>&1 | grep frac
echo "== pipe rej_bits window16"; MDPP_ZIG_WINDOW=16 MDPP_JIT_EXTRA="-DMDPP_ZIG_REJ_BITS" python tools/time_one.py fp64 2>&1 | grep frac
