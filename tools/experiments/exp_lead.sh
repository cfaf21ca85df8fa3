echo "== base"; python tools/time_one.py fp64 2>&1 | grep frac
for g in 2 4; do echo "== lead groups $g"; MDPP_JIT_EXTRA="-DMDPP_LEAD_GROUPS=$g" python tools/time_one.py fp64 2>&1 | grep frac; done
