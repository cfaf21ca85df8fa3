echo "== old (HEAD)"; (cd _oldpkg && python ../tools/time_one.py fp64 2>&1 | grep frac)
echo "== tree"; python tools/time_one.py fp64 2>&1 | grep frac
echo "== tree hetero 100"; python tools/time_hetero.py fp64 100 2>&1 | grep frac
