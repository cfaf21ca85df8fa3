# compile-option lottery on the fp64 headline (the schedule is what is tight)
for f in "--restrict" "--extra-device-vectorization" "--maxrregcount=120" "--maxrregcount=112" "--ptxas-options=--allow-expensive-optimizations=true" "--ptxas-options=-O4"; do
  echo "== $f"; MDPP_JIT_EXTRA="$f" python tools/time_one.py fp64 2>&1 | grep -E "frac|rror" | head -2
done
