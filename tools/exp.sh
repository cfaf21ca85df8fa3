for n in fp64 fast; do
python tools/time_one.py $n
MDPP_JIT_EXTRA="-DMDPP_EXP_SKIP=4" python tools/time_one.py $n
MDPP_JIT_EXTRA="-DMDPP_EXP_SKIP=7" python tools/time_one.py $n
done
