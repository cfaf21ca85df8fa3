for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
echo "variant $v"
MDPP_RENDER_VARIANT=$v python tools/time_render.py 2>&1 | tail -${TAIL:-2}
done
for v in ${TESTV:-7}; do
MDPP_RENDER_VARIANT=$v python -m pytest tests/test_cuda_images.py -m gpu -q 2>&1 | tail -5
done
