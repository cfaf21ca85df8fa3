#!/bin/bash
# Offline SASS of the BASELINE configs[2] (C3: D=6, order 2, dense reward, no
# noise, auto-reset) specialisation of the continuous rollout kernel.
#   tools/jit_sass_continuous.sh [out_dir] [extra -D ...]
set -e
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/sass}; shift || true
mkdir -p "$OUT"
cat > "$OUT/centry.cu" <<'SRC'
#include "continuous_kernels.cuh"
extern "C" __global__ void __launch_bounds__(mdpp::kCBlock)
mdpp_jit_continuous(const __grid_constant__ mdpp::ContinuousParams p) {
  mdpp::continuous_body<MDPP_C_REAL, MDPP_C_NOISE>(p);
}
SRC
bits() { python3 -c "import struct,sys; print('0x%xll' % struct.unpack('<Q', struct.pack('<d', float(sys.argv[1])))[0])" "$1"; }
DEFS="-DMDPP_JIT -DMDPP_C_REAL=float -DMDPP_C_NOISE=2 -DMDPP_C_DIM=6 -DMDPP_C_ORDER=2 -DMDPP_C_NREL=2
 -DMDPP_C_DELAY=0 -DMDPP_C_EVERY_N=1 -DMDPP_C_DENSE=true -DMDPP_C_PNOISE=false -DMDPP_C_RNOISE=false
 -DMDPP_C_IMAGE=false -DMDPP_C_TARGET64=false -DMDPP_C_NORMAL=0 -DMDPP_C_IMODE=0 -DMDPP_C_LINE=false -DMDPP_C_SEQ=1 -DMDPP_C_NBOX=0 -DMDPP_C_HORIZON=100
 -DMDPP_C_AUTORESET=true -DMDPP_C_N_ENVS=1048576ll -DMDPP_C_FAST=true -DMDPP_C_REL0=0 -DMDPP_C_REL1=1"
for k in 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do DEFS="$DEFS -DMDPP_C_REL$k=0"; done
DEFS="$DEFS -DMDPP_C_AMAX_BITS=$(bits 1.0) -DMDPP_C_SMAX_BITS=$(bits 10.0) -DMDPP_C_INERTIA_BITS=$(bits 1.0)
 -DMDPP_C_RADIUS_BITS=$(bits 0.05) -DMDPP_C_ALW_BITS=$(bits 0.0) -DMDPP_C_SCALE_BITS=$(bits 1.0)
 -DMDPP_C_SHIFT_BITS=$(bits 0.0) -DMDPP_C_TERM_ADD_BITS=$(bits 0.0) -DMDPP_C_P_STD_BITS=$(bits 0.0)
 -DMDPP_C_R_STD_BITS=$(bits 0.0) -DMDPP_C_TU1_BITS=$(bits 0.5) -DMDPP_C_TU2_BITS=$(bits 0.25)
 -DMDPP_C_TU3_BITS=$(bits 0.125) -DMDPP_C_TU4_BITS=$(bits 0.0625)"
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v \
  -I mdp_playground_b200/csrc -I include $DEFS "$@" -cubin -o "$OUT/cjit.cubin" "$OUT/centry.cu"
cuobjdump -sass "$OUT/cjit.cubin" > "$OUT/cjit.sass"
grep -cE '^\s+/\*[0-9a-f]{4}\*/' "$OUT/cjit.sass" | sed 's/^/sass instructions: /'
