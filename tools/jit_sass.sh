#!/bin/bash
# Offline SASS of the headline (BASELINE configs[1]) specialisation of the
# discrete rollout kernel: the same source and -D list jit.cu hands to NVRTC,
# compiled with nvcc so that cuobjdump / nvdisasm can be used without a GPU.
#   tools/jit_sass.sh [out_dir] [extra -D ...]
set -e
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/sass}; shift || true
mkdir -p "$OUT"
cat > "$OUT/entry.cu" <<'SRC'
#include "discrete_kernels.cuh"
using JitCfg = mdpp::Cfg<MDPP_CFG_NOISE, MDPP_CFG_NORMAL, MDPP_CFG_SMEM, MDPP_CFG_RING,
                         MDPP_CFG_FAST, MDPP_CFG_CDF, MDPP_CFG_SINGLE,
                         MDPP_CFG_RING_REGS>;
extern "C" __global__ void __launch_bounds__(mdpp::kBlock, mdpp::kMinBlocksPerSM)
mdpp_jit_rollout(const __grid_constant__ mdpp::RolloutParams p) {
  mdpp::rollout_body<JitCfg>(p);
}
SRC
DEFS="-DMDPP_JIT -DMDPP_S=8 -DMDPP_A=8 -DMDPP_L=3 -DMDPP_DELAY=2 -DMDPP_EVERY_N=1
 -DMDPP_LOOKUP=0 -DMDPP_KEY_BITS=3 -DMDPP_HASH_SHIFT=0 -DMDPP_HASH_MASK=0u -DMDPP_KEY_MASK=511ull
 -DMDPP_PNOISE=true -DMDPP_RNOISE=true -DMDPP_CDF_LOG2=3 -DMDPP_HAS_GUIDE=true -DMDPP_P8=true
 -DMDPP_R_STD=0.25 -DMDPP_SCALE=1.0 -DMDPP_SHIFT=0.0 -DMDPP_TERM_REWARD=0.0 -DMDPP_SHIFT_NEGZERO=0 -DMDPP_TERM_NEGZERO=0
 -DMDPP_PN_T=429496730ull -DMDPP_PN_M=2348810239u -DMDPP_PN_SHIFT=25 -DMDPP_CFG_RING_REGS=2
 -DMDPP_N_ENVS=65536ll -DMDPP_AUTORESET=1 -DMDPP_HORIZON=100 -DMDPP_CFG_NOISE=2
 -DMDPP_CFG_NORMAL=${NORMAL:-1} -DMDPP_CFG_FAST=true -DMDPP_CFG_RING=true -DMDPP_CFG_CDF=3
 -DMDPP_CFG_SINGLE=true -DMDPP_CFG_SMEM=${SMEM:-true} -DMDPP_CFG_STAGE=${STAGE:-true} -DMDPP_IRR=0 -DMDPP_OBS_DTYPE=0"
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v \
  -I mdp_playground_b200/csrc -I include $DEFS "$@" -cubin -o "$OUT/jit.cubin" "$OUT/entry.cu"
cuobjdump -sass "$OUT/jit.cubin" > "$OUT/jit.sass"
grep -cE '^\s+/\*[0-9a-f]{4}\*/' "$OUT/jit.sass" | sed 's/^/sass instructions: /'
