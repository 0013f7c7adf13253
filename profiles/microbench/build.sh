#!/bin/sh
# tcgen05 microbenchmarks and in-kernel phase traces of the policy kernels.  Build here (nvcc
# cross-compiles for sm_100a), run on a B200:   sh profiles/microbench/build.sh && gpurun -- ./profiles/microbench/mma_rate2
set -e
cd "$(dirname "$0")"
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17"
nvcc $F -o mma_rate mma_rate.cu          # cycles per tcgen05.mma issued under `if (threadIdx.x == 0)` (lane-election loop per MMA)
nvcc $F -o mma_rate2 mma_rate2.cu        # the same from warp-uniform code (elect.sync), several issuing warps; tcgen05.ld / st throughput
nvcc $F -fmad=false --expt-relaxed-constexpr -I ../../include -o tc_trace tc_trace.cu     # monolithic kernel: phase stamps of CTA 0
nvcc $F -fmad=false --expt-relaxed-constexpr -I ../../include -o ws_trace ws_trace.cu     # pipeline: one thread per role
nvcc $F -fmad=false -o divtest divtest.cu    # is the 3-FMA exact-remainder quotient IEEE-exact for non-integer divisors?  (no: r02_divtest.log)
nvcc $F -fmad=false -o f32x2 f32x2.cu        # packed fp32 (FFMA2 / FMUL2 / FADD2) against scalar: FMUL2 / FADD2 retire 2 ops per lane and cycle
