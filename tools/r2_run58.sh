#!/bin/bash
# compute-sanitizer on the round-2 kernels at small sizes: warp-specialised bulge chasing (mailboxes, slots, transfer
# warps), narrow Q2 variant, pipelined dense->band phase, doubling larft, factored D&C top level
mkdir -p gpurun_out
export BK_SY2SB_PIPE_MIN=256
( time timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_twostage.py 1100 40 ) > gpurun_out/r02e_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02e_sanitize_memcheck.log
grep -E "ERROR SUMMARY|residual|Invalid|rc=" gpurun_out/r02e_sanitize_memcheck.log | head
( time timeout 280 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_twostage.py 400 40 ) > gpurun_out/r02e_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02e_sanitize_racecheck.log
grep -E "RACECHECK SUMMARY|residual|hazard|rc=|real" gpurun_out/r02e_sanitize_racecheck.log | head
