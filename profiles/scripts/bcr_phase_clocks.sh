#!/bin/bash
# bcr_phase_clocks.sh: per-phase SM clocks of CTA 0 of bcr_eliminate (build with -DRSPL_BCR_CLOCKS; device printf)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/bcr
RSPL_BA_LIB=build/ab/bcrclk.so timeout 300 python profiles/scripts/r02_c5_target.py 2000 2>&1 | grep "phase clocks" | tail -4 | tee gpurun_out/bcr/phase_clocks.txt
