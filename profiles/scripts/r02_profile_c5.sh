#!/bin/bash
# DRAM traffic of the Schur class of C5 (kb_schur_prep, kb_schur_reduce, kb_assemble_bcr, the cyclic-reduction kernels):
# one `--set full` capture of one trial of the full-size problem, summarised on the box -> profiles/traffic_c5.json
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
ncu --clock-control none --set full -k regex:"kb_schur_prep|kb_schur_reduce|kb_assemble_bcr|bcr_" -s 60 -c 33 -o $O/r02_c5_schur python profiles/scripts/r02_c5_target.py 2000 > $O/r02_ncu_c5s.log 2>&1
python profiles/ncu_summary.py $O/r02_c5_schur.ncu-rep > $O/r02_c5_schur.txt
rm -f $O/r02_c5_schur.ncu-rep
grep -E "^kernel|dram__bytes|gpu__time" $O/r02_c5_schur.txt | head -120
