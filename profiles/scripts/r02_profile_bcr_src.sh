#!/bin/bash
# source-level capture of the cyclic-reduction kernels of one C5 reduced solve (first level: 67 odd blocks)
cd "$(dirname "$0")/../.."
O=gpurun_out/bcrsrc
mkdir -p $O
timeout 900 ncu --clock-control none --set full --import-source on -k regex:"bcr_eliminate|bcr_update_resident|bcr_backsub|bcr_root" -s 24 -c 4 -o $O/bcr_src python profiles/scripts/r02_c5_target.py 2000 > $O/ncu.log 2>&1
ncu -i $O/bcr_src.ncu-rep --page raw --csv > $O/bcr_raw.csv 2>/dev/null
ncu -i $O/bcr_src.ncu-rep --page source --csv --kernel-name regex:bcr_eliminate > $O/bcr_elim_source.csv 2>/dev/null
python profiles/scripts/ncu_brief.py $O/bcr_raw.csv | grep -E "kernel|time_duration|grid_size|issue_active|warps_active" 
