#!/bin/bash
# Profiling pass of the final K7 (after the range-test-free math / per-edge line residual changes): launch list of the
# default bench workload and `--set full` captures of both instantiations, summarised on the box.
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
NCU="ncu --clock-control none"
summ() {
  python profiles/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt
  ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$1.csv 2>/dev/null
  python profiles/scripts/ncu_lines.py /tmp/$1.csv "$2" 60 > $O/$1.lines.txt
}
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/r02_launches_c2_final.csv python bench.py --workload c2 --steps 2 --warmup 1 > $O/r02_ncu_c2_final.log 2>&1
$NCU --set full --import-source on -k regex:frame_opt_kernel -s 1 -c 1 -o $O/r02_frame_c2_final python profiles/scripts/r02_ncu_target.py frame 4096 60 > $O/r02_ncu_f1.log 2>&1
summ r02_frame_c2_final frame_opt_kernel
$NCU --set full --import-source on -k regex:frame_opt_kernel -s 1 -c 1 -o $O/r02_frame_c2p_final python profiles/scripts/r02_ncu_target.py frame 4096 0 > $O/r02_ncu_f2.log 2>&1
summ r02_frame_c2p_final frame_opt_kernel
rm -f $O/r02_frame_c2p_final.ncu-rep $O/r02_frame_c2_final.ncu-rep
du -sh $O
