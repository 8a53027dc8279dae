#!/bin/bash
# Round-2 profiling pass on the GPU box (run through gpurun): launch lists and `--set full` captures of the dominant
# kernels, summarised to text on the box (gpurun brings back at most 64 MiB). Local-BA kernels are captured with
# RSPL_BA_GRAPH=off: ncu cannot profile kernel nodes of a graph that holds conditional nodes (the default
# whole-schedule graph); the host-driven driver launches the same kernels one by one.
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
NCU="ncu --clock-control none"
summ() { # report name [kernel substring for the per-line view]
  python profiles/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt
  if [ -n "$2" ]; then
    ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$1.csv 2>/dev/null
    python profiles/scripts/ncu_lines.py /tmp/$1.csv "$2" 40 > $O/$1.lines.txt
  fi
}
# launch lists (cold-cache, serialised: the SHARE per kernel is what counts)
RSPL_BA_GRAPH=off $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/r02_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 > $O/r02_ncu_c2.log 2>&1
RSPL_BA_GRAPH=off $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $O/r02_launches_c4.csv python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_c4.log 2>&1
# full captures
$NCU --set full --import-source on -k regex:frame_opt_kernel -s 1 -c 1 -o $O/r02_frame_c2 python profiles/scripts/r02_ncu_target.py frame 4096 60 > $O/r02_ncu_f1.log 2>&1
summ r02_frame_c2 frame_opt_kernel
$NCU --set full --import-source on -k regex:frame_opt_kernel -s 1 -c 1 -o $O/r02_frame_c2p python profiles/scripts/r02_ncu_target.py frame 4096 0 > $O/r02_ncu_f2.log 2>&1
summ r02_frame_c2p frame_opt_kernel
rm -f $O/r02_frame_c2p.ncu-rep
RSPL_BA_GRAPH=off $NCU --set full --import-source on -k regex:"kt_schur_tile" -s 20 -c 2 -o $O/r02_tile_final python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_t.log 2>&1
summ r02_tile_final "kt_schur_tile<(int)0>"
RSPL_BA_GRAPH=off $NCU --set full --import-source on -k regex:"kt_backsub_rc|kb_pose_blocks|kb_linearize|kb_solve|kt_tile_sum" -s 12 -c 7 -o $O/r02_local_others python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_o.log 2>&1
summ r02_local_others
rm -f $O/r02_local_others.ncu-rep
$NCU --set full --import-source on -k regex:"bcr_eliminate|bcr_update" -s 4 -c 4 -o $O/r02_bcr python profiles/scripts/r02_c5_target.py 600 > $O/r02_ncu_b.log 2>&1
summ r02_bcr "bcr_eliminate"
ncu -i $O/r02_bcr.ncu-rep --page source --csv --print-source cuda,sass > /tmp/bcr.csv 2>/dev/null
python profiles/scripts/ncu_lines.py /tmp/bcr.csv "bcr_update" 30 > $O/r02_bcr_update.lines.txt
rm -f $O/r02_bcr.ncu-rep
du -sh $O
