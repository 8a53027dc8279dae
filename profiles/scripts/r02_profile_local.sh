#!/bin/bash
# Final profiling pass of the local-BA / global-BA kernels (round 2): launch lists of C4 and C5 and `--set full`
# captures of the tile kernel, the other per-phase kernels, the cyclic-reduction and the dense Cholesky kernels,
# summarised on the box. Local-BA kernels are captured with RSPL_BA_GRAPH=off (ncu cannot profile kernel nodes of a
# graph with conditional nodes; the host-driven driver launches the same kernels one by one).
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
NCU="ncu --clock-control none"
summ() {
  python profiles/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt
  if [ -n "$2" ]; then
    ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$1.csv 2>/dev/null
    python profiles/scripts/ncu_lines.py /tmp/$1.csv "$2" 40 > $O/$1.lines.txt
  fi
}
RSPL_BA_GRAPH=off $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $O/r02_launches_c4_final.csv python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_c4.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r02_launches_c5_final.csv python bench.py --workload c5 --steps 1 --warmup 1 > $O/r02_ncu_c5.log 2>&1
RSPL_BA_GRAPH=off $NCU --set full --import-source on -k regex:"kt_schur_tile" -s 20 -c 2 -o $O/r02_tile_final2 python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_t.log 2>&1
summ r02_tile_final2 "kt_schur_tile<(int)0>"
rm -f $O/r02_tile_final2.ncu-rep
RSPL_BA_GRAPH=off $NCU --set full --import-source on -k regex:"kt_backsub_rc|kb_pose_blocks|kb_linearize|kb_solve|kt_tile_sum" -s 12 -c 7 -o $O/r02_local_others2 python profiles/scripts/r02_ncu_target.py local 1024 > $O/r02_ncu_o.log 2>&1
summ r02_local_others2
rm -f $O/r02_local_others2.ncu-rep
$NCU --set full --import-source on -k regex:"bcr_eliminate|bcr_update" -s 4 -c 4 -o $O/r02_bcr2 python profiles/scripts/r02_c5_target.py 600 > $O/r02_ncu_b.log 2>&1
summ r02_bcr2 "bcr_eliminate"
rm -f $O/r02_bcr2.ncu-rep
RSPL_BA_DENSE_FULL=1 $NCU --set full --import-source on -k regex:"dc_update|dc_panel" -s 20 -c 2 -o $O/r02_dense python profiles/scripts/r02_c5_target.py 600 > $O/r02_ncu_d.log 2>&1
summ r02_dense "dc_update"
rm -f $O/r02_dense.ncu-rep
du -sh $O
