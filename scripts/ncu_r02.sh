#!/bin/bash
# Round-2 ncu evidence (one GPU): launch list of the bench command, then ONE --set full capture of the kernels of
# two consecutive iterations inside the replayed graph.  The plain run comes first and must exit 0.
set -u
CMD="python bench.py --steps 30 --warmup 5 --no-solve --no-cpu --no-pageable"
$CMD > gpurun_out/ncu_plain.log 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_v2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches_v2.csv)"
ncu --set full --clock-control none --import-source on \
    -k regex:'rbl_gather_kernel|pav_seg_merge_kernel|ss_partition_ranked_kernel|ss_partition_kernel|ss_bucket_kernel|gram_fista_persistent_kernel|chunk_prefix_kernel|scatter_active_kernel|sparse_dual_t_kernel' \
    -s 160 -c 16 -o gpurun_out/r02_full_v2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$? $(ls -la gpurun_out/r02_full_v2.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
