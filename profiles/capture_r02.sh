#!/bin/bash
# ncu captures of round 2 (run under gpurun, one GPU): one `--set full` capture per hot kernel at
# its SURVEY.md section 8d shape, plus the launch list of the bench step.  The reports are turned
# into text on the box (raw page summary + the hottest source lines) because gpurun_out/ is
# capped at 64 MiB; profiles/ncu_summary.py prints the metrics we steer by.
NCU="ncu --set full --clock-control none --import-source on"
B=profiles/microbench/bench_kernels.py
O=gpurun_out
summarise() {   # $1 = report stem
  python profiles/ncu_summary.py $O/$1.ncu-rep --stalls > $O/$1.txt 2>&1
  ncu -i $O/$1.ncu-rep --page source --csv > $O/$1.source.csv 2>/dev/null
  python profiles/hot_lines.py $O/$1.source.csv >> $O/$1.txt 2>&1
  rm -f $O/$1.ncu-rep $O/$1.source.csv
}
for what in "$@"; do
case $what in
qcp)   $NCU -k regex:qcp_stream -s 8 -c 2 -f -o $O/r02_ncu_qcp_stream_kernel python $B qcp --quick > $O/ncu_qcp.log 2>&1; summarise r02_ncu_qcp_stream_kernel ;;
fastns) $NCU -k regex:row_pairs -s 2 -c 1 -f -o $O/r02_ncu_fastns_row_pairs_kernel python $B fastns > $O/ncu_fastns.log 2>&1; summarise r02_ncu_fastns_row_pairs_kernel ;;
rdf)   $NCU -k regex:fine_walk -s 9 -c 1 -f -o $O/r02_ncu_rdf_fine_walk_kernel python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu > $O/ncu_rdf.log 2>&1; summarise r02_ncu_rdf_fine_walk_kernel
       $NCU -k regex:fine_plan -s 9 -c 1 -f -o $O/r02_ncu_rdf_fine_plan_kernel python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu >> $O/ncu_rdf.log 2>&1; summarise r02_ncu_rdf_fine_plan_kernel ;;
rdftri) $NCU -k regex:row_tile -s 1 -c 1 -f -o $O/r02_ncu_rdf_row_tile_kernel python $B rdf --quick > $O/ncu_rdftri.log 2>&1; summarise r02_ncu_rdf_row_tile_kernel ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu > $O/ncu_launches.log 2>&1 ;;
esac
done
