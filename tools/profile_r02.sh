#!/bin/bash
# Round-2 profiling pass on ONE B200 (run under gpurun): launch list of the default bench command, then one
# `ncu --set full` capture per kernel of interest, each after its plain run exited 0.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
L="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$L > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"clip_(fwd|bwd_np)_kernel" -s 8 -c 2 -o gpurun_out/r02_full_loss -f $L > gpurun_out/r02_ncu_loss.log 2>&1
echo "loss rc=$?"
N="python bench.py --steps 3 --warmup 3 --only l2norm"
$N > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:l2norm_cast_kernel -s 3 -c 1 -o gpurun_out/r02_full_l2norm -f $N > gpurun_out/r02_ncu_l2norm.log 2>&1
echo "l2norm rc=$?"
R="python bench.py --steps 4 --warmup 3 --only retrieval --no-cpu-baseline"
$R > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"topk_(floor|sweep|finalize)_kernel" -s 9 -c 3 -o gpurun_out/r02_full_topk -f $R > gpurun_out/r02_ncu_topk.log 2>&1
echo "topk rc=$?"
ls -la gpurun_out/*.ncu-rep
