#!/bin/bash
# Round-2 ncu captures, summarised on the GPU box (the reports themselves are too large to bring back).
# usage (on the box, from the repo root): bash tools/prof_r02.sh <tag>
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
cap() {  # name regex count -- command
  name=$1; regex=$2; count=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -c $count -o /tmp/$name -f "$@" > $out/${tag}_${name}_run.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > $out/${tag}_ncu_${name}.txt 2>&1
  for k in $(echo "$regex" | tr '|' ' '); do
    echo "===== per-line profile of $k (first launch) =====" >> $out/${tag}_ncu_${name}.txt
    python tools/ncu_lines.py /tmp/$name.ncu-rep $k 2>&1 | head -45 >> $out/${tag}_ncu_${name}.txt
  done
  rm -f /tmp/$name.ncu-rep
}
cap q5 "k_match_shallow|k_parse_pair|k_sort_scatter|k_trees|k_split_greedy|k_fin_write" 14 python tools/prof_run.py 100000000 1 5
cap q9json "k_parse_ondemand|k_rank_sig" 2 python tools/prof_run.py 50000000 1 9 json
cap q10 "k_match_all|k_match_level|k_zopfli|k_bs_forward" 4 python tools/prof_run.py 28000000 1 10
ls -la $out | tail -8
