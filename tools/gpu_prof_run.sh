#!/bin/bash
# Profiling run on the GPU box: warm phase traces, then ncu launch lists of one batched prove / verify call.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for B in 1024 512; do
  timeout 600 python tools/warm_trace.py 8 $B > $OUT/warm_$B.log 2> $OUT/warm_trace_$B.log
done
cat $OUT/warm_1024.log
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
for W in prove verify; do
  timeout 900 $NCU --log-file $OUT/ncu_${W}_1024.csv python tools/profile_protocol.py $W 1024 > $OUT/ncu_${W}_1024.log 2>&1
  python tools/ncu_shares.py $OUT/ncu_${W}_1024.csv "$W 1024" > $OUT/shares_${W}_1024.txt
done
head -12 $OUT/shares_prove_1024.txt; head -12 $OUT/shares_verify_1024.txt
