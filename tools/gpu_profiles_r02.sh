#!/bin/bash
# Round-2 profile artefacts: launch lists (per-launch device time) and full captures of the bucket accumulation at benchmark
# and protocol shapes. Each ncu run follows a plain run of the same command that exited 0.
OUT=gpurun_out/r02prof
mkdir -p $OUT
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
for W in prove verify; do
  for B in 1024 1; do
    python tools/profile_protocol.py $W $B > $OUT/plain_${W}_$B.log 2>&1 &&
    timeout 900 $NCU --log-file $OUT/ncu_launches_${W}_b$B.csv python tools/profile_protocol.py $W $B > $OUT/ncu_${W}_$B.log 2>&1
  done
done
BENCH="python bench.py --steps 2 --warmup 1 --no-blindbid --no-cpu --sustained-s 0 --msm-lanes 1"
$BENCH > $OUT/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches_bench_msm.csv $BENCH > $OUT/ncu_bench.log 2>&1
$BENCH > $OUT/plain_bench2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 4 -c 1 -o $OUT/full_k_accumulate_msm $BENCH > $OUT/ncu_full_msm.log 2>&1
python tools/profile_protocol.py prove 1024 > $OUT/plain_prove2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_accumulate -s 6 -c 2 -o $OUT/full_k_accumulate_prove python tools/profile_protocol.py prove 1024 > $OUT/ncu_full_prove.log 2>&1
# keep the raw metric pages (small), drop the reports themselves (tens of MB each; gpurun merges at most 64 MiB back)
for R in full_k_accumulate_msm full_k_accumulate_prove; do
  if [ -f $OUT/$R.ncu-rep ]; then
    ncu -i $OUT/$R.ncu-rep --page raw --csv > $OUT/$R.raw.csv 2>/dev/null
    ncu -i $OUT/$R.ncu-rep --page details --csv > $OUT/$R.details.csv 2>/dev/null
    rm -f $OUT/$R.ncu-rep
  fi
done
ls -la $OUT | tail -24
