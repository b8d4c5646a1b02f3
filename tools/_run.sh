OUT=gpurun_out/r02u; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout 600 python tools/warm_trace.py 8 1024 5 > $OUT/warm.log 2> $OUT/warm_trace.log; cat $OUT/warm.log; grep -E "bbp_blindbid_verify_batch|verify_prepare" $OUT/warm_trace.log
BBP_HOST_THREADS=4 taskset -c 0-3 timeout 600 python tools/verify_cuts.py 1024 "0:3" 2>&1 | tee $OUT/cuts_4c.log
timeout 600 python tools/verify_cuts.py 1024 "0:3" 2>&1 | tee $OUT/cuts_16c.log
timeout 600 python tools/verify_cuts.py 3072 "0:3" 2>&1 | tee $OUT/cuts_3072.log
