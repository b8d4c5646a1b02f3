OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log
timeout 600 python tools/warm_trace.py 8 1024 5 > $OUT/warm_1024.log 2> $OUT/warm_trace_1024.log; cat $OUT/warm_1024.log; grep -E "verify|timeline" $OUT/warm_trace_1024.log
timeout 600 python tools/verify_cuts.py 1024 "0:3,1024:1,342:3,256:4" 2>&1 | tee $OUT/cuts.log
timeout 600 python tools/verify_cuts.py 3072 "0:3,768:4" 2>&1 | tee $OUT/cuts3072.log
