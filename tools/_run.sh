OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rangeproof.py -m gpu -x -q > $OUT/pytest_rp.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_rp.log; tail -30 $OUT/pytest_rp.log
