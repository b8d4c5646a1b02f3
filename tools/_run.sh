OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_r1cs.py -m gpu -q > $OUT/pytest_r1cs.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r1cs.log; tail -30 $OUT/pytest_r1cs.log
