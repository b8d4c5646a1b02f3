OUT=gpurun_out/r02o; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_server.py -m gpu -x -q > $OUT/pytest_server.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_server.log; tail -30 $OUT/pytest_server.log
