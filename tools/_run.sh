OUT=gpurun_out/r02y; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_server.py tests/test_gpu_protocol.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -12 $OUT/pytest.log
