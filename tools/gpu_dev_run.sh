#!/bin/bash
# Development run on the GPU box: parity tests, then per-phase traces of the batched prover / verifier, then a short bench.
# Usage (from the repo root, through gpurun): bash tools/gpu_dev_run.sh [tag]
TAG=${1:-dev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
BBP_TRACE=1 timeout 600 python tools/protocol_perf.py 8 1024 > $OUT/perf_1024.log 2> $OUT/trace_1024.log
tail -2 $OUT/perf_1024.log
timeout 600 python tools/protocol_perf.py 8 1,16,256,1024,3072 > $OUT/perf_sweep.log 2>&1
cat $OUT/perf_sweep.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"
head -c 3000 $OUT/bench.json
