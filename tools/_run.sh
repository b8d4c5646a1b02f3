OUT=gpurun_out/r02v; mkdir -p $OUT
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/ref.err; echo "ref rc=$?"
python bench.py > $OUT/bench_n1.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
python __graft_entry__.py smoke 2>&1 | tail -2
