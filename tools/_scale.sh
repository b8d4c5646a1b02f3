N=$1
OUT=gpurun_out/r02w; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"
tail -c 300 $OUT/bench_n$N.json | head -c 200; echo
