#!/bin/bash
# multi-GPU pass (N from $1): numerical check N ranks == 1 GPU on the concatenated batch, then the bench line
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --check --out gpurun_out/check_n$N.json > gpurun_out/check_n$N.log 2>&1; echo "check exit $?"
grep -o '"ok": [a-z]*' gpurun_out/check_n$N.json | tr '\n' ' '; echo
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --quick > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n$N.log | head -2
exit 0
