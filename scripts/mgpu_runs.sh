#!/bin/bash
# Multi-GPU evidence run (one 8-GPU box): row-sharded parity under torchrun at 2/4/8 ranks, then the BASELINE
# configurations that shard — c2 (1M x 1000, rows), c3 (4M x 500 EHRM, rows, at 2/4/8), c4 (2M x 201 AoRR, rows) and
# c5 (256-point lambda grid, instances) — each through bench.py.  Outputs land in gpurun_out/.
set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5 > gpurun_out/n8_parity.log; tail -3 gpurun_out/n8_parity.log
run() {  # nproc port out args...
    local n=$1 port=$2 out=$3; shift 3
    timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
    echo "$out rc=$? $(head -c 160 gpurun_out/$out.json)"
}
run 8 29601 n8_c2 --steps 20 --warmup 5
run 8 29602 n8_c3 --config c3 --steps 20 --warmup 5 --no-cpu
run 2 29604 n2_c3 --config c3 --steps 20 --warmup 5 --no-cpu
run 8 29605 n8_c5 --config c5 --steps 20 --warmup 5 --no-cpu
run 8 29606 n8_c4 --config c4 --steps 20 --warmup 5 --no-cpu
