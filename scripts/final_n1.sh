#!/bin/bash
# Final single-GPU evidence of the round: GPU test suite, every bench line, warm kernel times.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/twin_flips.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in c1 c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/final_$c.json 2> gpurun_out/final_$c.err; echo "$c rc=$?"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "c2 rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_c2_ref.json 2> gpurun_out/final_c2_ref.err; echo "ref rc=$?"
python bench.py --storage fp32 --steps 20 --warmup 5 --no-pageable > gpurun_out/final_c2_fp32.json 2> gpurun_out/final_c2_fp32.err; echo "fp32 rc=$?"
python bench.py --impl reference --config c1 --steps 20 --warmup 5 > gpurun_out/final_c1_ref.json 2>/dev/null; echo "c1 ref rc=$?"
python scripts/kernel_times.py 1000000 1000 40 20 > gpurun_out/final_kernel_times_steady.txt 2>/dev/null
python scripts/kernel_times.py 1000000 1000 5 20 > gpurun_out/final_kernel_times_early.txt 2>/dev/null
python scripts/kernel_times.py 2000000 200 20 20 fp64 c4 > gpurun_out/final_kernel_times_c4.txt 2>/dev/null
python scripts/kernel_times.py 4000000 500 10 10 fp64 c3 > gpurun_out/final_kernel_times_c3.txt 2>/dev/null
head -8 gpurun_out/final_kernel_times_steady.txt
