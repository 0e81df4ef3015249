#!/bin/bash
# round 2, GPU call 21 (1 GPU): final numbers of the final build -- bench lines, configs, static solve, sanitizer
mkdir -p gpurun_out/r2
timeout 900 python bench.py > gpurun_out/r2/final_bench_256.json 2> gpurun_out/r2/final_bench_256.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2/final_bench_256.json
timeout 900 python bench.py --impl reference > gpurun_out/r2/final_reference_arm.json 2> gpurun_out/r2/final_reference_arm.err; echo "ref rc=$?"
tail -c 700 gpurun_out/r2/final_reference_arm.json
timeout 600 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/final_bench_c4.json 2> gpurun_out/r2/final_bench_c4.err; echo "c4 rc=$?"
tail -c 900 gpurun_out/r2/final_bench_c4.json
timeout 900 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 > gpurun_out/r2/final_configs.jsonl 2> gpurun_out/r2/final_configs.err
timeout 900 python tools/config_bench.py --config c2 --steps 2000 --warmup 200 >> gpurun_out/r2/final_configs.jsonl 2>> gpurun_out/r2/final_configs.err
timeout 900 python tools/config_bench.py --config c3 --steps 3000 >> gpurun_out/r2/final_configs.jsonl 2>> gpurun_out/r2/final_configs.err
cat gpurun_out/r2/final_configs.jsonl
timeout 900 python tools/linsolve_bench.py 64 16 16 128 32 32 256 64 64 384 128 128 256 256 256 > gpurun_out/r2/final_linsolve.jsonl 2> gpurun_out/r2/final_linsolve.err
cat gpurun_out/r2/final_linsolve.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lin_step -s 200 -c 2 -f -o gpurun_out/r2/final_linsolve_ncu \
  python tools/linsolve_bench.py 384 128 128 > gpurun_out/r2/final_linsolve_ncu.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/r2/final_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2/final_memcheck.log
