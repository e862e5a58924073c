#!/bin/bash
# The GPU-side commands behind the tracked profiles/ artefacts of a round (run under gpurun on one B200; outputs land in
# gpurun_out/ and tools/make_profiles.sh turns them into profiles/): bench lines, ncu launch list, one full ncu capture of
# the benchmark kernel, the shape and precision tables.
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_mix_kernel -s 3 -c 1 -o gpurun_out/prof_mix_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mix.log 2>&1; echo "ncu rc=$?"
python tools/shape_bench.py --precisions f16x3 --json gpurun_out/shapes_final.json > gpurun_out/shapes_final.log 2>&1; echo "shapes rc=$?"
python tools/precision_report.py --precisions fp32,tf32x3,f16x3,tf32,bf16 --json gpurun_out/precision_final.json > gpurun_out/precision_final.log 2>&1; echo "prec rc=$?"
