#!/bin/bash
# The GPU-side commands behind the tracked profiles/ artefacts of a round (run under gpurun on one B200; outputs land in
# gpurun_out/ and tools/make_profiles.sh turns them into profiles/): bench lines, ncu launch list, full ncu captures of
# every kernel family of DESIGN.md section 4, the shape and precision tables.
T="timeout -s KILL"
$T 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
$T 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; echo "ref rc=$?"
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
$T 400 ncu --set full --clock-control none --import-source on -k regex:rollout_mix_kernel -s 3 -c 1 -o gpurun_out/prof_mix_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/ncu_mix.log 2>&1; echo "ncu mix rc=$?"
tools/prof_report.sh gpurun_out/prof_mix_final.ncu-rep lrds_tc_mix_a MixCfgILb0ELi1ELi2ELb0ELb0 45 > gpurun_out/mix_summary.md 2>&1; rm -f gpurun_out/prof_mix_final.ncu-rep
$T 400 ncu --set full --clock-control none --import-source on -k regex:rollout_mix_small_kernel -s 2 -c 1 -o gpurun_out/prof_mix_small_final -f python tools/shape_bench.py --precisions f16x3 --only "small cfg2 many_modes d=50 M=16 EI K=200 B=8192" > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
tools/prof_report.sh gpurun_out/prof_mix_small_final.ncu-rep lrds_tc_mix_s rollout_mix_small_kernel 30 > gpurun_out/mix_small_summary.md 2>&1; rm -f gpurun_out/prof_mix_small_final.ncu-rep
$T 400 ncu --set full --clock-control none --import-source on -k regex:rollout_lin_kernel -s 2 -c 1 -o gpurun_out/prof_lin_final -f python tools/shape_bench.py --precisions f16x3 --only "cfg3 phi4 d=100 PIS" > gpurun_out/ncu_lin.log 2>&1; echo "ncu lin rc=$?"
tools/prof_report.sh gpurun_out/prof_lin_final.ncu-rep lrds_tc_f16x3 rollout_lin_kernelILi4ELb1ELb1 30 > gpurun_out/lin_summary.md 2>&1; rm -f gpurun_out/prof_lin_final.ncu-rep
$T 400 ncu --set full --clock-control none --import-source on -k regex:rollout_cmcd_tc_kernel -s 2 -c 1 -o gpurun_out/prof_cmcd_final -f python tools/shape_bench.py --precisions f16x3 --only "cfg4 logreg sonar" > gpurun_out/ncu_cmcd.log 2>&1; echo "ncu cmcd rc=$?"
tools/prof_report.sh gpurun_out/prof_cmcd_final.ncu-rep lrds_tc_f16x3 rollout_cmcd_tc_kernelILi4ELi0 30 > gpurun_out/cmcd_tc_summary.md 2>&1; rm -f gpurun_out/prof_cmcd_final.ncu-rep
$T 400 ncu --set full --clock-control none --import-source on -k regex:mala_kernel -s 1 -c 1 -o gpurun_out/prof_mala_final -f python tools/mala_bench.py > gpurun_out/ncu_mala.log 2>&1; echo "ncu mala rc=$?"
tools/prof_report.sh gpurun_out/prof_mala_final.ncu-rep lrds_capi mala_kernel 30 > gpurun_out/mala_summary.md 2>&1; rm -f gpurun_out/prof_mala_final.ncu-rep
$T 400 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_kernel -s 1 -c 1 -o gpurun_out/prof_mlp_grad_final -f python tools/mlp_grad_bench.py > gpurun_out/ncu_mlp_grad.log 2>&1; echo "ncu mlp_grad rc=$?"
tools/prof_report.sh gpurun_out/prof_mlp_grad_final.ncu-rep lrds_mlp_grad mlp_grad_kernel 30 > gpurun_out/mlp_grad_summary.md 2>&1; rm -f gpurun_out/prof_mlp_grad_final.ncu-rep
$T 300 ncu --set full --clock-control none --import-source on -k regex:score_cot_kernel -s 2 -c 1 -o gpurun_out/prof_sc_final -f python tools/score_cot_bench.py > gpurun_out/ncu_sc.log 2>&1; echo "ncu score_cot rc=$?"
tools/prof_report.sh gpurun_out/prof_sc_final.ncu-rep lrds_capi score_cot_kernel 25 > gpurun_out/score_cot_summary.md 2>&1; rm -f gpurun_out/prof_sc_final.ncu-rep
$T 200 python tools/mlp_grad_bench.py --json gpurun_out/mlp_grad_bench.json > gpurun_out/mlp_grad_bench.log 2>&1; echo "mlp_grad bench rc=$?"
$T 400 python tools/train_bench.py --cpu-batch 512 --json gpurun_out/train_bench_final.json > gpurun_out/train_bench_final.log 2>&1; echo "train bench rc=$?"
$T 200 python tools/train_phases.py --json gpurun_out/train_phases.json > gpurun_out/train_phases.log 2>&1; echo "train phases rc=$?"
$T 600 python tools/shape_bench.py --precisions f16x3 --json gpurun_out/shapes_final.json > gpurun_out/shapes_final.log 2>&1; echo "shapes rc=$?"
$T 600 python tools/precision_report.py --precisions fp32,tf32x3,f16x3,tf32,bf16 --json gpurun_out/precision_final.json > gpurun_out/precision_final.log 2>&1; echo "prec rc=$?"
$T 120 python tools/mala_bench.py > gpurun_out/mala_final.json 2> gpurun_out/mala_final.err; echo "mala rc=$?"
