#!/bin/bash
# Regenerates the tracked profiles/ artefacts of a round from gpurun_out/ (tools/run_round_profiles.sh):  tools/make_profiles.sh r02
set -e
R=${1:-r02}
hdr() {  # kernel title, command, summary file -> profiles file
  { echo "## $R $1 - B200"; echo; echo "Command: \`$2\` (one launch, cold-cache and serialised under the profiler; live timings are in ${R}_final_bench.json / ${R}_shapes.json)."; echo; tail -n +3 "$3"; } > "$4"
}
hdr "f16x3 benchmark kernel (rollout_mix_kernel: drift network, mixture logits and score contractions on tcgen05), bench workload" \
    "ncu --set full --clock-control none --import-source on -k regex:rollout_mix_kernel -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads" gpurun_out/mix_summary.md profiles/${R}_mix_summary.md
hdr "small-batch kernel (rollout_mix_small_kernel: four threads per particle), B = 8192, K = 200" \
    "ncu --set full ... -k regex:rollout_mix_small_kernel -s 2 -c 1 python tools/shape_bench.py --precisions f16x3 --only 'small cfg2 ... B=8192'" gpurun_out/mix_small_summary.md profiles/${R}_mix_small_summary.md
hdr "lattice kernel (rollout_lin_kernel<F16X3, EM, PHI4>), PIS d = 100, K = 256, B = 131072" \
    "ncu --set full ... -k regex:rollout_lin_kernel -s 2 -c 1 python tools/shape_bench.py --precisions f16x3 --only 'cfg3 phi4 d=100 PIS'" gpurun_out/lin_summary.md profiles/${R}_lin_summary.md
hdr "logistic-regression CMCD kernel (rollout_cmcd_tc_kernel), sonar shape d = 61, K = 100, B = 262144" \
    "ncu --set full ... -k regex:rollout_cmcd_tc_kernel -s 2 -c 1 python tools/shape_bench.py --precisions f16x3 --only 'cfg4 logreg sonar'" gpurun_out/cmcd_tc_summary.md profiles/${R}_cmcd_tc_summary.md
hdr "MALA kernel (mala_kernel), mcmc_sample defaults over ManyModes d = 50" \
    "ncu --set full ... -k regex:mala_kernel -s 1 -c 1 python tools/mala_bench.py" gpurun_out/mala_summary.md profiles/${R}_mala_summary.md
hdr "weight-gradient kernel (mlp_grad_kernel: forward recompute, backward-data and weight-gradient GEMMs on tcgen05), 13.1 M rows, d = 50" \
    "ncu --set full ... -k regex:mlp_grad_kernel -s 1 -c 1 python tools/mlp_grad_bench.py" gpurun_out/mlp_grad_summary.md profiles/${R}_mlp_grad_summary.md
hdr "score-cotangent reduction kernel (score_cot_kernel), 13.1 M rows, ManyModes d = 50" \
    "ncu --set full ... -k regex:score_cot_kernel -s 2 -c 1 python tools/score_cot_bench.py" gpurun_out/score_cot_summary.md profiles/${R}_score_cot_summary.md
python - "$R" <<'PY'
import csv, collections, sys
R=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/launches_final.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
cols=rows[hdr]; ki=cols.index("Kernel Name"); vi=cols.index("Metric Value"); ui=cols.index("Metric Unit")
agg=collections.OrderedDict(); tot=0
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(",","")); u=r[ui]
    ms = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else v)
    name=r[ki].split("(")[0][:110]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms; tot+=ms
out=[["kernel","launches","total_ms","ms_per_launch","share_pct"]]+[[k,n,f"{ms:.4f}",f"{ms/n:.4f}",f"{100*ms/tot:.2f}"] for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])]
with open(f'profiles/{R}_final_launches.csv','w') as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline (whole process: set-up, warm-up, timed and e2e loops, the other BASELINE shapes), aggregated per kernel\n")
    csv.writer(f).writerows(out)
PY
cp gpurun_out/bench_final.json profiles/${R}_final_bench.json
cp gpurun_out/bench_final_reference.json profiles/${R}_final_bench_reference.json
cp gpurun_out/shapes_final.json profiles/${R}_shapes.json
cp gpurun_out/mala_final.json profiles/${R}_mala_bench.json
cp gpurun_out/mlp_grad_bench.json profiles/${R}_mlp_grad_bench.json
cp gpurun_out/train_bench_final.json profiles/${R}_train_bench.json
cp gpurun_out/train_phases.json profiles/${R}_train_phases.json
python tools/sass_evidence.py > profiles/${R}_sass_evidence.txt
python tools/precision_md.py gpurun_out/precision_final.json ${R} > profiles/${R}_precision.md
