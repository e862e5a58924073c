#!/bin/bash
# Regenerates the tracked profiles/ artefacts of a round from gpurun_out/ (tools/run_round_profiles.sh):  tools/make_profiles.sh r01
set -e
R=${1:-r01}
tools/prof_report.sh gpurun_out/prof_mix_final.ncu-rep lrds_tc_mix_a rollout_mix_kernelILi4ENS_6MixCfgILb0ELi1ELi2ELb0 40 > /tmp/mixsum.md 2>/dev/null
{ echo "## $R f16x3 benchmark kernel (rollout_mix_kernel: drift network + mixture-score contractions on tcgen05) - bench workload, B200"; echo; echo "Command: \`ncu --set full --clock-control none --import-source on -k regex:rollout_mix_kernel -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline\` (one launch, cold-cache and serialised; the live timing of the same kernel is in ${R}_final_bench.json)."; echo; tail -n +3 /tmp/mixsum.md; } > profiles/${R}_mix_summary.md
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
    name=r[ki].split("(")[0][:90]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms; tot+=ms
out=[["kernel","launches","total_ms","share_pct"]]+[[k,n,f"{ms:.4f}",f"{100*ms/tot:.2f}"] for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])]
with open(f'profiles/{R}_final_launches.csv','w') as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline (whole process: set-up, warm-up, timed and e2e loops), aggregated per kernel\n")
    csv.writer(f).writerows(out)
PY
cp gpurun_out/bench_final.json profiles/${R}_final_bench.json
cp gpurun_out/bench_final_reference.json profiles/${R}_final_bench_reference.json
cp gpurun_out/shapes_final.json profiles/${R}_shapes.json
