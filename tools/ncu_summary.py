#!/usr/bin/env python
"""Markdown summary of one ncu capture (ncu -i X.ncu-rep --page raw --csv) for profiles/: duration, pipe utilisation,
stall reasons, memory traffic.   python tools/ncu_summary.py <raw.csv> [title]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.per_cycle_active", "active warps per SM"),
    ("smsp__inst_executed.sum", "warp-instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1/shared data-pipe wavefronts %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld (LDTM) instructions"),
    ("smsp__sass_inst_executed_op_tmem_stt.sum", "tcgen05.st (STTM) instructions"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"## {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
    print(f"kernel: `{d.get('Kernel Name', ('', '?'))[1]}`\n")
    print("| metric | value |\n|---|---|")
    for k, name in KEYS:
        if k in d:
            print(f"| {name} (`{k}`) | {d[k][1]} {d[k][0]} |")
    print("\nstall reasons (warps stalled per issued instruction):\n")
    st = [(float(v[1]), k.split("issue_stalled_")[1].split("_per_")[0]) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    print(", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True) if x >= 0.01))


if __name__ == "__main__":
    main()
