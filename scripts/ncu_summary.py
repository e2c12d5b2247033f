"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per captured launch + the hottest SASS lines.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--top 30] [--launch 1]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_static', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = ['Kernel Name'] + KEYS + [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k]
    lines = []
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            vals = [r[i] for r in data]
            if 'issue_stalled' in k and all(float(v or 0) < 0.05 for v in vals):
                continue
            lines.append(f"| {k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')} | {units[i]} | " + " | ".join(v[:60] for v in vals) + " |")
    return lines


def source(rep, launch, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f':::{launch}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:          # a report with a single launch
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith('0x')]
    data = data[:len(data) // 2] if len(data) > 1 and data[0][0] == data[len(data) // 2][0] else data
    iS, iSrc, iEx, iT = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
    iL = hdr.index('stall_long_sb')
    tot = sum(int(r[iS]) for r in data)
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top]
    lines = [f"total samples {tot}, SASS instructions {len(data)}, warp instructions executed {sum(int(r[iEx]) for r in data)}",
             "| # | SASS | samples | executed | avg threads | long_sb |", "|---|---|---|---|---|---|"]
    for i in sorted(idx):
        r = data[i]
        lines.append(f"| {i} | `{r[iSrc].strip()[:70]}` | {r[iS]} | {r[iEx]} | {r[iT]} | {r[iL]} |")
    return lines


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 30
    launch = int(sys.argv[sys.argv.index('--launch') + 1]) if '--launch' in sys.argv else 1
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(3)) + " |\n|---|---|---|---|---|")
    print("\n".join(raw(rep)))
    print()
    print("\n".join(source(rep, launch, top)))
