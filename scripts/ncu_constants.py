"""Regenerate profiles/ncu_constants.json — the per-unit constants bench.py's roofline record takes from ncu — from an
`ncu --set full` capture and the bench line of the SAME command run without ncu (read here, no GPU needed).

    python scripts/ncu_constants.py gpurun_out/prof_k_traverse_<tag>.ncu-rep gpurun_out/plain_<tag>.log k_traverse [--note "..."]

The capture holds consecutive launches of one kernel of one wavefront batch (depth 1, 2, 3, ...); the bench line's `stages`
give the units (grid rays, rays) of each depth per launch.  Written: DRAM bytes and warp instructions per unit (sums over
the captured launches / their units), and time-weighted issue-active, FP64-pipe, L1 / L2 hit rates, threads per instruction.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def to_bytes(v, unit):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep, bench_log, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    line = json.loads([l for l in open(bench_log).read().splitlines() if l.startswith("{")][-1])
    stages = [s for s in line["stages"] if s["kernel"] == kernel]
    hdr, units, data = ncu_raw(rep)
    col = lambda name: hdr.index(name)
    data = [r for r in data if kernel in r[col("Kernel Name")]]
    n = min(len(data), len(stages))
    if n == 0:
        raise SystemExit("no matching launches / stages")
    # launch j of the capture = the kernel's launch at depth stages[j]["depth"] of the first timed batch
    per_launch_units = [s["units"] / s["launches"] for s in stages[:n]]
    tot_units = sum(per_launch_units)
    f = lambda r, name: float(r[col(name)])
    dur = [f(r, "gpu__time_duration.sum") for r in data[:n]]
    dram = sum(to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]) +
               to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]) for r in data[:n])
    inst = sum(f(r, "smsp__inst_executed.sum") for r in data[:n])
    wavg = lambda name: sum(f(r, name) * t for r, t in zip(data[:n], dur)) / sum(dur)
    out = {}
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    if os.path.exists(path):
        out = json.load(open(path))
    out[kernel] = {
        "source": os.path.basename(rep), "command": line.get("config", {}).get("workload", ""), "note": note,
        "launches_captured": n, "depths": [s["depth"] for s in stages[:n]], "units_captured": tot_units,
        "dram_bytes_per_unit": dram / tot_units, "warp_inst_per_unit": inst / tot_units,
        "issue_active_ncu": wavg("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
        "fp64_pipe_active": wavg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
        "l1_hit": wavg("l1tex__t_sector_hit_rate.pct") / 100.0, "l2_hit": wavg("lts__t_sector_hit_rate.pct") / 100.0,
        "threads_per_inst": wavg("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "registers": int(f(data[0], "launch__registers_per_thread")),
    }
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out[kernel], indent=1))


if __name__ == "__main__":
    main()
