"""Extract the metrics DESIGN.md / profiles/ quote from .ncu-rep files (read with `ncu -i ... --page raw --csv`).
usage: python tools/ncu_summary.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
]


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    units = rows[1]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        m = {}
        for k in KEYS:
            if k in d and d[k] != "":
                u = units[hdr.index(k)]
                m[k] = f"{d[k]} {u}".strip()
        out.append(dict(kernel=d.get("Kernel Name"), id=d.get("ID"), metrics=m))
    return out


if __name__ == "__main__":
    res = {}
    for rep in sys.argv[2:]:
        res[rep.split("/")[-1]] = read(rep)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    for k, v in res.items():
        for e in v:
            print(k, e["kernel"][:60], e["metrics"].get("gpu__time_duration.sum"), e["metrics"].get("dram__bytes_read.sum"), e["metrics"].get("sm__warps_active.avg.pct_of_peak_sustained_active"))
