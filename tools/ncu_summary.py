"""Summarise an .ncu-rep: key metrics + stall breakdown.  usage: python tools/ncu_summary.py file.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_fp64_op_dmma.sum"]

def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name",):
                d["kernel"] = v[:80]
            if h in KEYS:
                d[h] = f"{v} {u}".strip()
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    fv = float(v)
                except ValueError:
                    continue
                if fv >= 0.15:
                    d.setdefault("stalls_per_issue", {})[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(fv, 2)
        res.append(d)
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)

main()
