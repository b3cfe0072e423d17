"""Condense an `ncu --page raw --csv` export into the per-kernel summary kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof_raw.csv profiles/ncu_full_rN_summary.json [traffic.json]"""
import csv, json, re, sys

KEYS = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_MB": "dram__bytes_read.sum",
    "dram_write_MB": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "dyn_smem_KB": "launch__shared_mem_per_block_dynamic",
    "warp_inst": "smsp__inst_executed.sum",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
}
SCALE = {"us": {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}, "MB": {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3},
         "KB": {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("(int)", "").replace("(bool)", "")
        d = {}
        for k, m in KEYS.items():
            if m not in hdr:
                continue
            i = hdr.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            for suffix, table in SCALE.items():
                if k.endswith("_" + suffix):
                    v *= table.get(units[i], 1.0)
            d[k] = round(v, 3)
        if "dram_read_MB" in d and "time_us" in d:
            d["dram_GBps"] = round((d["dram_read_MB"] + d.get("dram_write_MB", 0.0)) / d["time_us"] * 1e3, 1)
        key, n = name, 2
        while key in out:
            key = "%s #%d" % (name, n); n += 1
        out[key] = d
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    tot = sum(v["time_us"] for v in out.values())
    for k, v in out.items():
        print("%-46s %7.1f us %5.1f%%  %7.1f MB  %6.0f GB/s  regs %3d  warps %4.1f%%  issue %4.1f%%" % (
            k[:46], v["time_us"], 100 * v["time_us"] / tot, v.get("dram_read_MB", 0) + v.get("dram_write_MB", 0),
            v.get("dram_GBps", 0), v.get("regs", 0), v.get("warps_active_pct", 0), v.get("issue_active_pct", 0)))
    if len(sys.argv) > 3:       # DRAM bytes per half-pass of the two interior kernel families (roofline.traffic in bench.py)
        e = sum(v["dram_read_MB"] + v["dram_write_MB"] for k, v in out.items() if k.startswith("e_interior"))
        h = sum(v["dram_read_MB"] + v["dram_write_MB"] for k, v in out.items() if k.startswith("h_interior"))
        json.dump({"e_interior": round(e * 1e6), "h_interior": round(h * 1e6), "source": sys.argv[1]}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
