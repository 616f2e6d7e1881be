"""Summarise an `ncu -i <rep> --page raw --csv` dump: one JSON record per profiled launch with the metrics DESIGN.md
quotes (duration, DRAM bytes, throughputs, hit rates, occupancy, registers).  usage: ncu_summary.py raw.csv > out.json"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "lts_t_bytes",
    "lts__t_sectors_srcunit_tex.sum": "lts_sectors_from_tex",
    "lts__t_sectors_srcunit_tex_op_read.sum": "lts_sectors_from_tex_read",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_bytes",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "smsp__inst_executed.sum": "inst_executed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_of_elapsed",
}

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
out = []
for r in data:
    rec = {"kernel": r[col["Kernel Name"]].replace("rgcn::", "")}
    for m, name in KEEP.items():
        if m in col and r[col[m]] != "":
            try:
                rec[name] = float(r[col[m]].replace(",", ""))
            except ValueError:
                rec[name] = r[col[m]]
            if units[col[m]]:
                rec[name + "_unit"] = units[col[m]]
    out.append(rec)
json.dump(out, sys.stdout, indent=1)
