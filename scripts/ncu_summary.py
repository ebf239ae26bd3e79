"""Summarise an .ncu-rep (ncu --set full) into a small JSON: per launch the duration, DRAM
bytes, tensor-pipe / DRAM / L2 utilisation and launch geometry.  The binary reports stay out of
the tree (profiles/*.ncu-rep is git-ignored); the JSON is what gets committed.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json [kernel-regex]
"""
import csv
import json
import re
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__registers_per_thread": "regs", "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "sm__inst_executed.sum": "warp_inst",
    "lts__t_sectors_op_atom.sum": "l2_atom_sectors", "lts__t_sectors_op_red.sum": "l2_red_sectors",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rx = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        if rx and not rx.search(name):
            continue
        rec = dict(kernel=name.split("(")[0])
        for k, nm in KEEP.items():
            if k in d and d[k] != "":
                try:
                    rec[nm] = float(d[k].replace(",", ""))
                    u = units[hdr.index(k)]
                    if u and nm in ("duration_us", "dram_read", "dram_write"):
                        rec[nm + "_unit"] = u
                except ValueError:
                    pass
        launches.append(rec)
    json.dump(dict(source=rep.split("/")[-1], how="ncu --set full --clock-control none",
                   launches=launches), open(out, "w"), indent=1)
    print(out, len(launches), "launches")


if __name__ == "__main__":
    main()
