#!/usr/bin/env python
"""Summarise an ncu report of bench.py's top kernel into profiles/ (run in the build container; ncu reads the
.ncu-rep offline):   python scripts/ncu_summary.py gpurun_out/TAG_prof.ncu-rep TAG [workload] [pairs]

Writes profiles/r1_TAG_ncu_summary.txt (the metrics the README / DESIGN cite) and updates profiles/traffic.json
(dram bytes read + written per launch of the workload's kernel, consumed by bench.py's roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "8k_rot_poly_linear"
    pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = [f"Kernel Name {d.get('Kernel Name', ('', '?'))[1]}"]
    for w in WANT:
        if w in d:
            lines.append(f"{w} [{d[w][0]}] {d[w][1]}")
    for h in hdr:
        if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
            lines.append(f"{h} [{d[h][0]}] {d[h][1]}")
    out_dir = ROOT / (sys.argv[5] if len(sys.argv) > 5 else "profiles")
    (out_dir / f"{tag}_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    if out_dir.name != "profiles":
        print("\n".join(lines[:14]))
        return

    def gb(name):
        u, v = d[name]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        return float(v) * scale

    tp = ROOT / "profiles" / "traffic.json"
    traffic = json.loads(tp.read_text()) if tp.exists() else {}
    traffic[workload] = {"pairs_per_launch": pairs, "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                         "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
                         "source": f"profiles/{tag}_ncu_summary.txt (ncu --set full, one launch)"}
    tp.write_text(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
