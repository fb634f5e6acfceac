#!/usr/bin/env python3
"""Condense an ncu report (read here, no GPU needed) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep reads_per_launch [kernel-substring] > profiles/rNN_xxx.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    reads = float(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("report: %s" % rep)
        print("kernel: %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        print("reads per launch: %d" % reads)
        for k in KEYS:
            if k in d and d[k] not in ("", None):
                print("  %-62s %s %s" % (k, d[k], u.get(k, "")))
        try:
            inst = float(d["smsp__inst_executed.sum"])
            print("  derived: warp instructions per read = %.1f" % (inst / reads))
            rd = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
            print("  derived: DRAM traffic per launch = %.0f bytes (%.1f per read)" % (rd + wr, (rd + wr) / reads))
        except (KeyError, ValueError):
            pass
        print("  warp stall reasons per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
        st = []
        for k in hdr:
            if "issue_stalled" in k and "per_issue_active" in k:
                try:
                    st.append((float(d[k]), k.split("stalled_")[1].split("_per")[0]))
                except ValueError:
                    pass
        for v, name in sorted(st, reverse=True)[:9]:
            print("    %-22s %.3f" % (name, v))
        print()


if __name__ == "__main__":
    main()
