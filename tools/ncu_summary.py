#!/usr/bin/env python3
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the per-launch table committed under profiles/."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "dur_us"), ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd_MB"),
        ("dram__bytes_write.sum", "dram_wr_MB"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("kernel | grid | " + " | ".join(k for _, k in KEYS) + " | top stalls (warps per issue-active cycle)")
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    vals = []
    for h, k in KEYS:
        v = d.get(h, "")
        try:
            v = float(v)
            if k == "dur_us" and u.get(h) in ("ms", "msecond"): v *= 1e3
            if k == "dur_us" and u.get(h) in ("ns", "nsecond"): v /= 1e3
            if k.endswith("_MB"):
                v = v * {"Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6}.get(u.get(h), 1)
            vals.append(f"{v:.4g}")
        except ValueError:
            vals.append(str(v))
    st = []
    for h, v in d.items():
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try: st.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print(f"{d['Kernel Name'][:70]} | {d.get('Grid Size','')} | " + " | ".join(vals) + " | " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:5]))
