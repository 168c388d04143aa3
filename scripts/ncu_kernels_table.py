"""One line per captured launch of an .ncu-rep (ncu --set full): duration, DRAM bytes and rate, L2 hit rate,
issue-slot utilisation, occupancy, top stall reasons.  usage: ncu_kernels_table.py <rep> [peak_GBps]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6551.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = lambda r, k: r[hdr.index(k)] if k in hdr else ""
num = lambda s: float(s.replace(",", "")) if s not in ("", "n/a") else float("nan")
print(f"# {rep}: ncu --set full --clock-control none; DRAM rate against the measured copy peak {peak:.0f} GB/s")
print("# kernel | grid x block | regs | us | dram rd MB | dram wr MB | GB/s | of peak | L2 hit % | issue active % | warps/SM % | top stalls")
for r in rows[2:]:
    name = col(r, "Kernel Name").replace("<unnamed>::", "")
    name = name[:name.index("(")] if "(" in name and not name.startswith("void") else name.split("(CUtensorMap")[0][:60]
    us = num(col(r, "gpu__time_duration.sum"))
    us = us / 1000.0 if rows[1][hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else us
    unit_r = rows[1][hdr.index("dram__bytes_read.sum")]
    scale = {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(unit_r, 1.0)
    rd = num(col(r, "dram__bytes_read.sum")) * scale
    unit_w = rows[1][hdr.index("dram__bytes_write.sum")]
    wr = num(col(r, "dram__bytes_write.sum")) * {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(unit_w, 1.0)
    gbs = (rd + wr) / us * 1e3 if us else float("nan")      # MB / us = TB/s -> GB/s
    st = []
    for i, k in enumerate(hdr):
        if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued"):
            try:
                st.append((float(r[i].replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    stalls = ", ".join(f"{k} {100 * v / tot:.0f}%" for v, k in sorted(st, reverse=True)[:3])
    print(f"{name} | {col(r, 'launch__grid_size')} x {col(r, 'launch__block_size')} | {col(r, 'launch__registers_per_thread')} | "
          f"{us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {gbs / peak:.2f} | {col(r, 'lts__t_sector_hit_rate.pct')} | "
          f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | {col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | {stalls}")
