"""Print the handful of ncu metrics we read for every capture: python tools/ncu_brief.py file.ncu-rep"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pat = re.compile(r'^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_(read|write)\.sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|'
                 r'sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|lts__throughput.avg.pct_of_peak_sustained_elapsed|'
                 r'l1tex__m_xbar2l1tex_read_bytes.sum|l1tex__m_xbar2l1tex_read_bytes.sum.per_second|lts__t_sector_hit_rate.pct|sm__cycles_elapsed.max|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|'
                 r'sm__throughput.avg.pct_of_peak_sustained_elapsed|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active.ratio|sm__inst_executed.sum|launch__shared_mem_per_block_dynamic)$')
for i, h in enumerate(hdr):
    if pat.match(h):
        vals = [r[i][:48] for r in rows[2:]]
        if h.startswith("smsp__average") and all(float(v or 0) < 0.3 for v in vals):
            continue
        print("%-86s %-10s %s" % (h, units[i], vals))
