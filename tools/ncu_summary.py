"""Turn an .ncu-rep into the per-kernel summary kept under profiles/ (metric, unit, one column per captured launch).
usage: python tools/ncu_summary.py <report.ncu-rep> <out.csv>   (runs `ncu -i ... --page raw --csv`; no GPU needed)"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(gpu__time_duration|dram__bytes_(read|write)\.sum|gpu__dram_throughput|dram__throughput|sm__cycles_(active|elapsed)\.(avg|max|min)$|sm__inst_executed\.sum|smsp__inst_executed\.sum$|"
                  r"sm__issue_active|smsp__issue_active|smsp__average_warps_issue_stalled_.*_per_issue_active|sm__warps_active|launch__|sm__icc_request|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|"
                  r"sm__throughput|lts__throughput|l1tex__throughput|sm__pipe_fp64|smsp__warps_eligible|TPC\.TriageCompute\.sm__cycles_active)")
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
names = [f"{r[ki].split('(')[0]}#{n}" for n, r in enumerate(data)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + names)
    for i, h in enumerate(hdr):
        if KEEP.match(h):
            w.writerow([h, units[i]] + [r[i] for r in data])
print(out, len(names), "launches")
