"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> <launches.csv> <prof.ncu-rep> "<note>"
"""
import collections
import csv
import subprocess
import sys

tag, launches, rep, note = sys.argv[1:5]
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    k = r[ki].split('(')[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# {tag}: {note}", "# ncu --metrics gpu__time_duration.sum --clock-control none launch list; per-launch times are cold-cache and",
       "# serialised: compare SHARES, not absolutes", "kernel,launches,total_us,share_pct,avg_us"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot * 100:.1f},{v[1] / v[0]:.1f}")
open(f'profiles/{tag}_launch_shares.csv', 'w').write("\n".join(out) + "\n")
print("\n".join(out))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
idx = [hdr.index(w) for w in want if w in hdr]
with open(f'profiles/{tag}_ncu_full_summary.csv', 'w') as f:
    f.write(f"# {tag}: {note}\n# ncu --set full --clock-control none --import-source on, raw page, selected metrics (second row = units)\n")
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([rows[1][i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
print(open(f'profiles/{tag}_ncu_full_summary.csv').read())
