"""Per-source-line instruction counts, active lanes and stall samples of one kernel launch in an ncu report
(needs --import-source on and -lineinfo).  usage: ncu_regions.py report.ncu-rep kernel-regex [launch-skip] [top-n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; agg = {}; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == '': continue
    d = dict(zip(hdr, r))
    try: agg[(cur_file, int(r[0]))] = (int(d['Instructions Executed']), int(d['Thread Instructions Executed']), int(d['# Samples']), r[1].strip()[:100])
    except Exception: pass
tot = [sum(v[i] for v in agg.values()) for i in range(3)]
print(f"total warp-inst {tot[0]}  thread-inst {tot[1]}  lanes {tot[1]/max(tot[0],1):.2f}  samples {tot[2]}")
byfile = {}
for (f, l), v in agg.items():
    a = byfile.setdefault(f, [0, 0, 0])
    for i in range(3): a[i] += v[i]
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:32s} inst% {100*a[0]/tot[0]:5.1f} lanes {a[1]/max(a[0],1):5.1f} samples% {100*a[2]/max(tot[2],1):5.1f}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{f:18s} {l:4d} inst% {100*v[0]/tot[0]:5.2f} lanes {v[1]/max(v[0],1):5.1f} smp% {100*v[2]/max(tot[2],1):5.2f}  {v[3]}")
