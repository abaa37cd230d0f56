"""Top source lines by warp-stall samples / executed instructions from an ncu report.
usage: python scripts/ncu_lines.py <rep> <kernel regex> [launch-skip]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
cur = None; hdr = None
lines = []
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "":
        d = dict(zip(hdr, r))
        # duplicate header names: '# Samples' etc unique enough
        try:
            lines.append((cur, int(r[0]), r[1].strip()[:100], float(r[hdr.index('# Samples')]), float(r[hdr.index('Instructions Executed')]),
                          float(r[hdr.index('Thread Instructions Executed')])))
        except ValueError:
            pass
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines); tt = sum(l[5] for l in lines)
print(f"total samples {ts:.0f} warp-instr {ti:.0f} thread-instr {tt:.0f} avg lanes {tt/ti:.1f}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for l in lines:
    byfile[l[0]][0] += l[3]; byfile[l[0]][1] += l[4]; byfile[l[0]][2] += l[5]
for k, v in byfile.items(): print(f"  {k:20s} samples {v[0]/ts*100:5.1f}%  warp-instr {v[1]/ti*100:5.1f}%  lanes {v[2]/max(v[1],1):.1f}")
print("top lines by samples:")
import os
by = 4 if os.environ.get("BY") == "ins" else 3  # BY=ins sorts by executed warp instructions instead of stall samples
for l in sorted(lines, key=lambda l: -l[by])[:int(os.environ.get("TOP", "45"))]:
    print(f"{l[0]:18s}:{l[1]:4d} smp {l[3]/ts*100:5.1f}% ins {l[4]/ti*100:5.1f}% lanes {l[5]/max(l[4],1):5.1f} | {l[2]}")
