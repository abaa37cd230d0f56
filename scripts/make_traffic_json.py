"""profiles/extend_traffic.json from a profiles/<tag>_ncu_full_summary.csv (one whole wavefront iteration, --set full).

    python scripts/make_traffic_json.py r02_v4 <segments through k_extend in that iteration> <walk share of all segments> "<how it was captured>"
"""
import csv, json, subprocess, sys
tag, seg, walk_share, how = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
rows = [r for r in csv.reader(l for l in open(f"profiles/{tag}_ncu_full_summary.csv") if not l.startswith("#"))]
hdr, rows = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
tot_bytes = tot_ms = 0.0
ext = None
for r in rows:
    b = (float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])) * 1e9
    tot_bytes += b
    tot_ms += float(r[col["gpu__time_duration.sum"]])
    if "k_extend" in r[0]:
        ext = (r, b)
r, b = ext
all_seg = seg / (1.0 - walk_share)
out = {
    "kernel": r[0].split("(")[0].replace("void ", ""),
    "dram_bytes_per_launch": int(b),
    "segments_per_launch": seg,
    "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
    "source": f"profiles/{tag}_ncu_full_summary.csv: {how}; k_extend {float(r[col['gpu__time_duration.sum']]):.2f} ms, dram__bytes_read.sum + dram__bytes_write.sum = "
              f"{float(r[col['dram__bytes_read.sum']]):.3f} GB + {float(r[col['dram__bytes_write.sum']]):.3f} GB; algorithmic bytes of the same launch: 97 B x {seg:.4g} segments = {97 * seg / 1e9:.3f} GB "
              "(ray record 64 B in, hit record 16 B in - the medium incumbent - 16 B out and the class byte)",
    "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
    "active_lanes_per_instruction": float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
    "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
    "iteration": {
        "dram_bytes": int(tot_bytes), "kernel_ms_sum": tot_ms, "segments_through_extend": seg,
        "dram_bytes_per_extend_segment": tot_bytes / seg,
        "walk_share_of_all_segments": walk_share, "segments_incl_walk": int(all_seg), "dram_bytes_per_segment": tot_bytes / all_seg,
        "note": "sum over the launches of one wavefront iteration; the segments k_walk evaluates in registers move no stream bytes - their number in this "
                "iteration is taken from the frame average (rt_stats.reserved[0] / segments); round-2 build without the walk: 475 B per segment, round 1: 564",
    },
}
json.dump(out, open("profiles/extend_traffic.json", "w"), indent=1)
print(json.dumps(out["iteration"], indent=1))
