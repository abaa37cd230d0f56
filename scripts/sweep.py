"""Tuning sweep: stage times of the cfg2 render for several values of an environment knob.
usage: python scripts/sweep.py ENVVAR v1 v2 ...   (spp fixed at 64, 3 repeats, best)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-2025_b200"))
import rt2025 as rt
var, vals = sys.argv[1], sys.argv[2:]
hs = rt.named_scene("book2_final", seed=7, params=[800, 64, 40])
for v in vals:
    os.environ[var] = v
    sc = rt.Scene(hs)
    best = None
    for k in range(4):
        _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES)
        if k and (best is None or st.ms_total < best.ms_total):
            best = st
    print(f"{var}={v:>6s}: total {best.ms_total:7.2f} ms  gen {best.ms_raygen:6.2f} extend {best.ms_extend:7.2f} media {best.ms_other:6.2f} shade {best.ms_shade:7.2f}  "
          f"{best.paths / best.ms_total / 1e3:7.1f} Mpaths/s")
    sc.close()
