"""Stage times of the cfg2 render (spp 144) for several wavefront capacities."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-2025_b200"))
import rt2025 as rt
hs = rt.named_scene("book2_final", seed=7, params=[800, 144, 40])
sc = rt.Scene(hs)
for cap in [int(x) for x in sys.argv[1:]]:
  for nb in (False, True):
    best = None
    for k in range(3):
        _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES, max_paths_in_flight=cap, no_binning=nb)
        if k and (best is None or st.ms_total < best.ms_total):
            best = st
    print(f"capacity {cap:9d} no_binning={int(nb)}: total {best.ms_total:7.2f} ms  gen {best.ms_raygen:6.2f} extend {best.ms_extend:7.2f} media {best.ms_other:6.2f} "
          f"shade {best.ms_shade:7.2f} iters {best.iterations} launches {best.kernel_launches}  {best.paths / best.ms_total / 1e3:7.1f} Mpaths/s")
