"""A/B of library builds and tuning knobs on the stage times of a render (runs every variant in a process of its own).

usage: python scripts/ab_stages.py [--scene book2|book1|cornell|final|synth] [--spp N] variant ...
  variant = name[:lib=<path under raytracer-2025_b200/>][:ENV=value]...      e.g.  r1:lib=librt2025_r1.so  new  new64:RT2025_FIFO_SLOTS=64
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "oracle")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import orc
rt = orc.rt
scene, spp = %(scene)r, %(spp)d
if scene == "book2": hs = rt.named_scene("book2_final", seed=7, params=[800, spp, 40])
elif scene == "book1": hs = rt.named_scene("book1_final", seed=7, params=[1200, spp, 50])
elif scene == "cornell": hs = rt.named_scene("cornell_glass", seed=7, params=[600, spp, 50])
elif scene == "synth":
    import tempfile
    from scenes_util import write_synthetic_assets, synthetic_obj_scene
    hs = synthetic_obj_scene(rt, write_synthetic_assets(tempfile.mkdtemp(), n=72), width=1920, spp=spp, depth=30)
else:
    from scenes_util import final_reduced_scene
    hs = final_reduced_scene(rt, width=1920, spp=spp, depth=30)
sc = rt.Scene(hs)
best = None
for k in range(4):
    _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES)
    if k and (best is None or st.ms_total < best.ms_total): best = st
print(f"%(name)-22s total {best.ms_total:8.2f} ms  gen {best.ms_raygen:6.2f} extend {best.ms_extend:7.2f} media+bin {best.ms_other:6.2f} shade {best.ms_shade:7.2f}  "
      f"{best.paths / best.ms_total / 1e3:7.1f} Mpaths/s  {best.segments / best.paths:.3f} seg/path  {best.iterations} iterations  errors {best.errors}", flush=True)
'''

if __name__ == "__main__":
    args = sys.argv[1:]
    scene, spp = "book2", 144
    while args and args[0].startswith("--"):
        if args[0] == "--scene": scene = args[1]
        elif args[0] == "--spp": spp = int(args[1])
        args = args[2:]
    for v in args:
        parts = v.split(":")
        env = dict(os.environ)
        for p in parts[1:]:
            k, _, val = p.partition("=")
            if k == "lib": env["RT2025_LIB"] = os.path.join(ROOT, "raytracer-2025_b200", val)
            else: env[k] = val
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, scene=scene, spp=spp, name=parts[0])], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else f"{parts[0]}: FAILED rc={r.returncode}\n{r.stdout[-800:]}{r.stderr[-1500:]}\n")
        sys.stdout.flush()
