"""Throughput of every BASELINE.json config that can be built here (1, 2, 3, 4 on the shipped assets, a synthetic stand-in for 4), 1 GPU."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
from scenes_util import final_reduced_scene, write_synthetic_assets, synthetic_obj_scene
rt = orc.rt
cfgs = [("cfg1 book1_final 1200x675 spp 10->9 depth 50", lambda: rt.named_scene("book1_final", seed=7, params=[1200, 10, 50]), 1),
        ("cfg2 book2_final 800x800 spp 1000->961 depth 40", lambda: rt.named_scene("book2_final", seed=7, params=[800, 1000, 40]), 4),
        ("cfg3 cornell_glass 600x600 spp 1000->961 depth 50", lambda: rt.named_scene("cornell_glass", seed=7, params=[600, 1000, 50]), 4),
        ("cfg4 assets/Final reduced (13 of 15 meshes, 38 234 triangles, generated HDR environment) 1920x1080 spp 256 of 3000 depth 30",
         lambda: final_reduced_scene(rt, width=1920, spp=256, depth=30), 1),
        ("cfg4-synthetic obj scene 1920x1080 spp 256 depth 30 (11.5k faces, Disney, fog mesh, portal)",
         lambda: synthetic_obj_scene(rt, write_synthetic_assets(tempfile.mkdtemp(), n=72), width=1920, spp=256, depth=30), 1)]
pick = os.environ.get("CFGS")  # e.g. CFGS=3,4; NO_ORACLE=1 skips the CPU side
for name, make, strata in cfgs:
    if pick and name.split()[0][3:] not in pick.split(","):
        continue
    hs = make()
    sc = rt.Scene(hs)
    best = None
    for k in range(3):
        _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES)
        if best is None or st.ms_total < best.ms_total:
            best = st
    _, cst = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_COUNT)
    ext = max(1, cst.segments - cst.walk_segments)  # the counters see the segments that pass through k_extend / the media passes
    ctxt = f", {cst.node_visits/ext:.1f} nodes + {cst.prim_tests/ext:.1f} prims per segment through extend"
    otxt = ""
    if not os.environ.get("NO_ORACLE"):
        osc = orc.OracleScene(hs)
        t0 = time.perf_counter()
        _, ost = osc.render(seed=1, sample_begin=0, sample_end=strata)
        dt = time.perf_counter() - t0
        otxt = f"; oracle {ost.paths/dt/1e6:.2f} Mpaths/s on {orc.lib().orc_num_threads()} threads ({strata} strata)"
    print(f"{name}: {best.paths/1e6:.1f} M paths in {best.ms_total:.1f} ms = {best.paths/best.ms_total/1e3:.1f} Mpaths/s, {best.segments/best.paths:.2f} seg/path, "
          f"errors {best.errors}, stages gen/ext/media/shade {best.ms_raygen:.0f}/{best.ms_extend:.0f}/{best.ms_other:.0f}/{best.ms_shade:.0f} ms" + ctxt + otxt, flush=True)
    sc.close()
