"""One short render for ncu captures: book2 (default) at a given spp, 2 renders."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
rt = orc.rt
scene = sys.argv[1] if len(sys.argv) > 1 else "book2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
if scene == "book2": hs = rt.named_scene("book2_final", seed=7, params=[800, spp, 40])
elif scene == "cornell": hs = rt.named_scene("cornell_glass", seed=7, params=[600, spp, 50])
elif scene == "book1": hs = rt.named_scene("book1_final", seed=7, params=[1200, spp, 50])
else:
    from scenes_util import final_reduced_scene
    hs = final_reduced_scene(rt, width=1920, spp=spp, depth=30)
sc = rt.Scene(hs)
for k in range(2):
    _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32)
print("paths", st.paths, "ms", st.ms_total, "launches", st.kernel_launches)
