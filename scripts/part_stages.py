"""Stage times of ONE partition of the bench frame on one GPU (what a rank of an N-GPU run executes): python scripts/part_stages.py [N] [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
rt = orc.rt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
hs = rt.named_scene("book2_final", seed=7, params=[800, spp, 40])
sc = rt.Scene(hs)
best = None
for k in range(4):
    _, st = sc.render(seed=1, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES, part_index=0, part_count=n)
    if k and (best is None or st.ms_total < best.ms_total): best = st
print(f"{os.environ.get('TAG', ''):16s} 1/{n} of the frame: total {best.ms_total:7.2f} ms  gen {best.ms_raygen:5.2f} extend {best.ms_extend:6.2f} media+bin {best.ms_other:5.2f} shade {best.ms_shade:6.2f}  "
      f"sum {best.ms_raygen + best.ms_extend + best.ms_other + best.ms_shade:7.2f}  {best.iterations} iterations  {best.kernel_launches} launches  walk share {best.walk_segments / best.segments:.3f}", flush=True)
