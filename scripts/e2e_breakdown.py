"""Where the end-to-end step of bench.py spends its host time (1 GPU): scene create, render, D2H, tonemap, close."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("rt2025", os.path.join(ROOT, "raytracer-2025_b200", "rt2025.py"))
rt = importlib.util.module_from_spec(spec); spec.loader.exec_module(rt)
import torch
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
hs = rt.named_scene("book2_final", seed=7, params=[800, spp, 40])
cam = hs.camera
H, W = cam.image_height, cam.image_width
fb = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
host_fb = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
stream = torch.cuda.current_stream().cuda_stream
for it in range(4):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    sc = rt.Scene(hs); t.append(time.perf_counter())
    st = sc.render_device(fb.data_ptr(), stream=stream, seed=2025, accum_type=rt.RT_ACCUM_F32); t.append(time.perf_counter())
    host_fb.copy_(fb); t.append(time.perf_counter())
    rgb = rt.tonemap(host_fb.numpy(), cam.toon_map); t.append(time.perf_counter())
    sc.close(); t.append(time.perf_counter())
    names = ["scene_create", "render_device", "d2h", "tonemap", "close"]
    print(f"iter {it}: " + "  ".join(f"{n} {1e3*(b-a):.1f} ms" for n, a, b in zip(names, t, t[1:])) + f"  | total {1e3*(t[-1]-t[0]):.1f} ms, device ms_total {st.ms_total:.1f}", flush=True)
# rgb8 path: render + tonemap on the device, 8-bit image back
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sc = rt.Scene(hs)
    img, st = sc.render_rgb8(seed=2025)
    sc.close()
    print(f"render_rgb8 path: {1e3*(time.perf_counter()-t0):.1f} ms total, device {st.ms_total:.1f} ms")
# finer: the raw C calls, no numpy conversions
import ctypes as C
import numpy as np
L = rt.product_lib()
img = host_fb.numpy()
out = np.empty(img.shape, dtype=np.uint8)
for it in range(8):
    t0 = time.perf_counter()
    L.rt_tonemap(img.ctypes.data, rt.RT_ACCUM_F32, img.size // 3, 0, out.ctypes.data)
    t1 = time.perf_counter()
    sc = rt.Scene(hs)
    t2 = time.perf_counter()
    sc.close()
    t3 = time.perf_counter()
    print(f"raw rt_tonemap {1e3*(t1-t0):.1f} ms, scene create {1e3*(t2-t1):.1f} ms, close {1e3*(t3-t2):.1f} ms", flush=True)
