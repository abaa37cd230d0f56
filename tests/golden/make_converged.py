"""Oracle-converged reference image for the north-star PSNR check (BASELINE.json: "PSNR >= 40 dB at high spp").

Config 3 (Cornell box, quad light + glass sphere in the lights list, mixture pdf) at 64x64: the mean of N_RENDERS oracle
renders of 961 spp each with seeds SEED0.. (independent of every seed the GPU tests use).  Written once:
    python tests/golden/make_converged.py        (~1 min of CPU)
-> tests/golden/cornell_glass_converged.npz: image (64,64,3) float64 mean linear radiance, spp_total, seeds
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

N_RENDERS, SEED0 = 32, 900001

if __name__ == "__main__":
    rt = orc.rt
    hs = rt.named_scene("cornell_glass", seed=7, params=[64, 1000, 50])
    osc = orc.OracleScene(hs)
    acc = np.zeros((64, 64, 3))
    errors = 0
    for k in range(N_RENDERS):
        img, st = osc.render(seed=SEED0 + k)
        acc += img
        errors += st.errors
    acc /= N_RENDERS
    np.savez_compressed(os.path.join(HERE, "cornell_glass_converged.npz"), image=acc, spp_total=np.uint64(N_RENDERS * 961),
                        seeds=np.arange(SEED0, SEED0 + N_RENDERS, dtype=np.uint64), errors=np.uint64(errors))
    print("mean", acc.mean(), "errors", errors)
