"""CPU leg of the golden fixtures: the oracle must reproduce the committed vectors bit for bit
(closest hits) and to libm accuracy (images), so the fixtures the GPU tests read stay honest."""
import os

import numpy as np
import pytest

from test_gpu_parity import GOLDEN, golden_scene


@pytest.mark.parametrize("name", ["book2_final", "cornell_glass", "book1_final", "random_graph", "final_reduced"])
def test_oracle_reproduces_golden(rt, orc, name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    osc = orc.OracleScene(golden_scene(rt, name))
    hits = osc.closest_hit(fx["rays"], mode=0)
    assert np.array_equal(hits["prim_id"], fx["hits"]["prim_id"])
    assert np.array_equal(hits["inst_id"], fx["hits"]["inst_id"])
    assert np.array_equal(hits["t"], fx["hits"]["t"])
    brute = osc.closest_hit(fx["rays"], mode=1)
    assert np.array_equal(brute["prim_id"], fx["hits"]["prim_id"])
    img, st = osc.render(seed=int(fx["render_seed"]))
    assert st.paths == int(fx["paths"]) and st.errors == int(fx["errors"])
    assert np.allclose(img, fx["image"], rtol=1e-12, atol=1e-14)
