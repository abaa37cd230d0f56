"""The scene compiler's tie ranks (csrc/compile.cpp) against a plain restatement of BVH::from_vec
(bvh.rs:16-46) on child lists full of equal box-mins.  Host only: builds and runs a small C++ check."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_walk_matches_per_node_stable_sort():
    r = subprocess.run(["make", "-s", "check_tie_order"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("OK")
    assert "0 rank mismatches" in r.stdout
