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


def test_team_parallel_split_of_big_nodes_gives_the_same_order():
    """Nodes of a million children and more are split by all threads (compile.cpp, split_node_parallel) instead of one task each,
    and lists of more than 32768 children are sorted by the parallel radix sort; RT2025_TIE_PAR_MIN / RT2025_TIE_RADIX_MIN force
    both paths on the small, tie-ridden inputs of the check (signed zeros included)."""
    env = dict(os.environ, RT2025_TIE_PAR_MIN="8", RT2025_TIE_RADIX_MIN="2")  # (and the radix sort of the three lists)
    b = subprocess.run(["make", "-s", "build/check_tie_order"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stdout[-2000:] + b.stderr[-2000:]
    r = subprocess.run(["build/check_tie_order"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("OK") and "0 rank mismatches" in r.stdout
    assert "1 rank mismatches" not in r.stdout


def test_entry_leaves_of_thick_media_are_complete():
    """csrc/compile.cpp, Medium::entry: the leaves k_walk tests instead of traversing from the root are all the world leaves
    whose box meets the boundary ball (flat scan over every node), parents' boxes contain their children's, and only media
    with an optical radius >= 1 are flagged (book2_final, 12 seeds)."""
    r = subprocess.run(["make", "-s", "check_walk_entries"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "12 thick media checked, 0 problems" in r.stdout
