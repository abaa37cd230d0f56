"""world_size-2 `gloo` test of the multi-GPU plumbing used by bench.py: every rank renders its
interleaved-tile partition, one reduce(sum) of the float framebuffer reassembles the image.
On the CPU box the renderer standing in for the device is the oracle (test infrastructure); the
partition arithmetic and the collective are what is under test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = orc.rt.named_scene("cornell_glass", seed=3, params=[40, 4, 6])
    osc = orc.OracleScene(hs)
    part_index, part_count = bench.partition_for_rank(rank, world)
    img, st = osc.render(seed=9, part_index=part_index, part_count=part_count, threads=2)
    fb = torch.from_numpy(img.astype(np.float32))
    total_paths = bench.reduce_framebuffer_and_paths(fb, st.paths, dst=0)
    if rank == 0:
        np.save(out_path, fb.numpy())
        assert total_paths == 40 * 40 * 4
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_reduce(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    out = str(tmp_path / "fb.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    hs = orc.rt.named_scene("cornell_glass", seed=3, params=[40, 4, 6])
    whole, _ = orc.OracleScene(hs).render(seed=9)
    got = np.load(out)
    assert np.allclose(got, whole.astype(np.float32), rtol=1e-6, atol=1e-7)
