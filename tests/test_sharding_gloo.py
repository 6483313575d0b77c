"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: columns are split
into contiguous blocks by qcoh_partition_columns, every rank computes only its own block (no
halo, no data-path collective), and the build-defined diagnostic's float64 partial sums are
all-reduced.  The per-rank compute here is the CPU oracle standing in for the kernels; on the GPU
box bench.py runs the same logic with libqcoh + NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, model_path, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import cpu as oracle
    from quickchem_b200 import capi, synth

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fields = synth.raw_fields(4, seed=77)  # every rank can rebuild the global state (test only)
    km, ncol_g = fields["T"].shape
    c0, n = capi.partition_columns(ncol_g, world, rank)
    shard = {k: np.ascontiguousarray(v[..., c0 : c0 + n]) for k, v in fields.items()}
    m = oracle.Model(model_path)
    r = oracle.run1(m, shard, synth.MAPL, tropp_min=0.0)
    # local float64 partial sums of the diagnostic: sum(OH * w), sum(w)
    pl = (shard["PLE"][:-1] + shard["PLE"][1:]) * np.float32(0.5)
    w = (shard["PLE"][1:].astype(np.float64) - shard["PLE"][:-1]) * (pl > shard["TROPP"][None, :])
    part = torch.tensor([(r["OH"].astype(np.float64) * w).sum(), w.sum()], dtype=torch.float64)
    dist.all_reduce(part)
    gathered = [None] * world
    dist.all_gather_object(gathered, (c0, n, r["OH"]))
    if rank == 0:
        oh = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])], axis=1)
        np.savez(os.path.join(out_dir, "out.npz"), OH=oh, part=part.numpy(), cols=np.array([(g[0], g[1]) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_rank(oracle, small_model_path, tmp_path):
    import socket

    import torch.multiprocessing as mp

    from quickchem_b200 import synth

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, small_model_path, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "out.npz")
    fields = synth.raw_fields(4, seed=77)
    ref = oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL, tropp_min=0.0)
    # sharding by columns changes nothing: bit-identical OH, shards tile the grid exactly
    assert np.array_equal(got["OH"].view(np.uint32), ref["OH"].view(np.uint32))
    cols = got["cols"][np.argsort(got["cols"][:, 0])]
    assert cols[0, 0] == 0 and np.all(cols[1:, 0] == np.cumsum(cols[:-1, 1])) and cols[:, 1].sum() == 96
    pl = (fields["PLE"][:-1] + fields["PLE"][1:]) * np.float32(0.5)
    w = (fields["PLE"][1:].astype(np.float64) - fields["PLE"][:-1]) * (pl > fields["TROPP"][None, :])
    expect = np.array([(ref["OH"].astype(np.float64) * w).sum(), w.sum()])
    assert np.allclose(got["part"], expect, rtol=1e-10, atol=0)


def test_bench_fields_do_not_depend_on_the_sharding():
    """bench.py's synthetic state is a function of the global column only (synth.raw_fields_blocked), so that
    1, 2, 4 and 8 ranks compute the same global OH and the bench line's checksum can be compared across N."""
    import bench
    from quickchem_b200 import synth

    for grid, worlds in ((12, (1, 2, 4, 8)), (24, (1, 2, 3, 4, 6, 8))):
        whole = synth.raw_fields_blocked(grid, 7)
        assert whole["T"].shape == (72, 6 * grid * grid) and whole["PLE"].shape[0] == 73
        for world in worlds:
            parts = []
            for rank in range(world):
                j0, j1 = bench.shard_rows(grid, rank, world)
                parts.append(synth.raw_fields_blocked(grid, 7, j0, j1))
            assert bench.shard_rows(grid, 0, world)[0] == 0 and bench.shard_rows(grid, world - 1, world)[1] == 6 * grid
            for k in whole:
                assert np.array_equal(np.concatenate([p[k] for p in parts], axis=-1), whole[k]), (grid, world, k)
    assert not np.array_equal(synth.raw_fields_blocked(12, 7)["T"], synth.raw_fields_blocked(12, 8)["T"])
