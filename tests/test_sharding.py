"""Multi-rank path on CPU: world_size-2 gloo processes run `crosscorrelate` with the oracle-backed
test double; every rank counts its LPT share of the patch pairs and one sum-reduce to rank 0
reproduces the single-rank (golden) result exactly."""

import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases
import golden_io
from yet_another_wizz_b200.sharding import Shard, assign_pairs_lpt, pair_costs

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lpt_assignment_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    pi = np.repeat(np.arange(64), 7)[:400]
    pj = np.concatenate([np.arange(64), rng.integers(0, 64, 336)])
    costs = pair_costs(pi, pj, rng.integers(1000, 2000, 64), rng.integers(10000, 20000, 64))
    for world in (1, 2, 4, 8):
        shares = assign_pairs_lpt(costs, world)
        assert sorted(np.concatenate(shares).tolist()) == list(range(len(costs)))
        loads = np.array([costs[s].sum() for s in shares])
        assert loads.max() <= loads.mean() * 1.05 + costs.max()
    assert not Shard().active


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_engine import OracleEngine

    import yet_another_wizz_b200 as yb

    g = golden_io.load(name)
    config = golden_cases.config_from_golden(g)
    cats = {k: golden_cases.catalog_from_golden(g, k) for k in ("ref", "unk", "ref_rand", "unk_rand")}
    eng = OracleEngine()
    corrs = yb.crosscorrelate(config, cats["ref"], cats["unk"], ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"],
                              engine=eng)  # shard picked up from torch.distributed
    if rank == 0:
        np.savez(os.path.join(out_dir, "rank0.npz"), **{k: getattr(corrs[0], k).counts.counts for k in ("dd", "dr", "rd", "rr")})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduce_matches_golden(tmp_path):
    name = "cross_unweighted"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    g = golden_io.load(name)
    got = np.load(tmp_path / "rank0.npz")
    for kind in ("dd", "dr", "rd", "rr"):
        np.testing.assert_array_equal(got[kind], g[f"cross_{kind}_counts_s0"])


def test_contiguous_patch_groups():
    from yet_another_wizz_b200.sharding import assign_patches_contiguous

    # 8 x 8 grid of patch centres on the equator, equal cost: groups are compact and balanced
    ras, decs = np.meshgrid(np.deg2rad(np.arange(8) * 5.0 + 2.5), np.deg2rad(np.arange(8) * 3.0 - 10.5))
    xyz = np.column_stack([np.cos(ras.ravel()) * np.cos(decs.ravel()), np.sin(ras.ravel()) * np.cos(decs.ravel()),
                           np.sin(decs.ravel())])
    for world in (1, 2, 4, 8):
        groups = assign_patches_contiguous(np.ones(64), xyz, world)
        assert sorted(np.concatenate(groups).tolist()) == list(range(64))
        assert all(len(g) == 64 // world for g in groups)
        for g in groups:  # compact: bounding box of a group covers at most twice its share of the grid
            cols, rows = g % 8, g // 8
            assert (np.ptp(cols) + 1) * (np.ptp(rows) + 1) <= 2 * len(g)
    # unequal costs: loads within one patch of the mean
    costs = np.random.default_rng(1).uniform(1, 3, 64)
    groups = assign_patches_contiguous(costs, xyz, 4)
    loads = np.array([costs[g].sum() for g in groups])
    assert loads.max() - loads.min() <= 2 * costs.max()


def test_patch_fractions_balance_exactly():
    """shares cut at k / world of the cost: every patch is covered once, loads are equal up to the slivers that
    snap to a patch boundary; the rows of a shared patch are split into disjoint compact strips"""
    from yet_another_wizz_b200.sharding import assign_patch_fractions, split_rows

    rng = np.random.default_rng(3)
    ra, dec = rng.uniform(0.0, 0.7, 64), rng.uniform(-0.2, 0.2, 64)
    xyz = np.column_stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])
    costs = rng.uniform(0.5, 1.5, 64)
    costs[5] = 0.0  # an empty patch still belongs to exactly one rank
    for world in (1, 2, 3, 8):
        shares = assign_patch_fractions(costs, xyz, world)
        cover = np.zeros(64)
        for share in shares:
            for p, f0, f1 in share:
                assert 0.0 <= f0 < f1 <= 1.0
                cover[p] += f1 - f0
        np.testing.assert_allclose(cover, 1.0, atol=1e-12)
        loads = np.array([sum(costs[p] * (f1 - f0) for p, f0, f1 in share) for share in shares])
        assert loads.max() - loads.min() <= 2 * 0.03 * costs.max() + 1e-9
        assert sum(1 for share in shares for p, f0, f1 in share if (f0, f1) != (0.0, 1.0)) <= 2 * (world - 1)
    pts = rng.normal(size=(1000, 3))
    pts /= np.linalg.norm(pts, axis=1)[:, None]
    a, b, c = split_rows(pts, 0.0, 0.25), split_rows(pts, 0.25, 0.6), split_rows(pts, 0.6, 1.0)
    assert sorted(np.concatenate([a, b, c]).tolist()) == list(range(1000))
    assert (len(a), len(b), len(c)) == (250, 350, 400)
