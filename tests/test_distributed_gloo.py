"""CPU, world_size 2, gloo: the only collectives of the path -- the global mean of a
sharded loss map, int64 count all-reduce, '2d' row gather -- reproduce the un-sharded oracle result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dct_b200.distributed as D
    import oracle as O
    g = torch.Generator().manual_seed(1234)
    K, B, C, H, W = 2, 6, 4, 16, 16
    z = [3 * torch.randn(B, C, H, W, generator=g) for _ in range(K)]
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    lo, hi = D.shard_batch(B)
    zs = [t[lo:hi].numpy() for t in z]
    _, mp_local, _ = O.jsd_logits_fwdbwd(zs, 1.0, want_grad=False)
    n_local = mp_local.size
    mean = D.global_mean(torch.tensor(mp_local.sum(dtype=np.float64)), n_local)
    n_glob = D.global_pixel_count(n_local, torch.device("cpu"))
    counts, _ = O.dice_counts(zs[0], gt[lo:hi].numpy())
    c3 = D.all_reduce_counts(torch.from_numpy(counts.sum(0, keepdims=True)))
    rows = D.all_gather_rows(torch.from_numpy(O.dice_from_counts(counts)))
    # local gradients with the GLOBAL normaliser equal the matching slice of the un-sharded gradient
    _, _, gz = O.jsd_logits_fwdbwd(zs, float(n_local) / n_glob)
    q.put((rank, float(mean), n_glob, c3.numpy(), rows.numpy(), lo, hi, gz[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reductions_match_unsharded_oracle():
    import oracle as O
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(1234)
    K, B, C, H, W = 2, 6, 4, 16, 16
    z = [3 * torch.randn(B, C, H, W, generator=g) for _ in range(K)]
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    mean, _, gz = O.jsd_logits_fwdbwd([t.numpy() for t in z], 1.0)
    counts, _ = O.dice_counts(z[0].numpy(), gt.numpy())
    for rank, m, n_glob, c3, rows, lo, hi, g0 in res:
        assert n_glob == B * H * W
        assert abs(m - mean) <= 1e-6 * np.log(K)
        assert np.array_equal(c3, counts.sum(0, keepdims=True))          # '3d' counts: exact
        assert np.array_equal(rows, O.dice_from_counts(counts))           # '2d' rows in rank order
        assert np.abs(g0 - gz[0][lo:hi]).max() <= 1e-5 / (B * H * W)
    assert [r[5:7] for r in res] == [(0, 3), (3, 6)]


def test_shard_batch_covers_everything():
    import dct_b200.distributed as D
    for n in (1, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [D.shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _report_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    from dct_b200.cotrain import DeviceReport, init_distributed
    import oracle as O
    assert init_distributed("gloo") == (rank, world, 0)
    g = torch.Generator().manual_seed(1234)
    K, B, C, H, W = 2, 6, 4, 16, 16
    z = [3 * torch.randn(B, C, H, W, generator=g) for _ in range(K)]
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    lo, hi = (rank * B // world, (rank + 1) * B // world)
    rep = DeviceReport(K, C, torch.device("cpu"))
    for k in range(K):   # what CoTrainStep.step feeds it: the fused kernel's per-image counts of this rank's shard
        counts, _ = O.dice_counts(z[k][lo:hi].numpy(), gt[lo:hi].numpy())
        rep.add_counts("unlab", k, torch.from_numpy(counts))
    rep.add_losses([torch.tensor(1.0 + rank), torch.tensor(2.0)], torch.tensor(0.25 * (rank + 1)), None)
    out = rep.reduce()
    q.put((rank, out["unlab_dice"].numpy(), out["losses"].numpy(), out["world"]))
    dist.barrier()
    dist.destroy_process_group()


def test_device_report_reduces_to_the_unsharded_batch_dice():
    """CoTrainStep's reporting: per-rank integer counters + loss sums -> ONE all-reduce -> the Dice of
    the un-sharded batch ('3d' arithmetic, exact) and the mean of the per-rank losses."""
    import oracle as O
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_report_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(1234)
    K, B, C, H, W = 2, 6, 4, 16, 16
    z = [3 * torch.randn(B, C, H, W, generator=g) for _ in range(K)]
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    want = np.stack([O.dice_from_counts(O.dice_counts(z[k].numpy(), gt.numpy())[0].sum(0, keepdims=True))[0] for k in range(K)])
    for rank, dice, losses, w in res:
        assert w == 2 and np.array_equal(dice, want)
        assert np.allclose(losses, [1.5, 2.0, 0.375, 0.0])
