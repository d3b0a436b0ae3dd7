"""GPU: the fused exchange of the step's loss sums (include/dct_b200.h "Fused cross-rank exchange").

On one GPU the exchange is a loopback (world = 1: the own mailbox is the only peer), which exercises everything but
the NVLink hop: the descriptor, the publication from the last CTA of the step's last kernel (the *_pub tile kernel,
its non-tile fallback, the stand-alone publication kernel), sequence tags, the ring, CUDA-graph replay.  The cross-rank hop is checked by
`bench.py --gpus N` itself (its result must equal an NCCL all-reduce of the same sums, or the run fails).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _step(K, C, B, H, W, px, with_vat=True, mode="fused"):  # fused: exercises the *_pub kernel; "chained" is the product default
    from dct_b200.engine import ConsistencyStep
    return ConsistencyStep(K, C, B, H, W, cin=1, n_global=B * H * W, with_vat=with_vat, exchange=px, exchange_mode=mode)


def test_loopback_publication_matches_local_sums(dev):
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        px = PeerExchange(dev, n=4, nslots=4)
        K, C, B, H, W = 3, 4, 2, 64, 64
        step = _step(K, C, B, H, W, px)
        g = torch.Generator(device=dev).manual_seed(7)
        sets = [StepBuffers.allocate(K, C, B, H, W, 1, dev, g) for _ in range(3)]
        for i in range(7):
            s = sets[i % 3]
            step.run(s)
            torch.cuda.synchronize()
            assert px.published() == i + 1
            got = px.read(i + 1)
            assert torch.equal(got, s.sums[:4]), (got, s.sums)      # world == 1: the published bits themselves
            assert float(s.sums[0]) > 0 and float(s.sums[2]) > 0
        # the ring holds the last 4 publications only
        px.read(4)
        with pytest.raises(RuntimeError):
            px.read(3)
        with pytest.raises(RuntimeError):
            px.read(8)   # not made yet
        px.close()
    finally:
        dct_b200.set_check_mode(old)


def test_loopback_under_cuda_graph_and_jsd_only_trigger(dev):
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        px = PeerExchange(dev, n=4, nslots=8)
        K, C, B, H, W = 2, 4, 2, 64, 64
        step = _step(K, C, B, H, W, px, with_vat=False)           # the JSD kernel itself is the step's last kernel
        g = torch.Generator(device=dev).manual_seed(11)
        bufs = StepBuffers.allocate(K, C, B, H, W, 1, dev, g)
        graph = step.capture(bufs)                                  # warm-up run + capture: one publication so far
        torch.cuda.synchronize()
        base = px.published()
        assert base == 1
        for i in range(5):
            graph.replay()
        torch.cuda.synchronize()
        assert px.published() == base + 5
        assert torch.equal(px.read(base + 5), bufs.sums[:4])
        px.close()
    finally:
        dct_b200.set_check_mode(old)


def test_fallback_shapes_and_plain_entry_points(dev):
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        px = PeerExchange(dev, n=4, nslots=4)
        K, C, B = 2, 2, 2
        g = torch.Generator(device=dev).manual_seed(3)
        bufs = StepBuffers.allocate(K, C, B, 32, 32, 1, dev, g)
        _step(K, C, B, 32, 32, px).run(bufs)
        torch.cuda.synchronize()
        assert px.published() == 1
        # the plain entry points never publish (public autograd API, a step without an exchange)
        z = [t.clone().requires_grad_() for t in bufs.logits]
        dct_b200.jsd_consistency_from_logits(z, weight=1.0).backward()
        _step(K, C, B, 32, 32, None).run(bufs)
        torch.cuda.synchronize()
        assert px.published() == 1
        # odd HW: the *_pub entry point falls back to the non-tile kernels + the stand-alone publication
        bo = StepBuffers.allocate(K, C, B, 33, 31, 1, dev, g)
        _step(K, C, B, 33, 31, px).run(bo)
        torch.cuda.synchronize()
        assert px.published() == 2
        assert torch.equal(px.read(2), bo.sums[:4])
        # the chained mode: plain kernel + the one-thread publication kernel
        _step(K, C, B, 32, 32, px, mode="chained").run(bufs)
        torch.cuda.synchronize()
        assert px.published() == 3
        assert torch.equal(px.read(3), bufs.sums[:4])
        px.seq.zero_()   # (restart the numbering for the checks below)
        _step(K, C, B, 32, 32, px).run(bufs); _step(K, C, B, 33, 31, px).run(bo)
        torch.cuda.synchronize()
        assert px.published() == 2
        # 7 classes: no tile instantiation either
        b7 = StepBuffers.allocate(K, 7, B, 32, 32, 1, dev, g)
        _step(K, 7, B, 32, 32, px).run(b7)
        torch.cuda.synchronize()
        assert px.published() == 3
        assert torch.equal(px.read(3), b7.sums[:4])
        px.close()
    finally:
        dct_b200.set_check_mode(old)


@pytest.mark.parametrize("mode", ["deferred", "early"])
@pytest.mark.parametrize("graph", [False, True])
def test_deferred_publication_runs_beside_the_next_step(dev, graph, mode):
    """exchange_mode='deferred': step i publishes step i-1's sums on a forked branch (eager and captured);
    'early' (the product's choice at world > 1): the first CTA to finish step i's first kernel does (dct_jsd_fwdbwd_pub_f32)."""
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import ConsistencyStep, StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        px = PeerExchange(dev, n=4, nslots=8)
        K, C, B, H, W = 2, 4, 2, 64, 64
        step = ConsistencyStep(K, C, B, H, W, cin=1, n_global=B * H * W, exchange=px, exchange_mode=mode)
        g = torch.Generator(device=dev).manual_seed(5)
        sets = [StepBuffers.allocate(K, C, B, H, W, 1, dev, g) for _ in range(3)]
        step.run(sets[2])                                    # produces the sums the first publication carries
        torch.cuda.synchronize()
        assert px.published() == 0                           # deferred: a step never publishes its own sums
        want2 = sets[2].sums[:4].clone()
        if graph:
            graphs = [step.capture(sets[j], publish_prev=sets[(j - 1) % 3]) for j in range(3)]
            torch.cuda.synchronize()
            px.seq.zero_()                                   # (the captures' warm-up runs published too)
            for j in range(3):
                graphs[j].replay()
        else:
            for j in range(3):
                step.run(sets[j], publish_prev=sets[(j - 1) % 3])
        px.publish(sets[2].sums)                             # the loop's last step
        torch.cuda.synchronize()
        assert px.published() == 4
        assert torch.equal(px.read(2), sets[0].sums[:4]) and torch.equal(px.read(3), sets[1].sums[:4])
        assert torch.equal(px.read(4), sets[2].sums[:4])
        if not graph:
            assert torch.equal(px.read(1), want2)
        px.close()
    finally:
        dct_b200.set_check_mode(old)


@pytest.mark.parametrize("shape", [(3, 4, 2, 64, 64), (2, 2, 2, 64, 64), (2, 19, 1, 32, 64), (2, 4, 2, 37, 41), (5, 3, 1, 40, 52)])
def test_early_publication_every_kernel_family_and_round_graph(dev, shape):
    """The early publication through every JSD launch family -- 256-pixel tensor-map stages with fused Dice, the 4-row
    row-copy stages, C = 19, and shapes outside the tile pipeline (odd HW, runtime K: plain launch + stand-alone publication)
    -- and under the graph of R consecutive steps bench.py replays; losses and Dice counts must not notice the passenger."""
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import ConsistencyStep, StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        K, C, B, H, W = shape
        px = PeerExchange(dev, n=4, nslots=8)
        plain = ConsistencyStep(K, C, B, H, W, cin=1, n_global=B * H * W, with_vat=False, with_dice=C <= 4)
        early = ConsistencyStep(K, C, B, H, W, cin=1, n_global=B * H * W, with_vat=False, with_dice=C <= 4, exchange=px,
                                exchange_mode="early")
        g = torch.Generator(device=dev).manual_seed(11)
        sets = [StepBuffers.allocate(K, C, B, H, W, 1, dev, g) for _ in range(3)]
        want = []
        for s in sets:                                       # reference results without any exchange
            plain.run(s)
            torch.cuda.synchronize()
            want.append((s.sums.clone(), [t.clone() for t in s.grad_logits], s.dice_counts.clone()))
        graph = early.capture_many(sets, publish_chain=True)
        torch.cuda.synchronize()
        px.seq.zero_()
        for _ in range(2):                                   # two rounds: 6 steps, 6 publications (the first one of set 2's sums)
            graph.replay()
        px.publish(sets[2].sums)
        torch.cuda.synchronize()
        assert px.published() == 7
        for j, s in enumerate(sets):
            assert torch.equal(s.sums, want[j][0]) and torch.equal(s.dice_counts, want[j][2])
            assert all(torch.equal(a, b) for a, b in zip(s.grad_logits, want[j][1]))
        # ring of 8: publications 1..7 = sums of sets 2,0,1,2,0,1 and the final 2
        for q, j in zip(range(1, 8), [2, 0, 1, 2, 0, 1, 2]):
            assert torch.equal(px.read(q), sets[j].sums[:4]), (q, j)
        px.close()
    finally:
        dct_b200.set_check_mode(old)
