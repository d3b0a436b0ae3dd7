"""GPU: the fused exchange of the step's loss sums (include/dct_b200.h "Fused cross-rank exchange").

On one GPU the exchange is a loopback (world = 1: the own mailbox is the only peer), which exercises everything but
the NVLink hop: arming a workspace, the publication from the last CTA of the step's last kernel (tile pipeline and
register-tiled finishers), sequence tags, the ring, CUDA-graph replay.  The cross-rank hop is checked by
`bench.py --gpus N` itself (its result must equal an NCCL all-reduce of the same sums, or the run fails).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _step(K, C, B, H, W, px, with_vat=True):
    from dct_b200.engine import ConsistencyStep
    return ConsistencyStep(K, C, B, H, W, cin=1, n_global=B * H * W, with_vat=with_vat, exchange=px)


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


def test_unarmed_and_foreign_sum_pointers_do_not_publish(dev):
    import dct_b200
    from dct_b200.distributed import PeerExchange
    from dct_b200.engine import StepBuffers
    old = dct_b200.set_check_mode("deferred")
    try:
        px = PeerExchange(dev, n=4, nslots=4)
        K, C, B, H, W = 2, 2, 2, 32, 32
        g = torch.Generator(device=dev).manual_seed(3)
        bufs = StepBuffers.allocate(K, C, B, H, W, 1, dev, g)
        armed = _step(K, C, B, H, W, px)
        armed.run(bufs)
        torch.cuda.synchronize()
        assert px.published() == 1
        # same workspace, public autograd API: its sum outputs are other buffers -> no publication
        z = [t.clone().requires_grad_() for t in bufs.logits]
        dct_b200.jsd_consistency_from_logits(z, weight=1.0).backward()
        torch.cuda.synchronize()
        assert px.published() == 1
        # odd HW (register-tiled finisher) with an armed trigger publishes as well
        Ho = 33
        bo = StepBuffers.allocate(K, C, B, Ho, 31, 1, dev, g)
        so = _step(K, C, B, Ho, 31, px, with_vat=False)
        so.run(bo)
        torch.cuda.synchronize()
        assert px.published() == 2
        assert torch.equal(px.read(2), bo.sums[:4])
        st = dct_b200._runtime.state(dev)
        px.disarm(st.workspace)
        armed.exchange = None
        armed.run(bufs)
        torch.cuda.synchronize()
        assert px.published() == 2
        px.close()
    finally:
        dct_b200.set_check_mode(old)
