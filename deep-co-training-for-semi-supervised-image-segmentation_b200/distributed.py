"""Data-parallel glue for the consistency path (one process per GPU, torch.distributed).

The per-pixel path needs no exchange: the batch is sharded on dim 0 and
the only cross-rank coupling is a handful of scalars / small integer tensors:

  * the global mean of a loss map:  all_reduce(SUM) of [sum_local, n_local]
  * Dice '3d' counts (summed over the whole batch) and confusion matrices: all_reduce(SUM) of int64

These helpers are backend-agnostic (NCCL on the B200 box, gloo in the CPU tests) and are the
ONLY collectives this package issues; gradient all-reduce belongs to DDP around the networks.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_batch(n_items: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous slice of a batch of ``n_items`` (equal shards, remainder to
    the first ranks) -- the same split for the unlabeled batch and every labeled batch."""
    if rank is None:
        rank = dist.get_rank() if is_distributed() else 0
    if world is None:
        world = dist.get_world_size() if is_distributed() else 1
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_mean(local_sum: torch.Tensor, n_local: int, group=None) -> torch.Tensor:
    """sum_local / n_local over all ranks == the reference's ``.mean()`` over the un-sharded batch."""
    buf = torch.stack([local_sum.reshape(()).to(torch.float64),
                       torch.tensor(float(n_local), dtype=torch.float64, device=local_sum.device)])
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(torch.float32)


def global_pixel_count(n_local: int, device, group=None) -> int:
    """Total pixel count over ranks (the ``n_global`` of the fused losses)."""
    if not is_distributed():
        return int(n_local)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def all_reduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of integer count tensors (Dice '3d' [.,C,3] after a batch sum, confusion [C,C])."""
    if is_distributed():
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate per-image Dice rows ('2d') from all ranks in rank order (equal row counts)."""
    if not is_distributed():
        return rows
    out = [torch.empty_like(rows) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, rows.contiguous(), group=group)
    return torch.cat(out, 0)


# ----------------------------------------------------------------------------------------------------
# Fused exchange of the step's loss sums over NVLink peer memory (include/dct_b200.h, "Fused cross-rank exchange")
# ----------------------------------------------------------------------------------------------------
class PeerExchange:
    """Every rank's loss sums in every rank's HBM, written by the step's own last kernel.

    One process per GPU.  Each rank owns a small *mailbox* in device memory, shared with the other ranks of the
    node through CUDA IPC; the step's last kernel (``dct_kl_from_logits_fwdbwd_pub_f32``, handed the descriptor
    :meth:`descriptor` builds for the step's ``sums`` buffer) pushes all ``n`` sums into row ``rank`` of EVERY mailbox
    with plain NVLink stores (sequence-tagged 8-byte words, see ``peer_publish`` in csrc/dct_common.cuh).  There is no collective launch and no NCCL kernel
    competing with the persistent CTAs of the next step; :meth:`read` sums the rows in rank order, so all ranks get
    bit-identical totals.  ``world == 1`` (no process group) is a loopback on the own mailbox.

    The mailbox is a ring of ``nslots`` publications: read a publication before ``nslots`` newer ones have been
    made (a reporting interval shorter than the ring), or :meth:`read` reports a sequence mismatch.
    """

    def __init__(self, device, n: int = 4, nslots: int = 64, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        h = _lib.lib()
        assert 1 <= n <= _lib.PUB_MAX_VALUES and nslots >= 1
        self.device = torch.device(device)
        self.n, self.nslots, self.group = n, nslots, group
        self.world = dist.get_world_size(group) if is_distributed() else 1
        self.rank = dist.get_rank(group) if is_distributed() else 0
        if self.world > _lib.MAX_PEERS:
            raise ValueError(f"PeerExchange serves one node: at most {_lib.MAX_PEERS} ranks")
        self.words = nslots * self.world * _lib.PUB_ROW_WORDS
        with torch.cuda.device(self.device):
            own, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            _lib.check(h.dct_mailbox_create(self.words * 8, C.byref(own), handle), "dct_mailbox_create")
            self._own = own.value
            self._ptrs = [None] * self.world
            self._ptrs[self.rank] = self._own
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    _lib.check(h.dct_mailbox_open(handles[r], C.byref(p)), "dct_mailbox_open")
                    self._ptrs[r] = p.value
        self.seq = torch.zeros(1, dtype=torch.int64, device=self.device)   # device-side publication counter
        self._table = torch.tensor([_to_signed64(p) for p in self._ptrs], dtype=torch.int64, device=self.device)
        self._desc = {}      # sums pointer -> dct_peer_pub (host struct)
        self._mail = _as_int64_tensor(self._own, self.words, self.device)   # view of the own mailbox
        assert int(h.dct_peer_pub_bytes()) == C.sizeof(_lib.PeerPub) == 40

    def descriptor(self, sums: torch.Tensor):
        """``dct_peer_pub`` (a host-side ctypes struct, copied into the kernel parameters by the ``*_pub`` entry
        points) publishing ``sums[:n]``; cached per ``sums`` buffer.  Pass ``ctypes.byref`` of it to the C ABI."""
        assert sums.dtype == torch.float64 and sums.is_cuda and sums.numel() >= self.n and sums.is_contiguous()
        key = sums.data_ptr()
        desc = self._desc.get(key)
        if desc is None:
            desc = self._lib.PeerPub()
            desc.src, desc.seq = key, self.seq.data_ptr()
            desc.n, desc.rank, desc.world, desc.nslots = self.n, self.rank, self.world, self.nslots
            desc.mailbox_table = self._table.data_ptr()
            self._desc[key] = desc
        return desc

    def publish(self, sums: torch.Tensor) -> None:
        """Stand-alone publication (one small CTA) for steps whose last kernel has no fused variant."""
        self._lib.check(self._lib.lib().dct_exchange_publish(self._C.byref(self.descriptor(sums)),
                                                             torch.cuda.current_stream(self.device).cuda_stream),
                        "dct_exchange_publish")

    def published(self) -> int:
        """Publications this rank has made so far (host sync)."""
        return int(self.seq.item())

    def read(self, seq: int, check: bool = True, wait_s: float = 0.0) -> torch.Tensor:
        """Global sums ``[n]`` float64 (device) of publication number ``seq`` (1-based, < 2**32: the tag is the low 32
        bits and 0 is the freshly zeroed mailbox): the rows of all ranks added in rank order.

        The slot is copied ONCE and both the tags and the values are taken from that copy, so a peer's store landing
        in between cannot produce a validated-but-torn total.  ``check`` (one host sync per poll) verifies that every
        word of the copy carries ``seq``; peers' rows arrive asynchronously, so with ``wait_s > 0`` the copy is
        re-taken until they all have or the time is up (then, as with ``wait_s == 0``: RuntimeError)."""
        import time
        assert 1 <= seq < (1 << 32), "publication numbers are 1-based and the tag holds 32 bits"
        R = self._lib.PUB_ROW_WORDS
        slot = seq % self.nslots
        view = self._mail[slot * self.world * R:(slot + 1) * self.world * R].view(self.world, R)[:, :2 * self.n]
        deadline = time.monotonic() + max(float(wait_s), 0.0)
        while True:
            rows = view.clone()                                   # one snapshot: validated and decoded below
            if not check:
                break
            bad = int((((rows >> 32) & 0xffffffff) != seq).sum().item())
            if bad == 0:
                break
            if time.monotonic() >= deadline:
                raise RuntimeError(f"PeerExchange.read({seq}): {bad} word(s) of slot {slot} carry another sequence number "
                                   "(publication not arrived yet, or overwritten: ring of %d slots)" % self.nslots)
            time.sleep(50e-6)
        lo = rows & 0xffffffff
        bits = lo[:, 0::2] | (lo[:, 1::2] << 32)
        vals = bits.contiguous().view(torch.float64)          # [world, n]
        total = vals[0].clone()
        for r in range(1, self.world):                        # fixed rank order: bit-identical on every rank
            total = total + vals[r]
        return total

    def close(self) -> None:
        h = self._lib.lib()
        torch.cuda.synchronize(self.device)
        if is_distributed():
            dist.barrier(group=self.group)                    # nobody unmaps while a peer may still be storing
        for r, p in enumerate(self._ptrs):
            if p is not None and r != self.rank:
                h.dct_mailbox_close(p, 0)
        if is_distributed():
            dist.barrier(group=self.group)
        if self._own is not None:
            self._mail = None
            h.dct_mailbox_close(self._own, 1)
            self._own = None
        self._ptrs = []


def _to_signed64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


class _RawCuda:
    """Minimal ``__cuda_array_interface__`` holder: a torch view of device memory this library allocated."""

    def __init__(self, ptr: int, nwords: int):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _as_int64_tensor(ptr: int, nwords: int, device) -> torch.Tensor:
    with torch.cuda.device(device):
        return torch.as_tensor(_RawCuda(ptr, nwords), device=device)
