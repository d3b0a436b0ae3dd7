"""Per-(device, stream) scratch state and the assertion policy of the drop-ins.

The reference guards this path with Python ``assert``s (``simplex``, ``one_hot``,
label range), each of which synchronises the device and all of which vanish under
``python -O``.  The kernels instead raise device-side flags; this module decides
when those flags are read back:

  'eager'    (default when ``__debug__``)  check after every call -> same
             AssertionError at the same call site as the reference (one 16-byte D2H)
  'deferred' never sync inside the path; call ``raise_if_flagged()`` when convenient
  'off'      (default under ``python -O``, where the reference's asserts vanish too)
"""
import os

import torch

from . import _lib

_MODES = ("eager", "deferred", "off")
_mode = os.environ.get("DCT_B200_CHECK", "eager" if __debug__ else "off")
if _mode not in _MODES:
    raise ValueError(f"DCT_B200_CHECK must be one of {_MODES}")

_states = {}

_FLAG_MESSAGES = {
    _lib.FLAG_SIMPLEX: "input is not a probability simplex along dim 1 (utils.simplex)",
    _lib.FLAG_LABEL: "labels outside [0, C) (class2one_hot)",
    _lib.FLAG_PRED: "predictions outside [0, C) (ConfusionMatrix bincount size)",
    _lib.FLAG_ONEHOT: "tensor is not one-hot along dim 1 (utils.one_hot)",
}


def set_check_mode(mode: str) -> str:
    """Returns the previous mode."""
    global _mode
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {_MODES}")
    old, _mode = _mode, mode
    return old


def get_check_mode() -> str:
    return _mode


class check_mode:
    """``with check_mode('deferred'): ...`` -- scoped assertion policy ('off' is never overridden: under ``python -O``
    the reference's asserts are gone too)."""

    def __init__(self, mode):
        self.mode, self.old = mode, None

    def __enter__(self):
        if self.mode is not None and _mode != "off":
            self.old = set_check_mode(self.mode)
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            set_check_mode(self.old)
        return False


class StreamState:
    __slots__ = ("workspace", "flags")

    def __init__(self, device):
        nbytes = int(_lib.lib().dct_workspace_bytes())
        self.workspace = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.flags = torch.zeros(_lib.NUM_FLAGS, dtype=torch.int32, device=device)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got device '{t.device}'. "
                           "dct_b200 has no CPU fallback (the CPU oracle lives under oracle/ and is test-only).")


def state(device) -> StreamState:
    stream = torch.cuda.current_stream(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream.cuda_stream)
    st = _states.get(key)
    if st is None:
        st = _states[key] = StreamState(device)
    return st


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def flags_ptr(st: StreamState):
    return None if _mode == "off" else st.flags.data_ptr()


def after_call(st: StreamState) -> None:
    if _mode == "eager":
        _raise_from(st)


def _raise_from(st: StreamState) -> None:
    vals = st.flags.tolist()  # the only host sync on the path
    if any(vals):
        st.flags.zero_()
        msgs = [m for i, m in _FLAG_MESSAGES.items() if vals[i]]
        raise AssertionError("; ".join(msgs))


def raise_if_flagged() -> None:
    """Deferred mode: read back every stream's flags and raise the reference's AssertionError if set."""
    for st in list(_states.values()):
        _raise_from(st)
