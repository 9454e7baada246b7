"""Flat parameter plumbing: state_dict <-> one fp32 buffer (+ one int64 side buffer).

The reference passes Python lists of `state_dict`s around (main.py:130,181,196,219: three deep
copies per client per round).  On a B200 a client model is 28 MB of the 180 GB of HBM, so every
client's parameters live in ONE contiguous fp32 buffer (each tensor 16-byte aligned so 128-bit
loads work per tensor as well as across the whole buffer); `FlatStateDict` is an OrderedDict
whose values are views into it, so it still *is* a state_dict for `load_state_dict`, while
FedAvg sees K flat pointers.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import operator

import torch

_ALIGN = 4  # elements (16 bytes)


@dataclass(frozen=True)
class FlatLayout:
    keys: tuple
    shapes: tuple
    is_int: tuple          # True for int64 entries (BatchNorm num_batches_tracked)
    offsets: tuple         # element offset inside the f32 buffer (float keys) or the i64 buffer (int keys)
    numels: tuple
    n_f32: int             # padded length of the fp32 buffer
    n_i64: int

    @property
    def float_index(self):
        return [i for i, b in enumerate(self.is_int) if not b]

    @property
    def int_index(self):
        return [i for i, b in enumerate(self.is_int) if b]


_layout_cache: dict = {}


_DTYPE_OF = operator.attrgetter("dtype")
_layout_dtypes: dict = {}      # id(layout) -> per-entry dtype tuple (layouts live in _layout_cache for the process lifetime)


def layout_of(state_dict, like: "FlatLayout | None" = None) -> FlatLayout:
    """like: a layout the dict is expected to have (client 0's).  The match is then decided by three C-level passes
    (key tuple, numel and dtype of every entry — what the kernels' indexing depends on) instead of building and
    hashing the full (key, shape, dtype) signature, which costs 0.5 ms per 727-entry dict; entries whose shape
    differs from client 0's at equal numel are averaged element by element."""
    lay = getattr(state_dict, "layout", None)
    if isinstance(lay, FlatLayout) and not getattr(state_dict, "ints_as_float", False):
        return lay          # a FlatStateDict knows its layout (its keys / shapes cannot change)
    if like is not None and len(state_dict) == len(like.keys) and tuple(state_dict.keys()) == like.keys:
        vals = list(state_dict.values())
        dts = _layout_dtypes.get(id(like))
        if dts is None:
            dts = _layout_dtypes[id(like)] = tuple(torch.int64 if b else torch.float32 for b in like.is_int)
        if tuple(map(torch.Tensor.numel, vals)) == like.numels and tuple(map(_DTYPE_OF, vals)) == dts:
            return like
    sig = tuple((k, tuple(v.shape), v.dtype) for k, v in state_dict.items())
    lay = _layout_cache.get(sig)
    if lay is not None:
        return lay
    keys, shapes, is_int, offsets, numels = [], [], [], [], []
    off_f = off_i = 0
    for k, shp, dt in sig:
        n = 1
        for s in shp:
            n *= s
        keys.append(k); shapes.append(shp); numels.append(n)
        if dt == torch.float32:
            is_int.append(False); offsets.append(off_f)
            off_f += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        elif dt == torch.int64:
            is_int.append(True); offsets.append(off_i)
            off_i += n
        else:
            raise TypeError(f"fedmlp_b200 handles float32 and int64 state_dict entries; {k} is {dt}")
    lay = FlatLayout(tuple(keys), tuple(shapes), tuple(is_int), tuple(offsets), tuple(numels), off_f, off_i)
    _layout_cache[sig] = lay
    return lay


class FlatStateDict(OrderedDict):
    """state_dict whose tensors are views into `flat_f32` / `flat_i64` (device buffers)."""

    layout: FlatLayout
    flat_f32: torch.Tensor
    flat_i64: torch.Tensor | None

    @classmethod
    def empty(cls, layout: FlatLayout, device, ints_as_float: bool = False, lazy: bool = False, zero: bool = True):
        """ints_as_float: int64 entries are float32 (FedAvg's output dtype quirk) stored behind
        the fp32 parameters in the same buffer.
        lazy: the 727 views of a DenseNet121 cost milliseconds of Python to create; a lazy dict allocates only
        the flat buffers and builds its entries on first access as a mapping (FedAvg outputs: the flat round
        loop never looks at them, load_state_dict does).  zero=False skips the memset (every element that a
        view can see is about to be written)."""
        self = cls()
        self.layout = layout
        extra = layout.n_i64 if ints_as_float else 0
        alloc = torch.zeros if zero else torch.empty
        self.flat_f32 = alloc(layout.n_f32 + extra, dtype=torch.float32, device=device)
        self.flat_i64 = None if ints_as_float or layout.n_i64 == 0 else alloc(
            layout.n_i64, dtype=torch.int64, device=device)
        self.ints_as_float = ints_as_float
        self._pending = True
        if not lazy:
            self._materialise()
        return self

    def _materialise(self):
        if not getattr(self, "_pending", False):
            return
        self._pending = False
        layout, ints_as_float = self.layout, self.ints_as_float
        for i, k in enumerate(layout.keys):
            n, off = layout.numels[i], layout.offsets[i]
            if not layout.is_int[i]:
                v = self.flat_f32[off:off + n]
            elif ints_as_float:
                v = self.flat_f32[layout.n_f32 + off:layout.n_f32 + off + n]
            else:
                v = self.flat_i64[off:off + n]
            OrderedDict.__setitem__(self, k, v.view(layout.shapes[i]))

    # every way of looking at the mapping goes through _materialise (overriding __iter__ also keeps
    # dict(x) / OrderedDict(x) / update(x) off CPython's storage-level fast path)
    def __getitem__(self, key):
        self._materialise()
        return OrderedDict.__getitem__(self, key)

    def __iter__(self):
        self._materialise()
        return OrderedDict.__iter__(self)

    def __len__(self):
        return len(self.layout.keys) if hasattr(self, "layout") else OrderedDict.__len__(self)

    def __contains__(self, key):
        self._materialise()
        return OrderedDict.__contains__(self, key)

    def keys(self):
        self._materialise()
        return OrderedDict.keys(self)

    def values(self):
        self._materialise()
        return OrderedDict.values(self)

    def items(self):
        self._materialise()
        return OrderedDict.items(self)

    def get(self, key, default=None):
        self._materialise()
        return OrderedDict.get(self, key, default)

    def __reversed__(self):
        self._materialise()
        return OrderedDict.__reversed__(self)

    def __eq__(self, other):
        self._materialise()
        return OrderedDict.__eq__(self, other)

    __hash__ = None

    def __repr__(self):
        self._materialise()
        return OrderedDict.__repr__(self)

    def copy(self):
        self._materialise()
        return OrderedDict(self.items())

    @classmethod
    def from_state_dict(cls, state_dict, device=None):
        """Pack (copy) an ordinary state_dict into flat storage on `device`."""
        lay = layout_of(state_dict)
        first = next(iter(state_dict.values()))
        device = torch.device(device) if device is not None else first.device
        self = cls.empty(lay, device)
        torch._foreach_copy_(list(self.values()), [v.detach() for v in state_dict.values()])
        return self

    def __setitem__(self, key, value):
        """Assigning to an existing key copies INTO the flat view (the dict keeps aliasing its buffers, which is
        what FedAvg's flat path reads); a new key would break the layout and is refused."""
        self._materialise()
        if OrderedDict.__contains__(self, key) and isinstance(value, torch.Tensor):
            OrderedDict.__getitem__(self, key).copy_(value)
            return
        raise KeyError(f"FlatStateDict has a fixed layout; cannot add key {key!r}")

    def clone(self):
        new = FlatStateDict.empty(self.layout, self.flat_f32.device, getattr(self, "ints_as_float", False))
        new.flat_f32.copy_(self.flat_f32)
        if self.flat_i64 is not None:
            new.flat_i64.copy_(self.flat_i64)
        return new

    def __deepcopy__(self, memo):
        return self.clone()


def flatten_module_(module: torch.nn.Module) -> FlatStateDict:
    """Re-point every parameter/buffer of `module` into one flat buffer (in place) and return the
    FlatStateDict that aliases them; afterwards `module.state_dict()` tensors ARE the flat views,
    so training updates the flat buffer directly and FedAvg needs no packing copy."""
    sd = module.state_dict(keep_vars=True)
    flat = FlatStateDict.from_state_dict({k: v.detach() for k, v in sd.items()})
    for k, v in sd.items():
        v.data = flat[k]
    return flat


def flat_view_of(state_dict, lay=None):
    """If `state_dict`'s tensors are consecutive views of one flat buffer laid out as by
    FlatLayout, return (f32_base_ptr, i64_base_ptr); else None.  Lets FedAvg recognise models
    prepared with flatten_module_ even when handed a plain OrderedDict from net.state_dict().
    lay: the dict's layout if the caller already has it (layout_of is a pass over all entries)."""
    if isinstance(state_dict, FlatStateDict) and not getattr(state_dict, "ints_as_float", False):
        return (state_dict.flat_f32.data_ptr(),
                0 if state_dict.flat_i64 is None else state_dict.flat_i64.data_ptr())
    if lay is None:
        lay = layout_of(state_dict)
    base_f = base_i = None
    last_f = None
    for i, v in enumerate(state_dict.values()):
        if not v.is_contiguous():
            return None
        p = v.data_ptr()
        if not lay.is_int[i]:
            last_f = v
        if lay.is_int[i]:
            b = p - 8 * lay.offsets[i]
            if base_i is None:
                base_i = b
            elif b != base_i:
                return None
        else:
            b = p - 4 * lay.offsets[i]
            if base_f is None:
                base_f = b
            elif b != base_f:
                return None
    if base_f is not None:
        # the flat kernel reads n_f32 floats (padded to 4) with 128-bit loads from the base: the base must be
        # 16-byte aligned and the last tensor's storage must extend over the padding
        if base_f % 16:
            return None
        st = last_f.untyped_storage()
        if base_f + 4 * lay.n_f32 > st.data_ptr() + st.nbytes():
            return None
    return (base_f or 0, base_i or 0)
