"""``Module`` / ``Sequential`` / ``ModuleList`` with the reference's reflection-based behaviour
(neunet/nn/modules.py): parameters are found by walking ``__dict__``, ``to(device)`` replaces every
attribute that has a ``.to`` (subclasses commonly skip ``super().__init__()``), ``state_dict`` keys
are attribute paths."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .. import backend as _be
from ..autograd import Tensor


def _is_param(x):
    return isinstance(x, Tensor) and x.__class__.__name__ == "Parameter"


class Module:
    def __init__(self):
        self.training = True

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def backward(self, *args, **kwargs):
        raise NotImplementedError

    # -- children ----------------------------------------------------------------------------
    def _children(self):
        return list(self.__dict__.items())

    def parameters(self):
        params, seen = [], set()
        for _, item in self._children():
            if isinstance(item, Tensor):
                if item.requires_grad and _is_param(item) and id(item) not in seen:
                    params.append(item)
                    seen.add(id(item))
            if hasattr(item, "parameters"):
                params.extend(item.parameters())
        return params

    def eval(self):
        self.training = False
        for _, item in self._children():
            if hasattr(item, "eval"):
                item.eval()

    def train(self, mode: bool = True):
        self.training = mode
        for _, item in self._children():
            if hasattr(item, "train"):
                item.train(mode)

    def to(self, device):
        self.xp = _be.get_xp(device)
        self.device = device
        for name, item in list(self.__dict__.items()):
            if name in ("xp", "device"):
                continue
            if hasattr(item, "to") and not isinstance(item, type):
                self.__dict__[name] = item.to(device)
        return self

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self.to("cuda")

    # -- checkpoints: host NumPy arrays so pickles are interchangeable with the reference's cpu ones --
    def state_dict(self):
        sd = OrderedDict()
        for name, item in self._children():
            if _is_param(item):
                sd[name] = _be.to_host(item.data).copy()
            elif hasattr(item, "state_dict"):
                for k, v in item.state_dict().items():
                    sd[name + "." + k] = v
        return sd

    def load_state_dict(self, state_dict):
        for name, item in self._children():
            if _is_param(item):
                if name in state_dict:
                    item.data = _cast_like(state_dict[name], item)
            elif hasattr(item, "load_state_dict"):
                sub = {k.split(".", 1)[1]: v for k, v in state_dict.items() if k.startswith(name + ".")}
                item.load_state_dict(sub)


def _cast_like(value, param: Tensor):
    value = _be.to_host(value) if _be.is_device_array(value) else np.asarray(value)
    if param.device == "cpu":
        return np.array(value, dtype=param.dtype)
    return param.xp.array(value, dtype=param.dtype)


class _ModuleSeq(Module):
    """Shared list behaviour of Sequential and ModuleList (children live in ``self.modules``)."""

    def _children(self):
        return [(str(i), m) for i, m in enumerate(self.modules)]

    def forward(self, X):
        for m in self.modules:
            X = m(X)
        return X

    def to(self, device):
        for i, m in enumerate(self.modules):
            if hasattr(m, "to"):
                self.modules[i] = m.to(device)
        return self


class Sequential(_ModuleSeq):
    def __init__(self, *modules):
        self.modules = list(modules)
        self.training = True


class ModuleList(_ModuleSeq):
    def __init__(self, modules):
        self.modules = list(modules)

    def __getitem__(self, i): return self.modules[i]
    def __setitem__(self, i, m): self.modules[i] = m
    def __delitem__(self, i): del self.modules[i]
    def __len__(self): return len(self.modules)
    def __iter__(self): return iter(self.modules)
    def append(self, m): self.modules.append(m)
    def extend(self, ms): self.modules.extend(ms)
    def insert(self, i, m): self.modules.insert(i, m)
