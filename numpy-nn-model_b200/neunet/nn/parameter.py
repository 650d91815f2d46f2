"""``Parameter``: a Tensor that ``Module.parameters()`` collects (class-name check, like the
reference: neunet/nn/modules.py:23-39, neunet/nn/parameter.py:16-49)."""
from __future__ import annotations

from ..autograd import Tensor


class Parameter(Tensor):
    def __init__(self, data: Tensor, requires_grad=True):
        if not isinstance(data, Tensor):
            raise TypeError("Data must be a tensor")
        super().__init__(data=data.data, requires_grad=requires_grad, device=data.device, dtype=data.dtype)

    def to(self, device):
        """Always returns a NEW Parameter (so optimizers must be built after ``model.to``, as every
        reference example does)."""
        if device not in ("cpu", "cuda"):
            raise ValueError("Device must be 'cpu' or 'cuda'")
        return Parameter(Tensor(self.data, dtype=self.dtype, device=device), requires_grad=self.requires_grad)
