from .activations import *  # noqa: F401,F403
from .activations import GELU, LeakyReLU, LogSoftmax, ReLU, Sigmoid, Softmax, Softplus, Swish, Tanh  # noqa: F401
from .layers import *  # noqa: F401,F403
from .layers import (BatchNorm2d, Conv2d, ConvTranspose2d, Dropout, Embedding, Flatten, LayerNorm, Linear,  # noqa: F401
                     LinearSwish, MaxPool2d, RMSNorm)
from .losses import BCELoss, CrossEntropyLoss, L1Loss, MSELoss, NLLLoss  # noqa: F401
from .modules import Module, ModuleList, Sequential  # noqa: F401
from .parameter import Parameter  # noqa: F401

# Names the reference's `neunet.nn` exports that are outside this repository's scope (the dense
# forward/backward hot path and the layers its named configs need). Asking for one says so instead of
# failing with a bare AttributeError.
_OUT_OF_SCOPE = frozenset({
    "BatchNorm1d", "RNN", "LSTM", "GRU", "Bidirectional", "AvgPool2d", "ZeroPad2d", "KLDivLoss", "Mish", "ELU", "SELU",
    "Softsign", "TanhExp", "Softmax2d", "Swiglu", "Tanhshrink"})


class _OutOfScope(NotImplementedError, AttributeError):
    """Also an AttributeError, so ``hasattr`` / ``getattr(..., default)`` keep working."""


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise _OutOfScope(
            f"neunet.nn.{name} exists in the reference (AkiRusProd/numpy-nn-model) but is outside the hot path this package "
            "implements (see DESIGN.md section 8); use the reference package for it.")
    raise AttributeError(f"module 'neunet.nn' has no attribute {name!r}")
