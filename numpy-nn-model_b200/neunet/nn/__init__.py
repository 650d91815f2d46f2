from .activations import *  # noqa: F401,F403
from .activations import GELU, LeakyReLU, LogSoftmax, ReLU, Sigmoid, Softmax, Softplus, Swish, Tanh  # noqa: F401
from .layers import *  # noqa: F401,F403
from .layers import (BatchNorm2d, Conv2d, ConvTranspose2d, Dropout, Embedding, Flatten, LayerNorm, Linear,  # noqa: F401
                     LinearSwish, MaxPool2d, RMSNorm)
from .losses import BCELoss, CrossEntropyLoss, L1Loss, MSELoss, NLLLoss  # noqa: F401
from .modules import Module, ModuleList, Sequential  # noqa: F401
from .parameter import Parameter  # noqa: F401
