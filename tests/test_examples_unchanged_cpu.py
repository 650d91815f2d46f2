"""The named examples must run UNCHANGED on this package (north-star).

The model / training-loop code cells of the reference's own notebooks (examples/gpt.ipynb,
examples/ddpm.ipynb, examples/convolutional_digits_classifier.ipynb) are read from the reference tree
at test time (never copied into this repository) and exec'd verbatim against THIS `neunet`; the only
edit is the notebooks' hard-coded ``device = "cuda"`` line, which becomes "cpu" here (no GPU in the
CPU suite -- the same classes run on "cuda" in tests/test_models.py through examples/models.py).
Data-loading / tokenizer / plotting cells need the network and are not executed; synthetic batches of
the documented shapes take their place. Skipped when the reference tree is not mounted (GPU box)."""
import json
import math
import os
import re
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import pytest

import neunet
import neunet.nn as nn
from neunet import Tensor
from neunet.optim import Adam

EX = "/root/reference/examples"
pytestmark = pytest.mark.skipif(not os.path.isdir(EX), reason="reference tree not mounted")


def _code_cells(name):
    nb = json.load(open(os.path.join(EX, name)))
    return ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]


def _cpu(src):
    out, n = re.subn(r"""^device\s*=\s*['"]cuda['"]\s*$""", 'device = "cpu"', src, flags=re.M)
    return out


def _ns(**extra):
    class _NoTqdm:  # tqdm(iterable, ...) -> iterable with a no-op set_description
        def __init__(self, it=None, **kw):
            self.it = it
        def __iter__(self):
            return iter(self.it)
        def set_description(self, *a, **k):
            pass
    ns = dict(np=np, math=math, nn=nn, neunet=neunet, nnet=neunet, Tensor=Tensor, Adam=Adam, Optional=Optional,
              Tuple=Tuple, Path=Path, tqdm=_NoTqdm, device="cpu")
    ns.update(extra)
    return ns


def test_gpt_notebook_model_cells_train_unchanged():
    cells = _code_cells("gpt.ipynb")
    ns = _ns()
    exec(_cpu(cells[1]), ns)                       # device = ...
    assert ns["device"] == "cpu"
    for i in range(2, 8):                          # MultiHeadAttention ... GPT, verbatim
        exec(cells[i], ns)
    np.random.seed(0)
    V, PAD = 60, 0
    decoder = ns["Decoder"](tgt_vocab_size=V, d_model=32, n_heads=4, d_ff=64, n_layers=2, dropout=0.1, max_len=64)
    model = ns["GPT"](decoder, PAD).to(ns["device"])            # cell 11's construction, small sizes
    optimizer = Adam(model.parameters(), lr=1.5e-4, betas=(0.9, 0.98), eps=1e-9)
    loss_function = nn.CrossEntropyLoss(ignore_index=PAD)
    ns.update(model=model, optimizer=optimizer, loss_function=loss_function)
    exec(cells[12], ns)                            # train_step(dataset, epoch, epochs), verbatim
    rng = np.random.RandomState(1)
    dataset = [rng.randint(3, V, (4, 17)) for _ in range(3)]
    dataset[0][0, -5:] = PAD                       # padded tail: exercises get_pad_mask + ignore_index
    layer0 = model.decoder.layers[0]
    before_self = [p.data.copy() for p in layer0.self_attn.parameters()]
    before_cross = [p.data.copy() for p in layer0.cross_attn.parameters()]
    l1 = ns["train_step"](dataset, 0, 2)
    l2 = ns["train_step"](dataset, 1, 2)
    assert np.isfinite(l1) and np.isfinite(l2) and l1 > 0
    # the loop zeroes the gradients after every step; what shows that training happened is the parameters:
    # self-attention moved, the notebook's never-used cross_attn got no gradient and Adam skipped it (optim.py:21-22)
    assert all(not np.array_equal(a, p.data) for a, p in zip(before_self, layer0.self_attn.parameters()))
    assert all(np.array_equal(a, p.data) for a, p in zip(before_cross, layer0.cross_attn.parameters()))
    # eval() cell and state_dict round trip, as the notebook does
    exec(cells[13], ns)
    assert np.isfinite(ns["eval"](dataset))
    sd = model.state_dict()
    model.load_state_dict(sd)


def test_ddpm_notebook_model_cells_train_unchanged():
    cells = _code_cells("ddpm.ipynb")
    ns = _ns(Image=None)
    exec(_cpu(cells[2]), ns)
    for i in (3, 4, 5, 6, 7):                      # linear_schedule, Diffusion, ResBlock, PositionalEncoding, SimpleUNet
        exec(cells[i], ns)
    np.random.seed(0)
    unet = ns["SimpleUNet"](image_channels=3, image_size=8, down_channels=(8, 16, 32), up_channels=(32, 16, 8)).to(ns["device"])
    diffusion = ns["Diffusion"](model=unet, timesteps=30, beta_start=0.0001, beta_end=0.02, criterion=nn.MSELoss())
    rng = np.random.RandomState(2)
    losses = []
    for _ in range(3):                             # body of Diffusion.train's inner loop (cell 4 l.262-271)
        batch = rng.uniform(-1, 1, (5, 3, 8, 8)).astype(np.float32)
        output, noise = diffusion.forward(neunet.tensor(batch, requires_grad=True, device=ns["device"], dtype=neunet.float32))
        loss = diffusion.criterion(output, noise)
        losses.append(loss.item())
        diffusion.optimizer.zero_grad()
        loss.backward()
        diffusion.optimizer.step()
    assert tuple(output.shape) == (5, 3, 8, 8) and all(np.isfinite(losses))
    assert all(p.grad is not None for p in unet.up_layers[0].transform.parameters())   # ConvTranspose2d 4x4 s2 p1 trained


def test_conv_classifier_notebook_cells_train_unchanged():
    cells = _code_cells("convolutional_digits_classifier.ipynb")
    ns = _ns()
    exec(_cpu(cells[2]), ns)                       # device + Conv2dClassifier + classifier/loss_fn/optimizer, verbatim
    exec(cells[3], ns)                             # one_hot_encode
    rng = np.random.RandomState(3)
    ns.update(image_size=(1, 28, 28), training_dataset=rng.uniform(-1, 1, (200, 784)).astype(np.float32),
              training_targets=rng.randint(0, 10, 200))
    exec(cells[4].replace("epochs = 3", "epochs = 1"), ns)     # the training loop, verbatim but one epoch
    assert np.isfinite(ns["loss"].item())
    assert ns["classifier"].conv1.weight.grad is not None and ns["classifier"].fc1.weight.grad is not None
