"""Small versions of the reference's named example models, written against the public neunet API
only, so the SAME definitions run on the unmodified reference (oracle/make_golden.py) and on this
package (tests). Architectures follow examples/gpt.ipynb (cells 2-7), the README / notebook conv
digits classifier and examples/ddpm.ipynb (cells 5-7); sizes are scaled down so the reference
finishes in seconds. `nn` / `neunet` are passed in because both packages are called `neunet`."""
import math

import numpy as np


def build_gpt(neunet, nn, vocab=50, d_model=32, n_heads=4, d_ff=64, n_layers=2, pad_idx=0, device="cpu", dropout=0.1):
    class MultiHeadAttention(nn.Module):
        def __init__(self):
            self.scale = math.sqrt(d_model)
            self.dropout = nn.Dropout(dropout)
            self.depth = d_model // n_heads
            self.wq, self.wk, self.wv = nn.Linear(d_model, d_model), nn.Linear(d_model, d_model), nn.Linear(d_model, d_model)
            self.fc = nn.Linear(d_model, d_model)

        def forward(self, q, k, v, mask=None):
            b = q.shape[0]
            split = lambda t: t.contiguous().reshape(b, -1, n_heads, self.depth).transpose(0, 2, 1, 3)
            q, k, v = split(self.wq(q)), split(self.wk(k)), split(self.wv(v))
            scores = neunet.matmul(q, k.transpose(0, 1, 3, 2)) / self.scale
            if mask is not None:
                scores = neunet.where(mask[:, None, ...] == 0, -1e9, scores)
            attn = self.dropout(nn.Softmax(axis=-1)(scores))
            x = neunet.matmul(attn, v)
            x = x.contiguous().transpose(0, 2, 1, 3).reshape(b, -1, n_heads * self.depth)
            return self.fc(x), attn

    class FeedForward(nn.Module):
        def __init__(self):
            self.fc_1, self.fc_2 = nn.Linear(d_model, d_ff), nn.Linear(d_ff, d_model)
            self.dropout = nn.Dropout(dropout)
            self.activation = nn.Swish()

        def forward(self, x):
            return self.fc_2(self.dropout(self.activation(self.fc_1(x))))

    class DecoderLayer(nn.Module):
        def __init__(self):
            self.self_attn = MultiHeadAttention()
            self.cross_attn = MultiHeadAttention()  # constructed but never used, as in the notebook
            self.ffn = FeedForward()
            self.norm1, self.norm2 = nn.RMSNorm(d_model), nn.RMSNorm(d_model)
            self.dropout = nn.Dropout(dropout)

        def forward(self, x, mask):
            n1 = self.norm1(x)
            a, attn = self.self_attn(n1, n1, n1, mask)
            x = x + self.dropout(a)
            x = x + self.dropout(self.ffn(self.norm2(x)))
            return x, attn

    class PositionalEncoding(nn.Module):
        def __init__(self, max_len=64):
            pe = neunet.zeros(max_len, d_model, requires_grad=False)
            position = neunet.arange(0, max_len, dtype=neunet.float32)[:, None, ...]
            div_term = neunet.exp(neunet.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
            pe[:, 0::2] = neunet.sin(position * div_term)
            pe[:, 1::2] = neunet.cos(position * div_term)
            self.pe = pe[None, ...]

        def forward(self, x):
            return x + self.pe[:, : x.shape[1]]

    class Decoder(nn.Module):
        def __init__(self):
            self.token_embedding = nn.Embedding(vocab, d_model)
            self.position_embedding = PositionalEncoding()
            self.layers = nn.ModuleList([DecoderLayer() for _ in range(n_layers)])
            self.fc_out = nn.Linear(d_model, vocab)
            self.dropout = nn.Dropout(dropout)
            self.scale = math.sqrt(d_model)

        def forward(self, x, mask):
            x = self.dropout(self.position_embedding(self.token_embedding(x) * self.scale))
            for layer in self.layers:
                x, attn = layer(x, mask)
            return self.fc_out(x), attn

    class GPT(nn.Module):
        def __init__(self):
            self.decoder = Decoder()
            self.pad_idx = pad_idx

        def forward(self, ids):
            pad = (ids != self.pad_idx).astype(int)[:, np.newaxis, :]
            t = ids.shape[1]
            sub = np.logical_not(np.triu(np.ones((t, t)), k=1).astype(int))
            mask = pad & sub
            dev = self.decoder.fc_out.device
            return self.decoder(neunet.tensor(ids, dtype=neunet.int32, device=dev),
                                neunet.tensor(mask, dtype=neunet.int32, device=dev))

    return GPT().to(device)


def gpt_train_step(neunet, nn, model, optimizer, batch, pad_idx=0):
    """One step of examples/gpt.ipynb cell 12 (teacher forcing, CE over flattened logits)."""
    dev = model.decoder.fc_out.device
    loss_fn = nn.CrossEntropyLoss(ignore_index=pad_idx)
    out, _ = model.forward(batch[:, :-1])
    logits = out.reshape(out.shape[0] * out.shape[1], out.shape[2])
    loss = loss_fn(logits, neunet.tensor(batch[:, 1:].flatten(), device=dev, dtype=neunet.int32))
    loss.backward()
    optimizer.step()
    return loss, logits


def build_conv_classifier(neunet, nn, device="cpu", side=12, channels=(4, 6)):
    """Conv(1->c1) LeakyReLU MaxPool Conv(c1->c2) LeakyReLU MaxPool BatchNorm2d flatten Linear Sigmoid
    (README.md:227-258; the README / notebook model is side=28, channels=(8, 16); the default is a scaled-down copy
    the reference finishes in seconds)."""
    c1, c2 = channels

    class Net(nn.Module):
        def __init__(self):
            self.conv1 = nn.Conv2d(1, c1, 3, 1, 1)
            self.conv2 = nn.Conv2d(c1, c2, 3, 1, 1)
            self.act = nn.LeakyReLU()
            self.pool = nn.MaxPool2d(2, 2)
            self.bn = nn.BatchNorm2d(c2)
            self.fc = nn.Linear(c2 * (side // 4) ** 2, 10)
            self.out = nn.Sigmoid()

        def forward(self, x):
            x = self.pool(self.act(self.conv1(x)))
            x = self.pool(self.act(self.conv2(x)))
            x = self.bn(x)
            x = x.reshape(x.shape[0], -1)
            return self.out(self.fc(x))

    return Net().to(device)


def build_unet(neunet, nn, device="cpu", ch=(8, 16), temb=8):
    """Two-level version of examples/ddpm.ipynb's SimpleUNet: input conv, a down ResBlock
    (Conv2d 4x4 s2 p1 transform), an up ResBlock on the skip-concatenated input (ConvTranspose2d
    4x4 s2 p1 transform), ConvTranspose2d 3x3 output; time embedding through a Linear."""
    class ResBlock(nn.Module):
        def __init__(self, cin, cout, up):
            self.time = nn.Linear(temb, cout)
            self.conv1 = nn.Conv2d(2 * cin if up else cin, cout, 3, 1, 1)
            self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
            self.transform = nn.ConvTranspose2d(cout, cout, 4, 2, 1) if up else nn.Conv2d(cout, cout, 4, 2, 1)
            self.bn1, self.bn2 = nn.BatchNorm2d(cout), nn.BatchNorm2d(cout)
            self.act = nn.LeakyReLU()

        def forward(self, x, t):
            h = self.bn1(self.act(self.conv1.forward(x)))
            h = h + self.act(self.time.forward(t))[:, :, None, None]
            h = self.bn2(self.act(self.conv2.forward(h)))
            return self.transform.forward(h)

    class UNet(nn.Module):
        def __init__(self):
            self.time_mlp = nn.Sequential(nn.Linear(temb, temb), nn.LeakyReLU())
            self.input_conv = nn.Conv2d(3, ch[0], 3, 1, 1)
            self.down = ResBlock(ch[0], ch[1], up=False)
            self.up = ResBlock(ch[1], ch[0], up=True)
            self.output_conv = nn.ConvTranspose2d(ch[0], 3, 3, 1, 1)

        def forward(self, x, temb_in):
            t = self.time_mlp(temb_in)
            x = self.input_conv(x)
            d = self.down(x, t)
            u = self.up(neunet.concatenate(d, d, axis=1), t)
            return self.output_conv(u)

    return UNet().to(device)


def build_ddpm_unet(neunet, nn, device="cpu", image_channels=3, image_size=32, down_channels=(128, 256, 512, 1024),
                    up_channels=(1024, 512, 256, 128), time_emb_dim=32):
    """examples/ddpm.ipynb cells 5-7 verbatim in structure (ResBlock, sinusoidal time encoding, SimpleUNet): for a
    power-of-two image size the input layer is Conv2d 3x3 p1, every down block ends in Conv2d 4x4 s2 p1, every up block
    takes the skip-concatenated input and ends in ConvTranspose2d 4x4 s2 p1, the output layer is ConvTranspose2d 3x3 p1.
    Defaults = cell 8 (3x32x32, down (128,256,512,1024), up (1024,512,256,128))."""
    class ResBlock(nn.Module):
        def __init__(self, cin, cout, up=False):
            self.time_embedding = nn.Linear(time_emb_dim, cout)
            if up:
                self.conv1 = nn.Conv2d(2 * cin, cout, kernel_size=(3, 3), padding=(1, 1))
                self.transform = nn.ConvTranspose2d(cout, cout, kernel_size=(4, 4), stride=(2, 2), padding=(1, 1))
            else:
                self.conv1 = nn.Conv2d(cin, cout, kernel_size=(3, 3), padding=(1, 1))
                self.transform = nn.Conv2d(cout, cout, kernel_size=(4, 4), stride=(2, 2), padding=(1, 1))
            self.conv2 = nn.Conv2d(cout, cout, kernel_size=(3, 3), padding=(1, 1))
            self.relu1, self.relu2, self.relu3 = nn.LeakyReLU(alpha=0.01), nn.LeakyReLU(alpha=0.01), nn.LeakyReLU(alpha=0.01)
            self.bnorm1 = nn.BatchNorm2d(cout, momentum=0.1, eps=1e-5)
            self.bnorm2 = nn.BatchNorm2d(cout, momentum=0.1, eps=1e-5)

        def forward(self, x, t):
            x = self.conv1.forward(x)
            h = self.relu1.forward(x)
            h = self.bnorm1.forward(h)
            t = self.time_embedding.forward(t)
            time_emb = self.relu2.forward(t)
            time_emb = time_emb[(...,) + (None,) * 2]
            h = h + time_emb
            h = self.conv2.forward(h)
            h = self.relu3.forward(h)
            h = self.bnorm2.forward(h)
            return self.transform.forward(h)

    class TimeEncoding(nn.Module):
        def __init__(self, max_len, d_model):
            pe = np.zeros((max_len, d_model))
            position = np.arange(0, max_len)[:, np.newaxis]
            div_term = np.exp(np.arange(0, d_model, 2) * (-np.log(10000.0) / d_model))
            pe[:, 0::2] = np.sin(position * div_term)
            pe[:, 1::2] = np.cos(position * div_term)
            self.pe = neunet.tensor(pe[:, np.newaxis, :].astype(np.float32), requires_grad=False, device=device)

        def forward(self, x):
            return x + self.pe[: x.shape[0], :]

    class SimpleUNet(nn.Module):
        def __init__(self):
            self.time_embedding = nn.Sequential(TimeEncoding(1000, time_emb_dim), nn.Linear(time_emb_dim, time_emb_dim),
                                                nn.LeakyReLU())
            if image_size & (image_size - 1) != 0:
                self.input_conv = nn.ConvTranspose2d(image_channels, down_channels[0], kernel_size=(5, 5))
                self.output_conv = nn.Conv2d(up_channels[-1], image_channels, kernel_size=(5, 5))
            else:
                self.input_conv = nn.Conv2d(image_channels, down_channels[0], kernel_size=(3, 3), padding=(1, 1))
                self.output_conv = nn.ConvTranspose2d(up_channels[-1], image_channels, kernel_size=(3, 3), padding=(1, 1))
            self.down_layers = nn.ModuleList([ResBlock(down_channels[i], down_channels[i + 1])
                                              for i in range(len(down_channels) - 1)])
            self.up_layers = nn.ModuleList([ResBlock(up_channels[i], up_channels[i + 1], up=True)
                                            for i in range(len(up_channels) - 1)])

        def forward(self, x, t):
            if not isinstance(t, neunet.Tensor):
                t = neunet.tensor(np.asarray(t, dtype=np.float32)[:, None, None], requires_grad=False, device=x.device)
            t = self.time_embedding.forward(t)
            t = t.reshape(t.shape[0], -1)
            x = self.input_conv.forward(x)
            residual_inputs = []
            for down_layer in self.down_layers:
                x = down_layer.forward(x, t)
                residual_inputs.append(x)
            for up_layer in self.up_layers:
                residual_x = residual_inputs.pop()
                x = neunet.concatenate(*(x, residual_x), axis=1)
                x = up_layer.forward(x, t)
            return self.output_conv.forward(x)

    return SimpleUNet().to(device)


def ddpm_train_step(neunet, nn, model, optimizer, x0, noise, t_frac, a_bar_sqrt, one_minus_a_bar_sqrt):
    """One step of examples/ddpm.ipynb cell 4 (Algorithm 1): x_t = sqrt(a_bar_t) x0 + sqrt(1 - a_bar_t) eps (per-sample
    coefficients given as arrays broadcast over (B,1,1,1)), MSE between the predicted and the true noise."""
    x_t = a_bar_sqrt * x0 + one_minus_a_bar_sqrt * noise
    pred = model.forward(neunet.tensor(x_t, requires_grad=False, device=x0.device), t_frac)
    loss = nn.MSELoss()(pred, noise)
    loss.backward()
    optimizer.step()
    return loss, pred
