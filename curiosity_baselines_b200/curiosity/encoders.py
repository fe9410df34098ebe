"""Feature encoders of the curiosity models: parameter holders with the reference's module
structure / ``state_dict`` keys (rlpyt/models/curiosity/encoders.py:7-119) plus the ``Chain`` of
library layers that actually runs them.  ``forward`` of these modules is never called.
"""
import torch
from torch import nn

from .. import _lib as L
from .. import layers as LY
from ..layers import BatchNormLayer, Chain, DenseConvLayer, conv_layer, linear_layer


class BurdaHead(nn.Module):
    """encoders.py:68-119: conv8s4-32, conv4s2-64, conv3s1-64, each followed by LeakyReLU (and BatchNorm2d when
    ``batch_norm``), Linear(3136->out).  The Sequential indices -- hence the ``state_dict`` keys -- are the
    reference's in both variants."""

    def __init__(self, image_shape, output_size=512, conv_output_size=3136, batch_norm=False):
        super().__init__()
        c, h, w = image_shape
        self.image_shape = image_shape
        self.output_size = output_size
        self.conv_output_size = conv_output_size
        self.batch_norm = bool(batch_norm)
        seq = []
        for cin, cout, k, s in ((c, 32, 8, 4), (32, 64, 4, 2), (64, 64, 3, 1)):
            seq += [nn.Conv2d(cin, cout, kernel_size=(k, k), stride=(s, s)), nn.LeakyReLU()]
            if batch_norm:
                seq.append(nn.BatchNorm2d(cout))
        seq += [nn.Flatten(), nn.Linear(conv_output_size, output_size)]
        self.model = nn.Sequential(*seq)

    def build_chain(self, trainable=True, last_act="none"):
        c, h, w = self.image_shape
        m = self.model
        step = 3 if self.batch_norm else 2
        layers, shape = [], (h, w, c)
        for i in range(3):
            lay = conv_layer(m[step * i], shape, "lrelu", trainable)
            layers.append(lay)
            shape = (lay.out_h, lay.out_w, lay.out_c)
            if self.batch_norm:
                layers.append(BatchNormLayer(m[step * i + 2], shape, trainable))
        assert shape[0] == shape[1] and shape[0] * shape[1] * 64 == self.conv_output_size
        fc = linear_layer(m[3 * step + 1], last_act, spatial=(shape[0], 64), trainable=trainable)
        return Chain(layers + [fc]), None


class UniverseHead(nn.Module):
    """encoders.py:7-35: five conv3x3 s2 p1 -> 32 + ELU (+ BatchNorm2d when ``batch_norm``); features are the CHW
    flatten (288)."""

    def __init__(self, image_shape, batch_norm=False):
        super().__init__()
        c, h, w = image_shape
        self.image_shape = image_shape
        self.batch_norm = bool(batch_norm)
        seq = []
        for layer in range(5):
            seq += [nn.Conv2d(c if layer == 0 else 32, 32, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1)), nn.ELU()]
            if batch_norm:
                seq.append(nn.BatchNorm2d(32))
        self.model = nn.Sequential(*seq)

    def build_chain(self, trainable=True, last_act=None):
        c, h, w = self.image_shape
        step = 3 if self.batch_norm else 2
        layers, shape = [], (h, w, c)
        for i in range(5):
            lay = conv_layer(self.model[step * i], shape, "elu", trainable)
            layers.append(lay)
            shape = (lay.out_h, lay.out_w, 32)
            if self.batch_norm:
                layers.append(BatchNormLayer(self.model[step * i + 2], shape, trainable))
        hw = shape[0] * shape[1]
        # our activations are NHWC; the reference flattens CHW.  perm[j] = NHWC index of CHW feature j
        j = torch.arange(32 * hw)
        perm = (j % hw) * 32 + j // hw
        return Chain(layers), perm


class MazeHead(nn.Module):
    """encoders.py:37-66: conv3 s1 p1 -> 16, ReLU, conv3 s2 p2 -> 16, ReLU, Linear(256->256), ReLU."""

    def __init__(self, image_shape, output_size=256, conv_output_size=256, batch_norm=False):
        super().__init__()                # (the reference's MazeHead accepts batch_norm and ignores it, encoders.py:41-58)
        c, h, w = image_shape
        self.image_shape = image_shape
        self.batch_norm = False
        self.output_size = output_size
        self.conv_output_size = conv_output_size
        self.model = nn.Sequential(
            nn.Conv2d(c, 16, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1)), nn.ReLU(),
            nn.Conv2d(16, 16, kernel_size=(3, 3), stride=(2, 2), padding=(2, 2)), nn.ReLU(),
            nn.Flatten(), nn.Linear(conv_output_size, output_size), nn.ReLU())

    def build_chain(self, trainable=True, last_act="relu"):
        c, h, w = self.image_shape
        m = self.model
        if LY.get_engine() != L.ENGINE_SIMT and h * w * c <= 1024:
            # tiny grid: both convolutions as dense (block-Toeplitz) Linear layers on the tcgen05 GEMM
            l1 = DenseConvLayer(m[0], (h, w, c), "relu", trainable)
            l2 = DenseConvLayer(m[2], (l1.conv_out_h, l1.conv_out_w, 16), "relu", trainable, in_features=l1.out_c)
            assert l2.conv_out_h == l2.conv_out_w and l2.real_out == self.conv_output_size and l2.out_c == l2.real_out
            fc = linear_layer(m[5], "relu", spatial=(l2.conv_out_h, 16), trainable=trainable)
            return Chain([l1, l2, fc]), None
        l1 = conv_layer(m[0], (h, w, c), "relu", trainable)
        l2 = conv_layer(m[2], (l1.out_h, l1.out_w, 16), "relu", trainable)
        assert l2.out_h == l2.out_w and l2.out_h * l2.out_w * 16 == self.conv_output_size
        fc = linear_layer(m[5], "relu", spatial=(l2.out_h, 16), trainable=trainable)
        return Chain([l1, l2, fc]), None


def make_encoder(feature_encoding, image_shape, batch_norm=False):
    """(module, feature_size) as icm.py:77-86 / disagreement.py:79-88 / ndigo.py:65-74."""
    if feature_encoding == "idf":
        return UniverseHead(image_shape=image_shape, batch_norm=batch_norm), 288
    if feature_encoding == "idf_burda":
        return BurdaHead(image_shape=image_shape, output_size=512, batch_norm=batch_norm), 512
    if feature_encoding == "idf_maze":
        return MazeHead(image_shape=image_shape, output_size=256, batch_norm=batch_norm), 256
    raise ValueError(f"unsupported feature_encoding {feature_encoding!r}")


def frame_input(obs):
    """uint8 [N,C,H,W] frames read in place: (pointer, dtype code, element strides n,h,w,c)."""
    if not obs.is_cuda:
        raise L.CuriosityLibError("observations must be on a CUDA device (no CPU fallback)")
    obs = obs.contiguous()
    n, c, h, w = obs.shape
    if obs.dtype == torch.uint8 and c * h * w > 1024:
        return obs, obs.data_ptr(), L.U8, (c * h * w, w, 1, h * w)
    # float observations (PyColab {0,1} masks) and tiny uint8 grids: stage once as bf16 NHWC (exact for integers <= 256)
    x = obs.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
    return x, x, L.BF16, None
