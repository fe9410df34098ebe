"""Feature encoders of the curiosity models: parameter holders with the reference's module
structure / ``state_dict`` keys (rlpyt/models/curiosity/encoders.py:7-119) plus the ``Chain`` of
library layers that actually runs them.  ``forward`` of these modules is never called.
"""
import torch
from torch import nn

from .. import _lib as L
from .. import layers as LY
from ..layers import Chain, DenseConvLayer, conv_layer, linear_layer


class BurdaHead(nn.Module):
    """encoders.py:68-119 (batch_norm=False): conv8s4-32, conv4s2-64, conv3s1-64, Linear(3136->out)."""

    def __init__(self, image_shape, output_size=512, conv_output_size=3136, batch_norm=False):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("batch_norm=True is outside the hot-path scope (reference default is False)")
        c, h, w = image_shape
        self.image_shape = image_shape
        self.output_size = output_size
        self.conv_output_size = conv_output_size
        self.model = nn.Sequential(
            nn.Conv2d(c, 32, kernel_size=(8, 8), stride=(4, 4)), nn.LeakyReLU(),
            nn.Conv2d(32, 64, kernel_size=(4, 4), stride=(2, 2)), nn.LeakyReLU(),
            nn.Conv2d(64, 64, kernel_size=(3, 3), stride=(1, 1)), nn.LeakyReLU(),
            nn.Flatten(), nn.Linear(conv_output_size, output_size))

    def build_chain(self, trainable=True, last_act="none"):
        c, h, w = self.image_shape
        m = self.model
        l1 = conv_layer(m[0], (h, w, c), "lrelu", trainable)
        l2 = conv_layer(m[2], (l1.out_h, l1.out_w, 32), "lrelu", trainable)
        l3 = conv_layer(m[4], (l2.out_h, l2.out_w, 64), "lrelu", trainable)
        assert l3.out_h == l3.out_w and l3.out_h * l3.out_w * 64 == self.conv_output_size
        fc = linear_layer(m[7], last_act, spatial=(l3.out_h, 64), trainable=trainable)
        return Chain([l1, l2, l3, fc]), None


class UniverseHead(nn.Module):
    """encoders.py:7-35: five conv3x3 s2 p1 -> 32 + ELU; features are the CHW flatten (288)."""

    def __init__(self, image_shape, batch_norm=False):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("batch_norm=True is outside the hot-path scope")
        c, h, w = image_shape
        self.image_shape = image_shape
        seq = []
        for layer in range(5):
            seq += [nn.Conv2d(c if layer == 0 else 32, 32, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1)), nn.ELU()]
        self.model = nn.Sequential(*seq)

    def build_chain(self, trainable=True, last_act=None):
        c, h, w = self.image_shape
        layers, shape = [], (h, w, c)
        for i in range(5):
            lay = conv_layer(self.model[2 * i], shape, "elu", trainable)
            layers.append(lay)
            shape = (lay.out_h, lay.out_w, 32)
        hw = shape[0] * shape[1]
        # our activations are NHWC; the reference flattens CHW.  perm[j] = NHWC index of CHW feature j
        j = torch.arange(32 * hw)
        perm = (j % hw) * 32 + j // hw
        return Chain(layers), perm


class MazeHead(nn.Module):
    """encoders.py:37-66: conv3 s1 p1 -> 16, ReLU, conv3 s2 p2 -> 16, ReLU, Linear(256->256), ReLU."""

    def __init__(self, image_shape, output_size=256, conv_output_size=256, batch_norm=False):
        super().__init__()
        c, h, w = image_shape
        self.image_shape = image_shape
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
