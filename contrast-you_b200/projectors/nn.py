"""Parameter-free building blocks of the projector heads (reference: contrastyou/projectors/nn.py:24-85).

The heads are the *feeders* of the loss kernels and stay on stock torch (tiny GEMMs / 1x1 convolutions, SURVEY.md
§8a8-a10).  Module classes and their positions inside the ``nn.Sequential`` containers are kept so that checkpoints
written by the reference (keys like ``_header.2.weight``, ``_headers.0.0.bias``) load with ``strict=True``.
"""
from torch import Tensor, nn
from torch.nn import functional as F
from torch.nn.modules.utils import _pair  # noqa

HEAD_TYPES = ("mlp", "linear")
POOL_NAMES = ("adaptive_avg", "adaptive_max", "identical", "none")


class Flatten(nn.Module):
    def forward(self, features: Tensor) -> Tensor:
        return features.view(features.shape[0], -1)


class Identical(nn.Module):
    def forward(self, input):  # noqa
        return input


class Normalize(nn.Module):
    """L2-normalise along ``dim`` (nn.py:47-54)."""

    def __init__(self, dim=1) -> None:
        super().__init__()
        self._dim = dim

    def forward(self, input):  # noqa
        return F.normalize(input, p=2, dim=self._dim)


class SoftmaxWithT(nn.Softmax):
    """softmax(x / T); the division is in place like the reference (nn.py:36-44)."""

    def __init__(self, dim, T: float = 1.0) -> None:
        super().__init__(dim)
        self._T = T

    def forward(self, input: Tensor) -> Tensor:  # noqa
        input /= self._T
        return super().forward(input)


def make_pool(pool_name, spatial_size):
    if pool_name == "adaptive_avg":
        return nn.AdaptiveAvgPool2d(spatial_size)
    if pool_name == "adaptive_max":
        return nn.AdaptiveMaxPool2d(spatial_size)
    if pool_name in (None, "none", "identical"):
        return Identical()
    raise KeyError(pool_name)


class _ProjectorHeadBase(nn.Module):
    def __init__(self, *, input_dim: int, output_dim: int, head_type: str, normalize: bool, pool_name="adaptive_avg",
                 spatial_size=(1, 1)):
        super().__init__()
        assert head_type in HEAD_TYPES, head_type
        assert pool_name in POOL_NAMES, pool_name
        self._input_dim, self._output_dim = input_dim, output_dim
        self._head_type, self._normalize, self._pool_name = head_type, normalize, pool_name
        self._spatial_size = _pair(spatial_size)
        self._pooling_module = make_pool(pool_name, self._spatial_size)
