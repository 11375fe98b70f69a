"""Projector / cluster heads feeding the InfoNCE and IIC losses (reference: contrastyou/projectors/heads.py).

Same constructor keywords and the same module tree (hence the same ``state_dict`` keys) as the reference:
``ProjectionHead._header``, ``DenseProjectionHead._projector``, ``{Cluster,DenseCluster}Head._headers`` and
``CrossCorrelationProjector._headers``.  Stock torch on purpose — these are not the graft target.
"""
import typing as t

from torch import Tensor, nn

from .nn import _ProjectorHeadBase, Flatten, Identical, Normalize, SoftmaxWithT

__all__ = ["ProjectionHead", "DenseProjectionHead", "ClusterHead", "DenseClusterHead", "CrossCorrelationProjector"]


def _tail(normalize: bool):
    return Normalize() if normalize else Identical()


def _global_stack(pool, dims: t.Sequence[int], tail: t.Sequence[nn.Module]) -> nn.Sequential:
    """pool -> flatten -> Linear(d0,d1) [-> LeakyReLU -> Linear(d1,d2) ...] -> tail   (heads.py:12-28, :44-61)"""
    layers: t.List[nn.Module] = [pool, Flatten()]
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        if i:
            layers.append(nn.LeakyReLU(0.01, inplace=True))
        layers.append(nn.Linear(a, b))
    return nn.Sequential(*layers, *tail)


def _dense_stack(dims: t.Sequence[int], tail: t.Sequence[nn.Module] = ()) -> nn.Sequential:
    """Conv1x1(d0,d1) [-> LeakyReLU -> Conv1x1(d1,d2)] -> tail   (heads.py:31-41, :64-77)"""
    layers: t.List[nn.Module] = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        if i:
            layers.append(nn.LeakyReLU(0.01, inplace=True))
        layers.append(nn.Conv2d(a, b, 1, 1, 0))
    return nn.Sequential(*layers, *tail)


class ProjectionHead(_ProjectorHeadBase):
    """global contrastive projection (heads.py:81-95): pool -> flatten -> MLP -> L2 normalise."""

    def __init__(self, *, input_dim: int, hidden_dim=256, output_dim: int, head_type: str, normalize: bool,
                 pool_name="adaptive_avg", spatial_size=(1, 1)):
        assert pool_name in ("adaptive_avg", "adaptive_max")
        super().__init__(input_dim=input_dim, output_dim=output_dim, head_type=head_type, normalize=normalize,
                         pool_name=pool_name, spatial_size=spatial_size)
        dims = (input_dim, hidden_dim, output_dim) if head_type == "mlp" else (input_dim, output_dim)
        self._header = _global_stack(self._pooling_module, dims, [_tail(normalize)])

    def forward(self, features, skip_normalize: bool = False):
        """``skip_normalize=True`` (extension) stops before the parameter-free ``Normalize`` tail so that the criterion can
        fuse the normalisation into its load stage (``SupConLoss1(normalize_input=True)``); the module tree — and with it
        the checkpoint keys — is unchanged."""
        if skip_normalize and isinstance(self._header[-1], Normalize):
            for m in list(self._header)[:-1]:
                features = m(features)
            return features
        return self._header(features)


class DenseProjectionHead(_ProjectorHeadBase):
    """pixel-wise contrastive projection (heads.py:99-123): 1x1-conv MLP -> pool to spatial_size -> normalise over C."""

    def __init__(self, *, input_dim: int, hidden_dim=128, output_dim: int, head_type: str, normalize: bool,
                 pool_name="adaptive_avg", spatial_size=(16, 16)):
        super().__init__(input_dim=input_dim, output_dim=output_dim, head_type=head_type, normalize=normalize,
                         pool_name=pool_name, spatial_size=spatial_size)
        dims = (input_dim, hidden_dim, output_dim) if head_type == "mlp" else (input_dim, output_dim)
        self._projector = _dense_stack(dims)

    def forward(self, features):
        out = self._pooling_module(self._projector(features))
        return Normalize()(out) if self._normalize else out


class ClusterHead(_ProjectorHeadBase):
    """IIC clustering on pooled features (heads.py:127-147): ``num_subheads`` x (pool, flatten, MLP, softmax/T)."""

    def __init__(self, *, input_dim: int, num_clusters=5, num_subheads=10, head_type="linear", T=1, normalize=False):
        super().__init__(input_dim=input_dim, output_dim=num_clusters, head_type=head_type, normalize=normalize,
                         pool_name="none", spatial_size=(1, 1))
        self._num_clusters, self._num_subheads, self._T = num_clusters, num_subheads, T
        dims = (input_dim, num_clusters) if head_type == "linear" else (input_dim, 128, num_clusters)
        self._headers = nn.ModuleList(
            _global_stack(nn.AdaptiveAvgPool2d((1, 1)), dims, [_tail(normalize), SoftmaxWithT(1, T=T)])
            for _ in range(num_subheads))

    def forward(self, features):
        return [h(features) for h in self._headers]


class _DenseSubheads(_ProjectorHeadBase):
    def _build(self, input_dim, hidden_dim, num_clusters, num_subheads, head_type, normalize, T):
        dims = (input_dim, num_clusters) if head_type == "linear" else (input_dim, hidden_dim, num_clusters)
        self._T = T
        self._headers = nn.ModuleList(_dense_stack(dims, [_tail(normalize), SoftmaxWithT(1, T=T)])
                                      for _ in range(num_subheads))

    def forward(self, features, skip_softmax: bool = False) -> t.List[Tensor]:
        """``skip_softmax=True`` (extension) stops before the parameter-free ``SoftmaxWithT`` tail and returns the logits, for
        ``IIDSegmentationLoss.forward_logits`` / ``forward_heads(..., logits_T=T)`` which fuse the softmax backward into the
        adjoint kernel; the module tree — and with it the checkpoint keys — is unchanged."""
        if skip_softmax:
            out = []
            for h in self._headers:
                z = features
                for m in list(h)[:-1]:
                    z = m(z)
                out.append(z)
            return out
        return [h(features) for h in self._headers]

    @property
    def temperature(self) -> float:
        return self._T


class DenseClusterHead(_DenseSubheads):
    """IIC segmentation clustering (heads.py:151-172): ``num_subheads`` x (1x1-conv [MLP], softmax over K)."""

    def __init__(self, *, input_dim: int, num_clusters=10, hidden_dim=64, num_subheads=10, T=1, head_type: str = "linear",
                 normalize: bool = False):
        super().__init__(input_dim=input_dim, output_dim=num_clusters, head_type=head_type, normalize=normalize,
                         pool_name="none", spatial_size=(1, 1))
        self._build(input_dim, hidden_dim, num_clusters, num_subheads, head_type, normalize, T)


class CrossCorrelationProjector(_DenseSubheads):
    """over-segmentation projector of the cross-correlation hooks (heads.py:176-200)."""

    def __init__(self, *, input_dim: int, num_clusters: int, head_type: str, normalize: bool, T: float = 1.0,
                 num_subheads: int = 1, hidden_dim: int = 128):
        super().__init__(input_dim=input_dim, output_dim=num_clusters, head_type=head_type, normalize=normalize,
                         pool_name="none", spatial_size=None)
        self._build(input_dim, hidden_dim, num_clusters, num_subheads, head_type, normalize, T)
