from .heads import ProjectionHead, DenseProjectionHead, ClusterHead, DenseClusterHead, CrossCorrelationProjector  # noqa: F401
from .nn import Normalize, SoftmaxWithT, Flatten, Identical  # noqa: F401
