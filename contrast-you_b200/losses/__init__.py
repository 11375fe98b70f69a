from .contrastive import SupConLoss1, SelfPacedSupConLoss, is_normalized  # noqa: F401
from .discreteMI import IIDSegmentationLoss, IIDLoss, compute_joint_2D, compute_joint_2D_with_padding_zeros, compute_joint  # noqa: F401
from .siblings import RedundancyCriterion, PUISegLoss, IMSATLoss, IMSATDynamicWeight, imsat_loss, imsat_with_entropy  # noqa: F401
