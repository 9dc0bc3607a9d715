"""crdpn_b200 -- B200-native hot path of 3DAug-Pose's contrastive distillation step.

Drop-in modules and functions behind the reference's Python surface, both thin shims over the C ABI of
``libcrdpn_b200.so`` (``include/crdpn_b200.h``, hand-written sm_100a CUDA):

* ``CRDLoss(opt)(f_s, f_t, idx, contrast_idx)`` -- CRD memory-bank NCE step (``crd.py``)
* ``ShapeEncoderPC(feature_dim)(shapes[B,3,P]) -> [B,feature_dim]`` -- the teacher's PointNet encoder
  (``pointnet.py``; reference ``auxiliary/model.py:154-180``)
* ``PointCloudSampler`` -- the encoder's input producer (``pointcloud.py``; reference ``auxiliary/dataset.py:121-150``)
* ``FrozenPoseTail`` / ``PoseTail`` -- the ``PoseEstimator`` tail as one tcgen05 chain kernel, frozen and trainable
  (``pose_tail.py``; reference ``auxiliary/model.py:183-203, 238-272``)
* the loss code either side of them (``kd_losses.py``): ``infoNCE_KD`` / ``poseNCE_KD`` (``auxiliary/model_utils.py:225-285``),
  ``CELoss`` / ``DeltaLoss`` (``auxiliary/loss.py``), ``TemperatureScaledKLDivLoss`` / ``calculate_kd_loss_new``
  (``KD/vision/vanilla/vanilla_kd.py``)

The directory name is fixed by the build harness and is not a Python identifier; load it with
``__graft_entry__.load_package()`` (registers it as ``crdpn_b200``).
There is no CPU fallback anywhere in this package.
"""
from . import _native  # noqa: F401
from .crd import AliasMethod, ContrastLoss, ContrastMemory, CRDLoss, Embed, Normalize  # noqa: F401

from .sharded import ShardedContrastMemory, ShardedCRDLoss, shard_bounds  # noqa: F401

__all__ = ["AliasMethod", "ContrastLoss", "ContrastMemory", "CRDLoss", "Embed", "Normalize",
           "ShardedContrastMemory", "ShardedCRDLoss", "shard_bounds"]
from .pointnet import ShapeEncoderPC  # noqa: F401,E402
from .pointcloud import PointCloudSampler  # noqa: F401,E402
from .pose_tail import FrozenPoseTail, PoseTail  # noqa: F401,E402
from . import kd_losses  # noqa: F401,E402
from .kd_losses import (CELoss, DeltaLoss, TemperatureScaledKLDivLoss, calculate_kd_loss_new, infoNCE, infoNCE_KD,  # noqa: F401,E402
                        multiposeNCE_KD, poseNCE, poseNCE_KD, singleinfoNCE_KD, student_kd_step_loss)

from .pipeline import GraphedStep, StepPipeline  # noqa: F401,E402

__all__ += ["StepPipeline", "GraphedStep", "ShapeEncoderPC", "PointCloudSampler", "FrozenPoseTail", "PoseTail", "CELoss", "DeltaLoss", "TemperatureScaledKLDivLoss", "calculate_kd_loss_new", "infoNCE_KD",
            "poseNCE_KD", "student_kd_step_loss", "infoNCE", "poseNCE", "singleinfoNCE_KD", "multiposeNCE_KD"]
