"""tsmdet_b200 -- B200-native PointNet++ set-abstraction ops + rotated IoU/NMS.

Drop-in for the hot path of blindopen/TSM-Det-Pointcloud- (an OpenPCDet fork):

    pointnet2_batch_cuda / iou3d_nms_cuda   same function names as the reference's pybind modules
    pointnet2_utils / iou3d_nms_utils       same autograd.Function / helper API as the reference
    model_nms_utils                         post-processing NMS drivers
    pointnet2_modules                       fused layer-0 set-abstraction + feature propagation

Everything computes through ``libtsmdet_b200.so`` (hand-written sm_100a CUDA behind a C ABI).
There is no CPU or eager-PyTorch fallback: importing an op module without the built library
raises, loudly.
"""
from ._version import __version__  # noqa: F401
