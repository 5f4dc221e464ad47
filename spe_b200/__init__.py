"""spe_b200 -- B200-native (sm_100a) hot path of MingXiangL/SPE behind the reference's Python API.

`spe_b200.models` mirrors the reference's `models/` (cait, cait_backbone, position_encoding, attention,
transformer, matcher, conditional_detr) and `spe_b200.util` its `util/` (box_ops, misc); all compute is
in libspe_b200.so (C ABI in include/spe_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
