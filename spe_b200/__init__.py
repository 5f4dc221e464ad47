"""spe_b200 -- B200-native (sm_100a) hot path of MingXiangL/SPE behind the reference's Python API.

`spe_b200.models` mirrors the reference's `models/` (cait, cait_backbone, position_encoding, attention,
transformer, matcher, conditional_detr) and `spe_b200.util` its `util/` (box_ops, misc); all compute is
in libspe_b200.so (C ABI in include/spe_b200.h).  There is no CPU fallback.

Around the model: `spe_b200.engine.TrainStep` (one training step, optionally one CUDA graph incl. the bucketed NCCL all-reduce and the
optimizer), `spe_b200.dp.FlatGradBuffer`, `spe_b200.optim.FlatAdamW` / `clip_grad_norm_` (main.py:177-190, engine.py:163-164),
`spe_b200.pseudo_labels` (engine.get_pseudo_label* on the device, bit-exact with cv2), `spe_b200.refine_loop.refine_iteration`
(the body of engine.train_one_epoch_refine).
"""
__version__ = "0.1.0"
