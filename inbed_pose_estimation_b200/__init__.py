"""B200-native batched SMPLify + SMPL body model (hot path of Inbed_pose_estimation).

Public surface mirrors the reference modules:
    SMPL, ModelOutput            <- models/smpl.py
    SMPLify                      <- smplify/smplify.py
    MaxMixturePrior              <- smplify/prior.py
    batch_rodrigues, perspective_projection, rot6d_to_rotmat, rotmat_to_rot6d, estimate_translation  <- utils/geometry.py
    rotation_matrix_to_angle_axis <- torchgeometry, as train/trainer.py:702-706 uses it
    FitsDict                     <- train/fits_dict.py
    train_losses                 <- the Trainer loss methods and post-SMPLify bookkeeping, train/trainer.py:88-178, 735-748
    constants, config            <- constants.py, config.py
"""
from . import config, constants, train_losses
from .fits_dict import FitsDict
from .geometry import (batch_rodrigues, estimate_translation, perspective_projection, rot6d_to_rotmat,
                       rotation_matrix_to_angle_axis, rotmat_to_rot6d, weak_perspective_projection)
from .prior import MaxMixturePrior
from .smpl import SMPL, ModelOutput
from .smplify import SMPLify

__all__ = ['SMPL', 'ModelOutput', 'SMPLify', 'MaxMixturePrior', 'FitsDict', 'batch_rodrigues', 'perspective_projection',
           'rot6d_to_rotmat', 'rotmat_to_rot6d', 'weak_perspective_projection', 'rotation_matrix_to_angle_axis', 'estimate_translation', 'constants', 'config', 'train_losses']
