"""B200-native batched SMPLify + SMPL body model (hot path of Inbed_pose_estimation).

Public surface mirrors the reference modules:
    SMPL, ModelOutput            <- models/smpl.py
    SMPLify                      <- smplify/smplify.py
    MaxMixturePrior              <- smplify/prior.py
    batch_rodrigues, perspective_projection  <- utils/geometry.py
    constants, config            <- constants.py, config.py
"""
from . import config, constants
from .geometry import batch_rodrigues, perspective_projection
from .prior import MaxMixturePrior
from .smpl import SMPL, ModelOutput
from .smplify import SMPLify

__all__ = ['SMPL', 'ModelOutput', 'SMPLify', 'MaxMixturePrior', 'batch_rodrigues', 'perspective_projection',
           'constants', 'config']
