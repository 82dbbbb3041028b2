"""Data locations used by the hot path (reference config.py:96,99,101).

Paths are relative to the current working directory, as in the reference.
Only the three entries the SMPLify / SMPL path reads are mirrored.
"""
JOINT_REGRESSOR_TRAIN_EXTRA = 'data/J_regressor_extra.npy'
STATIC_FITS_DIR = 'data/static_fits'
SMPL_MODEL_DIR = 'data/smpl'
GMM_PRIOR_DIR = 'data'  # SMPLify loads data/gmm_08.pkl (reference smplify/smplify.py:32-34)
