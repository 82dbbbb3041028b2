"""Integer tables and scalar constants of the SMPLify / SMPL hot path.

Mirrors the names the reference exposes in ``constants.py`` (reference
constants.py:1-2 FOCAL_LENGTH / IMG_RES, :40-92 JOINT_NAMES, :95 JOINT_IDS,
:98-116 JOINT_MAP, :127-132 SMPL_POSE_FLIP_PERM, :134-137 J24/J49 flip
permutations).  The tables are integer data and must be bit-exact; they are
built here from one compact (name, smpl-index) list instead of three separate
literals so that the name order, the id dictionary and the map can never
drift apart.  The same tables are compiled into the CUDA library
(csrc/tables.h) and ``tests/test_tables.py`` checks that both agree.
"""

FOCAL_LENGTH = 5000.
IMG_RES = 224

# (joint name, index into the 54 "SMPL + selected-vertex + extra-regressor" joints)
# First 25 rows: OpenPose BODY_25 order.  Last 24 rows: the ground-truth superset.
_JOINT_TABLE = (
    ('OP Nose', 24), ('OP Neck', 12), ('OP RShoulder', 17), ('OP RElbow', 19),
    ('OP RWrist', 21), ('OP LShoulder', 16), ('OP LElbow', 18), ('OP LWrist', 20),
    ('OP MidHip', 0), ('OP RHip', 2), ('OP RKnee', 5), ('OP RAnkle', 8),
    ('OP LHip', 1), ('OP LKnee', 4), ('OP LAnkle', 7), ('OP REye', 25),
    ('OP LEye', 26), ('OP REar', 27), ('OP LEar', 28), ('OP LBigToe', 29),
    ('OP LSmallToe', 30), ('OP LHeel', 31), ('OP RBigToe', 32), ('OP RSmallToe', 33),
    ('OP RHeel', 34),
    ('Right Ankle', 8), ('Right Knee', 5), ('Right Hip', 45), ('Left Hip', 46),
    ('Left Knee', 4), ('Left Ankle', 7), ('Right Wrist', 21), ('Right Elbow', 19),
    ('Right Shoulder', 17), ('Left Shoulder', 16), ('Left Elbow', 18), ('Left Wrist', 20),
    ('Neck (LSP)', 47), ('Top of Head (LSP)', 48), ('Pelvis (MPII)', 49),
    ('Thorax (MPII)', 50), ('Spine (H36M)', 51), ('Jaw (H36M)', 52), ('Head (H36M)', 53),
    ('Nose', 24), ('Left Eye', 26), ('Right Eye', 25), ('Left Ear', 28), ('Right Ear', 27),
)

JOINT_NAMES = [name for name, _ in _JOINT_TABLE]
JOINT_IDS = {name: i for i, name in enumerate(JOINT_NAMES)}
JOINT_MAP = {name: smpl_idx for name, smpl_idx in _JOINT_TABLE}

NUM_JOINTS_OUT = len(JOINT_NAMES)            # 49
NUM_SMPL_JOINTS = 24
NUM_BETAS = 10
NUM_VERTS = 6890
NUM_POSE_FEATURES = 9 * (NUM_SMPL_JOINTS - 1)  # 207

# SMPL kinematic tree (kintree_table[0] of the model file; parent of the root is -1).
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# Vertices smplx's VertexJointSelector appends to the 24 chain joints
# (5 face, 6 feet, 10 finger tips) -> joints 24..44 of the 54.
SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583,
                         3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673,
                         6191, 5782, 5905, 6016, 6133]

# Joints ignored in the body-fitting stage (reference smplify/smplify.py:28-29).
SMPLIFY_IGNORED_JOINTS = [JOINT_IDS[n] for n in
                          ('OP Neck', 'OP RHip', 'OP LHip', 'Right Hip', 'Left Hip')]
# Torso joints of the camera-fitting stage (reference smplify/losses.py:72-75).
CAMERA_OP_JOINTS = [JOINT_IDS[n] for n in ('OP RHip', 'OP LHip', 'OP RShoulder', 'OP LShoulder')]
CAMERA_GT_JOINTS = [JOINT_IDS[n] for n in ('Right Hip', 'Left Hip', 'Right Shoulder', 'Left Shoulder')]
# body_pose entries of the knee / elbow angle prior (reference smplify/losses.py:24).
ANGLE_PRIOR_IDS = [55 - 3, 58 - 3, 12 - 3, 15 - 3]
ANGLE_PRIOR_SIGNS = [1., -1., -1., -1.]

# Left/right swap of the 24 SMPL joints and the derived 72-entry pose permutation.
SMPL_JOINTS_FLIP_PERM = [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16,
                         19, 18, 21, 20, 23, 22]
SMPL_POSE_FLIP_PERM = [3 * j + c for j in SMPL_JOINTS_FLIP_PERM for c in range(3)]
J24_FLIP_PERM = [5, 4, 3, 2, 1, 0, 11, 10, 9, 8, 7, 6, 12, 13, 14, 15, 16, 17, 18, 19,
                 21, 20, 23, 22]
J49_FLIP_PERM = ([0, 1, 5, 6, 7, 2, 3, 4, 8, 12, 13, 14, 9, 10, 11, 16, 15, 18, 17,
                  22, 23, 24, 19, 20, 21] + [25 + i for i in J24_FLIP_PERM])
