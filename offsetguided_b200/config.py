"""Skeleton / keypoint tables the decoding path depends on.

These are *data* the decoder consumes, restated from the reference's
``config/coco_data.py`` (keypoint names :56-74, skeleton tables :12-53, hflip
table :99-116) so that the drop-in package is self-contained.  The two helper
functions reproduce ``heatmap_hflip`` (:119-128) and ``offset_hflip`` (:131-153)
but are computed from name pairs with dictionaries instead of list scans.

The CrowdPose table is builder-supplied: the reference keeps its CrowdPose
configuration on a branch that is not part of the mounted tree (README.md:133-135),
so only the *algorithm* is pinned for that configuration.
"""

COCO_KEYPOINTS = [
    'nose', 'left_eye', 'right_eye', 'left_ear', 'right_ear',
    'left_shoulder', 'right_shoulder', 'left_elbow', 'right_elbow',
    'left_wrist', 'right_wrist', 'left_hip', 'right_hip',
    'left_knee', 'right_knee', 'left_ankle', 'right_ankle',
]

# limb order matters: greedy grouping walks the list front to back
COCO_PERSON_SKELETON = [
    (0, 1), (0, 2), (1, 2), (1, 3), (2, 4), (5, 6), (4, 6), (3, 5),
    (5, 7), (7, 9), (6, 8), (8, 10), (5, 11), (6, 12), (11, 12), (11, 13),
    (13, 15), (12, 14), (14, 16)]

_REDUNDANT_TAIL = [
    (1, 5), (2, 6), (5, 12), (6, 11), (11, 14), (12, 13),
    (5, 9), (6, 10), (11, 15), (12, 16),
    (5, 0), (6, 0)]
COCO_PERSON_WITH_REDUNDANT_SKELETON = COCO_PERSON_SKELETON + _REDUNDANT_TAIL

DENSER_COCO_PERSON_SKELETON = [
    (0, 1), (0, 2), (1, 2), (0, 3), (0, 4), (3, 4), (0, 5), (0, 6), (1, 5),
    (2, 6), (1, 3), (2, 4), (3, 5), (4, 6), (5, 6), (5, 11), (6, 12), (5, 12),
    (6, 11), (11, 12), (5, 7), (6, 8), (7, 9), (8, 10), (5, 9), (6, 10), (7, 8),
    (9, 10), (9, 11), (10, 12), (9, 13), (10, 14), (13, 11), (14, 12),
    (11, 14), (12, 13), (11, 15), (12, 16), (15, 13), (16, 14),
    (13, 16), (14, 15), (13, 14), (15, 16)]

REDUNDANT_CONNECTIONS = [c for c in DENSER_COCO_PERSON_SKELETON
                         if c not in COCO_PERSON_SKELETON]

KINEMATIC_TREE_SKELETON = [
    (0, 1), (1, 3), (0, 2), (2, 4), (0, 5), (5, 7), (7, 9), (0, 6),
    (6, 8), (8, 10), (5, 11), (11, 13), (13, 15), (6, 12), (12, 14), (14, 16)]

# builder-supplied (see module docstring)
CROWDPOSE_KEYPOINTS = [
    'left_shoulder', 'right_shoulder', 'left_elbow', 'right_elbow',
    'left_wrist', 'right_wrist', 'left_hip', 'right_hip',
    'left_knee', 'right_knee', 'left_ankle', 'right_ankle',
    'head', 'neck',
]
CROWDPOSE_PERSON_SKELETON = [
    (12, 13), (13, 0), (13, 1), (0, 1), (0, 2), (2, 4), (1, 3), (3, 5),
    (0, 6), (1, 7), (6, 7), (6, 8), (8, 10), (7, 9), (9, 11)]


def _mirror_name(name):
    if name.startswith('left_'):
        return 'right_' + name[5:]
    if name.startswith('right_'):
        return 'left_' + name[6:]
    return name


def heatmap_hflip(keypoints):
    """Channel permutation that maps a W-flipped heatmap stack back onto the
    original channel order (reference config/coco_data.py:119-128)."""
    pos = {name: i for i, name in enumerate(keypoints)}
    return [pos[_mirror_name(name)] for name in keypoints]


def offset_hflip(keypoints, skeleton):
    """Limb permutation for W-flipped offset maps and the list of self-mirrored
    limbs whose un-averaged originals are restored after fusion
    (reference config/coco_data.py:131-153).

    For limb ``i = (a, b)`` the partner is the first limb whose mirrored name pair
    equals ``(a, b)``; if a limb whose mirrored pair equals ``(b, a)`` exists it
    takes precedence and ``i`` is recorded as "reserved".
    """
    names = [(keypoints[a], keypoints[b]) for a, b in skeleton]
    mirrored = [(_mirror_name(a), _mirror_name(b)) for a, b in names]
    first = {}
    for i, pair in enumerate(mirrored):
        first.setdefault(pair, i)
    flips = list(range(len(skeleton)))
    reserve = []
    for i, (a, b) in enumerate(names):
        if (a, b) in first:
            flips[i] = first[(a, b)]
        if (b, a) in first:
            flips[i] = first[(b, a)]
            reserve.append(i)
    return flips, reserve
