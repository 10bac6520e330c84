"""Caller-side tail of the decoding path: pose back-projection into the
original image frame and COCO result rows, vectorised.

Replaces the per-image / per-person / per-keypoint Python loops of the reference
(transforms/preprocess.py:33-63 ``annotations_inverse`` and evaluate.py:227-265), which
become the bottleneck once decoding itself takes microseconds.  Host-side numpy: these are
a few hundred floats per image that already live in pinned host memory.
"""
import numpy as np


def annotations_inverse(keypoints, meta):
    """Back-project the poses of one image into the original image space
    (reference transforms/preprocess.py:33-63): undo the padding shift, then the resize;
    keypoint scales are divided by sqrt(sx * sy).  Returns a new array.

    x and y are handled together; each step is evaluated in float64 and rounded back to the
    array's dtype, exactly like the reference's four in-place updates."""
    if meta['hflip']:
        raise NotImplementedError('horizontally flipped inputs are not back-projected '
                                  '(the reference raises here as well)')
    out = np.array(keypoints, copy=True)
    shift = np.asarray(meta['offset'], dtype=np.float64)[:2]
    zoom = np.asarray(meta['scale'], dtype=np.float64)[:2]
    shifted = (out[:, :, :2] + shift).astype(out.dtype)
    out[:, :, :2] = shifted / zoom
    out[:, :, 3] = out[:, :, 3] / np.sqrt(np.prod(meta['scale']))
    return out


def coco_results(batch_poses, metas):
    """COCO keypoint result rows of a decoded batch (reference evaluate.py:227-265).

    For every image the poses are back-projected, x / y rounded to two decimals, and each
    person becomes {'image_id', 'category_id': 1, 'keypoints': [x, y, flag] * C, 'score'} with
    flag = 1 when x > 0 or y > 0 and score = mean keypoint score over all C keypoints; an
    image without persons gets one all-zero annotation with score 0.01.
    Returns (result_rows, image_ids, back_projected_poses)."""
    rows, image_ids, projected = [], [], []
    for poses, meta in zip(batch_poses, metas):
        subset = annotations_inverse(poses, meta)
        projected.append(subset)
        image_id = meta['image_id']
        image_ids.append(image_id)
        subset[:, :, :2] = np.around(subset[:, :, :2], 2)
        c = subset.shape[1]
        if len(subset):
            sub64 = subset.astype(float)
            xy = sub64[:, :, :2]
            flags = ((xy[:, :, 0] > 0) | (xy[:, :, 1] > 0)).astype(float)
            triples = np.concatenate((xy, flags[:, :, None]), axis=2).reshape(len(subset), 3 * c)
            # sum(v) / len(v) with Python's left-to-right float64 accumulation
            scores = np.zeros(len(subset))
            for j in range(c):
                scores = scores + sub64[:, j, 2]
            scores = scores / c
            for person, score in zip(triples, scores):
                kps = person.tolist()
                for j in range(2, 3 * c, 3):
                    kps[j] = int(kps[j])
                rows.append({'image_id': image_id, 'category_id': 1, 'keypoints': kps, 'score': float(score)})
        else:
            rows.append({'image_id': image_id, 'category_id': 1,
                         'keypoints': np.zeros((c * 3,)).tolist(), 'score': 0.01})
    return rows, image_ids, projected


def frames_of(metas):
    """(n, 4) float64 frames (offset_x, offset_y, scale_x, scale_y) of a batch's metas, the argument
    of DecoderEngine.stage_frames / the ``frames`` of decode_features."""
    out = np.empty((len(metas), 4), dtype=np.float64)
    for i, meta in enumerate(metas):
        if meta['hflip']:
            raise NotImplementedError('horizontally flipped inputs are not back-projected '
                                      '(the reference raises here as well)')
        out[i, 0:2] = np.asarray(meta['offset'], dtype=np.float64)[:2]
        out[i, 2:4] = np.asarray(meta['scale'], dtype=np.float64)[:2]
    return out


def result_arrays(gpu_rows, n_keypoints):
    """The batch's result rows as arrays, from what the grouping kernel wrote (engine.last_result_rows):
    rows ordered by image, then by person rank; an image without persons gets the reference's
    all-zero row with score 0.01 (evaluate.py:258-265).
    Returns (keypoints (T, 3C) float32, scores (T,) float64, image_index (T,) int64)."""
    kp, sc, im, counts = gpu_rows
    order = np.argsort(im, kind='stable')              # packed rows of one image are contiguous, in rank order
    out_counts = np.maximum(counts, 1)
    image_index = np.repeat(np.arange(len(counts)), out_counts)
    real = np.repeat(counts > 0, out_counts)
    keypoints = np.zeros((len(image_index), 3 * n_keypoints), dtype=np.float32)
    scores = np.full((len(image_index),), 0.01, dtype=np.float64)
    keypoints[real] = kp[order]
    scores[real] = sc[order]
    return keypoints, scores, image_index


def rows_from_arrays(keypoints, scores, image_index, metas):
    """COCO result dictionaries (evaluate.py:244-265) from result_arrays."""
    ids = [meta['image_id'] for meta in metas]
    kp64 = keypoints.astype(np.float64)
    rows = []
    for person, score, img in zip(kp64.tolist(), scores.tolist(), image_index.tolist()):
        for j in range(2, len(person), 3):
            person[j] = int(person[j])
        rows.append({'image_id': ids[img], 'category_id': 1, 'keypoints': person, 'score': score})
    return rows
