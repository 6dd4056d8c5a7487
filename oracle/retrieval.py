"""All-pairs distance + ranking of evaluate_retrieval.py:22-73 restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two flavours:
  * `pairwise_dist_ref32`: float32 numpy, the very expressions of the reference
    (evaluate_retrieval.py:56-63; numexpr's 'A + B - 2 * C' is elementwise fp32).
  * `pairwise_dist64`: float64, used as the exact answer when judging which GPU
    rank positions are tie-free (SURVEY.md section 7 hard part 5).
`rank_stable` is the ordering definition of the build: ascending distance, ties
broken by ascending index (np.argsort(kind='stable')); the reference's
`np.argsort` (introsort, evaluate_retrieval.py:67) agrees wherever distances are distinct.
"""
import numpy as np


def normalize_rows(f):
    return f / np.linalg.norm(f, axis=-1, keepdims=True)            # :58


def pairwise_dist_ref32(features, normalize=False):
    f = np.array(features, dtype=np.float32, copy=True)
    if normalize:
        f /= np.linalg.norm(f, axis=-1, keepdims=True)              # :58
        return -np.dot(f, f.T)                                      # :59
    sq = np.sum(f ** 2, axis=-1)                                    # :61
    return sq[:, None] + sq[None, :] - 2 * np.dot(f, f.T)           # :62


def pairwise_dist64(features, normalize=False):
    f = np.asarray(features, dtype=np.float32).astype(np.float64)
    if normalize:
        # the reference normalises in float32 (in place, :58) before the product
        f32 = np.array(features, dtype=np.float32, copy=True)
        f32 /= np.linalg.norm(f32, axis=-1, keepdims=True)
        f = f32.astype(np.float64)
        return -np.dot(f, f.T)
    sq = np.sum(f ** 2, axis=-1)
    return sq[:, None] + sq[None, :] - 2 * np.dot(f, f.T)


def rank_stable(pdist):
    return np.argsort(pdist, axis=-1, kind='stable')                # :67 + defined tie-break


def tie_free_prefix(pdist64, ranking, tol):
    """For each query the number of leading rank positions whose distance gap to the next
    position exceeds `tol` (positions where a kernel with abs error < tol/2 must agree)."""
    d = np.take_along_axis(pdist64, ranking, axis=-1)
    gaps = np.diff(d, axis=-1)
    ok = gaps > tol
    bad = ~ok
    first_bad = np.where(bad.any(axis=-1), bad.argmax(axis=-1), ok.shape[-1])
    return first_bad


def pairwise_retrieval(features, normalize=False):
    """dict id -> ranked id list, like evaluate_retrieval.pairwise_retrieval(..., return_generator=False)
    for ndarray / dict / {'feat': dict} inputs (:43-54, :69-73)."""
    ind2id = None
    if isinstance(features, dict):
        if 'feat' in features:
            features = features['feat']
        ind2id = np.array(list(features.keys()))
        features = np.stack(list(features.values()))
        if features.ndim > 2:
            raise ValueError('Feature matrix must be 2-dimensional. Actual shape: {}'.format(features.shape))
    pd = pairwise_dist_ref32(features, normalize)
    ranking = rank_stable(pd)
    if ind2id is not None:
        return dict((ind2id[i], ind2id[r].tolist()) for i, r in enumerate(ranking))
    return dict((i, r.tolist()) for i, r in enumerate(ranking))
