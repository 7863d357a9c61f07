"""Mirrors /root/reference/src/bucketing.py:7-36.  The kernels compact on the GPU and never pad; the
bucket sizes survive only where they are part of a result's contract (padded array sizes returned by
construct_uniform_unknown_levelset_tree, the N_evals counter of cast_rays)."""
import numpy as np

bucket_sizes = [2 ** s for s in range(7, 31)]


def get_next_bucket_size(s):
    for b in bucket_sizes:
        if s <= b:
            return b
    raise ValueError("max bucket size exceeded")


def fits_in_smaller_bucket(size, curr_bucket_size):
    return get_next_bucket_size(size) < curr_bucket_size


def compactify_and_rebucket_arrays(mask, bucket_size, *arrs):
    """:16-32 on host arrays: order-preserving compaction of the masked rows padded to bucket_size.
    -> (out_mask, N_in, *arrays).  Padding rows are unspecified in the reference; zero here."""
    mask = np.asarray(mask, bool)
    idx = np.nonzero(mask)[0]
    n_in = idx.shape[0]
    out_mask = np.arange(bucket_size) < n_in
    outs = []
    for a in arrs:
        a = np.asarray(a)
        o = np.zeros((bucket_size,) + a.shape[1:], a.dtype)
        o[:min(n_in, bucket_size)] = a[idx[:bucket_size]]
        outs.append(o)
    return (out_mask, n_in, *outs)
