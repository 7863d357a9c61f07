"""Mirrors /root/reference/src/geometry.py:6-22 (the vector helpers the hot path's callers use)."""
import numpy as np


def norm(x):
    x = np.asarray(x, np.float32)
    return np.sqrt((x * x).sum(axis=-1, dtype=np.float32)).astype(np.float32)


def norm2(x):
    x = np.asarray(x, np.float32)
    return np.inner(x, x)


def normalize(x):
    x = np.asarray(x, np.float32)
    return (x / norm(x)[..., None]).astype(np.float32)


def orthogonal_dir(x, remove_dir):
    remove_dir = normalize(remove_dir)
    x = np.asarray(x, np.float32)
    x = x - np.dot(x, remove_dir).astype(np.float32) * remove_dir
    return normalize(x)


def dot(x, y):
    return np.sum(np.asarray(x) * np.asarray(y), axis=-1)
