"""ctypes binding of the C ABI in include/niq.h (csrc/niq_api.cu -> libniq.so).

This is the only place the Python layer touches native code.  There is NO CPU fallback: importing
this module without the built library, or creating a context without a CUDA device, raises.

The reference has no native boundary (its backend is XLA); the functions bound here replace the
jitted bodies of src/affine.py, src/queries.py and src/kd_tree.py -- see include/niq.h for the
file:line each entry point stands in for.
"""
import collections
import ctypes as C
import hashlib
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NIQ_LIB") or os.path.join(_HERE, "libniq.so")   # NIQ_LIB: development builds (tools/)

NIQ_OK, NIQ_EINVAL, NIQ_ENOMEM, NIQ_ECUDA, NIQ_ECAPACITY, NIQ_EUNSUPPORTED = 0, -1, -2, -3, -4, -5
MEM_HOST, MEM_DEVICE = 0, 1
OP_DENSE, OP_RELU, OP_ELU, OP_SQUEEZE_LAST, OP_SPATIAL, OP_SIN, OP_POW2_ENCODE, OP_TANH = 0, 1, 2, 3, 4, 5, 6, 7
MODE_IDS = {"interval": 0, "affine_fixed": 1, "affine_truncate": 2, "affine_all": 3, "affine_append": 4, "sdf": 5,
            "slope_interval": 6}
TREE_INTERIOR, TREE_EXTERIOR = 1, 2

# every symbol include/niq.h declares (tests check the library exports all of them)
SYMBOLS = (
    "niq_last_error", "niq_version", "niq_ctx_create", "niq_ctx_destroy", "niq_ctx_sync", "niq_ctx_device_info",
    "niq_ctx_launch_count", "niq_ctx_timer_start", "niq_ctx_timer_stop", "niq_ctx_kernel_ms",
    "niq_ctx_kernel_timing", "niq_ctx_exec_macs", "niq_ctx_mc_points", "niq_dev_alloc", "niq_dev_free", "niq_dev_upload", "niq_dev_download",
    "niq_measure_fp32_peak", "niq_fingerprint128", "niq_mlp_create", "niq_mlp_destroy", "niq_mlp_macs", "niq_mlp_tie_rel", "niq_eval_points",
    "niq_classify_general_boxes", "niq_classify_boxes", "niq_slope_forward", "niq_cast_rays", "niq_cast_rays_frustum", "niq_tree_build", "niq_tree_build_roots", "niq_tree_build_dealt", "niq_tree_count",
    "niq_tree_copy", "niq_tree_stats", "niq_tree_level_info", "niq_tree_destroy", "niq_marching_cubes", "niq_marching_cubes_tree",
    "niq_mesh_count", "niq_mesh_copy", "niq_mesh_destroy", "niq_mc_tables", "niq_find_any_intersection",
    "niq_find_any_intersection_batch",
    "niq_closest_point",
)


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("A", C.c_void_p), ("b", C.c_void_p)]


class ModeCfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("truncate_count", C.c_int32), ("truncate_policy", C.c_int32),
                ("sdf_lipschitz", C.c_float)]


class CastOpts(C.Structure):
    _fields_ = [("hit_eps", C.c_float), ("max_dist", C.c_float), ("n_max_step", C.c_int32),
                ("n_substeps", C.c_int32), ("safety_factor", C.c_float), ("interval_grow_fac", C.c_float),
                ("interval_shrink_fac", C.c_float), ("interval_init_size", C.c_float)]


class Camera(C.Structure):
    """niq_camera (include/niq.h): cam_params of src/queries.py:197 with the float32 transcendental constants precomputed."""
    _fields_ = [("root", C.c_float * 3), ("look", C.c_float * 3), ("up", C.c_float * 3), ("left", C.c_float * 3),
                ("tan_half_fov_x", C.c_float), ("tan_half_fov_y", C.c_float), ("half_fov_x", C.c_float),
                ("half_fov_y", C.c_float), ("res_x", C.c_int32), ("res_y", C.c_int32)]


class NiqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libniq error {code}: {msg}")
        self.code = code


_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libniq.so (once).  Fails loudly when it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  This backend has no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            L.niq_last_error.restype = C.c_char_p
            L.niq_version.restype = C.c_char_p
            for name in SYMBOLS:
                fn = getattr(L, name)             # AttributeError here = header / library mismatch
                if name not in ("niq_last_error", "niq_version"):
                    fn.restype = C.c_int
            _lib = L
    return _lib


def check(code):
    if code == NIQ_OK:
        return
    msg = lib().niq_last_error().decode("utf-8", "replace")
    if code == NIQ_EINVAL:
        raise ValueError(msg)                      # the reference raises ValueError for bad arguments
    raise NiqError(code, msg)                      # RuntimeError subclass (reference: RuntimeError)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One CUDA device + stream (niq_ctx)."""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        self.device = device
        check(lib().niq_ctx_create(C.c_int(device), C.byref(self.handle)))
        self._mlp_cache = collections.OrderedDict()      # LRU: most recently used last
        self._mlp_retired = collections.deque()          # evicted handles, closed a few evictions later

    MLP_CACHE_SIZE = 64
    MLP_RETIRE_DEPTH = 8

    def close(self):
        if self.handle:
            for m in self._mlp_cache.values():
                m.close()
            self._mlp_cache.clear()
            for m in self._mlp_retired:
                m.close()
            self._mlp_retired.clear()
            lib().niq_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- small helpers -------------------------------------------------------------------------
    def device_info(self):
        info = (C.c_int32 * 4)()
        check(lib().niq_ctx_device_info(self.handle, info))
        return {"sm_count": info[0], "cc": (info[1], info[2]), "smem_optin": info[3]}

    def launch_count(self):
        out = C.c_int64()
        check(lib().niq_ctx_launch_count(self.handle, C.byref(out)))
        return out.value

    def sync(self):
        check(lib().niq_ctx_sync(self.handle))

    def timer_start(self):
        check(lib().niq_ctx_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().niq_ctx_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def kernel_timing(self, on):
        check(lib().niq_ctx_kernel_timing(self.handle, C.c_int(1 if on else 0)))

    def kernel_ms(self, which, reset=True):
        ms, n = C.c_float(), C.c_int64()
        check(lib().niq_ctx_kernel_ms(self.handle, C.c_int(which), C.byref(ms), C.byref(n), C.c_int(1 if reset else 0)))
        return ms.value, n.value

    def exec_macs(self, on=True, reset=False):
        """Executed multiply-adds counted by the network kernels so far (zero-skipping accounting); also
        switches the counter on/off for subsequent launches."""
        out = C.c_int64()
        check(lib().niq_ctx_exec_macs(self.handle, C.c_int(1 if on else 0), C.byref(out), C.c_int(1 if reset else 0)))
        return out.value

    def mc_points(self, reset=True):
        """(lattice points evaluated by marching cubes so far, the reference's count for the same leaves)."""
        ev, lat = C.c_int64(), C.c_int64()
        check(lib().niq_ctx_mc_points(self.handle, C.byref(ev), C.byref(lat), C.c_int(1 if reset else 0)))
        return ev.value, lat.value

    def fp32_peak_tflops(self):
        out = C.c_float()
        check(lib().niq_measure_fp32_peak(self.handle, C.byref(out)))
        return out.value

    def alloc(self, nbytes):
        p = C.c_void_p()
        check(lib().niq_dev_alloc(self.handle, C.c_int64(nbytes), C.byref(p)))
        return p

    def free(self, p):
        check(lib().niq_dev_free(self.handle, p))

    def upload(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        check(lib().niq_dev_upload(self.handle, dptr, ptr(arr), C.c_int64(arr.nbytes)))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.alloc(arr.nbytes)
        self.upload(p, arr)
        return p

    def download(self, dptr, shape, dtype):
        out = np.empty(shape, dtype)
        check(lib().niq_dev_download(self.handle, ptr(out), dptr, C.c_int64(out.nbytes)))
        return out

    # -- MLP handles ---------------------------------------------------------------------------
    def mlp(self, params):
        """Handle for a params dict.  The reference looks `params` up afresh on every call (the GUI
        mutates the spatial_transformation entries between calls, src/main_intersection.py:171-183),
        so handles are cached by CONTENT hash, never by object identity."""
        key = params_digest(params)
        m = self._mlp_cache.get(key)
        if m is not None:
            self._mlp_cache.move_to_end(key)             # a hit refreshes the entry (true LRU)
            return m
        if len(self._mlp_cache) >= self.MLP_CACHE_SIZE:
            # Evict the least recently used entry, but do not close it yet: a caller that fetched several handles for ONE
            # query (find_any_intersection, cast_rays with two funcs) may still hold it.  It is closed MLP_RETIRE_DEPTH
            # evictions later, far more lookups than any single call makes.
            _, old = self._mlp_cache.popitem(last=False)
            self._mlp_retired.append(old)
            while len(self._mlp_retired) > self.MLP_RETIRE_DEPTH:
                self._mlp_retired.popleft().close()
        m = Mlp(self, params)
        self._mlp_cache[key] = m
        return m


def _fingerprint(a):
    """128-bit content fingerprint of an array's bytes (niq_fingerprint128: one memory-bound pass in the library, host code).
    The digest is recomputed on EVERY query call; at 2 MB of weights blake2b over the raw bytes alone cost more than a whole
    depth-12 tree build."""
    b = np.ascontiguousarray(a)
    out = (C.c_uint64 * 2)()
    check(lib().niq_fingerprint128(C.c_void_p(b.ctypes.data), C.c_int64(b.nbytes), out))
    return bytes(out)


def params_digest(params):
    h = hashlib.blake2b(digest_size=16)
    for k in sorted(params):
        a = np.asarray(params[k])
        h.update(k.encode())
        h.update(str(a.shape).encode())
        h.update(a.dtype.str.encode())
        h.update(_fingerprint(a) if a.nbytes > 256 else np.ascontiguousarray(a).tobytes())
    return h.digest()


_OP_KINDS = {"dense": OP_DENSE, "relu": OP_RELU, "elu": OP_ELU, "squeeze_last": OP_SQUEEZE_LAST,
             "spatial_transformation": OP_SPATIAL, "sin": OP_SIN, "pow2_frequency_encode": OP_POW2_ENCODE,
             "tanh": OP_TANH}          # tanh: ours, parity unpinned (the reference has no tanh op)


def op_descs(params):
    """params dict -> (OpDesc array, keep-alive list), following the key grammar of src/mlp.py:117-144."""
    import mlp as mlp_mod
    n = mlp_mod.n_ops(params)
    descs = (OpDesc * n)()
    keep = []
    for i in range(n):
        name, args = mlp_mod.get_op_data(params, i)
        args.pop("_", None)
        if name not in _OP_KINDS:
            raise NiqError(NIQ_EUNSUPPORTED, f"op '{name}' is not an op of the reference's mlp format (dense, relu, elu, sin, "
                                             "pow2_frequency_encode, squeeze_last, spatial_transformation) nor 'tanh' (ours)")
        d = descs[i]
        d.kind = _OP_KINDS[name]
        if name == "dense":
            A = _f32(args["A"])
            if A.ndim != 2:
                raise ValueError("dense.A must be 2-D (in, out)")
            keep.append(A)
            d.in_dim, d.out_dim = A.shape
            d.A = A.ctypes.data
            if args.get("b") is not None:
                b = _f32(args["b"])
                if b.shape != (A.shape[1],):
                    raise ValueError("dense.b must have shape (out,)")
                keep.append(b)
                d.b = b.ctypes.data
        elif name == "pow2_frequency_encode":
            coefs = _f32(args["coefs"])
            if coefs.ndim != 1:
                raise ValueError("pow2_frequency_encode.coefs must be 1-D")
            keep.append(coefs)
            d.in_dim, d.out_dim = 3, coefs.shape[0]
            d.A = coefs.ctypes.data
            if args.get("shift") is not None:
                shift = _f32(args["shift"])
                if shift.shape != coefs.shape:
                    raise ValueError("pow2_frequency_encode.shift must match coefs")
                keep.append(shift)
                d.b = shift.ctypes.data
        elif name == "spatial_transformation":
            R, t = _f32(args["R"]), _f32(args["t"])
            if R.shape != (3, 3) or t.shape != (3,):
                raise ValueError("spatial_transformation needs R (3,3) and t (3,)")
            keep += [R, t]
            d.in_dim = d.out_dim = 3
            d.A = R.ctypes.data
            d.b = t.ctypes.data
    return descs, n, keep


class Mlp:
    def __init__(self, ctx, params):
        self.ctx = ctx
        self.handle = C.c_void_p()
        descs, n, keep = op_descs(params)
        check(lib().niq_mlp_create(ctx.handle, C.c_int32(n), descs, C.byref(self.handle)))
        del keep
        macs = C.c_int64()
        check(lib().niq_mlp_macs(self.handle, C.byref(macs)))
        self.macs = macs.value
        rel = C.c_float()
        check(lib().niq_mlp_tie_rel(self.handle, C.byref(rel)))
        self.tie_rel = rel.value

    def close(self):
        if self.handle:
            lib().niq_mlp_destroy(self.handle)
            self.handle = C.c_void_p()


def mode_cfg(ctx):
    """affine.AffineContext -> niq_mode_cfg"""
    if ctx.mode not in MODE_IDS:
        raise NiqError(NIQ_EUNSUPPORTED, f"mode '{ctx.mode}' is outside this backend (supported: {sorted(MODE_IDS)})")
    cfg = ModeCfg()
    cfg.mode = MODE_IDS[ctx.mode]
    cfg.truncate_count = int(ctx.truncate_count) if ctx.mode == "affine_truncate" else 0
    if ctx.mode == "affine_append":
        cfg.truncate_count = int(ctx.n_append)
    if ctx.mode == "sdf":
        cfg.sdf_lipschitz = float(ctx.lipschitz_bound)
    if ctx.mode == "affine_truncate" and ctx.truncate_policy != "absolute":
        if ctx.truncate_policy == "relative":
            cfg.truncate_policy = 1      # the library answers NIQ_EUNSUPPORTED with the reason
        else:
            raise RuntimeError("policy should be one of 'absolute' or 'relative'")   # src/affine.py:148
    return cfg


# ----------------------------------------------------------------------------------------------
# process-wide default context (one per device), created lazily
# ----------------------------------------------------------------------------------------------
_contexts = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("NIQ_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    c = _contexts.get(device)
    if c is None:
        c = Context(device)
        _contexts[device] = c
    return c
