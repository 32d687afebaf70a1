"""sagephy_b200 — Python (ctypes) binding of the B200-native SaGePhy hot path.

The product is ``libsagephy_b200.so`` (hand-written sm_100a kernels behind the C ABI of ``include/sagephy_b200.h``);
this module only loads it and exposes the reference's app surface (``HostTreeGen``, ``GuestTreeGen``, ... with the
reference's own argument grammar, see src/main/java/uc/sgp/sagephy/apps/SaGePhy/*Gen.java in the reference) to
Python callers such as the tests and ``bench.py``. There is no CPU fallback: if the library or a CUDA device is
missing, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsagephy_b200.so")

STREAM_NAMES = ("unpruned.tree", "pruned.tree", "unpruned.info", "pruned.info",
                "unpruned.guest2host", "pruned.guest2host", "unpruned.leafmap", "pruned.leafmap")
EVENT_NAMES = ("SPECIATION", "CODIVERGENCE", "LEAF", "UNSAMPLED_LEAF", "DUPLICATION", "LOSS", "REPLACING_LOSS",
               "ADDITIVE_TRANSFER", "REPLACING_TRANSFER") + tuple(
    f"{k}_TRANSFER_{g}_{s}" for k in ("ADDITIVE", "REPLACING") for g in ("INTRAGENE", "INTERGENE") for s in ("INTRASPECIES", "INTERSPECIES"))
RNG_PHILOX, RNG_MT_REPLAY = 0, 1
FLAG_MULTI_TREE_INPUT = 1
FLAG_KEEP_TREES = 2
FLAG_SERIAL = 4
FLAG_PIPELINE = 8
LAYOUT_PACKED, LAYOUT_PER_REPLICATE = 0, 1
STREAM_UNPRUNED_TREE, STREAM_PRUNED_TREE = 0, 1
STATUS_OK, STATUS_USAGE = 0, 6

# every symbol include/sagephy_b200.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = (
    "sgp_context_create", "sgp_context_destroy", "sgp_last_error", "sgp_app_run", "sgp_multi_create", "sgp_multi_size", "sgp_multi_context", "sgp_multi_destroy", "sgp_app_run_multi", "sgp_batches_write_files_layout", "sgp_relax_run_on_batch", "sgp_domain_run_on_batch", "sgp_batch_n", "sgp_batch_stream_data",
    "sgp_batch_stream_device_data", "sgp_batch_stream_offsets", "sgp_batch_copy_stream", "sgp_batch_copy_offsets", "sgp_batch_copy_stream_async", "sgp_batch_copy_offsets_async", "sgp_context_wait_copies", "sgp_batch_stream_mask", "sgp_batch_values", "sgp_batch_stream_bytes", "sgp_batch_stream_suffix",
    "sgp_batch_status", "sgp_batch_attempts", "sgp_batch_sampled_leaves", "sgp_batch_root_times", "sgp_batch_total_times",
    "sgp_batch_event_counts", "sgp_batch_stdout", "sgp_batch_stderr", "sgp_batch_out_prefix", "sgp_batch_summary",
    "sgp_batch_timings", "sgp_batch_write_files", "sgp_batch_write_files_layout", "sgp_batch_free", "sgp_probe_double_to_string", "sgp_probe_math",
    "sgp_probe_mt", "sgp_probe_seed", "sgp_probe_philox", "sgp_probe_issue_peaks", "sgp_probe_host_model", "sgp_probe_parse_app")


class BatchOpts(C.Structure):
    _fields_ = [("n_replicates", C.c_int64), ("replicate_offset", C.c_int64), ("rng_mode", C.c_int32),
                ("stream_mask", C.c_uint32), ("keep_on_device", C.c_int32), ("flags", C.c_int32)]


class SummaryStruct(C.Structure):
    _fields_ = [("n_ok", C.c_int64), ("n_failed", C.c_int64), ("attempts_total", C.c_int64), ("vertices_sampled", C.c_int64), ("vertices_abandoned", C.c_int64), ("attempts_completed", C.c_int64),
                ("event_counts", C.c_int64 * 17), ("leaf_hist", C.c_int64 * 1025), ("out_bytes", C.c_int64 * 8)]


class TimingsStruct(C.Structure):
    _fields_ = [("find_ms", C.c_double), ("build_ms", C.c_double), ("label_ms", C.c_double), ("emit_size_ms", C.c_double),
                ("emit_write_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double), ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("sub_batches", C.c_int64)]


_lib = None


def load_library() -> C.CDLL:
    """Loads libsagephy_b200.so (built by __graft_entry__.build() / sagephy_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(sagephy_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, cp = C.c_void_p, C.c_int64, C.c_char_p
    L.sgp_context_create.argtypes = [C.c_int, C.POINTER(vp)]; L.sgp_context_create.restype = C.c_int
    L.sgp_context_destroy.argtypes = [vp]
    L.sgp_last_error.argtypes = [vp]; L.sgp_last_error.restype = cp
    L.sgp_app_run.argtypes = [vp, cp, C.c_int, C.POINTER(cp), C.POINTER(BatchOpts), C.POINTER(vp)]; L.sgp_app_run.restype = C.c_int
    L.sgp_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]; L.sgp_multi_create.restype = C.c_int
    L.sgp_multi_size.argtypes = [vp]; L.sgp_multi_size.restype = C.c_int
    L.sgp_multi_context.argtypes = [vp, C.c_int]; L.sgp_multi_context.restype = vp
    L.sgp_multi_destroy.argtypes = [vp]
    L.sgp_app_run_multi.argtypes = [vp, cp, C.c_int, C.POINTER(cp), C.POINTER(BatchOpts), C.POINTER(vp), C.POINTER(SummaryStruct)]; L.sgp_app_run_multi.restype = C.c_int
    L.sgp_batches_write_files_layout.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int64]; L.sgp_batches_write_files_layout.restype = C.c_int
    L.sgp_relax_run_on_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(cp), C.POINTER(BatchOpts), C.POINTER(vp)]; L.sgp_relax_run_on_batch.restype = C.c_int
    L.sgp_domain_run_on_batch.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(cp), C.POINTER(BatchOpts), C.POINTER(vp)]; L.sgp_domain_run_on_batch.restype = C.c_int
    L.sgp_batch_n.argtypes = [vp]; L.sgp_batch_n.restype = i64
    L.sgp_batch_stream_data.argtypes = [vp, C.c_int]; L.sgp_batch_stream_data.restype = vp
    L.sgp_batch_stream_device_data.argtypes = [vp, C.c_int]; L.sgp_batch_stream_device_data.restype = vp
    L.sgp_batch_stream_offsets.argtypes = [vp, C.c_int]; L.sgp_batch_stream_offsets.restype = C.POINTER(C.c_uint64)
    L.sgp_batch_copy_stream.argtypes = [vp, C.c_int, vp, C.c_uint64]; L.sgp_batch_copy_stream.restype = C.c_int
    L.sgp_batch_copy_offsets.argtypes = [vp, C.c_int, vp]; L.sgp_batch_copy_offsets.restype = C.c_int
    L.sgp_batch_copy_stream_async.argtypes = [vp, C.c_int, vp, C.c_uint64]; L.sgp_batch_copy_stream_async.restype = C.c_int
    L.sgp_batch_copy_offsets_async.argtypes = [vp, C.c_int, vp]; L.sgp_batch_copy_offsets_async.restype = C.c_int
    L.sgp_context_wait_copies.argtypes = [vp]; L.sgp_context_wait_copies.restype = C.c_int
    L.sgp_batch_stream_mask.argtypes = [vp]; L.sgp_batch_stream_mask.restype = C.c_uint32
    L.sgp_batch_values.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]; L.sgp_batch_values.restype = C.POINTER(C.c_double)
    L.sgp_batch_stream_bytes.argtypes = [vp, C.c_int]; L.sgp_batch_stream_bytes.restype = C.c_uint64
    L.sgp_batch_stream_suffix.argtypes = [vp, C.c_int]; L.sgp_batch_stream_suffix.restype = cp
    for name, typ in (("sgp_batch_status", C.c_int32), ("sgp_batch_attempts", C.c_int32), ("sgp_batch_sampled_leaves", C.c_int32),
                      ("sgp_batch_root_times", C.c_double), ("sgp_batch_total_times", C.c_double), ("sgp_batch_event_counts", C.c_int32)):
        f = getattr(L, name); f.argtypes = [vp]; f.restype = C.POINTER(typ)
    for name in ("sgp_batch_stdout", "sgp_batch_stderr", "sgp_batch_out_prefix"):
        f = getattr(L, name); f.argtypes = [vp]; f.restype = cp
    L.sgp_batch_summary.argtypes = [vp, C.POINTER(SummaryStruct)]
    L.sgp_batch_timings.argtypes = [vp, C.POINTER(TimingsStruct)]
    L.sgp_batch_write_files.argtypes = [vp]; L.sgp_batch_write_files.restype = C.c_int
    L.sgp_batch_write_files_layout.argtypes = [vp, C.c_int, C.c_int64]; L.sgp_batch_write_files_layout.restype = C.c_int
    L.sgp_batch_free.argtypes = [vp]
    L.sgp_probe_double_to_string.argtypes = [vp, C.POINTER(C.c_double), i64, C.c_char_p]; L.sgp_probe_double_to_string.restype = C.c_int
    L.sgp_probe_math.argtypes = [vp, C.c_int, C.POINTER(C.c_double), i64, C.POINTER(C.c_double)]; L.sgp_probe_math.restype = C.c_int
    L.sgp_probe_mt.argtypes = [vp, cp, i64, i64, C.POINTER(C.c_uint32)]; L.sgp_probe_mt.restype = C.c_int
    L.sgp_probe_seed.argtypes = [cp, i64, C.POINTER(C.c_uint32), C.c_char_p, C.c_int32]; L.sgp_probe_seed.restype = C.c_int
    L.sgp_probe_philox.argtypes = [vp, C.c_int, vp, vp, i64, vp]; L.sgp_probe_philox.restype = C.c_int
    L.sgp_probe_issue_peaks.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]; L.sgp_probe_issue_peaks.restype = C.c_int
    L.sgp_probe_parse_app.argtypes = [cp, C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int32]; L.sgp_probe_parse_app.restype = C.c_int
    _lib = L
    return L


class SagephyError(RuntimeError):
    pass


def _philox(handle, ctr, key, variant):
    L = load_library()
    c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4); k = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
    out = np.empty_like(c)
    rc = L.sgp_probe_philox(handle, variant, c.ctypes.data, k.ctypes.data, len(c), out.ctypes.data)
    if rc != 0:
        raise SagephyError(f"sgp_probe_philox failed ({rc})")
    return out


def philox_host(ctr, key, variant: int = 0) -> np.ndarray:
    """Philox4x32-10 blocks from the host build of the library's generator (no GPU)."""
    return _philox(None, ctr, key, variant)


def seed_probe(seed, replicate: int = 0):
    """Host-side seed logic of the library (no GPU): (4 big-endian MT key words, decimal text) of seed + replicate."""
    L = load_library()
    key = (C.c_uint32 * 4)(); buf = C.create_string_buffer(4096)
    rc = L.sgp_probe_seed(str(seed).encode(), replicate, key, buf, 4096)
    if rc != 0:
        raise SagephyError(f"sgp_probe_seed failed ({rc})")
    return [int(k) for k in key], buf.value.decode()


def probe_parse_app(app, args):
    """Host-side argument handling of HostTreeGen / GuestTreeGen / DomainTreeGen without a GPU (`sgp_probe_parse_app`): returns
    (status, stderr text) as `sgp_app_run` would leave them before its first kernel. With `-hybrid` that is the whole run."""
    L = load_library()
    argv = (C.c_char_p * max(len(args), 1))(*[str(a).encode() for a in args])
    buf = C.create_string_buffer(1 << 16)
    rc = L.sgp_probe_parse_app(app.encode(), len(args), argv, buf, 1 << 16)
    return rc, buf.value.decode()


@dataclass
class Summary:
    n_ok: int
    n_failed: int
    attempts_total: int
    vertices_sampled: int
    vertices_abandoned: int
    attempts_completed: int
    event_counts: np.ndarray
    leaf_hist: np.ndarray
    out_bytes: np.ndarray

    def as_vector(self) -> np.ndarray:
        """Flat int64 vector (what the multi-GPU driver all-reduces)."""
        return np.concatenate([np.array([self.n_ok, self.n_failed, self.attempts_total, self.vertices_sampled, self.vertices_abandoned, self.attempts_completed], dtype=np.int64),
                               self.event_counts, self.leaf_hist, self.out_bytes])

    @staticmethod
    def from_vector(v: np.ndarray) -> "Summary":
        v = np.asarray(v, dtype=np.int64)
        return Summary(int(v[0]), int(v[1]), int(v[2]), int(v[3]), int(v[4]), int(v[5]), v[6:23].copy(), v[23:23 + 1025].copy(), v[23 + 1025:23 + 1025 + 8].copy())


class Batch:
    """Result of one app run over n replicates (wraps sgp_batch)."""

    def __init__(self, lib, handle, status):
        self._lib, self._h, self.status_code = lib, handle, status

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def free(self):
        if getattr(self, "_h", None):
            self._lib.sgp_batch_free(self._h)
            self._h = None

    @property
    def n(self) -> int:
        return int(self._lib.sgp_batch_n(self._h))

    @property
    def stdout(self) -> str:
        return self._lib.sgp_batch_stdout(self._h).decode()

    @property
    def stderr(self) -> str:
        return self._lib.sgp_batch_stderr(self._h).decode()

    @property
    def out_prefix(self) -> str:
        return self._lib.sgp_batch_out_prefix(self._h).decode()

    def stream_suffix(self, s: int) -> str:
        return self._lib.sgp_batch_stream_suffix(self._h, s).decode()

    def stream_bytes(self, s: int) -> int:
        return int(self._lib.sgp_batch_stream_bytes(self._h, s))

    def stream(self, s: int) -> bytes:
        """Whole stream s (all replicates concatenated)."""
        nbytes = self.stream_bytes(s)
        if nbytes == 0:
            return b""
        p = self._lib.sgp_batch_stream_data(self._h, s)
        if not p:
            raise SagephyError("stream copy failed")
        return C.string_at(p, nbytes)

    def copy_stream_into(self, s: int, dst_ptr: int, capacity: int) -> int:
        """Copies stream s from HBM into caller-owned host memory (e.g. a pinned torch tensor's data_ptr()). Returns bytes."""
        rc = self._lib.sgp_batch_copy_stream(self._h, s, C.c_void_p(dst_ptr), capacity)
        if rc != 0:
            raise SagephyError(f"sgp_batch_copy_stream failed ({rc})")
        return self.stream_bytes(s)

    def copy_stream_into_async(self, s: int, dst_ptr: int, capacity: int) -> int:
        """Enqueues the copy on the context's copy stream (overlaps the next run); call Context.wait_copies() before reading."""
        rc = self._lib.sgp_batch_copy_stream_async(self._h, s, C.c_void_p(dst_ptr), capacity)
        if rc != 0:
            raise SagephyError(f"sgp_batch_copy_stream_async failed ({rc})")
        return self.stream_bytes(s)

    def copy_offsets_into_async(self, s: int, dst_ptr: int) -> None:
        rc = self._lib.sgp_batch_copy_offsets_async(self._h, s, C.c_void_p(dst_ptr))
        if rc != 0:
            raise SagephyError(f"sgp_batch_copy_offsets_async failed ({rc})")

    def copy_offsets_into(self, s: int, dst_ptr: int) -> None:
        rc = self._lib.sgp_batch_copy_offsets(self._h, s, C.c_void_p(dst_ptr))
        if rc != 0:
            raise SagephyError(f"sgp_batch_copy_offsets failed ({rc})")

    def offsets(self, s: int) -> np.ndarray:
        p = self._lib.sgp_batch_stream_offsets(self._h, s)
        return np.ctypeslib.as_array(p, shape=(self.n + 1,)).copy()

    def replicate(self, s: int, i: int) -> bytes:
        off = self.offsets(s)
        return self.stream(s)[int(off[i]):int(off[i + 1])]

    def files(self, i: int = 0) -> dict:
        """{suffix: bytes} of replicate i — what the reference would have written for that run."""
        out = {}
        mask = int(self._lib.sgp_batch_stream_mask(self._h))
        for s in range(8):
            if mask & (1 << s):
                out[self.stream_suffix(s)] = self.replicate(s, i)
        return out

    def values(self, which: int) -> np.ndarray:
        """BranchRelaxer: flat per-branch doubles (0 = rates, 1 = relaxed lengths), replicate-major."""
        cnt = C.c_int64()
        p = self._lib.sgp_batch_values(self._h, which, C.byref(cnt))
        return np.ctypeslib.as_array(p, shape=(cnt.value,)).copy()

    def _arr(self, fn, shape, dtype):
        p = getattr(self._lib, fn)(self._h)
        return np.ctypeslib.as_array(p, shape=shape).astype(dtype, copy=True)

    @property
    def rep_status(self): return self._arr("sgp_batch_status", (self.n,), np.int32)
    @property
    def attempts(self): return self._arr("sgp_batch_attempts", (self.n,), np.int32)
    @property
    def sampled_leaves(self): return self._arr("sgp_batch_sampled_leaves", (self.n,), np.int32)
    @property
    def root_times(self): return self._arr("sgp_batch_root_times", (self.n,), np.float64)
    @property
    def total_times(self): return self._arr("sgp_batch_total_times", (self.n,), np.float64)
    @property
    def event_counts(self): return self._arr("sgp_batch_event_counts", (self.n, 17), np.int32)

    def summary(self) -> Summary:
        s = SummaryStruct()
        self._lib.sgp_batch_summary(self._h, C.byref(s))
        return Summary(s.n_ok, s.n_failed, s.attempts_total, s.vertices_sampled, s.vertices_abandoned, s.attempts_completed, np.array(s.event_counts, dtype=np.int64),
                       np.array(s.leaf_hist, dtype=np.int64), np.array(s.out_bytes, dtype=np.int64))

    def timings(self) -> dict:
        t = TimingsStruct()
        self._lib.sgp_batch_timings(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in TimingsStruct._fields_}

    def write_files(self, layout=None, first_index: int = 0):
        """layout None: the reference's own file names (n = 1); LAYOUT_PACKED / LAYOUT_PER_REPLICATE: see sagephy_b200.h."""
        if layout is not None:
            rc = self._lib.sgp_batch_write_files_layout(self._h, int(layout), int(first_index))
            if rc != STATUS_OK:
                raise SagephyError(f"sgp_batch_write_files_layout failed ({rc})")
            return
        rc = self._lib.sgp_batch_write_files(self._h)
        if rc != 0:
            raise SagephyError(f"sgp_batch_write_files failed ({rc})")


class MultiContext:
    """Several devices of one node behind one call (wraps sgp_multi): replicates are sharded by index range, one host thread per
    device. run() returns (list of shard batches in index order, merged Summary)."""

    def __init__(self, devices):
        self._lib = load_library()
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.sgp_multi_create(dev, len(devices), C.byref(h))
        if rc != 0:
            raise SagephyError("sgp_multi_create failed: " + self._lib.sgp_last_error(None).decode())
        self._h, self.size = h, len(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sgp_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, app: str, args, n: int = 1, rng: int = RNG_PHILOX, offset: int = 0, stream_mask: int = 0, keep_on_device: bool = False, flags: int = 0):
        argv = [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        opts = BatchOpts(n, offset, rng, stream_mask, 1 if keep_on_device else 0, flags)
        outs = (C.c_void_p * self.size)()
        s = SummaryStruct()
        rc = self._lib.sgp_app_run_multi(self._h, app.encode(), len(argv), arr, C.byref(opts), outs, C.byref(s))
        shards = [Batch(self._lib, C.c_void_p(outs[i]), rc) for i in range(self.size) if outs[i]]
        if rc not in (STATUS_OK, STATUS_USAGE):
            raise SagephyError(f"{app} failed on {self.size} devices ({rc}): {shards[0].stderr if shards else ''}")
        summary = Summary(s.n_ok, s.n_failed, s.attempts_total, s.vertices_sampled, s.vertices_abandoned, s.attempts_completed, np.array(s.event_counts, dtype=np.int64),
                          np.array(s.leaf_hist, dtype=np.int64), np.array(s.out_bytes, dtype=np.int64))
        return shards, summary

    def write_files(self, shards, layout: int = LAYOUT_PACKED, first_index: int = 0):
        arr = (C.c_void_p * len(shards))(*[b._h for b in shards])
        rc = self._lib.sgp_batches_write_files_layout(arr, len(shards), int(layout), int(first_index))
        if rc != STATUS_OK:
            raise SagephyError(f"sgp_batches_write_files_layout failed ({rc})")


class Context:
    """One CUDA device + stream (wraps sgp_context)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.sgp_context_create(device, C.byref(h))
        if rc != 0:
            raise SagephyError("sgp_context_create failed: " + self._lib.sgp_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sgp_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, app: str, args, n: int = 1, rng: int = RNG_PHILOX, offset: int = 0, stream_mask: int = 0,
            keep_on_device: bool = False, check: bool = True, flags: int = 0) -> Batch:
        """Runs ``<app> args...`` (reference CLI grammar) for n replicates."""
        argv = [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        opts = BatchOpts(n, offset, rng, stream_mask, 1 if keep_on_device else 0, flags)
        h = C.c_void_p()
        rc = self._lib.sgp_app_run(self._h, app.encode(), len(argv), arr, C.byref(opts), C.byref(h))
        b = Batch(self._lib, h, rc)
        if check and rc not in (STATUS_OK, STATUS_USAGE):
            raise SagephyError(f"{app} failed ({rc}): {self._lib.sgp_last_error(self._h).decode()} {b.stderr if h else ''}")
        return b

    def relax_on_batch(self, src: "Batch", args, which: int = 1, n: int = 0, rng: int = RNG_PHILOX, offset: int = 0, stream_mask: int = 0,
                       keep_on_device: bool = False, check: bool = True) -> "Batch":
        """Chained stage: ``BranchRelaxer args...`` over the trees of a generator batch made with FLAG_KEEP_TREES (no host
        round trip). which = STREAM_PRUNED_TREE | STREAM_UNPRUNED_TREE; n = 0 means one replicate per source tree."""
        argv = [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        opts = BatchOpts(n, offset, rng, stream_mask, 1 if keep_on_device else 0, 0)
        h = C.c_void_p()
        rc = self._lib.sgp_relax_run_on_batch(self._h, src._h, which, len(argv), arr, C.byref(opts), C.byref(h))
        b = Batch(self._lib, h, rc)
        if check and rc not in (STATUS_OK, STATUS_USAGE):
            raise SagephyError(f"chained BranchRelaxer failed ({rc}): {self._lib.sgp_last_error(self._h).decode()} {b.stderr if h else ''}")
        return b

    def domain_on_batch(self, src: "Batch", args, first: int = 0, count: int = 0, name_first_index: int = 0, n: int = 1, rng: int = RNG_PHILOX, offset: int = 0,
                        stream_mask: int = 0, keep_on_device: bool = False, check: bool = True) -> "Batch":
        """Chained stage: ``DomainTreeGen args...`` over the pruned trees [first, first + count) of a generator batch made with
        FLAG_KEEP_TREES as the gene-tree forest (no Newick text is written or parsed). count = 0: all trees of the batch."""
        argv = [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        opts = BatchOpts(n, offset, rng, stream_mask, 1 if keep_on_device else 0, 0)
        h = C.c_void_p()
        rc = self._lib.sgp_domain_run_on_batch(self._h, src._h, first, count or src.n, name_first_index, len(argv), arr, C.byref(opts), C.byref(h))
        b = Batch(self._lib, h, rc)
        if check and rc not in (STATUS_OK, STATUS_USAGE):
            raise SagephyError(f"chained DomainTreeGen failed ({rc}): {self._lib.sgp_last_error(self._h).decode()} {b.stderr if h else ''}")
        return b

    def wait_copies(self):
        self._lib.sgp_context_wait_copies(self._h)

    # probes (parity tests of the device-side Java-exact primitives)
    def double_to_string(self, values) -> list:
        v = np.ascontiguousarray(values, dtype=np.float64)
        buf = C.create_string_buffer(32 * len(v))
        rc = self._lib.sgp_probe_double_to_string(self._h, v.ctypes.data_as(C.POINTER(C.c_double)), len(v), buf)
        if rc != 0:
            raise SagephyError("probe failed")
        raw = buf.raw
        return [raw[32 * i:32 * i + 32].split(b"\0", 1)[0].decode() for i in range(len(v))]

    def math(self, fn: int, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty_like(v)
        rc = self._lib.sgp_probe_math(self._h, fn, v.ctypes.data_as(C.POINTER(C.c_double)), len(v), out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise SagephyError("probe failed")
        return out

    def mt_words(self, seed, n: int, replicate: int = 0) -> np.ndarray:
        """MT19937 words of replicate `replicate` of a run with `-s seed` (seed: any Python int or decimal string)."""
        out = np.empty(n, dtype=np.uint32)
        rc = self._lib.sgp_probe_mt(self._h, str(seed).encode(), replicate, n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        if rc != 0:
            raise SagephyError("probe failed")
        return out

    def philox(self, ctr, key, variant: int = 0) -> np.ndarray:
        """Philox4x32-10 blocks computed on the device (variant 1 = precomputed round keys, the sampler's form)."""
        return _philox(self._h, ctr, key, variant)

    def issue_peaks(self) -> dict:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        rc = self._lib.sgp_probe_issue_peaks(self._h, C.byref(a), C.byref(b), C.byref(c))
        if rc != 0:
            raise SagephyError("probe failed")
        return {"dfma_ginst_per_s": a.value, "imad_ginst_per_s": b.value, "mixed_int_ginst_per_s": c.value}
