"""ctypes access to the CPU oracle (oracle/liboracle.so) — test infrastructure only, never used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make", "-C", ORACLE_DIR])
        L = C.CDLL(ORACLE_LIB)
        L.oracle_run.argtypes = [C.c_int, C.POINTER(C.c_char_p)]; L.oracle_run.restype = C.c_void_p
        L.oracle_n_files.argtypes = [C.c_void_p]; L.oracle_n_files.restype = C.c_int
        L.oracle_file_name.argtypes = [C.c_void_p, C.c_int]; L.oracle_file_name.restype = C.c_char_p
        L.oracle_file_data.argtypes = [C.c_void_p, C.c_int]; L.oracle_file_data.restype = C.c_void_p
        L.oracle_file_len.argtypes = [C.c_void_p, C.c_int]; L.oracle_file_len.restype = C.c_longlong
        L.oracle_stdout.argtypes = [C.c_void_p]; L.oracle_stdout.restype = C.c_char_p
        L.oracle_stderr.argtypes = [C.c_void_p]; L.oracle_stderr.restype = C.c_char_p
        L.oracle_attempts.argtypes = [C.c_void_p]; L.oracle_attempts.restype = C.c_int
        L.oracle_status.argtypes = [C.c_void_p]; L.oracle_status.restype = C.c_int
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_no_recipient.argtypes = [C.c_void_p]; L.oracle_no_recipient.restype = C.c_longlong
        L.oracle_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_longlong, C.c_longlong, C.c_void_p]; L.oracle_batch.restype = C.c_double
        L.oracle_mt_words.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.oracle_mt_doubles.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.oracle_mt_ints.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_double_to_string.argtypes = [C.c_double, C.c_char_p, C.c_int]; L.oracle_double_to_string.restype = C.c_int
        for f in ("oracle_log", "oracle_log10", "oracle_exp"):
            getattr(L, f).argtypes = [C.c_double]; getattr(L, f).restype = C.c_double
        L.oracle_pow.argtypes = [C.c_double, C.c_double]; L.oracle_pow.restype = C.c_double
        L.oracle_normal_quantile.argtypes = [C.c_double]; L.oracle_normal_quantile.restype = C.c_double
        L.oracle_gamma_quantile.argtypes = [C.c_double, C.c_double, C.c_double]; L.oracle_gamma_quantile.restype = C.c_double
        L.oracle_round_sig.argtypes = [C.c_double, C.c_int]; L.oracle_round_sig.restype = C.c_double
        _lib = L
    return _lib


def _argv(app, args):
    toks = [app.encode()] + [str(a).encode() for a in args]
    return len(toks), (C.c_char_p * len(toks))(*toks)


def oracle_run(app, args):
    """Runs one reference-style invocation. Returns dict(files={path: bytes}, stdout, stderr, attempts, status)."""
    L = oracle()
    n, arr = _argv(app, args)
    h = L.oracle_run(n, arr)
    try:
        files = {}
        for i in range(L.oracle_n_files(h)):
            files[L.oracle_file_name(h, i).decode()] = C.string_at(L.oracle_file_data(h, i), L.oracle_file_len(h, i))
        return dict(files=files, stdout=L.oracle_stdout(h).decode(), stderr=L.oracle_stderr(h).decode(),
                    attempts=L.oracle_attempts(h), status=L.oracle_status(h), no_recipient=int(L.oracle_no_recipient(h)))
    finally:
        L.oracle_free(h)


def oracle_files(app, args):
    return oracle_run(app, args)["files"]


STAT_COLS = ["attempts", "status", "n_unpruned", "n_pruned", "sampled_leaves", "pruned_root_time", "total_time", "vertices_created", "out_bytes"]


def oracle_batch(app, args, first, n, want_stats=True):
    """Replicate i runs with seed + first + i. Returns (seconds, stats[n, 26])."""
    L = oracle()
    k, arr = _argv(app, args)
    st = np.zeros((n, 26), dtype=np.float64) if want_stats else None
    sec = L.oracle_batch(k, arr, first, n, st.ctypes.data if want_stats else None)
    return sec, st


def oracle_double_to_string(v: float) -> str:
    buf = C.create_string_buffer(64)
    oracle().oracle_double_to_string(v, buf, 64)
    return buf.value.decode()


def oracle_mt_words(seed, n):
    out = np.empty(n, dtype=np.uint32)
    oracle().oracle_mt_words(str(seed).encode(), n, out.ctypes.data)
    return out


def oracle_mt_doubles(seed, n):
    out = np.empty(n, dtype=np.float64)
    oracle().oracle_mt_doubles(str(seed).encode(), n, out.ctypes.data)
    return out


# ---- process-pool front ends (the oracle is single-threaded; the GPU boxes have 16+ host cores) -------------------------------------
def _run_worker(job):
    app, args = job
    return oracle_run(app, args)


def _batch_worker(job):
    app, args, first, n = job
    return oracle_batch(app, args, first, n)[1]


def _mp_context():
    # "spawn": the parent usually holds a CUDA context and worker threads by now; fork()ing such a process is not safe
    import multiprocessing
    return multiprocessing.get_context("spawn")


def pool_size(procs=None):
    return max(1, min(procs or (os.cpu_count() or 1), 32))


def oracle_run_many(jobs, procs=None):
    """jobs: [(app, args), ...] -> list of oracle_run results, computed on a process pool."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = list(jobs)
    if len(jobs) <= 1 or pool_size(procs) == 1:
        return [_run_worker(j) for j in jobs]
    with ProcessPoolExecutor(pool_size(procs), mp_context=_mp_context()) as ex:
        return list(ex.map(_run_worker, jobs, chunksize=max(1, len(jobs) // (4 * pool_size(procs)))))


def oracle_batch_parallel(app, args, first, n, procs=None):
    """oracle_batch over [first, first + n) split across a process pool. Returns stats[n, 26]."""
    from concurrent.futures import ProcessPoolExecutor
    p = pool_size(procs)
    if p == 1 or n < 4 * p:
        return oracle_batch(app, args, first, n)[1]
    chunks = 4 * p
    edges = [first + (n * k) // chunks for k in range(chunks + 1)]
    with ProcessPoolExecutor(p, mp_context=_mp_context()) as ex:
        parts = list(ex.map(_batch_worker, [(app, list(args), edges[k], edges[k + 1] - edges[k]) for k in range(chunks) if edges[k + 1] > edges[k]]))
    return np.concatenate(parts, axis=0)
