"""ctypes binding of the C ABI in include/cpprob_sis.h (libcpprob_sis.so).

Plumbing for the tests and bench.py: every compute call goes through the same C entry points the
C++14 host API (include/cpprob/cpprob.hpp) uses.  There is no Python or CPU implementation of the
path here; if the shared library is missing, importing this module's `lib()` raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPPROB_SIS_LIB") or os.path.join(_HERE, "lib", "libcpprob_sis.so")   # override: kernel-variant sweeps



def _prefer_bundled_nccl():
    """The library opens NCCL by name at its first multi-GPU call (libnccl.so.2, CPPROB_SIS_NCCL_LIB overrides).  A Python
    process that imports torch LATER needs the NCCL torch was built against (its wheel's nvidia/nccl/lib/libnccl.so.2): the
    loader keeps one object per soname, so if the system's older libnccl got in first, `import torch` fails on a missing
    symbol.  Point the library at the wheel's copy when there is one (no import of torch needed for that)."""
    if os.environ.get("CPPROB_SIS_NCCL_LIB"):
        return
    import sys
    for p in sys.path:
        cand = os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["CPPROB_SIS_NCCL_LIB"] = cand
            return


_prefer_bundled_nccl()
EMIT_NONE, EMIT_ALL = 0, 1
DIST = {"normal": 0, "uniform_real": 1, "uniform_smallint": 2, "discrete": 3, "poisson": 4, "gamma": 5, "beta": 6}
PATHS = ("fused", "staged", "rows")      # cpprob_sis_stats.path (CPPROB_SIS_PATH_*)
BASE_COLS = 8
CHUNK = 1 << 15

# every symbol include/cpprob_sis.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "cpprob_sis_abi_version", "cpprob_sis_last_error", "cpprob_sis_create", "cpprob_sis_destroy",
    "cpprob_sis_register_model", "cpprob_sis_model_count", "cpprob_sis_model_name", "cpprob_sis_find_model",
    "cpprob_sis_describe", "cpprob_sis_run", "cpprob_sis_infer_to_files", "cpprob_sis_run_shard",
    "cpprob_sis_merge", "cpprob_sis_replay", "cpprob_sis_reduce_records", "cpprob_sis_logpdf",
    "cpprob_sis_sample", "cpprob_sis_philox", "cpprob_sis_dmath", "cpprob_sis_measure_dfma_peak",
    "cpprob_sis_measure_store_peak", "cpprob_sis_plan_shard", "cpprob_sis_probe_issue", "cpprob_sis_probe_dfma_chains", "cpprob_sis_run_multi", "cpprob_sis_write_summary",
    "cpprob_sis_text_stage_stats", "cpprob_sis_plan_rows", "cpprob_sis_merge_padded",
    "cpprob_sis_set_seed", "cpprob_sis_infer_to_files_multi", "cpprob_sis_comm_get_id", "cpprob_sis_comm_init", "cpprob_sis_comm_init_local", "cpprob_sis_comm_destroy", "cpprob_sis_comm_exchange", "cpprob_sis_run_dist",
]
COMM_ID_BYTES = 128


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("seed", C.c_uint64), ("max_batch", C.c_uint64), ("blocks_per_sm", C.c_int)]


class Slot(C.Structure):
    _fields_ = [("is_int", C.c_int), ("id", C.c_int), ("k", C.c_int), ("row", C.c_int), ("width", C.c_int)]


class Structure(C.Structure):
    _fields_ = [("n_ids", C.c_int), ("n_slots", C.c_int), ("n_real", C.c_int), ("n_int", C.c_int),
                ("n_samples", C.c_int), ("ids", C.POINTER(C.c_char_p)), ("slots", C.POINTER(Slot))]


class Stats(C.Structure):
    _fields_ = [("n_particles", C.c_uint64), ("n_neg_inf", C.c_uint64), ("n_nan", C.c_uint64),
                ("m_ref", C.c_double), ("max_log_w", C.c_double), ("log_sum_exp", C.c_double),
                ("log_evidence", C.c_double), ("ess", C.c_double),
                ("n_real", C.c_int), ("n_int", C.c_int),
                ("real_mean", C.POINTER(C.c_double)), ("real_var", C.POINTER(C.c_double)),
                ("int_lo", C.c_longlong), ("int_bins", C.c_int),
                ("int_prob", C.POINTER(C.c_double)), ("int_map", C.POINTER(C.c_longlong)),
                ("n_cols", C.c_int), ("sums", C.POINTER(C.c_double)),
                ("device_ms", C.c_double), ("kernel_launches", C.c_uint64), ("passes", C.c_int), ("path", C.c_int),
                ("particle_ms", C.c_double)]


class Block(C.Structure):
    _fields_ = [("first_particle", C.c_uint64), ("n", C.c_uint64), ("stride", C.c_uint64),
                ("n_real", C.c_int), ("n_int", C.c_int),
                ("real_rows", C.POINTER(C.c_double)), ("int_rows", C.POINTER(C.c_int32)), ("log_w", C.POINTER(C.c_double))]


BLOCK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Block))


class RunOptions(C.Structure):
    _fields_ = [("emit", C.c_int), ("force_rows", C.c_int), ("on_block", BLOCK_FN), ("user", C.c_void_p)]


class Partials(C.Structure):
    _fields_ = [("device_ptr", C.c_void_p), ("n_chunks_local", C.c_uint32), ("n_chunks_total", C.c_uint32),
                ("chunk_first", C.c_uint32), ("rows_per_chunk", C.c_uint32), ("n_cols", C.c_int), ("m_ref", C.c_double),
                ("device_ms", C.c_double), ("kernel_launches", C.c_uint64)]


_lib = None
_lock = threading.Lock()


def lib():
    """Loads libcpprob_sis.so (built by __graft_entry__.build() / cpprob_b200/csrc/Makefile)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the SIS engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        dp, u64 = C.POINTER(C.c_double), C.c_uint64
        L.cpprob_sis_abi_version.restype = C.c_int
        L.cpprob_sis_last_error.restype = C.c_char_p
        L.cpprob_sis_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.cpprob_sis_destroy.argtypes = [C.c_void_p]
        L.cpprob_sis_destroy.restype = None
        L.cpprob_sis_model_name.argtypes = [C.c_int]
        L.cpprob_sis_model_name.restype = C.c_char_p
        L.cpprob_sis_find_model.argtypes = [C.c_char_p]
        L.cpprob_sis_describe.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, C.POINTER(Structure)]
        L.cpprob_sis_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, u64, C.POINTER(RunOptions), C.POINTER(Stats)]
        L.cpprob_sis_infer_to_files.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, u64, C.c_char_p, C.c_int, C.POINTER(Stats)]
        L.cpprob_sis_run_shard.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, u64, C.c_int, C.c_int, dp, C.POINTER(Partials)]
        L.cpprob_sis_merge.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, C.c_void_p, C.c_uint32, C.c_int, C.c_double, u64,
                                       C.POINTER(Stats)]
        L.cpprob_sis_merge_padded.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_double, u64,
                                              C.POINTER(Stats)]
        L.cpprob_sis_replay.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, dp, C.POINTER(C.c_int32), u64, u64, dp]
        L.cpprob_sis_reduce_records.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(C.c_int32), C.c_int, dp, u64, u64, C.POINTER(Stats)]
        L.cpprob_sis_logpdf.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, dp, u64, dp]
        L.cpprob_sis_sample.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, u64, u64, u64, dp]
        L.cpprob_sis_philox.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64, C.POINTER(C.c_uint32)]
        L.cpprob_sis_dmath.argtypes = [C.c_void_p, C.c_int, dp, u64, dp]
        L.cpprob_sis_measure_dfma_peak.argtypes = [C.c_void_p, dp, dp]
        L.cpprob_sis_measure_store_peak.argtypes = [C.c_void_p, dp]
        L.cpprob_sis_write_summary.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Stats)]
        L.cpprob_sis_run_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_size_t, u64, C.POINTER(Stats)]
        L.cpprob_sis_set_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.cpprob_sis_infer_to_files_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_size_t, u64, C.c_char_p, C.POINTER(Stats)]
        L.cpprob_sis_comm_get_id.argtypes = [C.c_void_p]
        L.cpprob_sis_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.cpprob_sis_comm_init_local.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.cpprob_sis_comm_destroy.argtypes = [C.c_void_p]
        L.cpprob_sis_comm_exchange.argtypes = [C.c_void_p]
        L.cpprob_sis_run_dist.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t, u64, C.POINTER(Stats)]
        L.cpprob_sis_text_stage_stats.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(u64), C.POINTER(u64)]
        L.cpprob_sis_probe_issue.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpprob_sis_probe_dfma_chains.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp]
        L.cpprob_sis_plan_rows.argtypes = [u64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.cpprob_sis_plan_shard.argtypes = [u64, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.POINTER(u64), C.POINTER(u64)]
        _lib = L
        return L


def plan_rows(n_total, rank, world, rows_per_chunk=1):
    """(row_first, n_rows_local, n_rows_total): the partial rows `rank` hands to the gather — host arithmetic only."""
    rf, nl, nt = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().cpprob_sis_plan_rows(int(n_total), rank, world, int(rows_per_chunk), C.byref(rf), C.byref(nl), C.byref(nt)))
    return rf.value, nl.value, nt.value


def plan_shard(n_total, rank, world):
    """(chunk_first, n_chunks_local, n_chunks_total, first_particle, n_local) of `rank` — host arithmetic only."""
    cf, ncl, nct = C.c_uint32(), C.c_uint32(), C.c_uint32()
    fp, nl = C.c_uint64(), C.c_uint64()
    _check(lib().cpprob_sis_plan_shard(int(n_total), rank, world, C.byref(cf), C.byref(ncl), C.byref(nct), C.byref(fp), C.byref(nl)))
    return cf.value, ncl.value, nct.value, fp.value, nl.value


class SisError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cpprob_sis error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc < 0:
        raise SisError(rc, lib().cpprob_sis_last_error().decode())
    return rc


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def stats_to_dict(st, structure=None):
    d = {k: getattr(st, k) for k in ("n_particles", "n_neg_inf", "n_nan", "m_ref", "max_log_w", "log_sum_exp",
                                      "log_evidence", "ess", "n_real", "n_int", "int_lo", "int_bins", "n_cols",
                                      "device_ms", "kernel_launches", "passes", "particle_ms")}
    d["path"] = PATHS[st.path] if 0 <= st.path < len(PATHS) else str(st.path)
    d["real_mean"] = np.array([st.real_mean[i] for i in range(st.n_real)])
    d["real_var"] = np.array([st.real_var[i] for i in range(st.n_real)])
    d["int_prob"] = np.array([st.int_prob[i] for i in range(st.n_int * st.int_bins)]).reshape(st.n_int, st.int_bins or 1)[:, :st.int_bins]
    d["int_map"] = np.array([st.int_map[i] for i in range(st.n_int)], dtype=np.int64)
    d["sums"] = np.array([st.sums[i] for i in range(st.n_cols)])
    return d


def comm_get_id():
    """cpprob_sis_comm_get_id: the 128 bytes rank 0 hands to every rank of a new communicator."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(lib().cpprob_sis_comm_get_id(buf))
    return bytes(buf)


def comm_init_local(engines):
    """cpprob_sis_comm_init_local: one process, one engine per GPU, rank r = engines[r]."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().cpprob_sis_comm_init_local(arr, len(engines)))


def infer_to_files_multi(engines, model, obs, n, prefix):
    """cpprob_sis_infer_to_files_multi: the posterior files of one inference written by several GPUs, rank r = engines[r]."""
    obs = _f64(obs)
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    st = Stats()
    _check(lib().cpprob_sis_infer_to_files_multi(arr, len(engines), engines[0].model_id(model), _dptr(obs), obs.size, int(n),
                                                 prefix.encode(), C.byref(st)))
    return stats_to_dict(st)


def run_multi(engines, model, obs, n):
    """cpprob_sis_run_multi: one inference over several GPUs of this process (engines[0] owns the results)."""
    obs = _f64(obs)
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    st = Stats()
    _check(lib().cpprob_sis_run_multi(arr, len(engines), engines[0].model_id(model), _dptr(obs), obs.size, int(n), C.byref(st)))
    return stats_to_dict(st)


class Engine:
    """One GPU's SIS engine (cpprob_sis_create / cpprob_sis_destroy)."""

    def __init__(self, device=0, seed=0x5EED, max_batch=0, blocks_per_sm=0):
        self._L = lib()
        self._h = C.c_void_p()
        cfg = Config(device, seed, max_batch, blocks_per_sm)
        _check(self._L.cpprob_sis_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.cpprob_sis_destroy(self._h)
            self._h = C.c_void_p()

    def set_seed(self, seed):
        _check(self._L.cpprob_sis_set_seed(self._h, int(seed)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- registry -----------------------------------------------------------------------------
    def model_id(self, name):
        return _check(self._L.cpprob_sis_find_model(name.encode()))

    def models(self):
        return [self._L.cpprob_sis_model_name(i).decode() for i in range(self._L.cpprob_sis_model_count())]

    def describe(self, model, obs):
        obs = _f64(obs)
        s = Structure()
        _check(self._L.cpprob_sis_describe(self._h, self.model_id(model), _dptr(obs), obs.size, C.byref(s)))
        return {"ids": [s.ids[i].decode() for i in range(s.n_ids)],
                "slots": [(s.slots[i].is_int, s.slots[i].id, s.slots[i].k, s.slots[i].row) for i in range(s.n_slots)],
                "widths": [s.slots[i].width for i in range(s.n_slots)],
                "n_real": s.n_real, "n_int": s.n_int, "n_samples": s.n_samples}

    # ---- inference ----------------------------------------------------------------------------
    def run(self, model, obs, n, emit=EMIT_NONE, force_rows=False, collect=False):
        """cpprob_sis_run.  With collect=True (implies EMIT_ALL) also returns the SoA trace as numpy arrays."""
        obs = _f64(obs)
        blocks = []

        def on_block(_user, blk):
            b = blk.contents
            n_, st = int(b.n), int(b.stride)
            real = np.ctypeslib.as_array(b.real_rows, shape=(b.n_real, st))[:, :n_].copy() if b.n_real else np.zeros((0, n_))
            ints = np.ctypeslib.as_array(b.int_rows, shape=(b.n_int, st))[:, :n_].copy() if b.n_int else np.zeros((0, n_), np.int32)
            lw = np.ctypeslib.as_array(b.log_w, shape=(n_,)).copy()
            blocks.append((int(b.first_particle), real, ints, lw))
            return 0

        cb = BLOCK_FN(on_block) if collect else C.cast(None, BLOCK_FN)
        opt = RunOptions(EMIT_ALL if collect else emit, 1 if force_rows else 0, cb, None)
        st = Stats()
        _check(self._L.cpprob_sis_run(self._h, self.model_id(model), _dptr(obs), obs.size, int(n), C.byref(opt), C.byref(st)))
        out = stats_to_dict(st)
        if collect:
            blocks.sort(key=lambda b: b[0])
            out["real_rows"] = np.concatenate([b[1] for b in blocks], axis=1)
            out["int_rows"] = np.concatenate([b[2] for b in blocks], axis=1)
            out["log_w"] = np.concatenate([b[3] for b in blocks])
        return out

    def infer_to_files(self, model, obs, n, prefix, emit=EMIT_ALL):
        obs = _f64(obs)
        st = Stats()
        _check(self._L.cpprob_sis_infer_to_files(self._h, self.model_id(model), _dptr(obs), obs.size, int(n), prefix.encode(), emit,
                                                 C.byref(st)))
        return stats_to_dict(st)

    def text_stage_stats(self):
        """Stage times of the last infer_to_files(EMIT_ALL): GPU text kernels, D2H copies, host file writes."""
        k, c, w = C.c_double(), C.c_double(), C.c_double()
        b, f = C.c_uint64(), C.c_uint64()
        _check(self._L.cpprob_sis_text_stage_stats(self._h, C.byref(k), C.byref(c), C.byref(w), C.byref(b), C.byref(f)))
        return {"kernel_ms": k.value, "copy_ms": c.value, "write_s": w.value, "bytes": b.value, "fixups": f.value}

    def comm_init(self, comm_id, rank, world):
        """cpprob_sis_comm_init (collective over the ranks of the new communicator)."""
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(comm_id))
        _check(self._L.cpprob_sis_comm_init(self._h, buf, rank, world))

    def comm_exchange(self):
        """How this engine's communicator moves the partial rows: "none" (one rank), "nccl" or "peer" (peer memory)."""
        return ("none", "nccl", "peer")[self._L.cpprob_sis_comm_exchange(self._h)]

    def run_dist(self, model, obs, n_total):
        """cpprob_sis_run_dist (collective): this rank's shard, one NCCL all-gather, the merge; bit-identical everywhere."""
        obs = _f64(obs)
        st = Stats()
        _check(self._L.cpprob_sis_run_dist(self._h, self.model_id(model), _dptr(obs), obs.size, int(n_total), C.byref(st)))
        return stats_to_dict(st)

    def prepared(self, model, obs, n_total, dist=False):
        """A zero-argument callable that makes exactly one cpprob_sis_run (dist=False) or cpprob_sis_run_dist call with
        everything marshalled beforehand and returns the raw Stats struct (valid until the engine's next call;
        stats_to_dict turns it into the usual dict).  For timing loops: nothing but the C-ABI call is inside."""
        obs = _f64(obs)
        obs_p, n_obs, mid, n = _dptr(obs), obs.size, self.model_id(model), C.c_uint64(int(n_total))
        st, h = Stats(), self._h
        st_ref = C.byref(st)
        if dist:
            fn = self._L.cpprob_sis_run_dist

            def call():
                rc = fn(h, mid, obs_p, n_obs, n, st_ref)
                if rc < 0:
                    _check(rc)
                return st
        else:
            fn = self._L.cpprob_sis_run
            opt = RunOptions(EMIT_NONE, 0, C.cast(None, BLOCK_FN), None)
            opt_ref = C.byref(opt)

            def call():
                rc = fn(h, mid, obs_p, n_obs, n, opt_ref, st_ref)
                if rc < 0:
                    _check(rc)
                return st
        call.keepalive = (obs, st)
        return call

    def run_shard(self, model, obs, n_total, rank, world, m_ref=None):
        obs = _f64(obs)
        p = Partials()
        mr = C.c_double(m_ref) if m_ref is not None else None
        _check(self._L.cpprob_sis_run_shard(self._h, self.model_id(model), _dptr(obs), obs.size, int(n_total), rank, world,
                                            C.byref(mr) if mr is not None else None, C.byref(p)))
        return p

    def merge(self, model, obs, gathered_ptr, n_chunks_total, n_cols, m_ref, n_total):
        """gathered_ptr: device address of the [n_chunks_total][n_cols] partials.  Returns (stats, needs_rebase)."""
        obs = _f64(obs)
        st = Stats()
        rc = _check(self._L.cpprob_sis_merge(self._h, self.model_id(model), _dptr(obs), obs.size, C.c_void_p(gathered_ptr),
                                             n_chunks_total, n_cols, m_ref, int(n_total), C.byref(st)))
        return stats_to_dict(st), rc == 1

    def merge_padded(self, model, obs, gathered_ptr, world, rows_per_rank, rows_per_chunk, n_cols, m_ref, n_total):
        """gathered_ptr: device address of the raw all-gather output, `world` segments of rows_per_rank rows."""
        obs = _f64(obs)
        st = Stats()
        rc = _check(self._L.cpprob_sis_merge_padded(self._h, self.model_id(model), _dptr(obs), obs.size, C.c_void_p(int(gathered_ptr)),
                                                    world, rows_per_rank, rows_per_chunk, n_cols, m_ref, int(n_total), C.byref(st)))
        return stats_to_dict(st), rc == 1

    def replay(self, model, obs, real_rows=None, int_rows=None):
        obs = _f64(obs)
        real = _f64(real_rows) if real_rows is not None and np.size(real_rows) else None
        ints = np.ascontiguousarray(np.asarray(int_rows, dtype=np.int32)) if int_rows is not None and np.size(int_rows) else None
        ref = real if real is not None else ints
        n = ref.shape[1]
        out = np.empty(n, dtype=np.float64)
        _check(self._L.cpprob_sis_replay(self._h, self.model_id(model), _dptr(obs), obs.size,
                                         _dptr(real) if real is not None else None,
                                         ints.ctypes.data_as(C.POINTER(C.c_int32)) if ints is not None else None,
                                         n, n, _dptr(out)))
        return out

    def reduce_records(self, log_w, real_rows=None, int_rows=None):
        lw = _f64(log_w)
        n = lw.size
        real = _f64(real_rows) if real_rows is not None and np.size(real_rows) else None
        ints = np.ascontiguousarray(np.asarray(int_rows, dtype=np.int32)) if int_rows is not None and np.size(int_rows) else None
        st = Stats()
        _check(self._L.cpprob_sis_reduce_records(self._h, _dptr(real) if real is not None else None, 0 if real is None else real.shape[0],
                                                 ints.ctypes.data_as(C.POINTER(C.c_int32)) if ints is not None else None,
                                                 0 if ints is None else ints.shape[0], _dptr(lw), n, n, C.byref(st)))
        return stats_to_dict(st)

    # ---- device distribution layer --------------------------------------------------------------
    def logpdf(self, kind, params, x):
        params, x = _f64(params), _f64(x)
        out = np.empty_like(x)
        _check(self._L.cpprob_sis_logpdf(self._h, DIST[kind], _dptr(params), params.size, _dptr(x), x.size, _dptr(out)))
        return out

    def sample(self, kind, params, n, seed=1, first=0):
        params = _f64(params)
        out = np.empty(int(n), dtype=np.float64)
        _check(self._L.cpprob_sis_sample(self._h, DIST[kind], _dptr(params), params.size, seed, first, int(n), _dptr(out)))
        return out

    def philox(self, ctr, key):
        ctr = np.ascontiguousarray(np.asarray(ctr, dtype=np.uint32)).reshape(-1, 4)
        key = np.ascontiguousarray(np.asarray(key, dtype=np.uint32)).reshape(-1, 2)
        out = np.empty_like(ctr)
        u32p = C.POINTER(C.c_uint32)
        _check(self._L.cpprob_sis_philox(self._h, ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), ctr.shape[0], out.ctypes.data_as(u32p)))
        return out

    def dmath(self, fn, x):
        x = _f64(x)
        n = x.size // 2 if fn == 5 else x.size
        out = np.empty(n, dtype=np.float64)
        _check(self._L.cpprob_sis_dmath(self._h, fn, _dptr(x), n, _dptr(out)))
        return out

    def dfma_peak(self):
        t, mhz = C.c_double(), C.c_double()
        _check(self._L.cpprob_sis_measure_dfma_peak(self._h, C.byref(t), C.byref(mhz)))
        return t.value, mhz.value

    def probe_issue(self, int_per_dfma):
        ms = C.c_double()
        _check(self._L.cpprob_sis_probe_issue(self._h, int_per_dfma, C.byref(ms)))
        return ms.value

    def probe_dfma_chains(self, chains, blocks_per_sm):
        ms, n = C.c_double(), C.c_double()
        _check(self._L.cpprob_sis_probe_dfma_chains(self._h, chains, blocks_per_sm, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def store_peak(self):
        g = C.c_double()
        _check(self._L.cpprob_sis_measure_store_peak(self._h, C.byref(g)))
        return g.value
