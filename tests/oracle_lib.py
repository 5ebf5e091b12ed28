"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE (the CPU restatement of the reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
KIND = {"normal": 0, "uniform_real": 1, "uniform_smallint": 2, "discrete": 3, "poisson": 4}


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    def __init__(self, path):
        L = C.CDLL(path)
        dp, u64 = C.POINTER(C.c_double), C.c_uint64
        L.oracle_logpdf.argtypes = [C.c_int, dp, C.c_int, dp, u64, dp]
        L.oracle_run.argtypes = [C.c_char_p, dp, C.c_int, u64, C.c_char_p, C.c_int, C.c_uint, C.c_int]
        L.oracle_run.restype = C.c_double
        L.oracle_replay_logw.argtypes = [C.c_char_p, dp, C.c_int, dp, u64, C.c_int, dp]
        L.oracle_replay_files.argtypes = [C.c_char_p, dp, C.c_int, dp, u64, u64, C.c_char_p, C.c_int]
        L.oracle_stats_text.argtypes = [C.c_char_p]
        L.oracle_stats_text.restype = C.c_char_p
        ip = C.POINTER(C.c_int)
        L.oracle_stats_real.argtypes = [C.c_char_p, C.c_int, ip, ip, dp, dp]
        L.oracle_stats_int.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, ip, ip, dp, ip, C.POINTER(C.c_uint64)]
        L.oracle_parse_records.argtypes = [C.c_char_p, C.c_int, C.c_int, u64, ip, dp, dp]
        L.oracle_parse_records.restype = C.c_longlong
        u32p = C.POINTER(C.c_uint32)
        L.oracle_philox.argtypes = [u32p, u32p, u64, u32p]
        L.oracle_philox.restype = None
        self.L = L

    def logpdf(self, kind, params, x):
        params, x = _f64(params), _f64(x)
        out = np.empty_like(x)
        assert self.L.oracle_logpdf(KIND[kind], _dp(params), params.size, _dp(x), x.size, _dp(out)) == 0
        return out

    def run(self, model, obs, n, prefix, how="fast", seed=1, progress=0):
        obs = _f64(obs)
        s = self.L.oracle_run(model.encode(), _dp(obs), obs.size, int(n), prefix.encode(), 0 if how == "faithful" else 1, seed, progress)
        assert s >= 0, "oracle does not know this model"
        return s

    def replay_files(self, model, obs, values, prefix, how="faithful", n_traces=None):
        """The restated inference loop on prescribed sampled values [n_traces][per trace] (or a flat list with n_traces given,
        for traces of different lengths): writes <prefix>.real/.int/.ids."""
        obs, values = _f64(obs), _f64(values)
        rc = self.L.oracle_replay_files(model.encode(), _dp(obs), obs.size, _dp(values), values.size,
                                        values.shape[0] if n_traces is None else int(n_traces), prefix.encode(), 0 if how == "faithful" else 1)
        assert rc == 0, f"oracle_replay_files: {rc}"

    def replay_logw(self, model, obs, values):
        """values: [n_traces][values_per_trace] sampled values in program order."""
        obs, values = _f64(obs), _f64(values)
        n, per = values.shape
        out = np.empty(n)
        assert self.L.oracle_replay_logw(model.encode(), _dp(obs), obs.size, _dp(values), n, per, _dp(out)) == 0
        return out

    def stats_text(self, prefix):
        return self.L.oracle_stats_text(prefix.encode()).decode()

    def stats_real(self, prefix, max_rows=4096):
        ids = np.zeros(max_rows, np.int32)
        ks = np.zeros(max_rows, np.int32)
        mean = np.zeros(max_rows)
        var = np.zeros(max_rows)
        ip = C.POINTER(C.c_int)
        n = self.L.oracle_stats_real(prefix.encode(), max_rows, ids.ctypes.data_as(ip), ks.ctypes.data_as(ip), _dp(mean), _dp(var))
        assert n >= 0
        return ids[:n], ks[:n], mean[:n], var[:n]

    def stats_int(self, prefix, lo, bins, max_rows=4096):
        ids = np.zeros(max_rows, np.int32)
        ks = np.zeros(max_rows, np.int32)
        prob = np.zeros((max_rows, bins))
        mp = np.zeros(max_rows, np.int32)
        npts = np.zeros(max_rows, np.uint64)
        ip = C.POINTER(C.c_int)
        n = self.L.oracle_stats_int(prefix.encode(), max_rows, lo, bins, ids.ctypes.data_as(ip), ks.ctypes.data_as(ip), _dp(prob),
                                    mp.ctypes.data_as(ip), npts.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert n >= 0
        return ids[:n], ks[:n], prob[:n], mp[:n], npts[:n]

    def parse_records(self, path, kind, per_record, max_records):
        """kind 'real' | 'int'.  Returns (ids[per_record], values[n][per_record], logw[n])."""
        ids = np.zeros(max(per_record, 1), np.int32)
        values = np.zeros((max_records, max(per_record, 1)))
        logw = np.zeros(max_records)
        n = self.L.oracle_parse_records(path.encode(), 0 if kind == "real" else 1, per_record, max_records,
                                        ids.ctypes.data_as(C.POINTER(C.c_int)), _dp(values), _dp(logw))
        assert n >= 0, f"cannot parse {path}"
        return ids[:per_record], values[:n, :per_record], logw[:n]

    def philox(self, ctr, key):
        ctr = np.ascontiguousarray(np.asarray(ctr, dtype=np.uint32)).reshape(-1, 4)
        key = np.ascontiguousarray(np.asarray(key, dtype=np.uint32)).reshape(-1, 2)
        out = np.empty_like(ctr)
        u32p = C.POINTER(C.c_uint32)
        self.L.oracle_philox(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), ctr.shape[0], out.ctypes.data_as(u32p))
        return out


_cached = None


def load():
    global _cached
    if _cached is None:
        if not os.path.exists(LIB):
            subprocess.run(["make"], cwd=os.path.join(ROOT, "oracle"), check=True)
        _cached = Oracle(LIB)
    return _cached
