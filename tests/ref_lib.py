"""ctypes wrapper of oracle/_ref/libcpprob_ref.so — TEST INFRASTRUCTURE.

That library is the REFERENCE'S OWN serialization / NDArray / EmpiricalDistribution / StatsPrinter headers and its
log-pdf headers (normal, uniform_real, uniform_smallint, discrete, poisson, diagonal multivariate normal), compiled from
/root/reference by oracle/Makefile (target _ref) behind the C driver oracle/ref_driver.cpp.  It is built in the container
that has /root/reference and travels to the GPU box as a built file; nothing here reads /root/reference at run time."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libcpprob_ref.so")
STATS_PRINTER = os.path.join(ROOT, "oracle", "_ref", "ref_stats_printer")
REF_SIS = os.path.join(ROOT, "oracle", "_ref", "ref_sis")


def ref_sis(model, obs, n, prefix, replay=None):
    """The REFERENCE'S OWN cpprob::inference(StateType::sis, ...) (oracle/ref_sis.cpp: its state.cpp, trace.cpp, utils.cpp,
    models ... linked unmodified).  `replay`: the values its sample statements return, [n][per trace] in program order
    (None: it draws by itself).  Writes <prefix>.real / .int / .any / .ids as the reference does; returns the wall seconds of
    the inference call."""
    args = [REF_SIS, model, str(int(n)), prefix]
    tmp = None
    if replay is not None:
        tmp = prefix + ".replay.f64"
        np.ascontiguousarray(replay, dtype=np.float64).tofile(tmp)
        args.append(tmp)
    else:
        args.append("-")
    args += [repr(float(x)) for x in obs]
    r = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if tmp:
        os.remove(tmp)
    assert r.returncode == 0, r.stderr[-2000:]
    return float(r.stderr.strip().split()[-1])


def available():
    if not os.path.exists(LIB) and os.path.isdir("/root/reference/include/cpprob"):
        subprocess.run(["make", "_ref"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB)


class Ref:
    def __init__(self):
        L = C.CDLL(LIB)
        dp, u64p, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int)
        L.ref_describe.restype = C.c_char_p
        L.ref_write_real_record.argtypes = [u64p, dp, C.c_int, C.c_double, C.c_char_p, C.c_int]
        L.ref_write_int_record.argtypes = [u64p, C.POINTER(C.c_longlong), C.c_int, C.c_double, C.c_char_p, C.c_int]
        L.ref_write_ndarray_record.argtypes = [u64p, ip, dp, C.c_int, C.c_double, C.c_char_p, C.c_int]
        L.ref_parse_real_record.argtypes = [C.c_char_p, u64p, dp, C.c_int, dp]
        L.ref_parse_int_record.argtypes = [C.c_char_p, u64p, ip, C.c_int, dp]
        L.ref_reprint_record.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_parse_real_record_ndarray.argtypes = [C.c_char_p]
        L.ref_empirical_real.argtypes = [dp, dp, C.c_uint64, dp, dp]
        L.ref_empirical_int.argtypes = [ip, dp, C.c_uint64, C.c_int, ip, dp, ip, u64p]
        L.ref_stats_text.argtypes = [C.c_char_p]
        L.ref_logpdf.argtypes = [C.c_int, dp, C.c_int, dp, C.c_uint64, dp]
        L.ref_logpdf_mvn.argtypes = [dp, dp, C.c_int, dp, C.c_uint64, dp]
        L.ref_stats_text.restype = C.c_char_p
        self.L = L

    def describe(self):
        return self.L.ref_describe().decode()

    # ---- writer (serialization.hpp operator<< under dump_predicts' stream state) ----
    def write_real(self, ids, vals, log_w):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        buf = C.create_string_buffer(64 + 40 * max(1, ids.size))
        n = self.L.ref_write_real_record(ids.ctypes.data_as(C.POINTER(C.c_uint64)), vals.ctypes.data_as(C.POINTER(C.c_double)),
                                         ids.size, log_w, buf, len(buf))
        assert n >= 0
        return buf.raw[:n]

    def write_int(self, ids, vals, log_w):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.int64)
        buf = C.create_string_buffer(64 + 48 * max(1, ids.size))
        n = self.L.ref_write_int_record(ids.ctypes.data_as(C.POINTER(C.c_uint64)), vals.ctypes.data_as(C.POINTER(C.c_longlong)),
                                        ids.size, log_w, buf, len(buf))
        assert n >= 0
        return buf.raw[:n]

    def write_ndarray(self, ids, widths, vals, log_w):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        widths = np.ascontiguousarray(widths, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        buf = C.create_string_buffer(64 + 40 * max(1, vals.size) + 8 * ids.size)
        n = self.L.ref_write_ndarray_record(ids.ctypes.data_as(C.POINTER(C.c_uint64)), widths.ctypes.data_as(C.POINTER(C.c_int)),
                                            vals.ctypes.data_as(C.POINTER(C.c_double)), ids.size, log_w, buf, len(buf))
        assert n >= 0
        return buf.raw[:n]

    # ---- parser (serialization.hpp operator>> as StatsPrinter::load_distr calls it) ----
    def parse(self, line, kind, cap=4096):
        """line: bytes without the newline.  Returns (ids, values, log_w) or None if the reference's parser rejects it."""
        ids = np.zeros(cap, np.uint64)
        lw = C.c_double()
        if kind == "real":
            vals = np.zeros(cap, np.float64)
            n = self.L.ref_parse_real_record(line, ids.ctypes.data_as(C.POINTER(C.c_uint64)), vals.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(lw))
        else:
            vals = np.zeros(cap, np.int32)
            n = self.L.ref_parse_int_record(line, ids.ctypes.data_as(C.POINTER(C.c_uint64)), vals.ctypes.data_as(C.POINTER(C.c_int)), cap, C.byref(lw))
        assert n != -2, "cap too small"
        if n < 0:
            return None
        return ids[:n].copy(), vals[:n].copy(), lw.value

    def reprint(self, line, kind):
        buf = C.create_string_buffer(2 * len(line) + 256)
        n = self.L.ref_reprint_record(line, 0 if kind == "real" else 1, buf, len(buf))
        return None if n < 0 else buf.raw[:n]

    def parse_ndarray_ok(self, line):
        return self.L.ref_parse_real_record_ndarray(line)

    def parse_file(self, path, kind, per_record):
        """All records of a posterior file through the reference's parser: (ids[per], values[n][per], log_w[n])."""
        vals, lws, ids0 = [], [], None
        with open(path, "rb") as f:
            for line in f:
                r = self.parse(line.rstrip(b"\n"), kind, cap=per_record + 1)
                assert r is not None, f"the reference's parser rejects {line[:80]!r}"
                ids, v, lw = r
                assert ids.size == per_record
                ids0 = ids if ids0 is None else ids0
                vals.append(v)
                lws.append(lw)
        return ids0, np.array(vals), np.array(lws)

    # ---- EmpiricalDistribution ----
    def empirical_real(self, x, log_w):
        x, log_w = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(log_w, np.float64)
        m, v = C.c_double(), C.c_double()
        self.L.ref_empirical_real(x.ctypes.data_as(C.POINTER(C.c_double)), log_w.ctypes.data_as(C.POINTER(C.c_double)), x.size, C.byref(m), C.byref(v))
        return m.value, v.value

    def empirical_int(self, x, log_w, cap=4096):
        x, log_w = np.ascontiguousarray(x, np.int32), np.ascontiguousarray(log_w, np.float64)
        values, probs = np.zeros(cap, np.int32), np.zeros(cap)
        mp, npts = C.c_int(), C.c_uint64()
        k = self.L.ref_empirical_int(x.ctypes.data_as(C.POINTER(C.c_int)), log_w.ctypes.data_as(C.POINTER(C.c_double)), x.size, cap,
                                     values.ctypes.data_as(C.POINTER(C.c_int)), probs.ctypes.data_as(C.POINTER(C.c_double)),
                                     C.byref(mp), C.byref(npts))
        assert k >= 0
        return dict(zip(values[:k].tolist(), probs[:k].tolist())), mp.value, npts.value

    # ---- logpdf<> of utils_normal_distribution.hpp / utils_uniform_real.hpp / utils_uniform_smallint.hpp / utils_discrete.hpp / utils_poisson.hpp ----
    def logpdf(self, kind, params, x):
        kinds = {"normal": 0, "uniform_real": 1, "uniform_smallint": 2, "discrete": 3, "poisson": 4}
        params, x = np.ascontiguousarray(params, np.float64), np.ascontiguousarray(x, np.float64)
        out = np.empty_like(x)
        dp = C.POINTER(C.c_double)
        assert self.L.ref_logpdf(kinds[kind], params.ctypes.data_as(dp), params.size, x.ctypes.data_as(dp), x.size, out.ctypes.data_as(dp)) == 0
        return out

    def logpdf_mvn(self, mean, covariance, x):
        """logpdf<multivariate_normal_distribution> (utils_multivariate_normal.hpp:20-33) at the rows of x[n][dim]; the second
        argument is the covariance diagonal (the reference takes its square root, multivariate_normal.hpp:178-186)."""
        mean, sigma = np.ascontiguousarray(mean, np.float64), np.ascontiguousarray(covariance, np.float64)
        x = np.ascontiguousarray(x, np.float64).reshape(-1, mean.size)
        out = np.empty(x.shape[0])
        dp = C.POINTER(C.c_double)
        assert self.L.ref_logpdf_mvn(mean.ctypes.data_as(dp), sigma.ctypes.data_as(dp), mean.size, x.ctypes.data_as(dp), x.shape[0], out.ctypes.data_as(dp)) == 0
        return out

    # ---- StatsPrinter ----
    def stats_text(self, prefix):
        """`std::cout << cpprob::StatsPrinter{prefix} << std::endl`, run as a process (StatsPrinter exits on a bad line)."""
        r = subprocess.run([STATS_PRINTER, prefix], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return r.stdout


def ref_log_w(ref, model, obs, values):
    """log_w of one trace: `log_w += logpdf(distr, x)` per observe statement in program order (state.cpp:212-223,
    cpprob.hpp:79-90), every term from the REFERENCE'S logpdf<> and added in IEEE double (Python floats)."""
    lw = 0.0
    if model in ("gaussian_unknown_mean", "gaussian_unknown_mean_mu"):
        sd = 2.0 if model == "gaussian_unknown_mean" else np.sqrt(2.0)
        for y in obs:
            lw += float(ref.logpdf("normal", [values[0], sd], [y])[0])
    elif model == "linear_gaussian_1d":
        for state, y in zip(values, obs):
            lw += float(ref.logpdf("normal", [state, 1.0], [y])[0])
    elif model == "hmm":
        for s, y in zip(values, obs):
            lw += float(ref.logpdf("normal", [[-1.0, 0.0, 1.0][int(s)], 1.0], [y])[0])
    elif model == "gaussian_2d_unk_mean":
        s2 = np.sqrt(2.0)
        lw += float(ref.logpdf_mvn(values, [s2, s2], obs)[0])
    return lw


_cached = None


def load():
    global _cached
    if _cached is None:
        assert available(), "oracle/_ref is not built (needs /root/reference; `make -C oracle _ref`)"
        _cached = Ref()
    return _cached
