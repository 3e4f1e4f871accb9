"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes view of oracle/liboracle.so (the plain-C restatement of the reference decode path, see
oracle/ldpc_oracle.h) plus helpers to run / parse the reference dump harness oracle/_ref/dump_ref.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
AWGN, BSC, BEC = 1, 2, 3
KIND = {"AWGN": AWGN, "BSC": BSC, "BEC": BEC}
ERASURE = 69


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


class _Code(ct.Structure):
    _fields_ = [("nc", ct.c_int), ("mc", ct.c_int), ("nnz", ct.c_int),
                ("n_punct", ct.c_int), ("n_short", ct.c_int),
                ("punct", ct.POINTER(ct.c_int)), ("shorten", ct.POINTER(ct.c_int)),
                ("nct", ct.c_int), ("mct", ct.c_int), ("kct", ct.c_int), ("kc", ct.c_int),
                ("max_degree", ct.c_int),
                ("bit_pos", ct.POINTER(ct.c_int)),
                ("e_row", ct.POINTER(ct.c_int)), ("e_col", ct.POINTER(ct.c_int)),
                ("row_ptr", ct.POINTER(ct.c_int)), ("row_edge", ct.POINTER(ct.c_int)),
                ("col_ptr", ct.POINTER(ct.c_int)), ("col_edge", ct.POINTER(ct.c_int))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ct.CDLL(LIB)
        P = ct.POINTER(_Code)
        L.orc_load.restype = P
        L.orc_load.argtypes = [ct.c_char_p]
        L.orc_load_matrix.restype = P
        L.orc_load_matrix.argtypes = [ct.c_char_p, ct.c_int]
        L.orc_free.argtypes = [P]
        dp = ct.POINTER(ct.c_double)
        bp = ct.POINTER(ct.c_uint8)
        L.orc_decode.restype = ct.c_int
        L.orc_decode.argtypes = [P, dp, ct.c_int, ct.c_int, ct.c_int, dp, bp]
        L.orc_decode_bec.restype = ct.c_int
        L.orc_decode_bec.argtypes = [P, bp, bp, ct.c_int, ct.c_int, ct.c_int, bp, bp]
        L.orc_is_codeword.restype = ct.c_int
        L.orc_is_codeword.argtypes = [P, bp]
        L.orc_count_bit_errors.restype = ct.c_int
        L.orc_count_bit_errors.argtypes = [P, bp, bp]
        L.orc_llr_awgn.argtypes = [P, dp, ct.c_double, dp]
        L.orc_llr_bsc.argtypes = [P, bp, ct.c_double, dp]
        L.orc_llr_bec.argtypes = [P, bp, bp, bp]
        L.orc_multiply_left.argtypes = [P, bp, bp]
        L.orc_multiply_right.argtypes = [P, bp, bp]
        L.orc_rank.restype = ct.c_int
        L.orc_rank.argtypes = [P]
        ip = ct.POINTER(ct.c_int)
        L.orc_layers_valid.restype = ct.c_int
        L.orc_layers_valid.argtypes = [P, ct.c_int, ip, ip]
        L.orc_decode_layered.restype = ct.c_int
        L.orc_decode_layered.argtypes = [P, ct.c_int, ip, ip, dp, ct.c_int, ct.c_int, ct.c_int, ct.c_double, dp, bp]
        L.orc_auto_layers.restype = ct.c_int
        L.orc_auto_layers.argtypes = [P, ip]
        L.orc_channel_frame_ask.argtypes = [P, ct.c_int, ip, ip, ct.c_double, ct.c_uint64, ct.c_uint32, ct.c_uint64, bp, dp]
        L.orc_philox4x32_10.argtypes = [ct.POINTER(ct.c_uint32)] * 3
        L.orc_channel_frame.argtypes = [P, P, ct.c_int, ct.c_double, ct.c_uint64, ct.c_uint32, ct.c_uint64, bp, dp, bp]
        L.orc_normal_block.argtypes = [ct.c_uint64, ct.c_uint32, ct.c_uint64, ct.c_uint32, dp]
        L.orc_normal_from_words.argtypes = [ct.c_uint32, ct.c_uint32, dp]
        L.orc_sim_point.argtypes = [P, P, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_double, ct.c_uint64,
                                    ct.c_uint32, ct.c_uint64, ct.c_uint64, ct.c_int, ct.POINTER(ct.c_uint64)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double))


def _bp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_uint8))


class Code:
    """Loaded code file (restates ldpc_code, /root/reference/src/core/ldpc.h:47-71)."""

    def __init__(self, path, matrix_only=False):
        L = lib()
        self._p = L.orc_load_matrix(path.encode(), 0) if matrix_only else L.orc_load(path.encode())
        if not self._p:
            raise FileNotFoundError(path)
        c = self._p.contents
        for k in ("nc", "mc", "nnz", "nct", "mct", "kct", "kc", "max_degree", "n_punct", "n_short"):
            setattr(self, k, getattr(c, k))
        ar = lambda p, n: np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()
        self.puncture = ar(c.punct, c.n_punct)
        self.shorten = ar(c.shorten, c.n_short)
        self.bit_pos = ar(c.bit_pos, c.nct)
        self.e_row = ar(c.e_row, c.nnz)
        self.e_col = ar(c.e_col, c.nnz)
        self.row_ptr = ar(c.row_ptr, c.mc + 1)
        self.row_edge = ar(c.row_edge, c.nnz)
        self.col_ptr = ar(c.col_ptr, c.nc + 1)
        self.col_edge = ar(c.col_edge, c.nnz)

    def __del__(self):
        try:
            lib().orc_free(self._p)
        except Exception:
            pass

    # -- decoders ------------------------------------------------------------------------------
    def decode(self, llr_in, iterations=50, early_term=True, minsum=False):
        """Full-length (nc) LLRs in -> (llr_out[nc], co[nc], iters). One frame or [n, nc]."""
        llr_in = np.ascontiguousarray(llr_in, dtype=np.float64)
        single = llr_in.ndim == 1
        x = llr_in.reshape(-1, self.nc)
        out = np.empty_like(x)
        co = np.empty(x.shape, dtype=np.uint8)
        its = np.empty(x.shape[0], dtype=np.int32)
        L = lib()
        for t in range(x.shape[0]):
            its[t] = L.orc_decode(self._p, _dp(x[t]), iterations, int(early_term), int(minsum), _dp(out[t]), _bp(co[t]))
        if single:
            return out[0], co[0], int(its[0])
        return out, co, its

    # -- layered schedule (opt-in mode; specification) -----------------------------------------
    def auto_layers(self):
        """First-fit layers (no two checks of a layer share a variable) -> list of check-index arrays."""
        lo = np.zeros(self.mc, dtype=np.int32)
        nl = lib().orc_auto_layers(self._p, lo.ctypes.data_as(ct.POINTER(ct.c_int)))
        return [np.nonzero(lo == l)[0].astype(np.int32) for l in range(nl)]

    @staticmethod
    def _layer_arrays(layers):
        ptr = np.zeros(len(layers) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(l) for l in layers])
        chk = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in layers]) if layers else np.zeros(0, np.int32), dtype=np.int32)
        return ptr, chk

    def layers_valid(self, layers):
        ptr, chk = self._layer_arrays(layers)
        ip = ct.POINTER(ct.c_int)
        return bool(lib().orc_layers_valid(self._p, len(layers), ptr.ctypes.data_as(ip), chk.ctypes.data_as(ip)))

    def decode_layered(self, llr_in, layers, iterations=50, early_term=True, minsum=True, ms_scale=1.0):
        llr_in = np.ascontiguousarray(llr_in, dtype=np.float64)
        x = llr_in.reshape(-1, self.nc)
        out = np.empty_like(x)
        co = np.empty(x.shape, dtype=np.uint8)
        its = np.empty(x.shape[0], dtype=np.int32)
        ptr, chk = self._layer_arrays(layers)
        ip = ct.POINTER(ct.c_int)
        for t in range(x.shape[0]):
            its[t] = lib().orc_decode_layered(self._p, len(layers), ptr.ctypes.data_as(ip), chk.ctypes.data_as(ip), _dp(x[t]), iterations,
                                              int(early_term), int(minsum), float(ms_scale), _dp(out[t]), _bp(co[t]))
        return out, co, its

    def decode_bec(self, llr_in, cw, iterations=50, early_term=True, deg1_compat=True):
        llr_in = np.ascontiguousarray(llr_in, dtype=np.uint8).reshape(-1, self.nc)
        cw = np.ascontiguousarray(cw, dtype=np.uint8).reshape(-1, self.nc)
        out = np.empty_like(llr_in)
        co = np.empty_like(llr_in)
        its = np.empty(llr_in.shape[0], dtype=np.int32)
        L = lib()
        for t in range(llr_in.shape[0]):
            its[t] = L.orc_decode_bec(self._p, _bp(llr_in[t]), _bp(cw[t]), iterations, int(early_term),
                                      int(deg1_compat), _bp(out[t]), _bp(co[t]))
        return out, co, its

    def is_codeword(self, co):
        co = np.ascontiguousarray(co, dtype=np.uint8)
        return bool(lib().orc_is_codeword(self._p, _bp(co)))

    def count_bit_errors(self, co, cw):
        return lib().orc_count_bit_errors(self._p, _bp(np.ascontiguousarray(co, np.uint8)), _bp(np.ascontiguousarray(cw, np.uint8)))

    def syndrome(self, word):
        word = np.ascontiguousarray(word, dtype=np.uint8)
        s = np.zeros(self.mc, dtype=np.uint8)
        lib().orc_multiply_right(self._p, _bp(word), _bp(s))
        return s

    def multiply_left(self, left):
        """left[mc] * M -> [nc] (M = this matrix; used with a generator matrix)."""
        left = np.ascontiguousarray(left, dtype=np.uint8)
        r = np.zeros(self.nc, dtype=np.uint8)
        lib().orc_multiply_left(self._p, _bp(left), _bp(r))
        return r

    def rank(self):
        return lib().orc_rank(self._p)

    # -- new counter-based channel (specification) ---------------------------------------------
    def channel_frames(self, kind, x, seed, point, frame0, nframes, gen=None):
        """-> (cw[n,nc] u8, llr[n,nc] f64 or u8 for BEC)"""
        k = KIND[kind] if isinstance(kind, str) else kind
        cw = np.zeros((nframes, self.nc), dtype=np.uint8)
        lf = np.zeros((nframes, self.nc), dtype=np.float64)
        lu = np.zeros((nframes, self.nc), dtype=np.uint8)
        L = lib()
        gp = gen._p if gen is not None else None
        for t in range(nframes):
            L.orc_channel_frame(self._p, gp, k, float(x), int(seed), int(point), int(frame0 + t), _bp(cw[t]), _dp(lf[t]), _bp(lu[t]))
        return cw, (lu if k == BEC else lf)

    def default_modulation(self, M):
        """Gray labels and the consecutive bit mapper (symbol i carries transmitted positions i*bits .. i*bits+bits-1)."""
        bits = int(M).bit_length() - 1
        labels = np.array([j ^ (j >> 1) for j in range(M)], dtype=np.int32)
        n_sym = self.nct // bits
        bm = np.ascontiguousarray(self.bit_pos[:n_sym * bits].reshape(n_sym, bits).T, dtype=np.int32)
        return labels, bm

    def channel_frames_ask(self, M, labels, bm, snr, seed, point, frame0, nframes):
        """-> (scrambling bits [n, nc] u8, llr [n, nc] f64) of the M-ASK / bit-metric channel specification"""
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        bm = np.ascontiguousarray(bm, dtype=np.int32)
        cw = np.zeros((nframes, self.nc), dtype=np.uint8)
        llr = np.zeros((nframes, self.nc), dtype=np.float64)
        ip = ct.POINTER(ct.c_int)
        for t in range(nframes):
            lib().orc_channel_frame_ask(self._p, int(M), labels.ctypes.data_as(ip), bm.ctypes.data_as(ip), float(snr), int(seed), int(point),
                                        int(frame0 + t), _bp(cw[t]), _dp(llr[t]))
        return cw, llr

    def sim_point(self, kind, x, seed=0, point=0, frame0=0, nframes=100, decoding="BP", iterations=50,
                  early_term=True, bec_deg1_compat=True, threads=1, gen=None):
        """-> dict(fec, bec, frames, iters) over frames [frame0, frame0+nframes)."""
        k = KIND[kind] if isinstance(kind, str) else kind
        cnt = (ct.c_uint64 * 4)()
        gp = gen._p if gen is not None else None
        lib().orc_sim_point(self._p, gp, k, int(decoding == "BP_MS"), int(iterations), int(early_term),
                            int(bec_deg1_compat), float(x), int(seed), int(point), int(frame0), int(nframes),
                            int(threads), cnt)
        return dict(fec=int(cnt[0]), bec=int(cnt[1]), frames=int(cnt[2]), iters=int(cnt[3]))


def philox4x32_10(ctr, key):
    c = (ct.c_uint32 * 4)(*ctr)
    k = (ct.c_uint32 * 2)(*key)
    o = (ct.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def normal_block(seed, point, frame, j):
    """The four standard normals of Philox block j (channel specification v2)."""
    z = np.zeros(4)
    lib().orc_normal_block(int(seed), int(point), int(frame), int(j), _dp(z))
    return z


def normal_from_words(wr, wa):
    z = np.zeros(2)
    lib().orc_normal_from_words(int(wr), int(wa), _dp(z))
    return z


# ---- reference dump harness (container only) -------------------------------------------------
def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "dump_ref"))


def _parse_dump(path, with_inputs):
    raw = np.fromfile(path, dtype=np.uint8)
    nc, mc, nnz, nframes, is_bec = np.frombuffer(raw[:20].tobytes(), dtype=np.int32)
    w = 1 if is_bec else 8
    rec = (nc + nc * w if with_inputs else 0) + nc * w + nc + 4
    body = raw[20:].reshape(nframes, rec)
    o = 0
    res = {}
    ty = np.uint8 if is_bec else np.float64

    def take(n, dt):
        nonlocal o
        a = np.ascontiguousarray(body[:, o:o + n]).view(dt)
        o += n
        return a
    if with_inputs:
        res["cw"] = take(nc, np.uint8)
        res["llr_in"] = take(nc * w, ty)
    res["llr_out"] = take(nc * w, ty)
    res["co"] = take(nc, np.uint8)
    res["iters"] = take(4, np.int32).reshape(-1)
    return res


def ref_sim_dump(h, channel, decoding, iters, early_term, x, seed, nframes, g=None, tmp="/tmp/orc_dump.bin"):
    """Runs the unmodified reference channel+decoder for nframes; returns cw/llr_in/llr_out/co/iters."""
    subprocess.run([os.path.join(REF_DIR, "dump_ref"), "sim", h, g or "-", channel, decoding, str(iters),
                    str(int(early_term)), repr(float(x)), str(seed), str(nframes), tmp], check=True)
    return _parse_dump(tmp, True)


def ref_decode(h, decoding, iters, early_term, llr_full, tmp="/tmp/orc_dec"):
    """Decodes caller-supplied full-length LLRs [n, nc] with the unmodified reference decoder."""
    llr_full = np.ascontiguousarray(llr_full, dtype=np.float64)
    llr_full.tofile(tmp + ".in")
    subprocess.run([os.path.join(REF_DIR, "dump_ref"), "decode", h, decoding, str(iters), str(int(early_term)),
                    tmp + ".in", str(llr_full.shape[0]), tmp + ".out"], check=True)
    return _parse_dump(tmp + ".out", False)
