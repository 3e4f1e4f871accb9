"""Drop-in for the reference's Python wrapper (pyLDPC/ldpc.py): same class, same methods, same ctypes
structs passed by value, bound to the six reference symbols of libldpc_b200/libldpc.so.

    from libldpc_b200 import ldpc              # instead of: from pyLDPC import ldpc
    code = ldpc.LDPC("codes/ref_h_n1152_m1024.txt")
    code.simulate(snr=[-6, -3, 0.5], iterations=50, decoding="BP_MS")
    code.stop_simulation(); print(code.get_results())

The unmodified reference wrapper also works against this library: LDPC(pc, gen, lib=<path to libldpc.so>)
(its `lib=` keyword, pyLDPC/ldpc.py:33).
"""
import ctypes as ct
import threading

import numpy as np

from .api import channel_param, decoder_param, lib_path, sim_results_t, simulation_param

LIB_PATH = lib_path()
MAX_POINTS = 50  # pyLDPC/ldpc.py:57-58


class LDPC:
    def __init__(self, pc_file: str, gen_file="", lib=LIB_PATH):  # pyLDPC/ldpc.py:33-81
        self.pc_file = pc_file
        self.gen_file = gen_file
        n, m, nct, mct = ct.c_int(0), ct.c_int(0), ct.c_int(0), ct.c_int(0)
        self.lib = ct.cdll.LoadLibrary(lib)
        self.lib.ldpc_setup(pc_file.encode("utf-8"), gen_file.encode("utf-8"), ct.byref(n), ct.byref(m), ct.byref(nct), ct.byref(mct))
        self.n, self.m, self.nct, self.mct = n.value, m.value, nct.value, mct.value
        self.kct = self.nct - self.mct
        self.k = self.n - self.m
        self.sim_stop_flag = ct.c_bool(False)
        vd, vu = ct.c_double * MAX_POINTS, ct.c_uint64 * MAX_POINTS
        self.sim_results_struct = sim_results_t(vd(), vd(), vd(), vd(), vu(), vu())
        self.results = {}
        self.sim_params = {"earlyTerm": True, "iterations": 50, "decoding": "BP", "seed": 0, "snr": [], "channel": "AWGN",
                           "threads": 1, "maxFrames": int(10e9), "fec": 50}
        self._thread = None

    def encode(self, info_word: np.array) -> np.array:  # pyLDPC/ldpc.py:84-105
        if not self.gen_file:
            raise RuntimeError("No generator matrix provided for encoding")
        in_arr = (ct.c_uint8 * self.kct)(*[int(v) for v in info_word])
        out_arr = (ct.c_uint8 * self.nct)()
        self.lib.encode(ct.byref(in_arr), ct.byref(out_arr))
        return np.array(out_arr[0:self.nct])

    def decode(self, llr_in: np.array, early_term=True, iters=50, dec_type="BP"):  # pyLDPC/ldpc.py:108-132
        dec_params = decoder_param(early_term, iters, dec_type.encode("utf-8"))
        vec = ct.c_double * self.nct
        in_arr = vec(*[float(v) for v in llr_in])
        out_arr = vec()
        self.lib.decode.restype = ct.c_int
        iter_req = self.lib.decode(dec_params, ct.byref(in_arr), ct.byref(out_arr))
        return np.array(out_arr[0:self.nct]), iter_req

    def simulate(self, **args):  # pyLDPC/ldpc.py:135-169
        self.sim_params = {**self.sim_params, **args}
        snr = (ct.c_double * 3)(*self.sim_params["snr"])
        dec_param = decoder_param(self.sim_params["earlyTerm"], self.sim_params["iterations"], self.sim_params["decoding"].encode("utf-8"))
        ch_param = channel_param(self.sim_params["seed"], snr, self.sim_params["channel"].encode("utf-8"))
        sim_param = simulation_param(self.sim_params["threads"], self.sim_params["maxFrames"], self.sim_params["fec"], "".encode("utf-8"))

        def sim_thread():
            self.sim_stop_flag.value = False
            self.lib.simulate(dec_param, ch_param, sim_param, ct.byref(self.sim_results_struct), ct.byref(self.sim_stop_flag))

        self._thread = threading.Thread(target=sim_thread)
        self._thread.start()

    def wait(self, timeout=None):
        """(extension) joins the simulation thread."""
        if self._thread is not None:
            self._thread.join(timeout)

    def stop_simulation(self):  # pyLDPC/ldpc.py:171-177
        if not self.sim_stop_flag.value:
            self.results = self.get_results()
            self.sim_stop_flag.value = 1

    def get_results(self):  # pyLDPC/ldpc.py:179-192
        if not self.sim_stop_flag.value:
            max_index = np.sum(np.array(self.sim_results_struct.frames[0:MAX_POINTS]) > 0)
            return dict([(x, getattr(self.sim_results_struct, x)[0:max_index]) for (x, _) in self.sim_results_struct._fields_])
        return self.results

    def rank(self):  # pyLDPC/ldpc.py:194-200
        return self.lib.calculate_rank()

    def syndrome(self, v: np.array) -> np.array:  # pyLDPC/ldpc.py:202-218
        vec = ct.c_uint8 * self.n
        word = vec(*[int(x) for x in v])
        synd = vec()
        self.lib.syndrome(ct.byref(word), ct.byref(synd))
        return np.array(synd[0:self.m])
