"""Loads the CPU oracle (oracle/libhhd_oracle.so) behind the same Python Engine class the product uses.

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libhhd_oracle.so")
_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "hhd_oracle.c")
        if (not os.path.exists(ORACLE_LIB)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(ORACLE_LIB)):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        _lib = C.CDLL(ORACLE_LIB)
    return _lib


class OracleEngine(Engine):
    def __init__(self, *a, **kw):
        kw["_lib"] = oracle_lib()
        kw["_prefix"] = "hhdo_"
        super().__init__(*a, **kw)

    def step_begin(self):
        """First half of a step (monitors, state update, thresholds) -> global ids of the local spikes."""
        buf = np.empty(self.n_neurons, dtype=np.int32)
        n = C.c_int32(0)
        f = self._lib.hhdo_step_begin
        f.restype = C.c_int
        self._check(f(self._h, buf.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return buf[:n.value].copy()

    def step_end(self, all_spikes_global):
        a = np.ascontiguousarray(all_spikes_global, dtype=np.int32)
        f = self._lib.hhdo_step_end
        f.restype = C.c_int
        self._check(f(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(a.size)))


def load_golden(name):
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
    return {k: d[k] for k in d.files}


def golden_groups(g):
    """-> group_distr[group][stripe] = array of neuron ids."""
    ns = int(g["nstripes"])
    offs, flat = g["group_offs"], g["group_flat"]
    ngroups = (len(offs) - 1) // ns
    return [[flat[offs[gi * ns + s]:offs[gi * ns + s + 1]] for s in range(ns)] for gi in range(ngroups)]


def default_idc(g, i_exc=250.0, i_inh=200.0):
    """Background current of codes/Main.py:109-112: 250 pA into PC groups (0 and 7), 200 pA into interneurons."""
    groups = golden_groups(g)
    n = g["NeuPar"].shape[1]
    idc = np.zeros(n)
    for gi, per_stripe in enumerate(groups):
        for ids in per_stripe:
            idc[ids] = i_exc if gi in (0, 7) else i_inh
    return idc
