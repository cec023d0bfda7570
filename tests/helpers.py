"""Shared builders for the parity tests: the same network fed to the CUDA engine and to the CPU oracle."""
import numpy as np

from oracle_binding import OracleEngine, default_idc, golden_groups, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, load_codes_network


def make_engine(kind, g, dt=0.05, with_synapses=True, idc=None, **kw):
    cls = Engine if kind == "gpu" else OracleEngine
    n = g["NeuPar"].shape[1]
    e = cls(n, dt, **kw)
    syn = g["SynPar"] if with_synapses else np.zeros((7, 0))
    src = g["source_arr"] if with_synapses else np.zeros(0, dtype=np.int32)
    tgt = g["target_arr"] if with_synapses else np.zeros(0, dtype=np.int32)
    load_codes_network(e, g["NeuPar"], g["V0"], g["STypPar"], syn, src, tgt, default_idc(g) if idc is None else idc)
    return e


def pair(g, **kw):
    return make_engine("gpu", g, **kw), make_engine("cpu", g, **kw)


def first_divergence(a_i, a_s, b_i, b_s):
    """Index of the first differing (step, neuron) entry of two sorted spike lists (min length if one is a prefix)."""
    n = min(len(a_i), len(b_i))
    neq = np.nonzero((a_i[:n] != b_i[:n]) | (a_s[:n] != b_s[:n]))[0]
    return int(neq[0]) if len(neq) else n
