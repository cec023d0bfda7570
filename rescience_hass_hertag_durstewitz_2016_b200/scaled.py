"""Scaled-up inputs for configs 3-5 of BASELINE.json: many columns / many independent instances in one engine.

`tile_columns` repeats a one-column template (the arrays `codes/NetworkSetup.NetworkSetup` returns for Nstripes=1)
K times with shifted neuron ids.  Every copy keeps the template's parameters and intra-column connectivity; what
differs between copies is the synapse index (hence every release-failure draw) and, optionally, the background
current.  With `inter_column=False` this is exactly BASELINE config 3 (independent instances batched on one GPU);
it is also the synthetic stand-in used by bench.py for the 100- and 1000-column workloads until the sparse
multi-column builder replaces it.
"""
import numpy as np


def tile_columns(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, n_columns, jitter_idc=0.0, seed=0):
    """-> dict(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, n_per_column) for n_columns copies."""
    NeuPar = np.asarray(NeuPar, dtype=np.float64)
    n1 = NeuPar.shape[1]
    k = int(n_columns)
    offs = (np.arange(k, dtype=np.int64) * n1)
    src = (np.asarray(source_arr, dtype=np.int64)[None, :] + offs[:, None]).reshape(-1).astype(np.int32)
    tgt = (np.asarray(target_arr, dtype=np.int64)[None, :] + offs[:, None]).reshape(-1).astype(np.int32)
    idc = np.tile(np.asarray(I_DC, dtype=np.float64), k)
    if jitter_idc:
        rng = np.random.default_rng(seed)
        idc = idc * (1.0 + jitter_idc * (rng.random(idc.size) - 0.5))
    return dict(NeuPar=np.tile(NeuPar, (1, k)), V0=np.tile(np.asarray(V0, dtype=np.float64), (1, k)),
                SynPar=np.tile(np.asarray(SynPar, dtype=np.float64), (1, k)), source_arr=src, target_arr=tgt,
                I_DC=idc, n_per_column=n1)
