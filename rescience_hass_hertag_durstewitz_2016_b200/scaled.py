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


def tile_synapses_device(SynPar, source_arr, target_arr, n_columns, n_per_column, device, id_offset=0):
    """The synapse arrays of `tile_columns`, built as torch CUDA tensors directly in HBM (PyTorch is used for device
    buffers only): -> (src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms) for Engine.set_synapses_dev.
    1000 columns of the 1003-neuron column are 2.9e8 synapses = 17 GB, which should not be staged through host memory."""
    import torch
    S = torch.as_tensor(np.ascontiguousarray(SynPar, dtype=np.float64), device=device)
    k = int(n_columns)
    offs = (torch.arange(k, device=device, dtype=torch.int64) * int(n_per_column) + int(id_offset))[:, None]
    src1 = torch.as_tensor(np.asarray(source_arr, dtype=np.int64), device=device)[None, :]
    tgt1 = torch.as_tensor(np.asarray(target_arr, dtype=np.int64), device=device)[None, :]
    src = (src1 + offs).reshape(-1).to(torch.int32).contiguous()
    tgt = (tgt1 + offs).reshape(-1).to(torch.int32).contiguous()
    chan = (S[0].to(torch.int64) - 1).to(torch.uint8).repeat(k).contiguous()
    rows = [S[r].repeat(k).contiguous() for r in (4, 1, 2, 3, 6, 5)]   # g_max, U, tau_rec, tau_fac, p_fail, delay
    return (src, tgt, chan) + tuple(rows)
