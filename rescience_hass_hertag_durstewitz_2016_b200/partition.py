"""Column partition of a network over the GPUs of one box.

Rank r owns a contiguous range of neurons and every synapse whose TARGET is in that range (presynaptic ids stay
global).  The only data-path exchange is the per-step all-gather of spike indices inside libhhd (NCCL); this module is
the host logic around it: slicing the arrays the reference's NetworkSetup returns, keeping synapse indices (RNG keys)
consistent with a single-GPU run, and the flags a rank needs about neurons it does not own.
"""
import numpy as np


def split_bounds(n_neurons, world, align=1):
    """Contiguous ranges [b[r], b[r+1]) with sizes as equal as `align` (e.g. neurons per column) allows."""
    units = n_neurons // align
    if units * align != n_neurons:
        raise ValueError("n_neurons is not a multiple of align")
    cuts = [(units * r) // world * align for r in range(world + 1)]
    return np.asarray(cuts, dtype=np.int64)


def order_by_target_rank(target_arr, bounds):
    """Stable permutation that groups synapses by the rank owning their target.  A single-GPU run must use the SAME
    permuted order for bit-parity with the partitioned run, because release failures are keyed by synapse index."""
    rank_of = np.searchsorted(bounds, np.asarray(target_arr), side="right") - 1
    perm = np.argsort(rank_of, kind="stable")
    counts = np.bincount(rank_of, minlength=len(bounds) - 1)
    return perm, np.concatenate([[0], np.cumsum(counts)])


def partition_network(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, world, align=1):
    """-> (global_net, [local_net per rank]).  `global_net` is the same network with synapses re-ordered by target rank;
    local_net[r] has the keys Engine needs: NeuPar, V0, I_DC (local slices), SynPar/source_arr/target_arr (the rank's
    synapses, global ids), syn_id_base, global_offset, n_local, n_global, has_out (global)."""
    NeuPar, V0, SynPar = np.asarray(NeuPar), np.asarray(V0), np.asarray(SynPar)
    source_arr, target_arr, I_DC = np.asarray(source_arr), np.asarray(target_arr), np.asarray(I_DC)
    n = NeuPar.shape[1]
    bounds = split_bounds(n, world, align)
    perm, syn_bounds = order_by_target_rank(target_arr, bounds)
    g = dict(NeuPar=NeuPar, V0=V0, I_DC=I_DC, SynPar=SynPar[:, perm], source_arr=source_arr[perm], target_arr=target_arr[perm])
    has_out = np.zeros(n, dtype=np.uint8)
    has_out[np.unique(source_arr)] = 1
    locals_ = []
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        s0, s1 = int(syn_bounds[r]), int(syn_bounds[r + 1])
        locals_.append(dict(NeuPar=NeuPar[:, lo:hi], V0=V0[:, lo:hi], I_DC=I_DC[lo:hi], SynPar=g["SynPar"][:, s0:s1],
                            source_arr=g["source_arr"][s0:s1], target_arr=g["target_arr"][s0:s1], syn_id_base=s0,
                            global_offset=lo, n_local=hi - lo, n_global=n, has_out=has_out))
    return g, locals_


def load_partition(engine, local, STypPar):
    """Feed one rank's slice into an Engine created with n_global / global_offset / rank / n_ranks."""
    from .engine import channels_from_STypPar, neuron_rows_from_NeuPar
    engine.set_neurons(neuron_rows_from_NeuPar(local["NeuPar"], local["I_DC"]), local["V0"][0], local["V0"][1])
    engine.set_channels(**channels_from_STypPar(STypPar))
    engine.set_has_outgoing(local["has_out"])
    S = local["SynPar"]
    if S.shape[1]:
        chan = (S[0].astype(np.int64) - 1).astype(np.uint8)
        engine.set_synapses(local["source_arr"], local["target_arr"], chan, S[4], S[1], S[2], S[3], S[6], S[5],
                            syn_id_base=local["syn_id_base"])
