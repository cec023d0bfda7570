"""Run directories and network persistence of codes/SimulationSetup.py (`set_dir`, `save_setup`, `set_setup`, `save_group_distr`,
`set_group_distr`; used by codes/Main.py:580-588 through `'Recover/Save'`).

The reference writes the network as text: NeuPar / V0 / STypPar / SynPar and the DENSE connectivity `SPMtx[:, :, 0..1]` (N x N x 2
floats with np.savetxt — 50 MB for the 1003-neuron column, 25 TB for 1 M neurons).  Here `save_setup` writes ONE `setup.npz` with the
synapse list in coordinate form (target, source — the order of `sourcetarget(SPMtx)`, i.e. the synapse index order the façade
uses), and `set_setup` reads either that file or a directory the reference wrote (the text files above), so existing
`Simulation_<k>` directories stay usable.  Same call signatures; `SPMtx` comes back dense while the network is small enough for
that (NetworkSetup.DENSE_LIMIT) and as the pair (target_arr, source_arr) beyond, exactly as this package's `NetworkSetup` returns it.
"""
import os

import numpy as np

from .NetworkSetup import DENSE_LIMIT, sourcetarget


def set_dir():
    """The first free `Simulation_<k>` directory, created (codes/SimulationSetup.py:23-33)."""
    k = 0
    while os.path.isdir("Simulation_{}".format(k)):
        k += 1
    os.mkdir("Simulation_{}".format(k))
    return "Simulation_{}".format(k)


def save_group_distr(group_distr, sim_dir):
    """`group_distr[group][stripe]` as the reference's one-line text: stripes split by ';', groups by '-', ids by ','."""
    stripes = []
    for s in range(len(group_distr[0])):
        stripes.append("-".join(",".join(str(int(x)) for x in group_distr[g][s]) for g in range(len(group_distr))))
    with open(os.path.join(sim_dir, "group_distr.txt"), "a") as f:
        f.write(";".join(stripes))


def _parse_group_distr(text):
    stripes = text.split(";")
    n_groups = len(stripes[0].split("-"))
    distr = [[[] for _ in stripes] for _ in range(n_groups)]
    for s, stripe in enumerate(stripes):
        for g, group in enumerate(stripe.split("-")[:n_groups]):
            if group != "":
                distr[g][s].extend(int(x) for x in group.split(","))
    return distr


def _sim_path(sim_dir):
    return sim_dir if isinstance(sim_dir, str) and os.path.isdir(sim_dir) else "Simulation_{}".format(sim_dir)


def set_group_distr(sim_dir):
    with open(os.path.join(_sim_path(sim_dir), "group_distr.txt")) as f:
        return _parse_group_distr(f.read())


def _pairs(SPMtx):
    if isinstance(SPMtx, (tuple, list)) and len(SPMtx) == 2 and np.ndim(SPMtx[0]) == 1:
        return np.asarray(SPMtx[0], dtype=np.int64), np.asarray(SPMtx[1], dtype=np.int64)
    return sourcetarget(np.asarray(SPMtx))


def _dense(target_arr, source_arr, n, n_types):
    """The N x N x 2 matrix of 1-based synapse indices the reference's façade indexes (codes/NetworkSetup.py:211-236)."""
    M = np.zeros((n, n, 2))
    k = np.arange(len(target_arr))
    first = np.ones(len(k), dtype=bool)
    seen = {}
    for q, (t, s) in enumerate(zip(target_arr.tolist(), source_arr.tolist())):
        first[q] = (t, s) not in seen
        seen[(t, s)] = True
    M[target_arr[first], source_arr[first], 0] = k[first] + 1
    M[target_arr[~first], source_arr[~first], 1] = k[~first] + 1
    return M


def save_setup(NeuPar, V0, STypPar, SynPar, SPMtx, group_distr, sim_dir):
    """One compressed `setup.npz` (+ the reference's `group_distr.txt`) in `sim_dir`."""
    target_arr, source_arr = _pairs(SPMtx)
    flat = [int(x) for g in group_distr for st in g for x in st]
    offs = np.cumsum([0] + [len(st) for g in group_distr for st in g])
    np.savez_compressed(os.path.join(sim_dir, "setup.npz"), NeuPar=np.asarray(NeuPar), V0=np.asarray(V0), STypPar=np.asarray(STypPar),
                        SynPar=np.asarray(SynPar), target_arr=target_arr, source_arr=source_arr, group_flat=np.asarray(flat, dtype=np.int64),
                        group_offs=offs, n_groups=len(group_distr), n_stripes=len(group_distr[0]))
    save_group_distr(group_distr, sim_dir)


def set_setup(sim_dir, dense_limit=DENSE_LIMIT):
    """-> NeuPar, V0, STypPar, SynPar, SPMtx, group_distr from `Simulation_<sim_dir>` (or the directory `sim_dir`), written by
    `save_setup` here or by the reference's text-file `save_setup`."""
    path = _sim_path(sim_dir)
    npz = os.path.join(path, "setup.npz")
    if os.path.exists(npz):
        d = np.load(npz)
        NeuPar, V0, STypPar, SynPar = d["NeuPar"], d["V0"], d["STypPar"], d["SynPar"]
        target_arr, source_arr = d["target_arr"], d["source_arr"]
        ng, ns, offs, flat = int(d["n_groups"]), int(d["n_stripes"]), d["group_offs"], d["group_flat"]
        group_distr = [[list(map(int, flat[offs[g * ns + s]:offs[g * ns + s + 1]])) for s in range(ns)] for g in range(ng)]
    else:                                                    # a directory written by the reference
        txt = lambda name: np.genfromtxt(os.path.join(path, name + ".txt"), delimiter=",")
        NeuPar, V0, STypPar, SynPar = txt("NeuPar"), txt("V0"), txt("STypPar"), txt("SynPar")
        SPMtx = np.stack([txt("SPMtx0"), txt("SPMtx1")], axis=2)
        group_distr = set_group_distr(sim_dir)
        if NeuPar.shape[1] <= dense_limit:
            return NeuPar, V0, STypPar, SynPar, SPMtx, group_distr
        target_arr, source_arr = sourcetarget(SPMtx)
    n = NeuPar.shape[1]
    SPMtx = _dense(target_arr, source_arr, n, 2) if n <= dense_limit else (target_arr, source_arr)
    return NeuPar, V0, STypPar, SynPar, SPMtx, group_distr
