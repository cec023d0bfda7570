"""Generate golden network fixtures by running the REFERENCE's own builder.

Runs only in the build container (needs /root/reference); the GPU box never
executes this.  It imports `codes/ParamSetup.py` + `codes/NetworkSetup.py`
(numpy + scipy only) unmodified, seeds numpy, calls

    NetworkSetup(N1, Nstripes, scales1, clustering=True)      (codes/NetworkSetup.py:9)
    sourcetarget(SPMtx)                                       (codes/NetworkSetup.py:405)

and stores what the simulation facade receives (`codes/Main.py:585-600`).

Usage:  python tests/golden/make_golden_network.py NAME N1 NSTRIPES SEED [--digest-only]
"""
import hashlib
import os
import sys
import time

import numpy as np

REF = "/root/reference/codes"


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    name, n1, nstripes, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    digest_only = "--digest-only" in sys.argv
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from NetworkSetup import NetworkSetup, sourcetarget  # the reference's own code

    scales1 = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))
    np.random.seed(seed)
    t0 = time.time()
    NeuPar, V0, STypPar, SynPar, SPMtx, group_distr = NetworkSetup(n1, nstripes, scales1, True)
    state_after = np.random.random()   # pins the RNG position after construction
    t1 = time.time()
    target_arr, source_arr = sourcetarget(SPMtx)
    nsyn = SynPar.shape[1]
    # group_distr[group][stripe] -> flat + offsets
    flat, offs = [], [0]
    for g in range(len(group_distr)):
        for s in range(nstripes):
            flat.extend(int(x) for x in group_distr[g][s])
            offs.append(len(flat))
    out = dict(
        n1=n1, nstripes=nstripes, seed=seed, nsyn=nsyn, build_seconds=t1 - t0,
        rng_after=state_after,
        STypPar=STypPar,
        group_flat=np.asarray(flat, dtype=np.int32), group_offs=np.asarray(offs, dtype=np.int64),
        sha_NeuPar=digest(NeuPar), sha_V0=digest(V0), sha_SynPar=digest(SynPar),
        sha_target=digest(target_arr.astype(np.int64)), sha_source=digest(source_arr.astype(np.int64)),
        n_target=len(target_arr),
    )
    if not digest_only:
        out.update(NeuPar=NeuPar, V0=V0, SynPar=SynPar,
                   target_arr=target_arr.astype(np.int32), source_arr=source_arr.astype(np.int32))
    else:
        # a thin, deterministic sample so failures can be localised
        idx = np.arange(0, nsyn, max(1, nsyn // 4096))
        out.update(NeuPar=NeuPar, V0=V0, sample_idx=idx, SynPar_sample=SynPar[:, idx],
                   target_sample=target_arr[idx].astype(np.int32), source_sample=source_arr[idx].astype(np.int32))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "N=", NeuPar.shape[1], "Nsyn=", nsyn, "len(target_arr)=", len(target_arr),
          "build %.1fs" % (t1 - t0), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
