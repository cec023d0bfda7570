"""The C-ABI library loads, exports every symbol include/hhd.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge
from rescience_hass_hertag_durstewitz_2016_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hhd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hhd_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("hhd_create", "hhd_set_neurons", "hhd_set_synapses", "hhd_add_spike_source", "hhd_add_state_monitor",
                 "hhd_run", "hhd_read_spikes", "hhd_read_monitor", "hhd_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol():
    ge.build()
    lib = engine.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.hhd_abi_version() == engine.ABI_VERSION
    from rescience_hass_hertag_durstewitz_2016_b200 import analysis
    assert sorted(["hhd_" + f for f in engine.FUNCTIONS] + ["hhd_" + f for f in analysis.FUNCTIONS]) == declared_functions()


def test_oracle_exports_the_same_abi():
    from oracle_binding import oracle_lib
    lib = oracle_lib()
    for name in declared_functions():
        if name.startswith("hhd_an_"):
            continue                      # the analyses' oracle is oracle/analysis_oracle.py (numpy)
        assert hasattr(lib, name.replace("hhd_", "hhdo_", 1)), name


def test_no_cpu_fallback():
    """Without a usable B200 the product must refuse to run rather than fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    with pytest.raises(engine.HhdError) as ei:
        engine.Engine(10, 0.05)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_config_struct_layout_matches_header():
    assert C.sizeof(engine.Config) == 4 * 4 + 8 + 3 * 4 + 4 + 3 * 8 + 8 + 4 * 4 + 7 * 4 + 4   # with natural padding
    assert C.sizeof(engine.Counters) == 9 * 8 + 8 + 6 * 8


def test_missing_library_is_an_import_error(tmp_path):
    with pytest.raises(ImportError):
        engine.load_library(str(tmp_path / "libhhd.so"))
