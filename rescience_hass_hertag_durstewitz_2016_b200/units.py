"""Stand-in for the Brian 2 unit names the reference scripts use (`x/ms`, `250*pA`, `Duration*ms` ...).

Brian 2 stores every quantity in SI base units and `quantity/unit` yields a plain number; the scripts only ever use
units that way (codes/Main.py:640-642, codes/CortexNetwork.py:749-754, 3480-3512, 4300-4301).  With Brian absent, units
are plain floats (ms = 1e-3 ...), quantities are plain ndarrays in SI, and the same expressions give the same numbers.
"""
second = 1.0
ms = 1e-3
us = 1e-6
volt = 1.0
mV = 1e-3
amp = 1.0
nA = 1e-9
pA = 1e-12
siemens = 1.0
nS = 1e-9
farad = 1.0
pF = 1e-12
Hz = 1.0
kHz = 1e3
coulomb = 1.0

# engine units (ms, mV, pA, nS, pF) -> SI, per monitored variable
SI_SCALE = {"V": mV, "w": pA, "last_spike": ms}
for _n in ("g_AMPA_on", "g_AMPA_off", "g_GABA_on", "g_GABA_off", "g_NMDA_on", "g_NMDA_off", "g_AMPA", "g_GABA", "g_NMDA"):
    SI_SCALE[_n] = nS
for _n in ("I_AMPA", "I_GABA", "I_NMDA", "I_syn", "I_inj", "I_tot", "I_EXC", "I_EXC_tot", "I_exp"):
    SI_SCALE[_n] = pA
