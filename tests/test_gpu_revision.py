"""codes_revision call path on the GPU: `Cortex.setup(Ncells, Nstripes, constant_stimuli, method, dt, transient, basics_scales, ...)` as
codes_revision/Tests.py:27-32, 249, 427 use it — the revision's own generator (bit-identical network, tests/test_revision_generator.py),
the revision's model variant (Mg factor 1, last_spike0 = -1000 ms), basics_scales selectors — against the same façade on the CPU oracle."""
import numpy as np
import pytest

from oracle_binding import OracleEngine
from rescience_hass_hertag_durstewitz_2016_b200 import units
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.BasicsSetup import GROUPS_SETS
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.CortexSetup import Cortex

pytestmark = pytest.mark.gpu
CONST = [[("PC", 0), 250], [("IN", 0), 200]]


@pytest.mark.parametrize("method,scales", [("rk4", None), ("rk2", {"gmax_mean": [(dict(target=GROUPS_SETS["PC"], source=GROUPS_SETS["PC"]), 1.5)]})])
def test_cortex_setup_gpu_equals_oracle(method, scales):
    out = {}
    for name, cls in (("gpu", None), ("cpu", OracleEngine)):
        c = Cortex.setup(150, 1, CONST, method, 0.05, transient=20, basics_scales=scales, seed=3, accum="fixed", _engine_cls=cls)
        c.set_neuron_monitors("V", "V", ("PC_L23", 0), start=30, stop=60)
        c.set_longrun_neuron_monitors("I_tot", "I_tot", ("ALL", 0), 5000, start=20, stop=None, population_agroupate="sum")
        c.run(400)
        out[name] = (np.array(c.spikemonitor.i), np.array(c.spikemonitor.t), np.array(c.recorded.var("V")), np.array(c.recorded.var("I_tot")),
                     c.network.channel_constants()["mg_fac"])
        assert c.neurons.t / units.ms == pytest.approx(400.0)
        c.engine.close()
    (gi, gt, gV, gI, gm), (ci, ct, cV, cI, cm) = out["gpu"], out["cpu"]
    assert gm == cm == 1.0 and len(ci) > 50
    assert np.array_equal(gi, ci) and np.array_equal(gt, ct)                 # spike lists bit-identical
    assert np.array_equal(gV, cV)                                            # recorded membrane potentials too
    assert np.allclose(gI, cI, rtol=1e-12, atol=0)                           # summed current: reduction order differs
