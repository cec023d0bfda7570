"""Look-alikes of the reference's `codes_revision/` modules (NetworkSetup.Network / network_setup, CortexSetup.Cortex)
on the B200 engine.  numpy-backed: the reference's xarray containers are replaced by small labelled holders that answer
the same queries (`.loc[dict(param=...)]`, `.values`, `.coords`)."""
