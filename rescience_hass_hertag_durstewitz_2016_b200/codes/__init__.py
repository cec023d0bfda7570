"""Modules with the names and call signatures of the reference's `codes/` directory (CortexNetwork, NetworkSetup,
ParamSetup), so that `sys.path.insert(0, <this directory>)` lets `codes/Main*.py` run against the B200 engine."""
