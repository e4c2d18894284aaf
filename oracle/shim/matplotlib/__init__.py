"""Empty stub (oracle shim): the reference's dataset modules import matplotlib at module scope."""
