"""Empty stub (oracle shim)."""
