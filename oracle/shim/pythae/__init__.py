"""TEST INFRASTRUCTURE ONLY (oracle shim).

Container-only stand-in for the third-party `pythae` package (unpinned in the reference's
setup.py:29, not installable here: no network).  It provides the *containers and base classes*
the reference imports on the training-step path and NO arithmetic.  It exists so that
/root/reference/src can be imported in the build container to validate the oracle port and to
generate the golden vectors under tests/golden/.  Nothing in multivae_b200/ imports it.
"""
