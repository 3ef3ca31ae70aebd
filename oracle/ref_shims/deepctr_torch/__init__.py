"""Shim of the deepctr-torch 0.2.9 *names* the reference imports (see ../README.md)."""
