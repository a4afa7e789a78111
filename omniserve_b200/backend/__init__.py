"""Host-side mirror of the reference's `omniserve_backend` extension package.

Same module names, function names, positional argument order and in-place / return conventions as the
13 pybind11 modules registered by /root/reference/kernels/setup.py:156-333, implemented by calling the C
ABI of libomniserve_b200.so through ctypes.  `import omniserve_backend.<mod>` resolves here via the shim
package `omniserve_backend/` at the repository root, so omniserve/modeling and omniserve/engine import
and call it unchanged.
"""
