"""Shim so that `import omniserve_backend.<module>` (what omniserve/modeling does at import time,
e.g. w4a8_linear.py:12-13) resolves to the B200-native implementation in omniserve_b200.backend."""
import importlib
import sys

_MODULES = (
    "qgemm_w4a8_per_chn", "qgemm_w4a8_per_group", "qgemm_w8a8", "fused_kernels", "layernorm_ops", "activation_ops",
    "fused_attention_pure_dense", "fused_attention_fine_grained_dense", "fused_attention_fine_grained_sparse",
    "fused_attention_per_tensor_dense", "fused_attention_per_tensor_sparse", "fused_attention_selector",
    "fused_attention_ctx_pool",
)
for _m in _MODULES:
    _mod = importlib.import_module(f"omniserve_b200.backend.{_m}")
    sys.modules[f"{__name__}.{_m}"] = _mod
    globals()[_m] = _mod
