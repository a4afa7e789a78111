"""omniserve_backend.fused_attention_ctx_pool (reference: sparse_utils/ContextPool/context_pool_kernel.cu:145-213)."""


def paged_min_max_pool(*a, **k):
    raise NotImplementedError("LServe min/max pool: SURVEY.md section 8 row a11, scheduled after the dense path")
