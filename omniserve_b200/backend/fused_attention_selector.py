"""omniserve_backend.fused_attention_selector (reference: sparse_utils/KVPageSelector/fused_kv_page_selector.cpp:171-334)."""


def single_query_page_selector(*a, **k):
    raise NotImplementedError("LServe page selector: SURVEY.md section 8 row a9, scheduled after the dense path")
