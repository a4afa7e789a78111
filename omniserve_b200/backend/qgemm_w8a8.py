"""omniserve_backend.qgemm_w8a8 -- out of scope for the W4A8KV4 hot path ("stub OK in v1")."""


def w8a8_gemm_forward_cuda(in_feats, kernel, wscales, ascales, out_feats):
    raise NotImplementedError("qgemm_w8a8 is outside the W4A8KV4 north-star path (SURVEY.md section 2a)")
