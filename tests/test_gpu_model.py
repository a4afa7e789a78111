"""-m gpu: the host-side model runner (call order of llama_w4a8_unpad.py) over the C ABI ops."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(ops=None, fuse=True, group=-1):
    from omniserve_b200.model import LlamaConfig, LlamaW4A8
    cfg = LlamaConfig.tiny(group_size=group)
    m = LlamaW4A8(cfg, "cuda", ops=ops, fuse_silu_quant=fuse)
    m.alloc(batch=3, max_ctx=192, max_tokens=3 * 70)
    return cfg, m


def _prompts(cfg):
    g = torch.Generator().manual_seed(1)
    return torch.randint(0, cfg.vocab_size, (70 + 33 + 64,), generator=g).cuda(), [70, 33, 64]


@pytest.mark.parametrize("group", [-1, 128])
def test_graph_replay_equals_eager(group):
    from omniserve_b200.model import DecodeGraph
    cfg, m1 = _mk(group=group)
    toks, lens = _prompts(cfg)
    f1 = m1.prefill(toks, lens)
    m1.prepare_decode()
    eager = []
    t = f1.clone()
    for _ in range(4):
        t = m1.decode_step(t, 192)
        eager.append(t.clone())
    cfg, m2 = _mk(group=group)
    f2 = m2.prefill(toks, lens)
    assert torch.equal(f1, f2)
    g = DecodeGraph(m2, 192)
    g.tokens.copy_(f2)
    for i in range(4):
        g.step()
        torch.cuda.synchronize()
        assert torch.equal(g.out, eager[i]), f"step {i}"
    assert m2.context_lens.tolist() == [74, 37, 68]
    for a, b in zip(m1.kv.k_pools, m2.kv.k_pools):
        assert torch.equal(a, b)


def test_fused_silu_quant_equals_two_kernel_chain():
    cfg, m1 = _mk(fuse=True)
    cfg, m2 = _mk(fuse=False)
    toks, lens = _prompts(cfg)
    assert torch.equal(m1.prefill(toks, lens), m2.prefill(toks, lens))
    assert torch.equal(m1.last_hidden, m2.last_hidden)


def test_whole_stack_vs_reference_kernels():
    """Same weights, same prompts through the reference's rebuilt kernels (oracle/_ref) and ours."""
    from tests.gpu_util import ref_module
    from omniserve_b200.model import Ops
    if ref_module("qgemm_w4a8_per_chn") is None:
        pytest.skip("oracle/_ref not shipped")
    cfg, ours = _mk(fuse=False)
    cfg, ref = _mk(ops=Ops(ref_module), fuse=False)
    toks, lens = _prompts(cfg)
    a, b = ours.prefill(toks, lens), ref.prefill(toks, lens)
    ha, hb = ours.last_hidden.float(), ref.last_hidden.float()
    assert (ha - hb).abs().max() <= 2e-2 * hb.abs().max()      # two layers of fp16 / int8-rounding drift
    # V pages byte-identical would need identical hidden states; compare the first layer's pages instead
    assert (ours.kv.v_pools[0] != ref.kv.v_pools[0]).float().mean() < 2e-3
    ours.prepare_decode(); ref.prepare_decode()
    t = a.clone()
    o = ours.decode_step(t, 192)
    ho = (ours.last_decode_state[0].float() + ours.last_decode_state[1].float()).clone()
    r = ref.decode_step(t, 192)
    hr = ref.last_decode_state[0].float() + ref.last_decode_state[1].float()
    assert (ho - hr).abs().max() <= 3e-2 * hr.abs().max()
