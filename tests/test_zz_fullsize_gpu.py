"""GPU, full BASELINE sizes (config 2: GPT-XL, B = 8, CFG 4 -> 16 sequences, 1024 tokens; config 4: 768 x 512, B = 4, 1536 tokens):
size-independent properties of the conditional-decoding loop — the oracle needs ~35 min per batch at this size
(SURVEY.md §8d), so parity here is through properties that must hold exactly:
  * run-to-run determinism (fixed reduction orders, counter-based sampler noise),
  * image independence: an image's token grid does not depend on the other images of the batch (rows of the M = 16 GEMMs,
    (b, h) pairs of the attention split and the per-image sampler never mix),
  * every token is a valid code."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(side):
    from controlar_b200.autoregressive.models.gpt_t2i import GPT_models
    torch.manual_seed(0)
    m = GPT_models["GPT-XL"](block_size=side * side, cls_token_num=120, model_type="t2i").eval()
    m.output.weight.data.normal_(0, 0.02)
    m = m.to("cuda", torch.bfloat16)
    m.adapter.forward = lambda x: x            # the control encoder has its own tests; feed adapter_mlp output directly
    m.adapter_mlp.forward = lambda x: x
    return m


@pytest.mark.parametrize("side,B,N", [(32, 8, 1024), (48, 4, 1536)])
def test_fullsize_determinism_and_image_independence(side, B, N):
    from controlar_b200.autoregressive.models.generate import generate
    m = _model(side)
    g = torch.Generator(device="cuda").manual_seed(1)
    cond = torch.randn(B, 120, 2048, device="cuda", generator=g).to(torch.bfloat16)
    masks = torch.ones(B, 120, dtype=torch.int64, device="cuda")
    masks[:, :17] = 0                          # left padding like sample_t2i.py:146-160
    cond[:, :17] = 0
    ctrl = (torch.randn(B, N, 1280, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    kw = dict(emb_masks=masks, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, seed=7)
    a = generate(m, cond, N, condition=ctrl, **kw)
    b = generate(m, cond, N, condition=ctrl, **kw)
    assert a.shape == (B, N) and int(a.min()) >= 0 and int(a.max()) < 16384
    assert torch.equal(a, b), "same inputs, same seed: token grids must be identical"
    cond2, ctrl2 = cond.clone(), ctrl.clone()
    j = B // 2
    cond2[j, 17:] = torch.randn(103, 2048, device="cuda", generator=g).to(torch.bfloat16)
    ctrl2[j] = (torch.randn(N, 1280, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    c = generate(m, cond2, N, condition=ctrl2, **kw)
    keep = [i for i in range(B) if i != j]
    assert torch.equal(a[keep], c[keep]), "changing one image's prompt / control must not touch the other images"
    assert not torch.equal(a[j], c[j])
