"""GPU: the batched generation engine (controlar_b200/autoregressive/serve/llm.py) against direct `generate()`
calls of this package: the reference's serve/sample_c2i.py calling convention (class prompts + unconditional half, CFG combined in
the sampler), queue admission in batches of <= 8 images, and the control path (t2i + control maps + VQ decode) the reference's
serve lacks."""
import pytest
import torch

from oracle.weights import GPTSpec, make_vq_state_dict
from oracle.inputs import control_map, text_inputs
from tests.helpers import build_product_gpt

pytestmark = pytest.mark.gpu
SMALL = dict(dim=256, n_layer=6, n_head=4, vocab_size=2048)


def test_c2i_reference_convention_equals_direct_generate():
    from controlar_b200.autoregressive.models.generate import generate
    from controlar_b200.autoregressive.serve.llm import LLM, SamplingParams
    spec = GPTSpec(**SMALL, cls_token_num=1, block_size=64, model_type="c2i")
    model, _ = build_product_gpt(spec, 3, torch.bfloat16)
    labels = [207, 360, 387, 974, 88, 979, 417, 279, 1, 2, 3]
    llm = LLM(model=model, cfg_scale=4.0, max_images_per_batch=8, seed=11)
    ids = [[c] for c in labels] + [[spec.num_classes] for _ in labels]
    outs = llm.generate(prompt_token_ids=ids, sampling_params=SamplingParams(temperature=0.0, top_k=-1, top_p=1.0, max_tokens=24), use_tqdm=False)
    assert len(outs) == 2 * len(labels)
    got = torch.tensor([o.outputs[0].token_ids for o in outs])
    assert torch.equal(got[: len(labels)], got[len(labels):])
    want = torch.cat([generate(model, torch.tensor(labels[a:b], device="cuda"), 24, cfg_scale=4.0, temperature=1.0, top_k=0, top_p=1.0,
                               sample_logits=False).cpu() for a, b in ((0, 8), (8, 11))])
    assert torch.equal(got[: len(labels)], want.to(got.dtype))
    # sampled: reproducible for a fixed engine seed, and different images get different grids
    sp = SamplingParams(temperature=1.0, top_k=200, max_tokens=24)
    a = LLM(model=model, cfg_scale=4.0, seed=5).generate(prompt_token_ids=ids, sampling_params=sp)
    b = LLM(model=model, cfg_scale=4.0, seed=5).generate(prompt_token_ids=ids, sampling_params=sp)
    ta, tb = torch.tensor([o.outputs[0].token_ids for o in a]), torch.tensor([o.outputs[0].token_ids for o in b])
    assert torch.equal(ta, tb) and not torch.equal(ta[0], ta[1])


def test_t2i_control_requests_through_the_queue():
    from controlar_b200.autoregressive.models.generate import generate
    from controlar_b200.autoregressive.serve.llm import LLM, SamplingParams
    from controlar_b200.tokenizer.tokenizer_image.vq_model import VQ_models
    spec = GPTSpec(**SMALL, cls_token_num=120, block_size=64, model_type="t2i")
    model, _ = build_product_gpt(spec, 0, torch.bfloat16)
    vq = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8)
    vq.load_state_dict(make_vq_state_dict(seed=3))
    vq = vq.cuda().eval()
    B = 5
    cond, masks = text_inputs(120, spec.caption_dim, B, 7, torch.bfloat16)
    cmap = control_map(B, 128, 128, 9, "canny", torch.bfloat16)
    llm = LLM(model=model, vq=vq, cfg_scale=4.0, max_images_per_batch=4, seed=2)
    sp = SamplingParams(temperature=0.0, max_tokens=64)
    prompts = [dict(cond=cond[i], emb_mask=masks[i], control=cmap[i], control_strength=0.6) for i in range(B)]
    outs = llm.generate(prompts=prompts, sampling_params=sp)
    assert len(outs) == B and all(o.image is not None and tuple(o.image.shape) == (3, 128, 128) for o in outs)
    got = torch.tensor([o.outputs[0].token_ids for o in outs])
    assert int(got.min()) >= 0 and int(got.max()) < spec.vocab_size
    want = torch.cat([generate(model, cond[a:b].cuda(), 64, emb_masks=masks[a:b].cuda(), cfg_scale=4.0, condition=cmap[a:b].cuda(), control_strength=0.6,
                               temperature=1.0, top_k=0, top_p=1.0, sample_logits=False).cpu() for a, b in ((0, 4), (4, 5))])
    assert torch.equal(got, want.to(got.dtype))
