"""Row b on the GPU: the reference's sampling call sequence (autoregressive/sample/sample_t2i.py:15-19 imports, :56-62 model
construction + `.to(device, dtype=precision)`, :146-160 left-padded prompt embeddings, :163-176 `generate(...)` then
`vq_model.decode_code(index_sample, qzshape)`) executed through the drop-in module names in a fresh interpreter with the
PYTHONPATH INTEGRATION.md documents.  bf16 (the reference's default precision, sample_t2i.py:197); `--precision fp16` is refused
loudly (include/controlar_b200.h: the kernels' tensor-core and packet formats are bf16)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys, torch
from tokenizer.tokenizer_image.vq_model import VQ_models
from autoregressive.models.gpt_t2i import GPT_models
from autoregressive.models.generate import generate
precision = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[1]]
device = "cuda"
torch.manual_seed(0)
H = W = 128
vq_model = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).to(device).eval()
latent = H // 16
gpt_model = GPT_models["GPT-B"](block_size=latent ** 2, cls_token_num=120, model_type="t2i", condition_type="canny",
                                adapter_size="small").to(device=device, dtype=precision).eval()
gpt_model.output.weight.data.normal_(0, 0.02)
condition_img = (2 * ((torch.rand(2, 1, H, W, device=device) < 0.1).float() - 0.5)).repeat(1, 3, 1, 1)
caption_embs = torch.randn(2, 120, 2048, device=device, dtype=precision)
emb_masks = torch.zeros(2, 120, dtype=torch.int64, device=device)
emb_masks[0, -17:] = 1; emb_masks[1, -40:] = 1                       # left padding (sample_t2i.py:146-156)
c_indices = caption_embs * emb_masks[:, :, None]
qzshape = [len(c_indices), 8, latent, latent]
index_sample = generate(gpt_model, c_indices, latent * latent, emb_masks, condition=condition_img.to(precision), cfg_scale=4.0,
                        temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, control_strength=1.0)
assert index_sample.dtype == torch.int32 and tuple(index_sample.shape) == (2, latent * latent)
assert int(index_sample.min()) >= 0 and int(index_sample.max()) < 16384
samples = vq_model.decode_code(index_sample, qzshape)
assert tuple(samples.shape) == (2, 3, H, W) and samples.dtype == torch.float32 and bool(torch.isfinite(samples).all())
import controlar_b200._lib as L
assert L._lib is not None, "the CUDA library was not loaded"
print("OK", sys.argv[1])
"""


def _run(precision, tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT])
    return subprocess.run([sys.executable, "-c", CODE, precision], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)


def test_reference_sampling_sequence_through_dropin_names(tmp_path):
    r = _run("bf16", tmp_path)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK bf16"), r.stdout[-2000:] + r.stderr[-4000:]


def test_fp16_checkpoints_are_refused_loudly(tmp_path):
    r = _run("fp16", tmp_path)
    assert r.returncode != 0 and "bf16 and fp32" in (r.stdout + r.stderr), r.stdout[-1000:] + r.stderr[-2000:]
