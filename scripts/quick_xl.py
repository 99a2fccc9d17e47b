"""Quick XL decode-path timing (dev tool, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from controlar_b200.autoregressive.models.gpt_t2i import GPT_models
from controlar_b200.autoregressive.models.generate import generate
from controlar_b200 import _lib

torch.manual_seed(0)
B = int(os.environ.get("B", 8)); N = int(os.environ.get("N", 1024)); cfg = float(os.environ.get("CFG", 4.0))
SIDE = int(os.environ.get("SIDE", 32))          # RoPE table side: 32 (512x512), 48 (--image-size 768, config 4)
m = GPT_models["GPT-XL"](block_size=SIDE * SIDE, cls_token_num=120, model_type="t2i").eval()
m.output.weight.data.normal_(0, 0.02)
m = m.to("cuda", torch.bfloat16)
m.adapter.forward = lambda x: x
m.adapter_mlp.forward = lambda x: x
cond = torch.randn(B, 120, 2048, device="cuda", dtype=torch.bfloat16)
masks = torch.ones(B, 120, dtype=torch.int64, device="cuda")
ctrl = torch.randn(B, N, 1280, device="cuda", dtype=torch.bfloat16) * 0.1
for it in range(int(os.environ.get("ITERS", 3))):
    torch.cuda.synchronize(); t0 = time.time()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    out = generate(m, cond, N, emb_masks=masks, cfg_scale=cfg, condition=ctrl, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, seed=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = m._car_state
    tot = sum(st.step_bytes(120 + i + 1) for i in range(1, N))
    print(f"iter {it}: generate {ms:.1f} ms  ({ms/N:.3f} ms/token)  decode bytes {tot/1e12:.3f} TB -> {tot/ms/1e6:.0f} GB/s if all decode; wall {time.time()-t0:.2f}s", flush=True)
# prefill alone
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
b_eff = 2*B if cfg > 1 else B
cc = torch.cat([cond, cond]) if cfg > 1 else cond
cic = torch.cat([ctrl, torch.zeros_like(ctrl)]) if cfg > 1 else ctrl
e0.record(); st.prefill(cc, cic, 1.0, all_rows=False); e1.record(); torch.cuda.synchronize()
print("prefill ms", e0.elapsed_time(e1))
print("tokens sample", out[0, :8].tolist())
