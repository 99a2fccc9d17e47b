"""dev: where does the persistent sampler disagree on the small golden?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_ar_gpu import _setup
from controlar_b200 import engine
from oracle.ar_oracle import cfg_combine
name = sys.argv[1] if len(sys.argv) > 1 else "t2i_small_bf16"
topk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g, spec, dt, model, sd, cond, masks = _setup(name)
dev = "cuda"
B, N, T = g["B"], g["greedy_tokens"].shape[1], spec.cls_token_num
use_cfg = g["cfg_scale"] > 1.0
b_eff = 2 * B if use_cfg else B
ctrl_in = g["ctrl_in"].to(dev); c = cond.to(dev)
cc = torch.cat([c, torch.zeros_like(c) + model.cls_embedding.uncond_embedding]) if use_cfg else c
cond_comb = torch.cat([ctrl_in, torch.zeros_like(ctrl_in)]) if use_cfg else ctrl_in
model.setup_caches(b_eff, T + N, dt, n_img_tokens=N)
st = model._car_state
st.set_emb_mask(None if masks is None else (torch.cat([masks, masks]).to(dev) if use_cfg else masks.to(dev)))
st.prefill(cc, cond_comb, g["control_strength"] if use_cfg else 1.0, all_rows=False)
sp = engine.make_sampling(temperature=1.0, top_k=topk, top_p=1.0, sample_logits=False, cfg_scale=g["cfg_scale"])
choice, trace = st.generate_forced(sp, g["greedy_tokens"].to(dev))
torch.cuda.synchronize()
got = trace.permute(1, 0, 2).float()
nf = (~torch.isfinite(got)).nonzero()
print("non-finite logits:", nf.shape[0], "first", nf[:5].tolist(), "steps", sorted(set(nf[:, 1].tolist()))[:10])
z = cfg_combine(got, g["cfg_scale"]) if use_cfg else got
am = z.argmax(-1).cpu()
ch = choice.cpu().long()
bad = (ch != am).nonzero().tolist()
print("choice != argmax(trace) at", bad[:20], "of", ch.numel())
for b, i in bad[:8]:
    zz = z[b, i]
    print(b, i, "choice", int(ch[b, i]), "argmax", int(am[b, i]), "max", float(zz.max()), "isfinite", bool(torch.isfinite(zz).all()))
