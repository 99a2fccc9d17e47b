"""dev: free-running small generate in a loop — catch the intermittent bad token"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_ar_gpu import _setup
from controlar_b200.autoregressive.models.generate import generate
name = sys.argv[1] if len(sys.argv) > 1 else "t2i_small_bf16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
# some chain-path work first, like the test session does
g0, spec0, dt0, model0, sd0, cond0, masks0 = _setup("t2i_small_fp32")
model0.adapter.forward = lambda x: x; model0.adapter_mlp.forward = lambda x: x
N0 = g0["greedy_tokens"].shape[1]
o = generate(model0, cond0.to("cuda"), N0, emb_masks=masks0.to("cuda"), cfg_scale=g0["cfg_scale"], condition=g0["ctrl_in"].to("cuda"),
             control_strength=g0["control_strength"], temperature=1.0, top_k=0, top_p=1.0, sample_logits=False)
print("fp32 chain ok", bool(torch.equal(o.cpu(), g0["greedy_tokens"])), flush=True)
del model0
first = None
for it in range(iters):
    g, spec, dt, model, sd, cond, masks = _setup(name)       # new model + state every iteration, like the tests
    model.adapter.forward = lambda x: x; model.adapter_mlp.forward = lambda x: x
    N = g["greedy_tokens"].shape[1]
    out = generate(model, cond.to("cuda"), N, emb_masks=None if masks is None else masks.to("cuda"), cfg_scale=g["cfg_scale"],
                   condition=g["ctrl_in"].to("cuda"), control_strength=g["control_strength"], temperature=1.0, top_k=0, top_p=1.0,
                   sample_logits=False)
    torch.cuda.synchronize()
    out = out.cpu()
    bad = ((out < 0) | (out >= spec.vocab_size)).nonzero().tolist()
    if first is None:
        first = out
    print(f"iter {it}: bad tokens {bad[:6]} same-as-first {bool(torch.equal(out, first))} match-ref {float((out == g['greedy_tokens']).float().mean()):.3f}", flush=True)
    del model
