"""GPU: fused multi-tensor AdamW (controlar_b200/optim.py, row f1 — the optimiser of reference train_c2i.py:28-50) against
torch.optim.AdamW(fused=True) on the same parameters / gradients, decayed and non-decayed groups, several steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_adamw_matches_torch_fused():
    from controlar_b200.optim import AdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(300, 129), (65536 * 2 + 7,), (5,), (64, 64, 3, 3), (1,)]
    base = [torch.randn(*s, generator=g) for s in shapes]

    def make():
        ps = [torch.nn.Parameter(b.clone().cuda()) for b in base]
        groups = [{"params": [p for p in ps if p.dim() >= 2], "weight_decay": 0.05}, {"params": [p for p in ps if p.dim() < 2], "weight_decay": 0.0}]
        return ps, groups
    pa, ga = make()
    pb, gb = make()
    ours = AdamW(ga, lr=1e-3, betas=(0.9, 0.95))
    ref = torch.optim.AdamW(gb, lr=1e-3, betas=(0.9, 0.95), fused=True)
    for it in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.1 + it)
            a.grad = gr.clone(); b.grad = gr.clone()
        ours.step(); ref.step()
        ours.zero_grad(set_to_none=True); ref.zero_grad(set_to_none=True)
    for a, b in zip(pa, pb):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel < 2e-6, rel
        sa, sb = ours.state[a], ref.state[b]
        assert float((sa["exp_avg"] - sb["exp_avg"]).abs().max()) <= 1e-6 * float(sb["exp_avg"].abs().max())
        assert float((sa["exp_avg_sq"] - sb["exp_avg_sq"]).abs().max()) <= 1e-6 * float(sb["exp_avg_sq"].abs().max())
