"""Fused multi-tensor AdamW on the GPU library — the optimiser the reference's train scripts build
(`autoregressive/train/train_c2i.py:28-50`: `torch.optim.AdamW(optim_groups, lr, betas, fused=True)`, 2-D tensors decayed, the rest
not).  Same constructor / `param_groups` / `state` layout as torch's (`exp_avg`, `exp_avg_sq`, `step`), one kernel launch per step for
all parameters, fp32 state, ATen's fused arithmetic.  CUDA fp32 parameters only; no CPU path."""
from __future__ import annotations

import struct

import torch

from . import _lib
from ._lib import check, cur_stream, _ptr

CHUNK = 65536


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **unused):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi, group, items):
        """Device tables of one param group: rebuilt when the (param, grad, state) pointers change (grads are re-allocated by
        zero_grad(set_to_none=True))."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in items)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2], cached[3]
        blob, chunks = bytearray(), []
        for t, p in enumerate(items):
            st = self.state[p]
            blob += struct.pack("<QQQQqfi", p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                                float(group["weight_decay"]), 0)
            chunks += [(t, c) for c in range((p.numel() + CHUNK - 1) // CHUNK)]
        dev = items[0].device
        tab = torch.frombuffer(blob, dtype=torch.uint8).to(dev)
        ck = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self._tables[gi] = (key, tab, ck, len(chunks))
        return tab, ck, len(chunks)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            items = [p for p in group["params"] if p.grad is not None]
            if not items:
                continue
            for p in items:
                if p.device.type != "cuda" or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("controlar_b200.optim.AdamW: contiguous fp32 CUDA parameters and gradients only (the train scripts keep fp32 masters)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            step = int(self.state[items[0]]["step"].item())
            if any(int(self.state[p]["step"].item()) != step for p in items):
                raise RuntimeError("controlar_b200.optim.AdamW: parameters of one group must share the step count")
            tab, ck, n = self._table(gi, group, items)
            b1, b2 = group["betas"]
            with torch.cuda.device(items[0].device):
                check(lib.car_adamw_step(_ptr(tab), _ptr(ck), n, float(group["lr"]), float(b1), float(b2), float(group["eps"]), step, cur_stream()),
                      "car_adamw_step")
        return loss
