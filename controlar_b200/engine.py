"""Host-side handles over the C ABI (include/controlar_b200.h): packed model + per-generate() state.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libcontrolar_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
import functools

from ._lib import CarModelDesc, CarSampling, CarWeights, check, cur_stream, dtype_code, _ptr, _ptr_array


def _on_own_device(fn):
    """Run a handle method with the handle's device current: the library launches on the current device and `cur_stream()` is that
    device's current stream — a model on cuda:1 works without torch.cuda.set_device(1) (the reference wraps its calls
    in `with torch.device(device)`, generate.py:179-182)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = getattr(self, "device", None)
        if dev is None or dev.type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped


class ARModelHandle(_lib.NativeHandle):
    """CarModel: GEMM-ready copies of the transformer weights (re-packed when the module's weights change)."""

    def __init__(self, module):
        self.lib = _lib.lib()
        self.module = module
        self.handle = C.c_void_p()
        self.version = None
        self._keep = None
        self._dirty = 0
        self.device = module.tok_embeddings.weight.device
        self.generation = 0          # bumped whenever the CarModel handle is re-created: states compare THIS, not the raw pointer
        self._build()

    # the tensors whose storage the library borrows
    def _weights(self):
        m = self.module
        layers = list(m.layers)
        w = CarWeights()
        keep = []

        def P(t):
            keep.append(t)
            return _ptr(t.detach())
        w.tok_embeddings = P(m.tok_embeddings.weight)
        w.norm = P(m.norm.weight)
        w.output = P(m.output.weight)
        arrs = {}
        for name, get in [("attention_norm", lambda b: b.attention_norm.weight), ("wqkv", lambda b: b.attention.wqkv.weight),
                          ("wo", lambda b: b.attention.wo.weight), ("ffn_norm", lambda b: b.ffn_norm.weight),
                          ("w1", lambda b: b.feed_forward.w1.weight), ("w3", lambda b: b.feed_forward.w3.weight),
                          ("w2", lambda b: b.feed_forward.w2.weight)]:
            ts = [get(b).detach() for b in layers]
            keep.extend(ts)
            arrs[name] = _ptr_array(ts)
            setattr(w, name, C.cast(arrs[name], C.POINTER(C.c_void_p)))
        if m.model_type == "t2i":
            w.cap_fc1 = P(m.cls_embedding.cap_proj.fc1.weight)
            w.cap_fc2 = P(m.cls_embedding.cap_proj.fc2.weight)
        else:
            w.label_table = P(m.cls_embedding.embedding_table.weight)
        w.cond_fc1 = P(m.condition_mlp.cap_proj.fc1.weight)
        w.cond_fc2 = P(m.condition_mlp.cap_proj.fc2.weight)
        for j in range(3):
            w.ctl_fc1[j] = P(m.condition_layers[j].fc1.weight)
            w.ctl_fc2[j] = P(m.condition_layers[j].fc2.weight)
        keep.append(arrs)
        return w, keep

    def _signature(self):
        """Changes when parameters are replaced (data_ptr / dtype / device) or updated through autograd-visible in-place ops
        (`_version`).  Writes through `p.data` do NOT bump `_version`: call `invalidate()` after such updates."""
        m = self.module
        ps = [m.tok_embeddings.weight, m.output.weight, m.layers[0].attention.wqkv.weight]
        return tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in ps) + \
            (sum(p._version for p in m.parameters()), self._dirty)

    def _layout(self):
        """What the packed buffers were sized for: a change here needs a new CarModel, anything else a repack in place."""
        m = self.module
        w = m.tok_embeddings.weight
        return (w.dtype, str(w.device), tuple(w.shape), len(m.layers), tuple(m.layers[0].feed_forward.w1.weight.shape))

    def invalidate(self):
        """Force a repack at the next use (after `param.data` writes, which PyTorch's version counters do not see)."""
        self._dirty += 1

    @_on_own_device
    def _build(self):
        m = self.module
        self.device = m.tok_embeddings.weight.device
        cfg = m.config
        dt = m.tok_embeddings.weight.dtype
        d = CarModelDesc(dtype=dtype_code(dt), dim=cfg.dim, n_layer=cfg.n_layer, n_head=cfg.n_head,
                         ffn_dim=m.layers[0].feed_forward.w1.weight.shape[0], vocab_size=cfg.vocab_size,
                         cls_token_num=cfg.cls_token_num, block_size=cfg.block_size,
                         caption_dim=cfg.caption_dim if m.model_type == "t2i" else 0,
                         model_type=1 if m.model_type == "t2i" else 0, norm_eps=cfg.norm_eps, rope_base=cfg.rope_base)
        w, keep = self._weights()
        if self.handle:
            check(self.lib.car_model_destroy(self.handle), "car_model_destroy")
            self.handle = C.c_void_p()
        check(self.lib.car_model_create(C.byref(d), C.byref(w), cur_stream(), C.byref(self.handle)), "car_model_create")
        self._keep = keep
        self.version = self._signature()
        self.layout = self._layout()
        self.dtype = dt
        self.desc = d
        self.generation += 1

    @_on_own_device
    def refresh(self):
        """Bring the packed copies up to date if parameters were replaced / updated since the last pack.  Same layout:
        `car_model_repack` rewrites the library-owned buffers IN PLACE (the CarModel and every device pointer a live CarState holds
        stay valid — ADVICE r1: re-creating the model under a live state was a use-after-free).  Different dtype / device / shapes:
        a new CarModel (generation bumps; `setup_caches` then rebuilds the state)."""
        if self._signature() == self.version:
            return
        if self._layout() != self.layout:
            self._build()
            return
        w, keep = self._weights()
        check(self.lib.car_model_repack(self.handle, C.byref(w), cur_stream()), "car_model_repack")
        self._keep = keep
        self.version = self._signature()

    def close(self):
        if self.handle:
            self.lib.car_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ARTrainHandle(_lib.NativeHandle):
    """CarTrain: the teacher-forced training forward (reference gpt_t2i.py:420-431,451-484) on the module's fp32 parameters
    under bf16-autocast numerics.  Weights are borrowed (re-cast to bf16 inside every forward, like autocast does)."""

    def __init__(self, module, max_batch: int, max_img_tokens: int):
        from ._lib import CarTrainWeights
        self.lib = _lib.lib()
        m = module
        cfg = m.config
        if m.tok_embeddings.weight.dtype != torch.float32:
            raise NotImplementedError("controlar_b200 training forward: fp32 parameters (bf16 autocast is applied inside), as the train scripts keep them")
        w, keep = ARModelHandle._weights(self_like(m))
        tw = CarTrainWeights()
        tw.w = w
        tw.adapter_fc1 = _ptr(m.adapter_mlp.fc1.weight.detach()); tw.adapter_fc2 = _ptr(m.adapter_mlp.fc2.weight.detach())
        tw.adapter_dim = m.adapter_mlp.fc1.weight.shape[1]
        tw.num_classes = cfg.num_classes
        if m.model_type == "t2i":
            unc = m.cls_embedding.uncond_embedding.detach().to(torch.float32).contiguous()
            keep.append(unc)
            tw.cap_uncond = _ptr(unc)
        if not getattr(m, "zero_uncond_on_drop", False):
            cunc = m.condition_mlp.uncond_embedding.detach().to(torch.float32).contiguous()   # buffer, zeros unless a state dict says otherwise
            keep.append(cunc)
            tw.cond_uncond = _ptr(cunc)
        # else: the legacy gpt.py class gives dropped samples literal zeros (gpt.py:118-119): NULL = zeros in the library
        d = CarModelDesc(dtype=_lib.CAR_F32, dim=cfg.dim, n_layer=cfg.n_layer, n_head=cfg.n_head,
                         ffn_dim=m.layers[0].feed_forward.w1.weight.shape[0], vocab_size=cfg.vocab_size,
                         cls_token_num=cfg.cls_token_num, block_size=cfg.block_size,
                         caption_dim=cfg.caption_dim if m.model_type == "t2i" else 0,
                         model_type=1 if m.model_type == "t2i" else 0, norm_eps=cfg.norm_eps, rope_base=cfg.rope_base)
        dev = m.tok_embeddings.weight.device
        self.rope = m.freqs_cis.to(device=dev, dtype=torch.float32).contiguous()
        self.handle = C.c_void_p()
        check(self.lib.car_train_create(C.byref(d), C.byref(tw), max_batch, max_img_tokens, _ptr(self.rope), cur_stream(), C.byref(self.handle)),
              "car_train_create")
        self._keep = (keep, tw)
        self.max_batch, self.max_img_tokens = max_batch, max_img_tokens
        self.V, self.T = cfg.vocab_size, cfg.cls_token_num
        self.key = tuple(p.data_ptr() for p in m.parameters())

    def forward(self, idx, cond, feat, drop_ids, mask, targets, valid):
        B, n = idx.shape
        n_img = n + 1
        dev = idx.device
        idx = idx.to(torch.int32).contiguous()
        cond = cond.to(torch.int32).contiguous() if cond.dtype in (torch.int64, torch.int32) else cond.to(torch.float32).contiguous()
        feat = None if feat is None else feat.to(torch.bfloat16).contiguous()
        drop = drop_ids.to(torch.uint8).contiguous()
        S = self.T + n
        m8 = None
        if mask is not None:
            m8 = mask.reshape(B, S, S).to(torch.uint8).contiguous()
        tg = None if targets is None else targets.to(torch.int32).contiguous()
        vf = None if valid is None else valid.to(torch.float32).contiguous()
        logits = torch.empty(B, n_img, self.V, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32) if tg is not None else None
        check(self.lib.car_train_forward(self.handle, B, n_img, _ptr(idx), _ptr(cond), None if feat is None else _ptr(feat), _ptr(drop),
                                         None if m8 is None else _ptr(m8), None if tg is None else _ptr(tg),
                                         None if vf is None else _ptr(vf), _ptr(logits), None if loss is None else _ptr(loss), cur_stream()),
              "car_train_forward")
        self._last = (idx, cond, feat, drop, m8, tg, vf)      # car_train_backward reads them again
        self.generation = getattr(self, "generation", 0) + 1  # a backward belongs to the forward that produced its loss
        return logits, (None if loss is None else loss[0])

    # ---- backward ---------------------------------------------------------------------------------------------------
    @staticmethod
    def grad_params(m):
        """The parameters `car_train_backward` produces gradients for, in a fixed order (name, parameter)."""
        out = [("tok_embeddings.weight", m.tok_embeddings.weight), ("norm.weight", m.norm.weight), ("output.weight", m.output.weight)]
        for i, b in enumerate(m.layers):
            pre = f"layers.{i}."
            out += [(pre + "attention_norm.weight", b.attention_norm.weight), (pre + "attention.wqkv.weight", b.attention.wqkv.weight),
                    (pre + "attention.wo.weight", b.attention.wo.weight), (pre + "ffn_norm.weight", b.ffn_norm.weight),
                    (pre + "feed_forward.w1.weight", b.feed_forward.w1.weight), (pre + "feed_forward.w3.weight", b.feed_forward.w3.weight),
                    (pre + "feed_forward.w2.weight", b.feed_forward.w2.weight)]
        if m.model_type == "t2i":
            out += [("cls_embedding.cap_proj.fc1.weight", m.cls_embedding.cap_proj.fc1.weight),
                    ("cls_embedding.cap_proj.fc2.weight", m.cls_embedding.cap_proj.fc2.weight)]
        else:
            out += [("cls_embedding.embedding_table.weight", m.cls_embedding.embedding_table.weight)]
        out += [("condition_mlp.cap_proj.fc1.weight", m.condition_mlp.cap_proj.fc1.weight),
                ("condition_mlp.cap_proj.fc2.weight", m.condition_mlp.cap_proj.fc2.weight)]
        for j in range(3):
            out += [(f"condition_layers.{j}.fc1.weight", m.condition_layers[j].fc1.weight),
                    (f"condition_layers.{j}.fc2.weight", m.condition_layers[j].fc2.weight)]
        out += [("adapter_mlp.fc1.weight", m.adapter_mlp.fc1.weight), ("adapter_mlp.fc2.weight", m.adapter_mlp.fc2.weight)]
        return out

    def backward(self, module, loss_grad=None, want_feat_grad=True):
        """Gradients of the last forward(targets=...) on this handle -> ({name: fp32 grad}, d_feat bf16 or None).
        loss_grad: 0-dim / [1] fp32 CUDA tensor (d / d loss) or None = 1."""
        from ._lib import CarTrainWeights
        if getattr(self, "_last", None) is None:
            raise RuntimeError("controlar_b200: backward() needs a preceding training forward with targets")
        idx, cond, feat, drop, m8, tg, vf = self._last
        if tg is None:
            raise RuntimeError("controlar_b200: the last training forward had no targets / loss")
        names = self.grad_params(module)
        has_feat = feat is not None
        skip_wo_feat = ("condition_mlp.", "condition_layers.", "adapter_mlp.")
        G = {k: torch.empty_like(p, dtype=torch.float32) for k, p in names if has_feat or not k.startswith(skip_wo_feat)}
        L = len(module.layers)
        gw = CarTrainWeights()
        keep = []

        def P(key):
            t = G.get(key)
            return None if t is None else _ptr(t)
        gw.w.tok_embeddings = P("tok_embeddings.weight"); gw.w.norm = P("norm.weight"); gw.w.output = P("output.weight")
        for field, suffix in [("attention_norm", "attention_norm.weight"), ("wqkv", "attention.wqkv.weight"), ("wo", "attention.wo.weight"),
                              ("ffn_norm", "ffn_norm.weight"), ("w1", "feed_forward.w1.weight"), ("w3", "feed_forward.w3.weight"),
                              ("w2", "feed_forward.w2.weight")]:
            arr = _ptr_array([G[f"layers.{i}.{suffix}"] for i in range(L)])
            keep.append(arr)
            setattr(gw.w, field, C.cast(arr, C.POINTER(C.c_void_p)))
        if module.model_type == "t2i":
            gw.w.cap_fc1 = P("cls_embedding.cap_proj.fc1.weight"); gw.w.cap_fc2 = P("cls_embedding.cap_proj.fc2.weight")
        else:
            gw.w.label_table = P("cls_embedding.embedding_table.weight")
        gw.w.cond_fc1 = P("condition_mlp.cap_proj.fc1.weight"); gw.w.cond_fc2 = P("condition_mlp.cap_proj.fc2.weight")
        for j in range(3):
            gw.w.ctl_fc1[j] = P(f"condition_layers.{j}.fc1.weight"); gw.w.ctl_fc2[j] = P(f"condition_layers.{j}.fc2.weight")
        gw.adapter_fc1 = P("adapter_mlp.fc1.weight"); gw.adapter_fc2 = P("adapter_mlp.fc2.weight")
        dfeat = torch.empty_like(feat) if (has_feat and want_feat_grad) else None
        lg = None
        if loss_grad is not None:
            lg = loss_grad.detach().to(device=idx.device, dtype=torch.float32).reshape(1).contiguous()
        check(self.lib.car_train_backward(self.handle, C.byref(gw), None if dfeat is None else _ptr(dfeat), None if lg is None else _ptr(lg),
                                          cur_stream()), "car_train_backward")
        self._last = None
        return G, dfeat

    def close(self):
        if self.handle:
            self.lib.car_train_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ModuleView:
    def __init__(self, module):
        self.module = module


def self_like(module):
    """ARModelHandle._weights only reads ``self.module``."""
    return _ModuleView(module)


class ARStateHandle(_lib.NativeHandle):
    """CarState: KV caches (PyTorch-owned, reference layout), control tokens, scratch, the persistent decode kernel's packet
    buffers (and the CUDA graph of the per-kernel fallback chain)."""

    def __init__(self, model: ARModelHandle, b_eff: int, S: int, N: int, k_caches, v_caches, rope: torch.Tensor):
        self.lib = model.lib
        self.model = model
        self.device = rope.device
        self.b_eff, self.S, self.N = b_eff, S, N
        self._keep = (list(k_caches), list(v_caches), rope)
        self.handle = C.c_void_p()
        ka, va = _ptr_array(self._keep[0]), _ptr_array(self._keep[1])
        with torch.cuda.device(self.device):
            check(self.lib.car_state_create(model.handle, b_eff, S, N, C.cast(ka, C.POINTER(C.c_void_p)),
                                            C.cast(va, C.POINTER(C.c_void_p)), _ptr(rope), C.byref(self.handle)),
                  "car_state_create")
        self.V = model.desc.vocab_size
        self.T = model.desc.cls_token_num

    @_on_own_device
    def set_emb_mask(self, emb_mask: Optional[torch.Tensor]):
        if emb_mask is None:
            check(self.lib.car_state_set_emb_mask(self.handle, None, cur_stream()), "car_state_set_emb_mask")
            return
        em = (emb_mask != 0).to(torch.int32).contiguous()
        assert em.shape == (self.b_eff, self.T), (em.shape, self.b_eff, self.T)
        check(self.lib.car_state_set_emb_mask(self.handle, _ptr(em), cur_stream()), "car_state_set_emb_mask")
        self._mask_keep = em

    @_on_own_device
    def prefill(self, cond: torch.Tensor, condition: Optional[torch.Tensor], control_strength: float,
                all_rows: bool) -> torch.Tensor:
        dev = cond.device
        if cond.dtype in (torch.int64, torch.int32):
            cond = cond.to(torch.int32).contiguous()
        else:
            cond = cond.to(self.model.dtype).contiguous()
        if condition is not None:
            condition = condition.to(self.model.dtype).contiguous()
            assert condition.shape[0] == self.b_eff and condition.shape[1] == self.N, condition.shape
        shape = (self.b_eff, self.T, self.V) if all_rows else (self.b_eff, self.V)
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        check(self.lib.car_prefill(self.handle, _ptr(cond), _ptr(condition), float(control_strength), _ptr(out),
                                   1 if all_rows else 0, cur_stream()), "car_prefill")
        self._io_keep = (cond, condition)
        return out

    @_on_own_device
    def decode_step(self, tok: torch.Tensor, pos: int) -> torch.Tensor:
        tok = tok.reshape(-1).to(torch.int32).contiguous()
        assert tok.numel() == self.b_eff
        out = torch.empty((self.b_eff, self.V), dtype=torch.float32, device=tok.device)
        check(self.lib.car_decode_step(self.handle, _ptr(tok), int(pos), _ptr(out), cur_stream()), "car_decode_step")
        self._tok_keep = tok
        return out

    @_on_own_device
    def generate(self, sp: CarSampling, n_tokens: int, noise: Optional[torch.Tensor], device) -> torch.Tensor:
        B = self.b_eff // 2 if sp.cfg_scale > 1.0 else self.b_eff
        out = torch.empty((B, n_tokens), dtype=torch.int32, device=device)
        if noise is not None:
            noise = noise.to(torch.float32).contiguous()
            assert noise.shape == (n_tokens, B, self.V), noise.shape
        check(self.lib.car_generate(self.handle, C.byref(sp), int(n_tokens), _ptr(noise), _ptr(out), cur_stream()),
              "car_generate")
        self._noise_keep = noise
        return out

    @_on_own_device
    def generate_forced(self, sp: CarSampling, forced: torch.Tensor, trace: bool = True, noise: Optional[torch.Tensor] = None):
        """Teacher-forced device-side loop (car_generate_forced): returns (sampler choices int32 [B, n], logits fp32 [n, b_eff, V])."""
        B = self.b_eff // 2 if sp.cfg_scale > 1.0 else self.b_eff
        forced = forced.to(torch.int32).contiguous()
        assert forced.shape[0] == B, forced.shape
        n = forced.shape[1]
        out = torch.empty((B, n), dtype=torch.int32, device=forced.device)
        tr = torch.empty((n, self.b_eff, self.V), dtype=torch.float32, device=forced.device) if trace else None
        if noise is not None:
            noise = noise.to(torch.float32).contiguous()
            assert noise.shape == (n, B, self.V), noise.shape
        check(self.lib.car_generate_forced(self.handle, C.byref(sp), int(n), _ptr(noise), _ptr(forced), _ptr(tr), _ptr(out),
                                           cur_stream()), "car_generate_forced")
        self._noise_keep = (noise, forced)
        return out, tr

    def set_step_timer(self, buf: Optional[torch.Tensor]):
        """int64 [N] device tensor that receives the GPU globaltimer (ns) at the start of every decode iteration (None = off)."""
        if buf is not None:
            assert buf.dtype == torch.int64 and buf.numel() >= self.N and buf.is_cuda
        check(self.lib.car_state_set_step_timer(self.handle, _ptr(buf)), "car_state_set_step_timer")
        self._timer_keep = buf

    def step_bytes(self, n_context: int) -> int:
        return int(self.lib.car_decode_step_bytes(self.handle, int(n_context)))

    def close(self):
        if self.handle:
            self.lib.car_state_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_sampling(temperature=1.0, top_k=0, top_p=1.0, sample_logits=True, cfg_scale=1.0, cfg_interval=-1,
                  seed=0) -> CarSampling:
    return CarSampling(temperature=float(temperature), top_k=int(top_k or 0), top_p=float(top_p),
                       sample_logits=1 if sample_logits else 0, cfg_scale=float(cfg_scale),
                       cfg_interval=int(cfg_interval), seed=int(seed) & 0xFFFFFFFFFFFFFFFF)


def sample(logits: torch.Tensor, sp: CarSampling, cfg_on: bool = True, step: int = 0,
           noise: Optional[torch.Tensor] = None, return_probs: bool = False):
    """generate.sample() + CFG combine on [b_eff, V] fp32 logits (car_sample)."""
    lib = _lib.lib()
    logits = logits.to(torch.float32).contiguous()
    b_eff, V = logits.shape
    B = b_eff // 2 if sp.cfg_scale > 1.0 else b_eff
    idx = torch.empty((B,), dtype=torch.int32, device=logits.device)
    probs = torch.empty((B, V), dtype=torch.float32, device=logits.device) if return_probs else None
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
    with torch.cuda.device(logits.device):
        check(lib.car_sample(_ptr(logits), b_eff, V, C.byref(sp), 1 if cfg_on else 0, int(step), _ptr(noise), _ptr(idx),
                             _ptr(probs), cur_stream()), "car_sample")
    return (idx, probs) if return_probs else idx


def op_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """y = act(x @ w.T + bias) through the library's GEMM (car_op_linear)."""
    lib = _lib.lib()
    K = x.shape[-1]
    x2 = x.reshape(-1, K).contiguous()
    N = w.shape[0]
    y = torch.empty((x2.shape[0], N), dtype=x.dtype, device=x.device)
    check(lib.car_op_linear(dtype_code(x.dtype), _ptr(x2), _ptr(w.detach().contiguous()),
                            _ptr(bias.detach().contiguous()) if bias is not None else None, _ptr(y), x2.shape[0], N, K,
                            int(act), cur_stream()), "car_op_linear")
    return y.reshape(*x.shape[:-1], N)


def op_dense_linear(x: torch.Tensor, w: torch.Tensor, resid: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """y = act(x @ w.T) (+ resid) on the dense tcgen05 path (car_op_dense_linear); bf16."""
    lib = _lib.lib()
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.car_op_dense_linear(_ptr(x.contiguous()), _ptr(w.detach().contiguous()), _ptr(resid.contiguous()) if resid is not None else None,
                                      _ptr(y), M, N, K, int(act), cur_stream()), "car_op_dense_linear")
    return y


def op_rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.lib()
    K = x.shape[-1]
    x2 = x.reshape(-1, K).contiguous()
    y = torch.empty_like(x2)
    check(lib.car_op_rmsnorm(dtype_code(x.dtype), _ptr(x2), _ptr(w.detach().contiguous()), _ptr(y), x2.shape[0], K,
                             float(eps), cur_stream()), "car_op_rmsnorm")
    return y.reshape(x.shape)
