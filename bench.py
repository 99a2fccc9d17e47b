#!/usr/bin/env python
"""bench.py — images/sec of the ControlAR conditional-decoding hot path (BASELINE.json metric) on N GPUs.

One "step" = one pass of the hot path over one batch of synthetic input: control map -> DINOv2 + adapter_mlp ->
prefill -> N-1 KV-cache decode steps with CFG + top-k sampling -> VQGAN decode of the token grids, i.e. exactly
`generate()` + `decode_code()` of autoregressive/sample/sample_t2i.py:163-176, through this repo's drop-in modules.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 8] [--size 512]

N > 1 is launched with torchrun (one rank per GPU); every rank generates its own `--batch` images (weak scaling)
and the int32 token grids are all-gathered once over NCCL.  `--impl reference` times the CPU restatement of the
reference path (oracle/, kind "port": the reference is Python and /root/reference does not exist on the GPU box)
on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--height", type=int, default=0, help="non-square generation (config 4: --height 512 --width 768); default --size")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--adapter-size", default="small", choices=["small", "base"], help="DINOv2-small (config 2) or -base (config 3)")
    ap.add_argument("--condition-type", default="canny", help="canny | depth | hed | lineart | seg (config 3: depth)")
    ap.add_argument("--model", default="GPT-XL")
    ap.add_argument("--cfg-scale", type=float, default=4.0)
    ap.add_argument("--top-k", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


WORKLOAD = "configs[1]: LlamaGen-XL t2i + DINOv2-small canny, 512x512 (1024 tokens), batch=8/GPU, cfg 4.0, top-k 2000"


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (recipe: /opt/skills/guides/B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0.0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline = the oracle port, bounded sample extrapolated to images/sec
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_images_per_sec(args, n_decode_steps=4, threads=None):
    """Times prefill + a few decode steps + one VQ decode + DINOv2 of the CPU port of the reference on the host cores
    and extrapolates to a full batch (the reference's decode-step cost is position independent: it attends over all S
    cache slots every step, gpt_t2i.py:276-286).  Returns (images_per_sec, cores, sample description)."""
    from oracle.weights import GPTSpec, make_gpt_state_dict, make_vq_state_dict, dinov2_shapes, _fill
    from oracle.ar_oracle import AROracle, cfg_combine, sample_from_logits
    from oracle.vision_oracle import dinov2_adapter_oracle, vq_decode_oracle
    from oracle.inputs import text_inputs, control_map
    cores = threads or min(os.cpu_count() or 1, 32)     # more threads than this only thrash on the shared GPU hosts
    torch.set_num_threads(cores)
    dims = {"GPT-XL": (1280, 36, 20), "GPT-L": (1024, 24, 16), "GPT-B": (768, 12, 12)}[args.model]
    g = args.size // 16
    spec = GPTSpec(dim=dims[0], n_layer=dims[1], n_head=dims[2], vocab_size=16384, cls_token_num=120, block_size=g * g,
                   model_type="t2i")
    sd = make_gpt_state_dict(spec, 0, with_adapter=False)
    orc = AROracle(spec, sd, torch.bfloat16)
    B, N, T = min(args.batch, 2), g * g, 120          # bounded sample: 2 images, extrapolated per image
    cond, masks = text_inputs(T, spec.caption_dim, B, 1)
    cmap = control_map(B, args.size, args.size, 2, "canny")
    dsd = _fill(dinov2_shapes(384, prefix="model."), 0, 0.02)
    t0 = time.perf_counter()
    feat = dinov2_adapter_oracle(dsd, cmap[:1], "canny", torch.bfloat16, heads=6).float()
    t_dino = (time.perf_counter() - t0) * B
    t0 = time.perf_counter()
    ctrl = orc.mlp(orc.r(feat), "adapter_mlp").repeat(B, 1, 1)
    cc = torch.cat([cond, torch.zeros_like(cond) + orc.w["cls_embedding.uncond_embedding"]])
    cic = torch.cat([ctrl, torch.zeros_like(ctrl)])
    orc.setup_caches(2 * B, T + N)
    orc.apply_emb_masks(torch.cat([masks, masks]))
    lg = orc.prefill(cc, cic, 1.0)[:, -1]
    tok, _ = sample_from_logits(cfg_combine(lg, args.cfg_scale), top_k=args.top_k)
    t_prefill = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(n_decode_steps):
        t = tok.view(-1)
        lg = orc.decode(torch.cat([t, t]), T + i)
        tok, _ = sample_from_logits(cfg_combine(lg, args.cfg_scale), top_k=args.top_k)
    t_step = (time.perf_counter() - t0) / n_decode_steps
    vsd = make_vq_state_dict(0)
    codes = torch.randint(0, 16384, (1, N))
    t0 = time.perf_counter()
    vq_decode_oracle(vsd, codes, [1, 8, g, g])
    t_vq = (time.perf_counter() - t0) * B
    total = t_dino + t_prefill + t_step * (N - 1) + t_vq
    sample = (f"CPU port of the reference (oracle/), bf16 AR + fp32 VQ, batch {B} (B_eff {2 * B}): DINOv2 1 img {t_dino / B:.2f}s x{B}, "
              f"prefill {t_prefill:.1f}s, {n_decode_steps} decode steps {t_step:.2f}s/step extrapolated x{N - 1}, "
              f"VQ decode 1 img {t_vq / B:.1f}s x{B}; images/s = {B} / total")
    return B / total, cores, sample


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step = the bounded CPU sample of the workload (cpu_reference_images_per_sec: ~30-70 s of host work); the whole arm is
    # capped at ~4 minutes: at most `warmup` untimed samples while they are cheap, then up to `steps` timed ones
    vals = []
    t_start = time.perf_counter()
    budget = 240.0
    n_warm = 0
    while n_warm < args.warmup and time.perf_counter() - t_start < 0.25 * budget:
        cpu_reference_images_per_sec(args, n_decode_steps=2)
        n_warm += 1
    while len(vals) < args.steps:
        ips, cores, sample = cpu_reference_images_per_sec(args, n_decode_steps=2)
        vals.append(ips)
        if time.perf_counter() - t_start > budget:
            break
    v = sum(vals) / len(vals)
    line = {"metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": n_warm,
            "ms_per_step": 1000.0 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "global_batch": args.gpus * args.batch, "tokens_per_image": (args.size // 16) ** 2,
                       "parallelism": f"dp{args.gpus} (batch sharded, one all-gather of token grids)",
                       "l2": "working set larger than L2 (weights 1.5 GB + KV cache up to 3.4 GB stream every decode step)",
                       "sampling": {"cfg_scale": args.cfg_scale, "top_k": args.top_k, "temperature": 1.0, "top_p": 1.0}},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from controlar_b200.build import build
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()
    from controlar_b200 import _lib
    from controlar_b200.autoregressive.models.gpt_t2i import GPT_models
    from controlar_b200.autoregressive.models.generate import generate
    from controlar_b200.tokenizer.tokenizer_image.vq_model import VQ_models
    from controlar_b200.parallel import gather_token_grids, rank_seed
    from oracle.inputs import text_inputs, control_map     # seeded synthetic inputs only (no oracle compute here)

    torch.manual_seed(0)
    H_img, W_img = (args.height or args.size), (args.width or args.size)
    gh, gw = H_img // 16, W_img // 16
    g = max(gh, gw)                                   # RoPE table side = image_size / 16 (sample_t2i_MR.py:72-74)
    N, T, B = gh * gw, 120, args.batch
    gpt = GPT_models[args.model](block_size=g * g, cls_token_num=T, model_type="t2i", condition_type=args.condition_type,
                                 adapter_size=args.adapter_size).eval()
    gpt.output.weight.data.normal_(0, 0.02)          # the reference zero-inits the head (gpt_t2i.py:377)
    for blk in gpt.adapter.model.encoder.layer:      # HF init has layerscale 1.0
        blk.layer_scale1.lambda1.data.fill_(1.0); blk.layer_scale2.lambda1.data.fill_(1.0)
    gpt = gpt.to(dev, torch.bfloat16)
    vq = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).to(dev).eval()
    cond_h, masks_h = text_inputs(T, 2048, B, 1000 + rank, torch.bfloat16)
    cmap_h = control_map(B, H_img, W_img, 2000 + rank, args.condition_type, torch.bfloat16)
    cond_h, masks_h, cmap_h = cond_h.pin_memory(), masks_h.pin_memory(), cmap_h.pin_memory()
    img_h = torch.empty((B, 3, H_img, W_img), dtype=torch.float32).pin_memory()
    cond_d, masks_d, cmap_d = cond_h.to(dev), masks_h.to(dev), cmap_h.to(dev)
    kw = dict(cfg_scale=args.cfg_scale, temperature=1.0, top_k=args.top_k, top_p=1.0, sample_logits=True)
    lib = _lib.lib()
    t_decode = []

    def one_step(step_idx, host_io):
        if host_io:
            c, m, x = cond_h.to(dev, non_blocking=True), masks_h.to(dev, non_blocking=True), cmap_h.to(dev, non_blocking=True)
        else:
            c, m, x = cond_d, masks_d, cmap_d
        # rank seed mirrors sample_c2i_ddp.py:47 (global_seed * world + rank), advanced per step
        toks = generate(gpt, c, N, emb_masks=m, condition=x, seed=rank_seed(step_idx, world, rank), **kw)
        allt = gather_token_grids(toks)                  # the single NCCL all-gather of finished token grids (world > 1)
        img = vq.decode_code(toks, [B, 8, gh, gw])
        if host_io:
            img_h.copy_(img, non_blocking=True)
        return toks, img

    def timed(n_steps, host_io, first_idx):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            one_step(first_idx + i, host_io)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for i in range(args.warmup):
        one_step(i, False)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.car_launch_count(1)
    ms = timed(args.steps, False, args.warmup)
    launches = int(lib.car_launch_count(0))
    ms_e2e = timed(args.steps, True, args.warmup + args.steps)
    # decode-loop roofline: time the device-side decode loop alone (prefill excluded) with CUDA events on the stream
    # the kernels are launched on (torch's current stream is the stream handed to the library)
    st = gpt._car_state
    cc = torch.cat([cond_d, torch.zeros_like(cond_d) + gpt.cls_embedding.uncond_embedding])
    ctrl = gpt._car_encoder.forward(cmap_d, apply_mlp=True)
    cic = torch.cat([ctrl, torch.zeros_like(ctrl)])
    from controlar_b200.engine import make_sampling
    sp = make_sampling(1.0, args.top_k, 1.0, True, args.cfg_scale, -1, 7)
    dec_ms = []
    for _ in range(max(args.steps, 3)):
        st.prefill(cc, cic, 1.0, all_rows=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.generate(sp, N, None, dev); e1.record()
        torch.cuda.synchronize()
        dec_ms.append(e0.elapsed_time(e1))
    dec_ms = sorted(dec_ms)[len(dec_ms) // 2]
    step_bytes = sum(st.step_bytes(T + i + 1) for i in range(1, N))
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = step_bytes / (dec_ms * 1e-3) / 1e9
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.model, H_img, W_img, args.batch, args.adapter_size, args.condition_type) == ("GPT-XL", 512, 512, 8, "small", "canny") else
                   f"{args.model} t2i + DINOv2-{args.adapter_size} {args.condition_type}, {W_img}x{H_img}, batch={args.batch}/GPU",
                   "global_batch": world * B, "tokens_per_image": N, "parallelism": f"dp{world} (batch sharded, one all-gather of token grids)",
                   "l2": "working set larger than L2 (weights 1.5 GB + KV cache up to 3.4 GB stream every decode step)",
                   "sampling": {"cfg_scale": args.cfg_scale, "top_k": args.top_k, "temperature": 1.0, "top_p": 1.0}},
        "e2e": {"value": e2e, "unit": "images/s",
                "h2d_bytes_per_step": int(cond_h.numel() * 2 + masks_h.numel() * 8 + cmap_h.numel() * 2),
                "d2h_bytes_per_step": int(img_h.numel() * 4)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "traffic_note": "ncu --set full of a 24-token launch (profiles/r1_pk_decode_final_ncu.csv): dram read+write = 1.02 x algorithmic bytes", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                     "kernel": "pk_decode_kernel: persistent decode loop (per token 36 x {qkv | attention | wo | w1w3 | w2} + head + CFG/top-k sampler), one launch per generate()",
                     "algorithmic_bytes": step_bytes, "decode_ms": dec_ms, "ms_per_token": dec_ms / (N - 1)},
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_reference_images_per_sec(args, n_decode_steps=3)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
