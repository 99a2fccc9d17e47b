"""dev: one DINOv2-small forward + one VQ decode_code + one VQ encode at 512 x 512 (profiling target for the vision kernels)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from controlar_b200.tokenizer.tokenizer_image.vq_model import VQ_models
from controlar_b200.autoregressive.models.dinov2_adapter import Dinov2_Adapter

B = int(os.environ.get("B", 8))
vq = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).cuda().eval()
ad = Dinov2_Adapter(adapter_size="small", condition_type="canny").to("cuda", torch.bfloat16).eval()
codes = torch.randint(0, 16384, (B, 1024), device="cuda")
cmap = torch.randn(B, 3, 512, 512, device="cuda", dtype=torch.bfloat16)
for it in range(3):
    e = [torch.cuda.Event(True) for _ in range(4)]
    e[0].record(); f = ad(cmap); e[1].record(); img = vq.decode_code(codes, [B, 8, 32, 32]); e[2].record()
    _, _, (_, _, idx) = vq.encode(img[:1].clamp(-1, 1)); e[3].record()
    torch.cuda.synchronize()
    print(f"iter {it}: DINOv2-small {B} img {e[0].elapsed_time(e[1]):.2f} ms | VQ decode {B} img {e[1].elapsed_time(e[2]):.2f} ms | VQ encode (fp32 grade) 1 img {e[2].elapsed_time(e[3]):.2f} ms", flush=True)
