import sys, torch
sys.path.insert(0, "bodyct-dram_b200")
import models
from dram_native import ops
B, G = 5, 64
f = ops.new_volume(B, 17, G, G, G, "cuda"); f.normal_()
cam = torch.randn(B, 1, G, G, G, device="cuda")
pcm = models.PCM((G, G, G), 17, 1, 8, 0, 8, 1, 3, "scaled_dot_product_relu", False, p_enc_dim=0).cuda()
with torch.no_grad():
    for _ in range(3):
        out = pcm(cam, f)
torch.cuda.synchronize()
print("ok", out.shape)
