"""Bandwidth-bound kernels at the cfg4 size (vitl14 @518: 32 x 1369 tokens x 1024; images 32 x 518^2; student map 32 x 768 x 37^2),
one or two launches each, for an `ncu --set full` capture (profiles/r02_hbm_ncu_full.md)."""
import sys
import torch
sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops
from dinov2_distillation_b200.scalekd import ScaleKD
B, HW, D = 32, 1369, 1024
M = B * HW
w, b = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda")
x = torch.randn(M, D, device="cuda")
dy = torch.randn(M, D, device="cuda")
dres = torch.randn(M, D, device="cuda")
mean, rstd = torch.randn(M, device="cuda"), torch.rand(M, device="cuda") + 0.5
for _ in range(2):
    ops.layernorm_fwd(x, w, b, 1e-6, want_f32=False, want_bf16=True)
    ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_wgrad=False, want_bf16=True)
    ops.layernorm_bwd(dy, x, w, mean, rstd, want_wgrad=True, want_bf16=True)
S = torch.randn(B, HW, D, device="cuda")
T = torch.randn(B, HW + 1, D, device="cuda")
g = torch.tensor([1.0, 0.0], device="cuda")
for _ in range(2):
    _, ws = ops.kd_loss_fwd(S, T, 1, False, 0.08)
    ops.kd_loss_bwd(S, T, 1, False, 0.08, g, ws)
    ops.cast_bf16(x)
img = torch.randn(B, 3, 518, 518, device="cuda")
for _ in range(2):
    ops.patch_im2col(img)
m = ScaleKD(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=768, teacher_dims=D, query_hw=[37, 37], pos_hw=[37, 37],
            pos_dims=D, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=16).cuda().train()
fm = torch.randn(B, 768, 37, 37, device="cuda")
for _ in range(2):
    m.projector_0.tokenize(fm)
torch.cuda.synchronize()
print("done")
