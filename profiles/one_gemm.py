import sys, torch
sys.path.insert(0, '/root/repo')
from video_diffusion_b200 import ops
M, K, N = 1024, 128, 128
a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
out = torch.empty(M, N, device='cuda')
ops.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, out_f32=out)
torch.cuda.synchronize()
print('ok', float((out - a.float() @ w.float().t()).abs().max()))
