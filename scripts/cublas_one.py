import sys, torch
M, N, K = (int(x) for x in sys.argv[1:4])
x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): torch.matmul(x, w.t(), out=out)
torch.cuda.synchronize(); print("done")
