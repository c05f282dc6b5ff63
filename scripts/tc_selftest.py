import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
torch.zeros(1, device="cuda")   # context
def bf16(x): return torch.tensor(x).to(torch.bfloat16).to(torch.float64).numpy()
for (M, N, K) in [(128, 128, 64), (128, 128, 256), (300, 200, 150), (1000, 512, 785), (77, 16, 17)]:
    rng = np.random.default_rng(0)
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((N, M), np.float32)
    rc = m.lib.icnf_tc_gemm_selftest(M, N, K, A.ctypes.data, B.ctypes.data, D.ctypes.data)
    ref = bf16(B) @ bf16(A).T
    err = np.abs(D - ref).max() / np.abs(ref).max()
    print((M, N, K), "rc", rc, "max rel err %.3e" % err, flush=True)
