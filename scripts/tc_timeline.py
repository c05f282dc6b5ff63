"""Timeline of CTA 0 of one tensor-core GEMM (development aid): where the main loop and the epilogue spend their cycles."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cnf_b200 as m
torch.zeros(1, device="cuda")
fn = m.lib.icnf_tc_gemm_timeline
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 5 + [C.c_void_p]
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 512, 512)))
split = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ep = int(sys.argv[5]) if len(sys.argv) > 5 else 0
buf = np.zeros(3 * 8192, np.int64)
assert fn(M, N, K, split, ep, buf.ctypes.data) == 0
ev = []
for r in range(3):
    b = buf[r * 8192:(r + 1) * 8192]
    n = int(b[0])
    ev += [(int(b[2 + 2 * i]), int(b[1 + 2 * i])) for i in range(n)]
ev.sort()
t0 = ev[0][0]
print(f"GEMM {M}x{N}x{K} split={split} ep={ep} mode={os.environ.get('ICNF_TC_MODE','auto')}: {len(ev)} events, span {ev[-1][0]-t0} cycles")
names = {1: "tma", 2: "landed", 3: "mma", 4: "acc_ready", 5: "epi_done", 6: "chunk_begin", 7: "chunk_end"}
for t, tag in ev:
    print(f"{t - t0:8d}  {names[tag // 1000]:10s} {tag % 1000}")
