"""Debug: per-role barrier-wait cycle accounting of CTA (0,0,0) of the attention kernels."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200 import ringattention as ra, _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _lib.load()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.lwm_debug_set_prof.argtypes = [ctypes.c_void_p]
lib.lwm_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
B, H, D = 1, 32, 128
q, k, v, do = [torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16) for _ in range(4)]
out = torch.empty_like(q); lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
for _ in range(2):
    ra.fwd_step(q, k, v, out, lse, None, None, None, 0, 0, True, None, None, True, True)
delta = torch.empty_like(lse); ra.bwd_prep(out, do, delta)
dq = torch.zeros(B, S, H, D, dtype=torch.float32, device="cuda"); dk = torch.zeros_like(dq); dv = torch.zeros_like(dq)
for _ in range(2):
    ra.bwd_step(q, k, v, do, ra.lse_for_bwd(lse), delta, dq, dk, dv, 0, 0, True, None, None)
torch.cuda.synchronize()
b = buf.cpu().tolist()
n_it = S // 128
print("S=%d; bwd CTA(0,0,0): %d Q tiles" % (S, n_it))
names = ["do_full", "dq_drained", "p_ready", "q_full_next", "ds_ready", "-"]
tot = b[6]
print(" MMA issuer total %d cyc = %.0f cyc/iter" % (tot, tot / n_it))
for i, nme in enumerate(names):
    print("   wait %-12s %9d  (%.0f/iter, %.1f%%)" % (nme, b[i], b[i] / n_it, 100.0 * b[i] / max(tot, 1)))
tot = b[20]
print(" compute thread0 total %d = %.0f/iter; waits q_full %.0f s_full %.0f dp_full %.0f per iter" % (tot, tot / n_it, b[8] / n_it, b[9] / n_it, b[10] / n_it))
print("   phase A (exp) %.0f/iter of which tmem-ld+wait %.0f, pack+tmem-st+arrive %.0f ; phase B (dS) %.0f/iter of which tmem-ld+wait %.0f" % (
    b[13] / n_it, b[11] / n_it, b[12] / n_it, b[15] / n_it, b[14] / n_it))
print(" drain issuer total %d = %.0f/iter; wait dq_full %.0f, ld+stage->drained %.0f, TMA-read wait #1 %.0f, #2 %.0f per iter" % (
    b[22], b[22] / n_it, b[16] / n_it, b[17] / n_it, b[18] / n_it, b[19] / n_it))
# fwd: CTA (0,0,0) = last q pair => n_kv = S/128 tiles
tot = b[36]
print("fwd CTA(0,0,0): MMA issuer total %d = %.0f cyc/kv-tile" % (tot, tot / n_it))
for i, nme in enumerate(["kv_full(V)", "p_ready0", "kv_full(Knext)", "p_ready1"]):
    print("   wait %-14s %.0f/iter (%.1f%%)" % (nme, b[32 + i] / n_it, 100.0 * b[32 + i] / max(tot, 1)))
print(" softmax0 thread0 total %.0f/iter; wait s_full %.0f/iter" % (b[41] / n_it, b[40] / n_it))
