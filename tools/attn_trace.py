import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib
L = ctypes.CDLL(os.environ["MASKBIT_B200_LIB"])
n_seq = 512
qkv = torch.randn((n_seq * 257, 3072), device="cuda").to(torch.bfloat16)
out = torch.empty((n_seq * 257, 1024), dtype=torch.bfloat16, device="cuda")
tr = torch.zeros((8, 8, 12), dtype=torch.int64, device="cuda")
L.mb_test_attention_trace(ctypes.c_void_p(tr.data_ptr()))
for _ in range(2):
    L.mb_test_attention(ctypes.c_void_p(qkv.data_ptr()), ctypes.c_void_p(out.data_ptr()), n_seq, 257, 1024, 16, None)
torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
names = {0: ["tma_issue"], 1: ["S0_issue", "S1_issue", "PV0_issue", "PV1_issue"], 2: ["cls_start", "cls_done"],
         3: ["wg0 s_full", "wg0 pass1", "wg0 p_full", "wg0 o_full", "wg0 done"], 4: ["wg1 s_full", "wg1 pass1", "wg1 p_full", "wg1 o_full", "wg1 done"],
         5: ["mma0 qk_in", "mma0 t_free", "mma0 p_full", "mma0 v_in"], 6: ["mma1 qk_in", "mma1 t_free", "mma1 p_full", "mma1 v_in"]}
if len(sys.argv) > 1:
    q, k, v = qkv.double().view(n_seq, 257, 3, 16, 64)[:4].permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).permute(0, 2, 1, 3).reshape(4 * 257, 1024)
    print("max err", (out[:4 * 257].double() - ref).abs().max().item())
for role, evs in names.items():
    for e, nm in enumerate(evs):
        print(f"{nm:12s}", " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in t[role, e]))
