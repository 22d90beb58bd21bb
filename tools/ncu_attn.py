import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib()
from chatllm_cpp_b200 import session as S
cfg = S.make_config("llama3-8b", pkg.Q4_K, layers=2, max_len=4352)
s = S.DecodeSession(cfg, seed=0); s.fill_kv_random(4097); c = cfg
for i in range(6):
    W = s.layers[i % 2]
    L.b200_attn_decode(s.q.data_ptr(), W.kc.data_ptr(), W.vc.data_ptr(), s.att.data_ptr(), s.scratch.data_ptr(), c.heads, c.kv_heads, c.head_dim, 4097, c.kv_hidden, c.max_len, 1.0 / math.sqrt(c.head_dim), 0)
torch.cuda.synchronize()
