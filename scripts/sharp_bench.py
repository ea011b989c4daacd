"""sharp()/flat() per-kernel timing for beta = 0 and beta != 0 (library chosen by LGM_LIB_PATH)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
from lagomorph_b200 import _lib as L
dev = torch.device("cuda")
N, n = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 128
m = torch.randn(N, 3, n, n, n, device=dev)
for params in ([0.1, 0.0, 0.01], [0.1, 0.01, 0.001]):
    met = lm.FluidMetric(params)
    for _ in range(3): met.sharp(m)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    L.check(L.lib.lgm_profile_begin(L.stream_ptr(dev)))
    for _ in range(5): met.sharp(m)
    L.check(L.lib.lgm_profile_end(buf, len(buf)))
    ks = json.loads(buf.value.decode())
    print(os.environ.get("LGM_LIB_PATH", "default").split("/")[-1], params, " ".join("%s %.4f" % (k, v["ms"] / v["launches"]) for k, v in sorted(ks.items())))
