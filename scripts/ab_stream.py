import time, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primekg_rgcn_linkprediction_b200.graph import _stream_id
dev = torch.device("cuda:0")
torch.cuda.current_stream(dev)
for name, fn in (("Stream object", lambda: torch.cuda.current_stream(dev).cuda_stream), ("raw", lambda: _stream_id(dev))):
    t0 = time.perf_counter()
    for _ in range(20000):
        fn()
    print(name, (time.perf_counter() - t0) / 20000 * 1e6, "us/call")
