"""Single-GPU emulation of the peer path's cost (both 'peers' are local buffers; wait disabled by QGEMM_PEER_DBG=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
from quant_gemm import _lib
dev = torch.device("cuda")
L = _lib.lib()
F, K, T, n = 4096, 4096, 1, 64
w = bench_detail.make_weights(torch, 2, F, K, n, dev)
aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
outA = torch.zeros((2 * F, T), device=dev); outB = torch.zeros((2 * F, T), device=dev)
flags = torch.zeros(64, dtype=torch.int32, device=dev); done = torch.zeros(1, dtype=torch.int32, device=dev); step = torch.zeros(1, dtype=torch.int32, device=dev)
def mk(i):
    ps = _lib.QgemmPeers(); ps.world, ps.rank = 2, 0
    ps.C[0], ps.C[1] = outA.data_ptr(), outB.data_ptr()
    ps.flag[0], ps.flag[1] = flags.data_ptr(), flags.data_ptr() + 128
    ps.done, ps.step, ps.launches_per_step, ps.launch_index = done.data_ptr(), step.data_ptr(), n, i
    return ps
pss = [mk(i) for i in range(n)]
stream = torch.cuda.Stream()
def run(peer):
    with torch.cuda.stream(stream):
        def body():
            for i in range(n):
                if peer:
                    rc = L.qgemm_gemm_peers(2, aq.data_ptr(), w[i].data_ptr(), pss[i], T, F, K, 1, T, 0x10, stream.cuda_stream)
                else:
                    rc = L.qgemm_gemm(2, aq.data_ptr(), w[i].data_ptr(), outA.data_ptr(), T, F, K, 1, T, 0x10, None, 0, stream.cuda_stream)
                assert rc == 0, rc
        body(); stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            body()
        for _ in range(3): g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): g.replay()
        e1.record(stream); stream.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 5 / n
print("dbg", os.environ.get("QGEMM_PEER_DBG"), "plain %.2f us   peer-mode %.2f us" % (run(False), run(True)))
