"""Single-GPU timing of the tcgen05 peer path with local stand-in peers (isolates kernel-side costs from NVLink)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
from quant_gemm import _lib
L = _lib.lib()
dev = torch.device("cuda")
T, F, K = 4096, 14336, 8192
w = bench_detail.make_weights(torch, 2, F, K, 1, dev)[0]
aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
world = int(os.environ.get("WORLD", "2"))
bufs = [torch.empty((F, T), device=dev) for _ in range(world)]
flag = torch.zeros(32, dtype=torch.int32, device=dev)
done = torch.zeros(1, dtype=torch.int32, device=dev)
step = torch.zeros(1, dtype=torch.int32, device=dev)
ps = _lib.QgemmPeers()
ps.world, ps.rank = world, 0
for r in range(world):
    ps.C[r] = bufs[r].data_ptr(); ps.flag[r] = flag.data_ptr()
ps.done, ps.step, ps.launches_per_step, ps.launch_index, ps.wait_index = done.data_ptr(), step.data_ptr(), 1, 0, 0
st = torch.cuda.current_stream().cuda_stream
def run():
    assert L.qgemm_gemm_peers(2, aq.data_ptr(), w.data_ptr(), ps, T, F, K, 1, T, 0x90, st) == 0
    assert L.qgemm_peer_wait(ps, st) == 0
    assert L.qgemm_peer_step_advance(step.data_ptr(), st) == 0
def plain():
    quant_gemm.gemm(w, aq, F, T, K, 2, 0x10, out=bufs[0])
for name, fn in (("plain", plain), ("peers", run)):
    best = 1e9
    for i in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(name, "world", world, "copyout", not os.environ.get("QGEMM_MMQ_NO_COPYOUT"), "ms", round(best, 3), flush=True)
print("equal", all(torch.equal(bufs[0], b) for b in bufs))
