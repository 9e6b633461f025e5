import os, subprocess, sys
code = '''
import sys; sys.path[:0]=["llama.cpp-quant-gemm_b200","."]
import torch, quant_gemm, bench_detail
r = bench_detail.time_prefill(torch, quant_gemm, 2, 4096, 8192, 8192, reps=3)
print("us=%.0f tops=%.0f" % (r["us"], r["tops"]))
'''
for dbg in ["0", "1", "2", "3"]:
    e = dict(os.environ); e["QGEMM_MMQ_DBG"] = dbg
    out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print("dbg", dbg, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
