"""Tuning aid: time one GEMV shape under different env knobs (each in a fresh process)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, json
sys.path[:0] = [%r, %r]
import torch, quant_gemm, bench_detail
wt, T, F, K, flags = %s
print(json.dumps(bench_detail.time_shape(torch, quant_gemm, wt, T, F, K, flags, pool_bytes=%d)))
'''
def run(shape, env, pool=768 << 20):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", code % (os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT, shape, pool)],
                         env=e, capture_output=True, text=True)
    try:
        r = json.loads(out.stdout.strip().splitlines()[-1])
        print(shape, env, "us=%.2f gbs=%.0f" % (r["us"], r["gbs"]), flush=True)
    except Exception:
        print(shape, env, "FAILED", out.stderr[-300:], flush=True)
if __name__ == "__main__":
    for shape in [(2, 1, 32768, 4096, 0x10), (2, 1, 11008, 4096, 0x10), (2, 1, 4096, 4096, 0x10), (2, 1, 4096, 11008, 0x10)]:
        for env in [{}, {"QGEMM_GEMV_STAGES": "3"}, {"QGEMM_GEMV_STAGES": "8"}, {"QGEMM_GEMV_RT": "16"}, {"QGEMM_GEMV_RT": "16", "QGEMM_GEMV_STAGES": "8"}, {"QGEMM_GEMV_RT": "32", "QGEMM_GEMV_STAGES": "4"}]:
            run(shape, env)
