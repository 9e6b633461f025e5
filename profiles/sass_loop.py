"""Offline: instruction mix of the MMQ epilogue's per-block loop body (from LDTM to the back-edge)."""
import re, subprocess, sys, collections
so = "llama.cpp-quant-gemm_b200/lib/libqgemm_sm100.so"
kern = sys.argv[1] if len(sys.argv) > 1 else "mmq_kernelILi2ELb0"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
lines, on = [], False
for l in out.splitlines():
    if "Function :" in l:
        on = kern in l
    elif on and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m: lines.append((int(m.group(1), 16), m.group(2).strip()))
idx = [i for i, (a, t) in enumerate(lines) if "LDTM" in t]
if not idx: sys.exit("no LDTM")
i0 = idx[-1] if len(idx) < 3 else idx[0]
# find the FFMA2/FFMA region following the LDTM of the non-dump path: take the LDTM that is followed by I2FP/FFMA
for i in idx:
    seg = [t for a, t in lines[i:i + 200]]
    if any("FFMA" in t for t in seg): i0 = i
# loop body: from the TRYWAIT before LDTM to the backward branch after the last FFMA
j = i0
while j > 0 and "TRYWAIT" not in lines[j][1]: j -= 1
k = i0
last_f = max(i for i in range(i0, min(len(lines), i0 + 400)) if "FFMA" in lines[i][1])
k = last_f
while k < len(lines) and not lines[k][1].split()[-1].startswith("0x"): k += 1
body = [t for a, t in lines[j:k + 1]]
mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in body)
print(kern, "loop body instructions:", len(body))
print(", ".join(f"{k}:{v}" for k, v in mix.most_common()))
