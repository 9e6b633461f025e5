import sys
sys.path.insert(0,'profiles')
import sweep_gemv as s
for shape in [(2, 1, 4096, 11008, 0x10), (8, 1, 4096, 11008, 0x10), (2, 1, 4096, 14336, 0x10), (2, 1, 8192, 8192, 0x10)]:
    s.run(shape, {})
