import sys
sys.path.insert(0,'profiles')
import sweep_gemv as s
for shape in [(2, 4, 4096, 11008, 0x10), (2, 8, 4096, 11008, 0x10), (8, 8, 4096, 11008, 0x10), (7, 8, 4096, 14336, 0x10), (2, 2, 4096, 11008, 0x10)]:
    s.run(shape, {})
