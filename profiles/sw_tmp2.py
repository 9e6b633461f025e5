import sys
sys.path.insert(0,'profiles')
import sweep_gemv as s
for shape in [(2, 2, 4096, 4096, 0x310), (2, 2, 11008, 4096, 0x310), (7, 4, 11008, 4096, 0x10), (7, 8, 11008, 4096, 0x10), (2, 8, 11008, 4096, 0x10), (8, 8, 11008, 4096, 0x10)]:
    s.run(shape, {})
