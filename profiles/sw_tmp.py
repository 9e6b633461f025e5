import sys
sys.path.insert(0,'profiles')
import sweep_gemv as s
for env in [{}, {"QGEMM_GEMV_NOCOMPUTE":"1"}]:
  for shape in [(2, 1, 32768, 4096, 0x10), (2, 1, 11008, 4096, 0x10), (2, 1, 4096, 4096, 0x10), (2, 1, 4096, 11008, 0x10), (2, 1, 4096, 4096, 0), (8, 1, 11008, 4096, 0x10),(7, 1, 11008, 4096, 0x10), (2, 2, 11008, 4096, 0x10), (2, 4, 11008, 4096, 0x10), (2, 8, 11008, 4096, 0x10)]:
    s.run(shape, env)
