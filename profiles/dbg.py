import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np, torch, quant_gemm, datagen
for wt in (2,3,6,7,8):
  for T in (1,2,3,4,5,6,7,8):
    for nb in (8, 64, 128, 344):
      for fl in (0, 1):
        F=200
        wq = torch.from_numpy(datagen.fuzz_weight_blocks(wt, F, nb, seed=1)).cuda()
        aq = torch.from_numpy(datagen.fuzz_act_blocks(T, nb, seed=1, const_ds=False)).cuda()
        try:
            quant_gemm.gemm(wq, aq, F, T, nb*32, wt, 0x200|fl); torch.cuda.synchronize()
        except Exception as e:
            print("FAIL", wt, T, nb, fl, str(e)[-120:]); 
print("done")
