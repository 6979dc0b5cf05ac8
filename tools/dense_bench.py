"""Times the dense FP64 tensor-pipe kernels (DMMA) at the sizes the hot path uses them:
Gram product Z^T Z (compression), tall product Z T, and the whole compression."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == '__main__':
    import torch
    from optconpy_b200 import device as dv
    peak = dv.fp64_peak('dmma')
    out = []
    for n, K, kc in ((4802, 1296, 50), (4802, 1716, 50), (97146, 3072, 192)):
        Z = torch.randn((n, K), dtype=torch.float64, device='cuda')
        T = torch.randn((K, kc), dtype=torch.float64, device='cuda')
        # low numerical rank as an ADI factor has, so that the compression does what it does in the DRE
        Zl = (torch.randn((n, 96), dtype=torch.float64, device='cuda') @
              torch.randn((96, K), dtype=torch.float64, device='cuda')).contiguous()

        def timed(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)/reps
        ms_g = timed(lambda: dv.gram(Z, Z))
        ms_t = timed(lambda: dv.tall_gemm(Z, T))
        ms_c = timed(lambda: dv.compress(Zl, thresh=1e-9, k=kc), reps=3)
        fl_g = n*K*K          # symmetric half: 2 n K^2 / 2
        fl_t = 2.0*n*K*kc
        out.append(dict(n=n, K=K, kc=kc, gram_ms=round(ms_g, 3), gram_TFs=round(fl_g/ms_g/1e9, 2),
                        gram_frac_dmma_peak=round(fl_g/ms_g/1e9/peak, 3), tall_ms=round(ms_t, 3),
                        tall_TFs=round(fl_t/ms_t/1e9, 2), tall_GBs=round(8.0*n*(K + kc)/ms_t/1e6, 1),
                        compress_ms=round(ms_c, 3)))
        print(json.dumps(out[-1]), flush=True)
    print(json.dumps(dict(dmma_peak_TFs=peak)))
