"""Measured FP64 peaks of the GPU this runs on (roofline denominators the driver's
MEASURED_PEAKS.json does not hold): DMMA.8x8x4 chains (FP64 tensor pipe) and DFMA chains."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == '__main__':
    import torch
    from optconpy_b200 import device as dv
    out = dict(gpu=torch.cuda.get_device_name(0), dmma_tflops=dv.fp64_peak('dmma'),
               dfma_tflops=dv.fp64_peak('dfma'),
               how='register-only chains, 8 independent accumulator pairs per thread, 256 threads per '
                   'CTA, best of 1/2/4 CTAs per SM, 20000 iterations, CUDA events (ocb_fp64_peak)')
    print(json.dumps(out))
