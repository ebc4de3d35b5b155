"""IMAD.WIDE.U32.X issue rate of k_imad_peak as a function of resident warps per scheduler (GPU box only)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bn_b200
bn_b200.init(0)
lib = bn_b200.load()
scratch = torch.zeros(1024, dtype=torch.int32, device="cuda:0")
sms = lib.bn_b200_sm_count()
stream = torch.cuda.Stream()
sp = ctypes.c_void_p(stream.cuda_stream)
for mult in (1, 2, 3, 4, 6, 8):
    blocks, iters = sms * mult, 8192
    best = 1e30
    for _ in range(4):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record(stream)
        rc = lib.bn_b200_imad_peak_dev(ctypes.c_void_p(scratch.data_ptr()), blocks, iters, sp)
        a1.record(stream)
        torch.cuda.synchronize()
        assert rc == 0
        best = min(best, a0.elapsed_time(a1))
    rate = blocks * 256 * iters * 32 / (best * 1e-3)
    print("warps/scheduler %2d  %.2f T IMAD.WIDE/s  = %.2f cycles per warp-instruction per scheduler" % (
        2 * mult, rate / 1e12, 1.965e9 * sms * 4 * 32 / rate))
