#!/bin/bash
# usage: tools/gpu_gmul.sh "<variants>" -- 2^16 G1 scalar multiplications per library variant (device resident, CUDA events)
for v in $1; do
  if [ "$v" = "default" ]; then unset BN_B200_SO; else export BN_B200_SO=$PWD/bn_b200/libv_$v.so; fi
  python - "$v" <<'PY'
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import bn_b200, bench
lib = bn_b200.init(0)
dev = torch.device('cuda', 0)
n = 1 << 16
g1gen, g2gen = bench.generators()
k = torch.from_numpy(bench.splitmix_scalars(0xB2000003, 4096).view(np.int64)).to(dev).repeat(n // 4096, 1)
base = torch.from_numpy(np.tile(g1gen, (n, 1)).view(np.int64)).to(dev)
base2 = torch.from_numpy(np.tile(g2gen, (n // 4, 1)).view(np.int64)).to(dev)
p = torch.empty_like(base); o = torch.empty_like(base); o2 = torch.empty_like(base2)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = ctypes.c_void_p(st.cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
lib.bn_b200_g1_mul_batch_dev(P(base), P(k), P(p), ctypes.c_size_t(n), sp)
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
ms1 = t(lambda: lib.bn_b200_g1_mul_batch_dev(P(p), P(k), P(o), ctypes.c_size_t(n), sp))
ms2 = t(lambda: lib.bn_b200_g2_mul_batch_dev(P(base2), P(k), P(o2), ctypes.c_size_t(n // 4), sp))
print('%-10s g1_mul 2^16: %.3f ms = %.2f M/s   g2_mul 2^14: %.3f ms = %.2f M/s' % (sys.argv[1], ms1, n / ms1 / 1e3, ms2, n / 4 / ms2 / 1e3))
PY
done
