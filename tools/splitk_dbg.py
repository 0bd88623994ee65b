"""Role timers of one convolution with and without split-K (diagnosis).  python tools/splitk_dbg.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fav
from fav import _lib
p, h, w, cin, cout, k = 1, 15, 20, 512, 512, 3
x = torch.randn((p, h, w, cin)).to(torch.bfloat16).cuda()
wt = (torch.randn((cout, k, k, cin)) / (k * k * cin) ** 0.5).to(torch.bfloat16).cuda()
bias = torch.randn(cout).cuda()
y = torch.empty((p, h, w, cout), dtype=torch.bfloat16, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for split in (0, 1):
    hd = _lib.Handle(0)
    _lib.check(hd.lib.fav_set_option(hd.h, b"splitk", split), "opt")
    for _ in range(3):
        _lib.check(hd.lib.fav_conv2d(hd.h, P(x), P(wt), P(bias), None, P(y), p, h, w, cin, cout, k, k, 1, 1, 1, 0, 0, st), "conv")
    hd.lib.fav_conv_timing_enable(hd.h, 1)
    _lib.check(hd.lib.fav_conv2d(hd.h, P(x), P(wt), P(bias), None, P(y), p, h, w, cin, cout, k, k, 1, 1, 1, 0, 0, st), "conv")
    s8 = (C.c_uint64 * (8 * 8))(); ns = C.c_int()
    _lib.check(hd.lib.fav_conv_stats_read(hd.h, s8, 8, C.byref(ns)), "stats")
    ms = (C.c_float * 8)(); gf = (C.c_float * 8)(); n = C.c_int()
    _lib.check(hd.lib.fav_conv_timing_read_all(hd.h, ms, gf, 8, C.byref(n)), "read")
    v = list(s8[:8]); nc = max(1, v[7])
    print(f"split={split}: {ms[0]*1e3:.1f} us, ctas {nc}; per-CTA cycles: tma total {v[1]/nc:.0f} (wait-empty {v[0]/nc:.0f}), mma total {v[4]/nc:.0f} "
          f"(wait-full {v[2]/nc:.0f}, wait-tmem {v[3]/nc:.0f}), epi total {v[6]/nc:.0f} (wait-acc {v[5]/nc:.0f})")
