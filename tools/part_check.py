"""SM-partition pipeline vs the plain call: bit-identical audio (run under gpurun)."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/python-liquiddsp_b200"); sys.path.insert(0, "/root/repo/tests")
import liquiddsp as L
from util import am_iq
C, n = int(sys.argv[1]) if len(sys.argv) > 1 else 300, 65536 + 4096
x = np.stack([am_iq(n, seed=40 + c, f_off=100.0 + 3 * c) for c in range(C)])
outs = []
for part in (False, True):
    for k in ("LQB_TIMEPIPE_PARTITION", "LQB_TIMEPIPE_MAX"):
        os.environ.pop(k, None)
    if part:
        os.environ["LQB_TIMEPIPE_PARTITION"] = "1"; os.environ["LQB_TIMEPIPE_MAX"] = "100000"
    else:
        os.environ["LQB_NO_TIMEPIPE"] = "1"
    iir = L.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / 2e6, channels=C)
    rs = L.ComplexResampler(rate=48e3 / 2e6, Fc=48e3 / 2e6, channels=C)
    agc = L.AGC(channels=C); agc.lock = False; agc.scale = 0.01
    am = L.AmpModem(modulation=0.5, type="dsb", carrier=True, channels=C)
    de = L.DeemphasisFilter(48000, channels=C)
    ch = L.Chain(iir, rs, agc, am, de)
    y = np.concatenate([ch(np.ascontiguousarray(x[:, :65536])), ch(np.ascontiguousarray(x[:, 65536:]))], axis=1)
    outs.append(y); print(part, ch.last_kernels(), flush=True)
    os.environ.pop("LQB_NO_TIMEPIPE", None)
print("bit-identical" if np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)) else "DIFFERENT", outs[0].shape)
