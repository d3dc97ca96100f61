"""Front kernel time versus arithmetic per sample: IIR order 2/4/6/8 (1..4 sections) in front of the resampler, 65536 x 65536."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "python-liquiddsp_b200"))
import torch, liquiddsp as L
C, n = 65536, 65536
x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
L.synth_fill(0, x.data_ptr(), C, n, stream=s)
y = torch.empty((C, 1600), dtype=torch.complex64, device="cuda")
for order in (2, 4, 6, 8):
    ch = L.Chain(L.ComplexIIRFilter("cheby2", order=order, Fc=0.0075, channels=C), L.ComplexResampler(0.024, Fc=0.024, channels=C))
    for _ in range(3):
        ch.execute_dev(x.data_ptr(), n, y.data_ptr(), 1600, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ch.execute_dev(x.data_ptr(), n, y.data_ptr(), 1600, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("sections %d  plan %-20s %.3f ms  %.0f GB/s (%.1f %% of 6461)" % (order // 2, ch.plan(), ms, C * n * 8.192 / ms / 1e6, C * n * 8.192 / ms / 1e6 / 64.615))
