#!/usr/bin/env python
"""Config 1 of BASELINE.json on the GPU: the README's AMRadio, ONE channel, 64K-sample blocks from a pageable numpy
array, state carried -- object by object (five calls per block, README.md:41-58) and as one Chain call.
Prints one JSON line; `--blocks` bounds the run (config 1 proper is 20 M samples = 306 blocks)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-liquiddsp_b200"))

BLOCK, FS, PCM = 65536, 2.0e6, 48.0e3


def synth(nblocks, seed=0xB200):
    import numpy as np
    rng = np.random.default_rng(seed)
    t = np.arange(BLOCK * nblocks) / FS
    a = 0.6 * np.sin(2 * np.pi * 1000 * t) + 0.4 * np.sin(2 * np.pi * 2500 * t)
    x = 0.1 * (1 + 0.5 * a) * np.exp(1j * (2 * np.pi * 200 * t)) + 0.05 * np.exp(2j * np.pi * 60e3 * t)
    return (x + 0.02 / np.sqrt(2) * (rng.standard_normal(t.size) + 1j * rng.standard_normal(t.size))).astype(np.complex64)


def build(L):
    iir = L.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / FS)
    rs = L.ComplexResampler(rate=PCM / FS, Fc=PCM / FS)
    agc = L.AGC(); agc.lock = False; agc.scale = 0.01
    am = L.AmpModem(modulation=0.5, type="dsb", carrier=True)
    de = L.DeemphasisFilter(PCM)
    return iir, rs, agc, am, de


def run(nblocks=64, warm=4):
    import liquiddsp as L
    x = synth(nblocks + warm)
    out = {}
    for form in ("objects", "chain"):
        st = build(L)
        if form == "chain":
            ch = L.Chain(*st)
            call = ch
        else:
            def call(v, st=st):
                for s in st:
                    v = s(v)
                return v
        for b in range(warm):
            call(x[b * BLOCK:(b + 1) * BLOCK])
        L.synchronize()
        t0 = time.perf_counter()
        for b in range(warm, warm + nblocks):
            y = call(x[b * BLOCK:(b + 1) * BLOCK])
        L.synchronize()
        dt = time.perf_counter() - t0
        out[form] = {"msamples_per_s": nblocks * BLOCK / dt / 1e6, "ms_per_block": dt / nblocks * 1e3, "audio_per_block": int(y.size)}
        if form == "chain":
            out[form]["plan"] = ch.plan()
    out["blocks"] = nblocks
    out["input"] = "pageable numpy complex64, 1 channel x 64K-sample blocks, state carried"
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=64)
    a = ap.parse_args()
    print(json.dumps(run(a.blocks)))
