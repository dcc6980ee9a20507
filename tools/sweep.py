#!/usr/bin/env python3
"""BASELINE configs[3]: synthetic NTT / MSM sweep with the library's phase timers
(CUDA events on the launching streams).  Prints one JSON line per case.

    python tools/sweep.py [--ntt 16,18,20,22,24] [--msm 16,18,20,22] [--g2 16,20]
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b = importlib.import_module("zksnark-finalproject_b200")
PH = ["ntt_pass", "wm_pointwise", "msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce", "r1cs_eval"]


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    a = (a[:, 0::2] | (a[:, 1::2] << np.uint64(32))).astype(np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return np.ascontiguousarray(a)


def read_profile(ctx):
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_uint64 * 8)()
    units = (ctypes.c_uint64 * 8)()
    ctx.check(ctx._lib.b2z_profile_read(ctx.handle, ms, cnt, units, 1))
    return {PH[i]: (ms[i], int(cnt[i]), int(units[i])) for i in range(len(PH))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--msm", default="16,18,20,22")
    ap.add_argument("--g2", default="16,20")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    ctx = b.Context(0)
    L = ctx._lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    imad, imadw = ctypes.c_double(), ctypes.c_double()
    ctx.check(L.b2z_measure_int_peak(ctx.handle, ctypes.byref(imad), ctypes.byref(imadw)))
    print(json.dumps({"int_peak_imad_T": imad.value / 1e12, "int_peak_wide_mac_T": imadw.value / 1e12, "hbm_gbs": hbm}),
          flush=True)

    for lg in [int(x) for x in args.ntt.split(",") if x]:
        n = 1 << lg
        data = rand_fr(n, lg)
        dom = b.Radix2EvaluationDomain(ctx, n)
        for kind, fn in (("forward", dom.fft), ("inverse", dom.ifft), ("coset_forward", dom.get_coset(7).fft)):
            fn(data)                                    # warm-up: builds twiddle tables
            fn(data)
            L.b2z_profile_enable(ctx.handle, 1)
            read_profile(ctx)
            for _ in range(args.reps):
                fn(data)
            prof = read_profile(ctx)
            L.b2z_profile_enable(ctx.handle, 0)
            ms = prof["ntt_pass"][0] / args.reps
            passes = prof["ntt_pass"][1] / args.reps
            butterflies = n / 2 * lg
            print(json.dumps({
                "case": "ntt", "kind": kind, "log_n": lg, "ms_passes": ms, "passes": passes,
                "hbm_gbs_algorithmic": 64.0 * n / (ms * 1e-3) / 1e9, "hbm_frac": 64.0 * n / (ms * 1e-3) / 1e9 / hbm,
                "int_frac": 136.0 * butterflies / (ms * 1e-3) / imadw.value,
                "note": "algorithmic bytes 64n; integer work 136 wide multiply-adds per butterfly (SURVEY 8(d))"}),
                flush=True)

    def msm_case(group, lg):
        n = 1 << lg
        ks = rand_fr(n, 100 + lg)
        fb = b.FixedBase.msm_g1 if group == 1 else b.FixedBase.msm_g2
        bases, inf = fb(ctx, ks)
        sc = rand_fr(n, 200 + lg)
        msm = b.VariableBaseMSM.msm_bigint_g1 if group == 1 else b.VariableBaseMSM.msm_bigint_g2
        msm(ctx, bases, sc, inf)
        L.b2z_profile_enable(ctx.handle, 1)
        read_profile(ctx)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            msm(ctx, bases, sc, inf)
        wall = (time.perf_counter() - t0) / args.reps
        prof = read_profile(ctx)
        L.b2z_profile_enable(ctx.handle, 0)
        acc = prof["msm_accum_g1" if group == 1 else "msm_accum_g2"]
        dev_ms = (prof["msm_sort"][0] + acc[0] + prof["msm_reduce"][0]) / args.reps
        adds = acc[2] / args.reps
        macs = adds * (3000.0 if group == 1 else 9000.0)
        print(json.dumps({
            "case": "msm_g%d" % group, "log_n": lg, "device_ms": dev_ms, "sort_ms": prof["msm_sort"][0] / args.reps,
            "accum_ms": acc[0] / args.reps, "reduce_ms": prof["msm_reduce"][0] / args.reps,
            "points_per_s_device": n / (dev_ms * 1e-3), "wall_ms_incl_pcie_and_base_upload": wall * 1e3,
            "mixed_adds": adds, "accum_int_frac": macs / (acc[0] / args.reps * 1e-3) / imadw.value,
            "whole_msm_int_frac": macs / (dev_ms * 1e-3) / imadw.value,
            "note": "generic (non-precomputed) bases as b2z_msm_g%d receives them" % group}), flush=True)

    for lg in [int(x) for x in args.msm.split(",") if x]:
        msm_case(1, lg)
    for lg in [int(x) for x in args.g2.split(",") if x]:
        msm_case(2, lg)


if __name__ == "__main__":
    main()
