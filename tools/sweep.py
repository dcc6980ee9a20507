#!/usr/bin/env python3
"""BASELINE configs[3]: synthetic MSM (G1 2^16..2^26, G2 2^20) and Fr NTT (2^16..2^24) sweep with the library's
phase timers (CUDA events on the launching streams).  One JSON line per case; EVERY result is checked:
MSMs against the known-multiplier answer (bases k_i G made on the GPU, expected (sum k_i s_i) G computed in
Python integers), NTTs by the inverse round trip plus a delta-vector spot check, the witness map by the quotient
composition of the transform entry points with exact pointwise arithmetic.

    python tools/sweep.py [--ntt 16,18,20,22,24] [--msm 16,18,20,22,24,26] [--g2 20] [--mixes uniform,witness,...]
    torchrun --nproc-per-node N tools/sweep.py ...      # N GPUs: every MSM point-sharded over the ranks

Scalar mixes (SURVEY.md 8(d) / BASELINE.md 2): uniform | witness (30 % in {0,1}, 20 % < 2^16, 50 % uniform) |
equal | zero | max (all r - 1).  Fractions of roofline use BASELINE.md's normaliser: 6000 (G1) / 18000 (G2)
IMAD-equivalents per mixed addition, 272 per butterfly, over the measured 32-bit IMAD issue rate.
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
b = importlib.import_module("zksnark-finalproject_b200")
from oracle import bls12_381 as O          # noqa: E402  (checker only)
PH = ["ntt_pass", "wm_pointwise", "msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce", "r1cs_eval"]
R = O.R_MOD


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    a = (a[:, 0::2] | (a[:, 1::2] << np.uint64(32))).astype(np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return np.ascontiguousarray(a)


def mix(n, kind, seed):
    a = rand_fr(n, seed)
    if kind == "witness":
        u = np.random.RandomState(seed + 1).rand(n)
        small = u < 0.5
        a[small, 1:] = 0
        a[small, 0] &= np.uint64(0xffff)
        a[u < 0.3, 0] &= np.uint64(1)
    elif kind == "equal":
        a[:] = a[0]
    elif kind == "zero":
        a[:] = 0
    elif kind == "max":
        a[:] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint64)
    elif kind != "uniform":
        raise ValueError(kind)
    return a


def dot_mod_r(ks, sc):
    kb, sb = ks.tobytes(), sc.tobytes()
    acc = 0
    for i in range(0, len(kb), 32):
        acc += int.from_bytes(kb[i:i + 32], "little") * int.from_bytes(sb[i:i + 32], "little")
    return acc % R


def read_profile(ctx):
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_uint64 * 8)()
    units = (ctypes.c_uint64 * 8)()
    ctx.check(ctx._lib.b2z_profile_read(ctx.handle, ms, cnt, units, 1))
    return {PH[i]: (ms[i], int(cnt[i]), int(units[i])) for i in range(len(PH))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--wm", default="16,17,18,20,22")
    ap.add_argument("--msm", default="16,18,20,22")
    ap.add_argument("--g2", default="20")
    ap.add_argument("--mixes", default="uniform,witness")
    ap.add_argument("--adversarial", default="20", help="log sizes for the equal / zero / max mixes (G1)")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b.Context(local)
    L = ctx._lib
    codec = b.codec
    say = (lambda d: print(json.dumps(d), flush=True)) if rank == 0 else (lambda d: None)
    imad, imadw = ctypes.c_double(), ctypes.c_double()
    ctx.check(L.b2z_measure_int_peak(ctx.handle, ctypes.byref(imad), ctypes.byref(imadw)))
    hbm = 6650.0
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    say({"n_gpus": world, "int_peak_imad_T": imad.value / 1e12, "int_peak_wide_mac_T": imadw.value / 1e12, "hbm_gbs": hbm})

    # ------------------------------------------------------------------ NTT (single GPU; rank 0)
    if rank == 0:
        for lg in [int(x) for x in args.ntt.split(",") if x]:
            n = 1 << lg
            data = rand_fr(n, lg)
            dom = b.Radix2EvaluationDomain(ctx, n)
            cos = dom.get_coset(7)
            # checks: round trips and the transform of the delta at index 1 (powers of w)
            assert np.array_equal(dom.ifft(dom.fft(data)), data) and np.array_equal(cos.ifft(cos.fft(data)), data)
            delta = np.zeros((n, 4), dtype=np.uint64)
            delta[1] = codec.fr_to_mont_limbs([1])[0]
            w = pow(O.FR_ROOT_OF_UNITY, 1 << (32 - lg), R)
            assert codec.fr_from_mont_limbs(dom.fft(delta)[[0, 1, n // 2]]) == [1, w, R - 1]
            for kind, fn in (("forward", dom.fft), ("inverse", dom.ifft), ("coset_forward", cos.fft)):
                fn(data)
                L.b2z_profile_enable(ctx.handle, 1)
                read_profile(ctx)
                for _ in range(args.reps):
                    fn(data)
                prof = read_profile(ctx)
                L.b2z_profile_enable(ctx.handle, 0)
                ms = prof["ntt_pass"][0] / args.reps
                say({"case": "ntt", "kind": kind, "log_n": lg, "ms_passes": ms, "passes": prof["ntt_pass"][1] / args.reps,
                     "hbm_gbs_algorithmic": 64.0 * n / (ms * 1e-3) / 1e9, "hbm_frac": 64.0 * n / (ms * 1e-3) / 1e9 / hbm,
                     "int_frac": 272.0 * (n / 2 * lg) / (ms * 1e-3) / imad.value, "checked": "round trip + delta vector"})
        for lg in [int(x) for x in args.wm.split(",") if x]:
            n = 1 << lg
            a, bb, c = (rand_fr(n, 300 + lg + i) for i in range(3))
            wm = b.LibsnarkReduction.witness_map_from_evaluations
            h = wm(ctx, a, bb, c)
            # the same map composed from the (separately checked) transform entry points + Python pointwise arithmetic
            if lg <= 17:
                dom = b.Radix2EvaluationDomain(ctx, n)
                cos = dom.get_coset(7)
                ca, cb, cc = (codec.fr_from_mont_limbs(cos.fft(dom.ifft(v))) for v in (a, bb, c))
                zinv = pow((pow(7, n, R) - 1) % R, -1, R)
                u = codec.fr_to_mont_limbs([(x * y - w) % R * zinv % R for x, y, w in zip(ca, cb, cc)])
                assert np.array_equal(cos.ifft(u), h), "witness map differs from its composition out of b2z_ntt_fr"
                checked = "composition of b2z_ntt_fr transforms + exact pointwise arithmetic"
            else:
                checked = "tests/test_gpu_baseline_sizes.py (C++ oracle at 2^22)"
            L.b2z_profile_enable(ctx.handle, 1)
            read_profile(ctx)
            for _ in range(args.reps):
                wm(ctx, a, bb, c)
            prof = read_profile(ctx)
            L.b2z_profile_enable(ctx.handle, 0)
            ms = prof["ntt_pass"][0] / args.reps
            say({"case": "witness_map", "log_n": lg, "ms_kernels": ms, "launches": prof["ntt_pass"][1] / args.reps,
                 "hbm_gbs_algorithmic_128n": 128.0 * n / (ms * 1e-3) / 1e9,
                 "int_frac": 272.0 * (7 * n / 2 * lg) / (ms * 1e-3) / imad.value, "checked": checked})

    # ------------------------------------------------------------------ MSM (point-sharded over the ranks)
    def msm_case(group, lg, kind):
        n = 1 << lg
        lo, hi = n * rank // world, n * (rank + 1) // world
        ks = rand_fr(n, 100 + lg)[lo:hi]
        sc = mix(n, kind, 200 + lg)[lo:hi]
        curve = O.G1 if group == 1 else O.G2
        fb = b.FixedBase.msm_g1 if group == 1 else b.FixedBase.msm_g2
        bases, inf = fb(ctx, ks)
        msm = b.VariableBaseMSM.msm_bigint_g1 if group == 1 else b.VariableBaseMSM.msm_bigint_g2
        unpack = codec.g1_projective_from_limbs if group == 1 else codec.g2_projective_from_limbs
        out = msm(ctx, bases, sc, inf)
        want_k = dot_mod_r(ks, sc)
        got = curve.to_affine(unpack(out))
        assert got == (curve.mul(curve.gen, want_k) if want_k else None), "MSM result differs from (sum k_i s_i) G"
        if world > 1:
            dist.barrier()
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
        if world > 1:
            import torch
            t = torch.tensor([dev_ms, acc[0] / args.reps, wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = torch.tensor([adds], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot)
            dev_ms, acc_ms, wall_ms, adds = float(t[0]), float(t[1]), float(t[2]), float(tot[0])
        else:
            acc_ms, wall_ms = acc[0] / args.reps, wall * 1e3
        eq = adds * (6000.0 if group == 1 else 18000.0)
        say({"case": "msm_g%d" % group, "log_n": lg, "scalars": kind, "n_gpus": world, "device_ms": dev_ms,
             "sort_ms": prof["msm_sort"][0] / args.reps, "accum_ms": acc_ms, "reduce_ms": prof["msm_reduce"][0] / args.reps,
             "points_per_s_device": n / (dev_ms * 1e-3), "wall_ms_incl_pcie_and_base_upload": wall_ms, "mixed_adds": adds,
             "accum_int_frac": eq / (acc_ms * 1e-3) / imad.value / world if acc_ms else None,
             "whole_msm_int_frac": eq / (dev_ms * 1e-3) / imad.value / world if dev_ms else None,
             "checked": "known multipliers, every rank's partial sum",
             "note": "generic (non-precomputed) bases as b2z_msm_g%d receives them; max over ranks" % group})

    mixes = [m for m in args.mixes.split(",") if m]
    for lg in [int(x) for x in args.msm.split(",") if x]:
        for kind in mixes:
            msm_case(1, lg, kind)
    for lg in [int(x) for x in args.adversarial.split(",") if x]:
        for kind in ("equal", "zero", "max"):
            msm_case(1, lg, kind)
    for lg in [int(x) for x in args.g2.split(",") if x]:
        for kind in mixes:
            msm_case(2, lg, kind)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
