#!/usr/bin/env python3
"""Benchmark of the Groth16 prove path (BASELINE.json metric: Groth16 prove ms & proofs/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2]

One "step" = one full proof of the workload circuit: witness map (7 NTTs) + 4 G1 MSMs +
1 G2 MSM + combine + serialization, from the constraint-row evaluations / assignment to
the 192 proof bytes.  Synthesis, key generation and key upload are outside the timed
region.

  value      proofs/s with inputs resident in HBM (b2z_groth16_prove_device)
  e2e        proofs/s through b2z_groth16_prove with pinned HOST buffers: H2D of a, b, c, z
             and D2H of the proof inside the timed region
  roofline   the dominant kernel (G1 bucket accumulation), timed live with CUDA events
             on its own stream by the library's phase timers
  cpu_baseline  oracle/cpu (arkworks-algorithm restatement) on the host cores, N=1 only

N > 1: one process per GPU (torchrun), every rank proves its own proof of the same circuit
(independent proofs batch one per GPU, no data-path collective) -> "scaling": "weak".
`--impl reference` times the CPU restatement (the reference itself is Rust on un-vendored
crates and cannot be built here); rank 0 only.
"""
import argparse
import ctypes
import importlib
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "zksnark-finalproject_b200"
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEED = 0xB2000004

WORKLOADS = {
    # name: (description, builder)
    "c1": "Fibonacci n=1000 (BASELINE configs[0]): domain 2^10, 5 variables",
    "c2": "matrix-multiplication 16x16 with Poseidon-shaped hashes (BASELINE configs[1]): "
          "109955 constraints, domain 2^17",
    "m8": "matrix-multiplication 8x8 (development size): domain 2^15",
    "c3": "prime-SNARK shape (BASELINE configs[2]): 7 SHA-256-sized Boolean blocks + 3 Fermat modpows, NUM_BITS=20",
    "c5": "matrix-multiplication 64x64 with Poseidon-shaped hashes (BASELINE configs[4] shape): "
          "2152451 constraints, domain 2^22",
}


class Instance:
    """What the prover consumes: constraint matrices (CSR) + assignment."""

    def __init__(self, cm, z):
        self.cm, self.z = cm, z
        self.num_constraints, self.num_instance = cm.num_constraints, cm.num_instance_variables
        self.num_variables, self.domain_size = cm.num_variables, cm.domain_size


def build_instance(name):
    pkg = importlib.import_module(PKG)
    if name == "c1":
        inst = importlib.import_module(PKG + ".circuits").fibonacci_circuit(0, 1, 1000)
        cm = pkg.ConstraintMatrices.from_rows(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
        return Instance(cm, inst.z)
    if name == "c3":
        inst = importlib.import_module(PKG + ".circuits").prime_circuit(5, num_bits=20, k_bases=3, sha_blocks=7)
        cm = pkg.ConstraintMatrices.from_rows(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
        return Instance(cm, inst.z)
    n = {"c2": 16, "m8": 8, "c5": 64}[name]
    ones = [[1] * n for _ in range(n)]                 # bench/matrix.py:10-11 posts all-ones matrices
    cm, z = importlib.import_module(PKG + ".circuits_fast").matrix_circuit_fast(ones, ones)
    return Instance(cm, z)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def toxic_waste():
    rnd = random.Random(SEED)
    return [rnd.randrange(1, R_MOD) for _ in range(5)]


def cpu_prove_setup(pkg, inst, pk):
    from oracle import cpu_oracle
    cpk = cpu_oracle.CpuProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                                   pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                                   pk.beta_g2, pk.delta_g2)
    return cpu_oracle, cpk


def run_reference(args, rank, world):
    """CPU arm: oracle/cpu (arkworks-algorithm restatement) with all host threads."""
    if rank != 0:
        return
    pkg = importlib.import_module(PKG)
    codec = pkg.codec
    inst = build_instance(args.workload)
    # the key must be valid for the circuit; group elements are made on the GPU when there is one,
    # else (no GPU on this host) by the Python oracle -- either way outside the timed region
    try:
        ctx = pkg.Context(0)
        pk, _ = pkg.Groth16.generate_parameters_with_qap(ctx, inst.cm, inst.num_constraints, inst.num_instance,
                                                         inst.num_variables, *toxic_waste())
        a, b, c = pkg.LibsnarkReduction.constraint_evaluations_device(ctx, inst.cm, codec.fr_to_mont_limbs(inst.z))
        ctx.close()
    except Exception:
        from oracle import groth16 as OG
        ra, rb, rc = inst.cm.rows()
        a, b, c = pkg.LibsnarkReduction.constraint_evaluations((ra, rb, rc), inst.num_instance, inst.num_constraints,
                                                               inst.z)
        opk = OG.setup(OG.R1CS(inst.num_instance, inst.num_variables - inst.num_instance, ra, rb, rc),
                       toxic=toxic_waste())
        q1, q2 = codec.g1_to_limbs, codec.g2_to_limbs
        pk = pkg.ProvingKey(opk.num_variables, opk.num_instance, opk.domain_size, q1(opk.a_query), q1(opk.b_g1_query),
                            q2(opk.b_g2_query), q1(opk.h_query), q1(opk.l_query), q1([opk.alpha_g1])[0][0],
                            q1([opk.beta_g1])[0][0], q1([opk.delta_g1])[0][0], q2([opk.beta_g2])[0][0],
                            q2([opk.delta_g2])[0][0])
    cpu_oracle, cpk = cpu_prove_setup(pkg, inst, pk)
    cores = cpu_oracle.hardware_threads()
    cpu_oracle.set_threads(cores)
    z = codec.fr_to_mont_limbs(inst.z)
    rnd = random.Random(SEED ^ 1)
    rs = codec.fr_to_mont_limbs([rnd.randrange(R_MOD), rnd.randrange(R_MOD)])
    for _ in range(args.warmup):
        cpk.prove(a, b, c, z, rs[0], rs[1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpk.prove(a, b, c, z, rs[0], rs[1])
    dt = (time.perf_counter() - t0) / args.steps
    val = 1.0 / dt
    line = {
        "impl": "reference", "metric": "groth16_proofs_per_sec", "value": val, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (Fr 255-bit / Fq 381-bit integers)",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "num_constraints": inst.num_constraints,
                   "domain": inst.domain_size, "num_variables": inst.num_variables},
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": "%d full proofs of the same workload; arkworks-algorithm C++ restatement "
                                   "(the Rust reference cannot be built here)" % args.steps},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter)
    was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=2, help="proofs timed for cpu_baseline (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3                      # bounded sample: a CPU proof takes seconds
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module(PKG)
    codec = pkg.codec
    ctx = pkg.Context(local_rank)
    L = ctx._lib

    # ---- workload (outside the timed region)
    inst = build_instance(args.workload)
    pk, vk = pkg.Groth16.generate_parameters_with_qap(ctx, inst.cm, inst.num_constraints, inst.num_instance,
                                                      inst.num_variables, *toxic_waste())
    pk.upload(ctx)
    inst.cm.upload(ctx)
    z = codec.fr_to_mont_limbs(inst.z)
    a, b, c = pkg.LibsnarkReduction.constraint_evaluations_device(ctx, inst.cm, z)
    n, m = inst.domain_size, inst.num_variables
    rnd = random.Random(SEED ^ 1 ^ (rank << 8))      # each rank proves with its own r, s
    r, s = rnd.randrange(R_MOD), rnd.randrange(R_MOD)
    rs = codec.fr_to_mont_limbs([r, s])
    p = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    proof = np.zeros(192, dtype=np.uint8)

    def as_torch(x):
        return torch.from_numpy(x.view(np.int64).copy())
    # pinned host copies for the end-to-end arm, device copies for the resident arm
    host = [as_torch(x).pin_memory() for x in (a, b, c, z)]
    dev0 = [t.cuda() for t in host]
    dev = [torch.empty_like(t) for t in dev0[:3]] + [dev0[3]]
    hp = lambda t: ctypes.c_void_p(t.data_ptr())

    def step_resident():
        for d, s0 in zip(dev[:3], dev0[:3]):
            d.copy_(s0, non_blocking=True)             # the witness map clobbers a, b, c
        st = L.b2z_groth16_prove_device(ctx.handle, pk._handle, hp(dev[0]), hp(dev[1]), hp(dev[2]), hp(dev[3]),
                                        p(rs[0:1]), p(rs[1:2]), p(proof))
        ctx.check(st)

    def step_e2e():
        # the call a user of the reference makes: assignment in host memory -> 192 proof bytes; the
        # constraint rows are evaluated on the GPU from the uploaded matrices (b2z_groth16_prove_r1cs)
        st = L.b2z_groth16_prove_r1cs(ctx.handle, pk._handle, inst.cm._handle, hp(host[3]), p(rs[0:1]), p(rs[1:2]),
                                      p(proof))
        ctx.check(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        dev_s = e0.elapsed_time(e1) * 1e-3
        t = torch.tensor([max(dev_s, 0.0), wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    # ---- correctness of what is being timed (not in the timed region)
    step_resident()
    first = proof.tobytes()
    step_e2e()
    assert proof.tobytes() == first, "resident and host-buffer paths disagree"

    for _ in range(args.warmup):
        step_resident()
    launches0 = L.b2z_kernel_launches(ctx.handle)
    print("[bench] library kernels launched before the timed region: %d" % launches0, file=sys.stderr, flush=True)
    L.b2z_profile_enable(ctx.handle, 1)
    clocks = ClockSampler(local_rank)
    clocks.start()
    dev_s, wall_s = timed(step_resident, args.steps)
    clk = clocks.stop()
    launches = L.b2z_kernel_launches(ctx.handle) - launches0
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_uint64 * 8)()
    units = (ctypes.c_uint64 * 8)()
    ctx.check(L.b2z_profile_read(ctx.handle, ms, cnt, units, 1))
    L.b2z_profile_enable(ctx.handle, 0)
    for _ in range(args.warmup):
        step_e2e()
    e2e_dev_s, e2e_wall_s = timed(step_e2e, args.steps)

    # ---- point-sharded mode (N > 1): ONE proof computed by all ranks -- reported next to
    # the weak-scaling line; every rank holds 1/N of every base set, partial sums travel over NCCL
    sharded = None
    if world > 1:
        spk = pkg.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                             pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                             pk.beta_g2, pk.delta_g2).upload(ctx, rank=rank, world=world)
        r0, s0 = 0x1234567, 0x7654321             # the same r, s on every rank
        rs0 = codec.fr_to_mont_limbs([r0, s0])
        part = np.zeros(pkg._ffi.PARTIAL_BYTES, dtype=np.uint8)
        gathered = [torch.empty(pkg._ffi.PARTIAL_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
        out = {}

        def step_sharded():
            ctx.check(L.b2z_groth16_prove_partial_r1cs(ctx.handle, spk._handle, inst.cm._handle, hp(host[3]),
                                                       p(rs0[0:1]), p(rs0[1:2]), p(part)))
            dist.all_gather(gathered, torch.from_numpy(part).cuda())
            out["proof"] = pkg.Groth16.combine([bytes(g.cpu().numpy().tobytes()) for g in gathered])

        step_sharded()
        ref = np.zeros(192, dtype=np.uint8)
        ctx.check(L.b2z_groth16_prove_r1cs(ctx.handle, pk._handle, inst.cm._handle, hp(host[3]), p(rs0[0:1]),
                                           p(rs0[1:2]), p(ref)))
        assert out["proof"] == ref.tobytes(), "sharded proof differs from the single-GPU proof"
        for _ in range(args.warmup):
            step_sharded()
        sh_dev_s, sh_wall_s = timed(step_sharded, args.steps)
        sharded = {"ms_per_proof": sh_dev_s / args.steps * 1e3, "proofs_per_s": args.steps / sh_dev_s,
                   "scaling": "strong", "collective": "all_gather of %d B per rank (NCCL)" % pkg._ffi.PARTIAL_BYTES,
                   "bytes_equal_single_gpu": True}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    imad = ctypes.c_double()
    imadw = ctypes.c_double()
    ctx.check(L.b2z_measure_int_peak(ctx.handle, ctypes.byref(imad), ctypes.byref(imadw)))
    G1ACC = 3
    acc_ms = ms[G1ACC] / max(1, cnt[G1ACC])
    acc_adds = units[G1ACC] / max(1, cnt[G1ACC])
    # algorithmic bytes of one G1 accumulation launch: one 96 B affine base + one 4 B reference per mixed addition
    acc_bytes = acc_adds * (96 + 4)
    # algorithmic integer work: XYZZ mixed addition = 10 Fq products = 10 * 300 32x32+64 multiply-adds
    acc_macs = acc_adds * 3000.0
    roofline = {
        "kernel": "msm_accum_kernel<G1> (bucket accumulation, XYZZ mixed additions)",
        "bound": "hbm", "achieved": acc_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms else None, "peak": hbm_peak,
        "unit": "GB/s", "frac": (acc_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak) if acc_ms else None,
        # dram__bytes_read+write of one G1 accumulation launch from the committed ncu --set full capture
        # (profiles/r01_ncu/prof_accum_g1.raw.csv: 245.8 MB read + 8.6 MB written by the 1.63 M-addition launch of
        # the A query = 156 B per mixed addition, 1.56x the algorithmic 100 B), scaled to this run's average launch
        "traffic": acc_adds * 156.0,
        "peak_source": hbm_src,
        "note": "this kernel is integer-pipe bound, not HBM bound (SURVEY.md App. C): see int_pipe",
        "int_pipe": {
            "achieved": acc_macs / (acc_ms * 1e-3) / 1e12 if acc_ms else None, "peak": imadw.value / 1e12,
            "unit": "T multiply-add/s (32x32+64, carry-chained IMAD.WIDE)",
            "frac": (acc_macs / (acc_ms * 1e-3) / imadw.value) if acc_ms and imadw.value else None,
            "peak_source": "b2z_measure_int_peak on this GPU, same run", "imad_32_peak": imad.value / 1e12},
        "launch_ms": acc_ms, "mixed_adds_per_launch": acc_adds, "launches_timed": int(cnt[G1ACC]),
    }
    phase_names = ["ntt_pass", "wm_pointwise", "msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce", "r1cs_eval"]
    phases = {nm: {"ms_per_step": ms[i] / args.steps, "launches_per_step": cnt[i] / args.steps,
                   "units_per_step": units[i] / args.steps} for i, nm in enumerate(phase_names)}
    ntt_el = units[0] / args.steps
    phases["ntt_pass"]["hbm_gbs_algorithmic"] = (ntt_el * 64 / (ms[0] / args.steps * 1e-3) / 1e9) if ms[0] else None

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and args.cpu_steps > 0:
        cpu_oracle, cpk = cpu_prove_setup(pkg, inst, pk)
        cores = cpu_oracle.hardware_threads()
        cpu_oracle.set_threads(cores)
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            cpu_proof = cpk.prove(a, b, c, z, rs[0], rs[1])
        cdt = (time.perf_counter() - t0) / args.cpu_steps
        assert cpu_proof == first, "GPU proof bytes differ from the CPU oracle"
        cpu = {"value": 1.0 / cdt, "unit": "proofs/s", "cores": cores, "kind": "port",
               "sample": "%d full proofs of the same workload (%.2f s each); proof bytes equal the GPU's"
                         % (args.cpu_steps, cdt)}

    def copies(cnt_pts):
        c_bits = L.b2z_host_msm_window_bits(cnt_pts, 1)
        w = (255 + c_bits - 1) // c_bits
        if 255 - c_bits * (w - 1) > c_bits - 1:
            w += 1
        return w
    key_bytes = (3 * (m + 2) * 96 * copies(m + 2) + (m + 2) * 192 * copies(m + 2) + n * 96 * copies(n))

    per_step = dev_s / args.steps
    value = world / per_step
    e2e_val = world / (e2e_dev_s / args.steps)
    line = {
        "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (Fr 255-bit / Fq 381-bit integers)", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "num_constraints": inst.num_constraints, "domain": n,
                   "num_variables": m, "parallelism": "one independent proof per GPU" if world > 1 else "1 GPU",
                   "cache_policy": "inputs larger than L2: every step streams %.0f MB of key bases "
                                   "(> 126 MB L2) plus %.1f MB of a/b/c/z; no explicit flush" %
                                   (key_bytes / 1e6, (3 * n + m) * 32 / 1e6)},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": "proofs/s", "ms_per_step": e2e_dev_s / args.steps * 1e3,
                "h2d_bytes_per_step": int(m * 32 + 64), "d2h_bytes_per_step": 1344,
                "call": "b2z_groth16_prove_r1cs (row evaluation + witness map + 4 MSMs + host epilogue)"},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall_s / args.steps * 1e3,
        "roofline": roofline, "phases": phases, "cpu_baseline": cpu, "sharded_single_proof": sharded,
        "proof_sha": __import__("hashlib").sha256(first).hexdigest()[:16],
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
