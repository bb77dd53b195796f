#!/usr/bin/env python3
"""bench.py -- headline measurement of the G1 MSM hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 20] [--curve bls12381]

A "step" is one complete MSM (g1m_multiexpAffine semantics) over one batch of synthetic input:
  N = 1 : 2^log2n BLS12-381 G1 points (default 2^20 -- the size BASELINE.json's target is quoted on).
  N > 1 : one MSM of N * 2^log2n points, sharded by point range (weak scaling): rank g owns slice g, computes a partial
          G1 point, the N partials are exchanged with one NCCL all_gather (N * 144 B) and summed on every rank.
`value`  = points of the whole job / second with bases and scalars already resident in HBM (device pointers through the
           C ABI, result left on the device);  `e2e` = the same through b200msm_g1_multiexp_affine with pinned HOST
           buffers for bases, scalars and result (H2D and D2H inside the timed region).
`--impl reference` times the reference's own WASM MSM (compiled natively, oracle/_ref) on the host cores, on the SAME workload
(one MSM of 2^log2n points per step, one WASM instance per host core; jobs above 2^20 points are sampled at 2^20 points per step and say so).
With N > 1 the line also carries `strong_2p24` (BASELINE config 4: one MSM of 2^24 points on 1 and on N GPUs, both results checked) and
`abi_multi` (the same workload through ONE multi-device context, b200msm_create_multi, driven by rank 0).
Prints ONE JSON line (rank 0).
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200")):
    if p not in sys.path: sys.path.insert(0, p)

METRIC = "bls12381_g1_msm_points_per_s"
SEED = 0xB2000000
LIMB_PRODUCTS_PER_FQMUL = {"bls12381": 300, "bn128": 136,      # 2*n32^2 + n32 (build_f1m.js:575-660)
                           "bls12381_g2": 900, "bn128_g2": 408}   # G2: one Fq2 multiplication = 3 Fq multiplications (f2m_mul, build_f2m.js:152-194)
CURVE_ID = {"bls12381": 0, "bn128": 1, "bls12381_g2": 2, "bn128_g2": 3}
CURVE_LABEL = {"bls12381": "BLS12-381 G1", "bn128": "BN254 G1", "bls12381_g2": "BLS12-381 G2", "bn128_g2": "BN254 G2"}


def metric_name(cname):
    return METRIC if cname == "bls12381" else {"bn128": "bn254_g1", "bls12381_g2": "bls12381_g2", "bn128_g2": "bn254_g2"}[cname] + "_msm_points_per_s"
FQMUL_PER_AFFINE_ADD = 6                                       # build_multiexp_opt.js:1207-1233 + build_batchinverse.js


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--log2n", type=int, default=0, help="points per GPU = 2^log2n (default 20; 18 for --workload batched)")
    ap.add_argument("--log2n-total", type=int, default=0, help="strong scaling: total points 2^K split over the GPUs (overrides --log2n)")
    ap.add_argument("--curve", default="bls12381", choices=["bls12381", "bn128", "bls12381_g2", "bn128_g2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-log2n", type=int, default=0, help="points per reference step (0 = the workload itself, bounded at 2^20 points per step)")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="wall-clock cap of the reference arm's timed run (steps are clamped to fit, and the clamp is reported)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling block (one MSM of 2^--strong-log2n points on 1 and on N GPUs)")
    ap.add_argument("--strong-log2n", type=int, default=24)
    ap.add_argument("--no-abi-multi", action="store_true", help="N > 1: skip the measurement of the same workload through ONE multi-device context (b200msm_create_multi) on rank 0")
    ap.add_argument("--workload", default="single", choices=["single", "batched", "ntt"],
                    help="single: one MSM of 2^log2n points per GPU per step (default, the headline); batched: BASELINE config 5, --batch independent MSMs of 2^log2n points (default 2^18) spread over the GPUs per step")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-stream", action="store_true", help="skip the extra measurement of a stream of MSMs (b200msm_g1_multiexp_batch)")
    ap.add_argument("--no-window-table", action="store_true", help="skip the extra measurement with precomputed window tables")
    ap.add_argument("--table-window-bits", type=int, default=0, help="window width of the precomputed table (0 = auto)")
    a = ap.parse_args()
    if a.log2n == 0: a.log2n = {"batched": 18, "ntt": 24}.get(a.workload, 20)
    return a


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []; self.proc = None; self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc: return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try: self.proc.wait(timeout=2)
        except Exception: self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def _splitmix64(x):
    """the engine's base-point stream (b200msm_g1_generate_bases): P_i = splitmix64(seed + first + i) * G"""
    M = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    return x ^ (x >> 31)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p)); return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ workload description shared by both arms
def workload_config(cname, log2n_per_gpu, world, strong_total=0):
    """`config` of the JSON line: identical for `--impl ours` and `--impl reference` on the same flags"""
    n = (1 << strong_total) // world if strong_total else 1 << log2n_per_gpu
    total = (1 << strong_total) if strong_total else n * world
    n8 = {"bls12381": 48, "bn128": 32, "bls12381_g2": 96, "bn128_g2": 64}[cname]
    return {"workload": "%s MSM (g1m_multiexpAffine semantics), %d points per GPU, one MSM of %d points per step, uniform 256-bit scalars, bases P_i = k_i*G"
                        % (CURVE_LABEL[cname], n, total),
            "curve": cname, "log2n_per_gpu": log2n_per_gpu, "points_per_step": total,
            "cache": "no cache flush: the per-step working set (bases %d MiB + scalars %d MiB per GPU, plus sort/tree scratch > 1 GiB) exceeds the 126 MB L2 (and any host cache); 2 scalar sets alternate"
                     % (n * 2 * n8 >> 20, n * 32 >> 20)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_proc(conn, cname, seed, lo, cnt):
    """One host process = one instance of the reference's WASM module (as one ffjavascript worker) owning the point range [lo, lo + cnt):
    bases P_i = splitmix64(seed + i) * G generated once, two scalar sets; every command runs g1m_multiexpAffine on the slice."""
    try:
        import numpy as np
        import pyref, coracle, refwasm
        cv = pyref.CURVES[cname]
        bases = coracle.generate_bases(cv.cid, pyref.affine_to_bytes(cv, cv.G), seed, lo, cnt)
        sets = [np.random.default_rng([seed & 0xffffffff, lo, k]).integers(0, 256, size=cnt * 32, dtype=np.uint8).tobytes() for k in range(2)]
        pb = refwasm.RefModule(cname)
        conn.send(("ready", cnt))
        while True:
            cmd = conn.recv()
            if cmd is None: break
            conn.send(pb.msm_affine_raw(bases, sets[cmd % 2], 32, cnt))
    except Exception as ex:      # surfaced by the parent
        conn.send(("error", repr(ex)))


class ReferencePool:
    """n points over `procs` host processes, point-range slices, partial results added with the reference's own g1m_add"""
    def __init__(self, cname, n, procs, seed):
        import multiprocessing as mp
        import refwasm
        ctx = mp.get_context("fork")
        self.cname = cname; self.n = n; self.workers = []
        per = (n + procs - 1) // procs
        for k in range(procs):
            lo, hi = k * per, min(n, (k + 1) * per)
            if lo >= hi: break
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_proc, args=(b, cname, seed, lo, hi - lo), daemon=True); p.start()
            self.workers.append((p, a))
        for p, c in self.workers:
            msg = c.recv()
            if msg[0] != "ready": raise RuntimeError("reference worker failed: %r" % (msg,))
        self.pb = refwasm.RefModule(cname); self.n8 = self.pb.n8

    def step(self, i):
        for p, c in self.workers: c.send(i)
        parts = [c.recv() for p, c in self.workers]
        for part in parts:
            if isinstance(part, tuple): raise RuntimeError("reference worker failed: %r" % (part,))
        pb = self.pb; n8 = self.n8
        mark = pb.heap_mark()
        pacc = pb.alloc(3 * n8); pt = pb.alloc(3 * n8)
        pb.write(pacc, parts[0])
        for part in parts[1:]:
            pb.write(pt, part); pb.g1m_add(pacc, pt, pacc)
        out = pb.normalize_read(pacc)
        pb.heap_release(mark)
        return out

    def close(self):
        for p, c in self.workers:
            try: c.send(None)
            except Exception: pass
        for p, c in self.workers: p.join(timeout=5)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def time_reference_g2(cname, log2n, steps, warmup):
    """G2: the reference's own g2m_multiexpAffine on ONE host core (one WASM instance), 2^log2n points built from 256 distinct multiples of G2"""
    import random
    import refwasm
    base = cname[:-3]
    if not refwasm.available(base): raise RuntimeError("oracle/_ref not built")
    g2 = refwasm.RefG2(refwasm.RefModule(base)); n = 1 << log2n
    G = g2.generator_affine()
    distinct = b"".join(g2.times_scalar_affine(G, (0x9E3779B97F4A7C15 * (i + 1) & ((1 << 64) - 1)).to_bytes(8, "little")) for i in range(256))
    bases = (distinct * (n // 256 + 1))[: n * 2 * g2.e8]
    scalars = random.Random(log2n).getrandbits(256 * n).to_bytes(32 * n, "little")
    for _ in range(warmup): g2.msm_affine(bases, scalars, 32, n)
    t0 = time.perf_counter()
    for _ in range(steps): g2.msm_affine(bases, scalars, 32, n)
    dt = (time.perf_counter() - t0) / steps
    return {"pps": n / dt, "ms": dt * 1e3, "cores": 1, "kind": "reference", "steps": steps, "warmup": warmup, "n": n,
            "sample": "reference WASM (upstream wasmcurves g2m_multiexpAffine) AOT-compiled via C; one MSM of 2^%d points per step on one host core" % log2n}


def time_reference(cname, n, steps, warmup, budget_s, procs=0):
    """The reference's own MSM on the host cores: one MSM of n points per step, one WASM instance per core on n/cores points each
    (ffjavascript's worker sharding), steps clamped so that the run fits budget_s.  -> dict"""
    import pyref, coracle, refwasm
    if cname.endswith("_g2"): return time_reference_g2(cname, min(max(1, n.bit_length() - 1), 12), max(1, min(steps, 3)), min(warmup, 1))
    cv = pyref.CURVES[cname]
    cores = procs or host_cores()
    seed = SEED + 977
    if not refwasm.available(cname):      # the reference did not compile here: time the C port of its algorithm on one core
        bases = coracle.generate_bases(cv.cid, pyref.affine_to_bytes(cv, cv.G), seed, 0, n)
        import random
        scalars = random.Random(n).getrandbits(256 * n).to_bytes(32 * n, "little")
        t0 = time.perf_counter(); coracle.multiexp_affine(cv.cid, bases, scalars, 32, n); dt = time.perf_counter() - t0
        return {"pps": n / dt, "ms": dt * 1e3, "cores": 1, "kind": "port", "steps": 1, "warmup": 0, "n": n,
                "sample": "C port of the upstream algorithm (oracle/msm_oracle.c); one MSM of %d points on one host core" % n}
    pool = ReferencePool(cname, n, cores, seed)
    try:
        t0 = time.perf_counter(); first = pool.step(0); t1 = time.perf_counter() - t0          # first (warm-up) step also sizes the run
        w_done = 1
        steps_fit = max(1, int((budget_s - t1 * max(warmup, 1)) / max(t1, 1e-9)))
        k = max(1, min(steps, steps_fit))
        for i in range(1, warmup): pool.step(i); w_done += 1
        t0 = time.perf_counter()
        for i in range(k): res = pool.step(i)
        dt = (time.perf_counter() - t0) / k
        # the two runs of scalar set 0 agree (deterministic), and the result is a point: normalize_read returned canonical (x, y)
        if k >= 1 and pool.step(0) != first: raise RuntimeError("reference arm is not deterministic")
    finally:
        pool.close()
    return {"pps": n / dt, "ms": dt * 1e3, "cores": len(pool.workers), "kind": "reference", "steps": k, "warmup": w_done, "n": n,
            "steps_clamped": k < steps,
            "sample": ("reference WASM (upstream wasmcurves g1m_multiexpAffine) AOT-compiled via C; one MSM of %d points per step, one instance per host core on "
                       "%d points each (%d processes), partial results added with g1m_add" % (n, (n + len(pool.workers) - 1) // len(pool.workers), len(pool.workers)))}


def time_reference_single(cname, log2n=16):
    """ONE instance, the whole N, one host core -- the recipe of wasmcurves/benchmarks/multiexp.js:7-42"""
    import pyref, coracle, refwasm, random
    cv = pyref.CURVES[cname]; n = 1 << log2n
    bases = coracle.generate_bases(cv.cid, pyref.affine_to_bytes(cv, cv.G), SEED + log2n, 0, n)
    scalars = random.Random(log2n).getrandbits(256 * n).to_bytes(32 * n, "little")
    pb = refwasm.RefModule(cname)
    t0 = time.perf_counter(); pb.msm_affine_raw(bases, scalars, 32, n); dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "points/s", "cores": 1, "ms_per_sample": dt * 1e3,
            "sample": "one reference WASM instance, one MSM of the whole 2^%d points on one host core (wasmcurves/benchmarks/multiexp.js:7-42)" % log2n}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0: return
    cname = a.curve
    total = (1 << a.log2n_total) if a.log2n_total else (1 << a.log2n) * max(1, a.gpus)
    cap = 1 << (a.ref_log2n if a.ref_log2n else 20)
    n = min(total, cap)                                    # a bounded sample of the workload when the job is larger than 2^20 points
    r = time_reference(cname, n, a.steps, a.warmup, a.ref_budget_s)
    cfg = workload_config(cname, a.log2n, max(1, a.gpus), a.log2n_total)
    line = {"impl": "reference", "metric": metric_name(cname), "value": r["pps"], "unit": "points/s",
            "n_gpus": a.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "strong" if a.log2n_total else "weak",
            "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq)", "data": "synthetic",
            "config": cfg,
            "sample_points_per_step": r["n"], "sample_is_whole_workload": r["n"] == total, "steps_requested": a.steps, "steps_clamped_to_budget": bool(r.get("steps_clamped")),
            "cpu_baseline": {"value": r["pps"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["pps"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------ our arm
R_ORDER = {0: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 1: 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001}


def scalar_dot(seed, first, sc_dev, n, cid):
    """sum_i s_i * k_i mod r for this rank's points: k_i = splitmix64(seed + first + i) is the multiplier of base i (P_i = k_i * G), s_i the
    256-bit scalars in the device tensor.  Host big integers (numpy object arrays), chunked: the oracle-free known answer of the MSM."""
    import numpy as np
    r = R_ORDER[cid & 1]; total = 0; CH = 1 << 19

    def sm(x):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    with np.errstate(over="ignore"):
        for lo in range(0, n, CH):
            m = min(CH, n - lo)
            k = sm(np.uint64(seed) + np.arange(first + lo, first + lo + m, dtype=np.uint64)); k[k == 0] = 1
            w = sc_dev[lo * 32:(lo + m) * 32].cpu().numpy().view("<u8").reshape(m, 4).astype(object)
            sv = w[:, 0] + (w[:, 1] << 64) + (w[:, 2] << 128) + (w[:, 3] << 192)
            total = (total + int((sv * k.astype(object)).sum())) % r
    return total


def known_answer_point(eng, cid, gen_base_dev, k0, total):
    """canonical affine bytes of total * G, computed as (total / k0) * P_0 by a ONE-point MSM on the first base of the stream (P_0 = k0 * G)"""
    r = R_ORDER[cid & 1]; n8 = {0: 48, 1: 32, 2: 96, 3: 64}[cid]
    t = total * pow(k0, -1, r) % r
    return eng.normalize(cid, eng.multiexp_affine(cid, gen_base_dev[: 2 * n8], t.to_bytes(32, "little"), 32, 1))


def profile_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full capture of the CURRENT
    kernels (profiles/r2_ncu_traffic.json, written by tools/ncu_extract.py); None when no capture of this build exists"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        d = json.load(open(p)); e = d.get(kernel_key)
        return (float(e["dram_bytes_read"]) + float(e["dram_bytes_write"])) if e else None
    except Exception:
        return None


def kernel_roofline(eng, handle, scal, n, cid, cname, a, out_dev):
    """per-phase CUDA-event timings inside the engine (a separate single-lane stats loop over the same inputs, NOT the timed loop) -> (roofline object, averaged stats)"""
    NSETS = len(scal); n8 = {0: 48, 1: 32, 2: 96, 3: 64}[cid]
    agg = {}; reps = max(3, min(a.steps, 10))
    eng.multiexp_resident(handle, scal[0], 32, n, cid, out=out_dev, want_stats=True)      # untimed: grows the single-lane scratch
    for i in range(reps):
        _, st = eng.multiexp_resident(handle, scal[i % NSETS], 32, n, cid, out=out_dev, want_stats=True)
        for k, v in st.items(): agg[k] = agg.get(k, 0) + v
    st = {k: v / reps for k, v in agg.items()}
    imad = eng.probe_imad(); fq = eng.probe_fqmul(cid)
    hbm_peak, hbm_src = measured_peaks()
    lp = LIMB_PRODUCTS_PER_FQMUL[cname]
    adds = st["affine_adds"]
    # dominant kernel: k_tree_bwd of tree round 0 (the batch-affine backward pass): 5 of the 6 field multiplications of every addition of that round
    # (the sixth is the forward pass's running product, k_tree_fwd).  A single-launch form of the whole round was measured slower and is not shipped.
    mults = 5
    r0_ms = st["ms_k_tree_bwd_round0"]; adds0 = st["affine_adds_round0"]
    alg_lp = adds0 * mults * lp                               # algorithmic limb products of that launch
    achieved = alg_lp / (r0_ms * 1e-3) if r0_ms > 0 else 0.0
    # algorithmic HBM bytes of that launch per addition: forward pass 2 x-coordinates + prefix product written; backward pass 2 points + prefix + result
    bytes_per_add = 2 * 2 * n8 + n8 + 8 + 2 * n8           # 2 operand points + prefix product + operand record read, 1 point written
    kname = "k_tree_bwd<FIRST=1>"
    traffic = profile_traffic("%s|%s|2^%d" % (kname, cname, a.log2n))
    roof = {"bound": "imad", "kernel": kname + " (batch-affine backward pass, tree round 0; one launch per window group per step)",
            "achieved": achieved / 1e12, "peak": imad / 1e12, "unit": "T limb-products/s (32x32+64 IMAD.WIDE.U32)",
            "frac": (achieved / imad) if imad else None, "traffic": traffic,
            "avg_launch_ms": r0_ms, "algorithmic_units_per_launch": alg_lp, "additions_per_launch": adds0, "field_multiplications_per_addition": mults,
            "measured_in": "a separate single-lane stats loop of the engine (CUDA events around every kernel group on the launching stream), same inputs, not the timed loop",
            "peak_source": "measured in this run by b200msm_probe_imad: register-resident IMAD.WIDE.U32 carry chains on all SMs (the instruction the field multiplier is made of); plain 32-bit IMAD runs at twice this rate",
            "hbm": {"achieved": adds0 * bytes_per_add / (r0_ms * 1e-3) / 1e9 if r0_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (adds0 * bytes_per_add / (r0_ms * 1e-3) / 1e9 / hbm_peak) if r0_ms > 0 else None,
                    "algorithmic_bytes_per_launch": adds0 * bytes_per_add, "peak_source": hbm_src},
            "all_rounds": {"kernel_group": "k_tree_bwd, all rounds", "limb_products": adds * mults * lp, "ms": st["ms_k_tree_bwd"],
                           "frac_of_imad_peak": (adds * mults * lp / (st["ms_k_tree_bwd"] * 1e-3) / imad) if imad and st["ms_k_tree_bwd"] > 0 else None},
            "whole_accumulate": {"limb_products": adds * FQMUL_PER_AFFINE_ADD * lp, "ms": st["ms_accumulate"],
                                 "frac_of_imad_peak": (adds * FQMUL_PER_AFFINE_ADD * lp / (st["ms_accumulate"] * 1e-3) / imad) if imad and st["ms_accumulate"] > 0 else None},
            "fqmul_per_s_measured": fq, "fqmul_frac_of_imad_peak": fq * lp / imad if imad else None}
    return roof, st


class Timer:
    """K steps between two CUDA events on the launching stream, bracketed by barrier + synchronize, max over ranks"""
    def __init__(self, torch, dist, dev, stream, world):
        self.torch, self.dist, self.dev, self.stream, self.world = torch, dist, dev, stream, world
        self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)

    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1: self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def run(self, fn, steps, warmup, wall=False):
        for i in range(warmup): fn(i)
        self.sync_all()
        t0 = time.perf_counter(); self.e0.record(self.stream)
        for i in range(steps): fn(i)
        self.e1.record(self.stream)
        self.sync_all()
        ms = self.e0.elapsed_time(self.e1)
        if wall: ms = max(ms, (time.perf_counter() - t0) * 1e3)
        t = self.torch.tensor([ms / steps], dtype=self.torch.float64, device=self.dev)
        if self.world > 1: self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def run_ours(a):
    import torch
    import torch.distributed as dist
    import b200msm
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))      # a mismatched collective fails in minutes, not after the default 10
        cpu_group = dist.new_group(backend="gloo")            # host-side barriers / object exchange (an NCCL barrier would spin on the GPUs while rank 0 measures alone)
    cname = a.curve; cid = CURVE_ID[cname]; n8 = b200msm.N8[cid]
    strong = a.log2n_total > 0
    from b200msm.sharded import shard_range
    if strong:
        lo_, hi_ = shard_range(1 << a.log2n_total, rank, world); n = hi_ - lo_; first_pt = lo_
    else:
        n = 1 << a.log2n; first_pt = rank * n
    total_points = (1 << a.log2n_total) if strong else n * world
    seed = SEED + a.log2n
    eng = b200msm.Engine(local)
    if world * 8 > (os.cpu_count() or 1): eng.set_option("batch_blocking", 1)      # (stream_of_msms block) the ranks' worker threads outnumber the host cores
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    T = Timer(torch, dist, dev, stream, world)
    warm = max(3, a.warmup)

    # ---- synthetic inputs, generated on the device: bases P_i = k_i * G (global index range of this rank), uniform 256-bit scalars
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
    eng.generate_bases(cid, seed, first_pt, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    NSETS = 2
    scal = [torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g) for _ in range(NSETS)]
    handle = eng.upload_bases(cid, bases, n)        # device -> device copy: the engine's own resident copy
    out_dev = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world * 3 * n8, dtype=torch.uint8, device=dev) if world > 1 else None
    total_dev = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)

    def combine(src_dev, dst):
        """N > 1: one all_gather of the N partial points (N * 3*n8 bytes over NCCL / NVLink), summed on every rank (b200msm_g1_sum)"""
        dist.all_gather_into_tensor(gathered, src_dev)
        rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, dst.data_ptr())
        assert rc == 0, rc

    def step_device(i):
        eng.multiexp_resident(handle, scal[i % NSETS], 32, n, cid, out=out_dev)
        if world > 1: combine(out_dev, total_dev)

    # ---- correctness of the EXACT timed path on the FULL workload, oracle-free (the oracle is test infrastructure; in this file only the cpu_baseline leg
    # touches it): bases are P_i = k_i*G with a known splitmix64 stream, so sum_i s_i*P_i = (sum_i s_i*k_i mod r)*G.  Every rank adds up its own
    # s_i*k_i with host integers, the N sums are exchanged (gloo), and rank 0 compares the all-gathered + summed device result of one step with a one-point MSM.
    step_device(0); T.sync_all()
    mine = scalar_dot(seed, first_pt, scal[0], n, cid)
    sums = [mine]
    if world > 1:
        sums = [None] * world; dist.all_gather_object(sums, mine, group=cpu_group)
    k0 = _splitmix64(seed) or 1
    checked = None
    if rank == 0:
        g0 = torch.empty(2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, seed, 0, 1, g0)
        expect = known_answer_point(eng, cid, g0, k0, sum(sums) % R_ORDER[cid & 1])
        got = eng.normalize(cid, total_dev if world > 1 else out_dev)
        assert got == expect and any(got), "GPU MSM (all ranks combined) fails the known-answer identity sum_i s_i*k_i*G on the full workload"
        checked = True

    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    launches0 = eng.counter("launches")
    ms = T.run(step_device, a.steps, warm)
    launches = (eng.counter("launches") - launches0) // max(1, a.steps + warm)

    # ---- same workload against resident bases WITH the precomputed window table (b200msm_upload_bases_windowed): the table is a
    # function of the fixed bases only, built once outside the timed region like the upload itself; reported beside `value`, never as it
    win = None
    if not a.no_window_table:
        t0 = time.perf_counter()
        hwin = eng.upload_bases_windowed(cid, bases, n, 32, a.table_window_bits)
        build_ms = (time.perf_counter() - t0) * 1e3
        outw = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)

        def step_win(i):
            eng.multiexp_resident(hwin, scal[i % NSETS], 32, n, cid, out=outw)
            if world > 1: combine(outw, total_dev)
        l0 = eng.counter("launches")
        wms = T.run(step_win, a.steps, warm)
        wl = (eng.counter("launches") - l0) // max(1, a.steps + warm)
        eng.multiexp_resident(hwin, scal[0], 32, n, cid, out=outw); eng.multiexp_resident(handle, scal[0], 32, n, cid, out=out_dev)
        T.sync_all()
        same = eng.normalize(cid, outw) == eng.normalize(cid, out_dev)
        assert same, "window-table result differs from the ordinary pipeline"
        _, wst = eng.multiexp_resident(hwin, scal[0], 32, n, cid, out=outw, want_stats=True)
        win = {"ms_per_step": wms, "value": total_points / (wms * 1e-3), "unit": "points/s",
               "window_bits": int(wst["window_bits"]), "windows": int(wst["windows"]), "affine_adds": int(wst["affine_adds"]),
               "table_bytes_per_gpu": int(wst["windows"]) * n * 2 * n8, "table_build_ms": build_ms,
               "gpu_launches": int(wl), "result_equals_ordinary_path": bool(same),
               "api": "b200msm_upload_bases_windowed + b200msm_g1_multiexp_resident (rows 2^(window offset) * P_i precomputed once per base set; all windows share one bucket array)"}
        eng.free_bases(hwin)

    # ---- a STREAM of independent MSMs of the same size over the same resident bases (b200msm_g1_multiexp_batch: 8 worker contexts, one lane each):
    # MSMs in different phases overlap better than the lanes of one MSM, so the per-MSM time of a stream is below the latency of a single MSM.
    # Reported beside `value` (which stays the latency-bound single-MSM loop), never as it.
    stream_blk = None
    if not a.no_stream and cid < 2 and not strong and a.log2n <= 21:
        cnt = 8 if a.log2n >= 18 else 32
        scb = torch.cat([scal[k % NSETS] for k in range(cnt)]); outb = torch.zeros(cnt * 3 * n8, dtype=torch.uint8, device=dev)

        def step_stream(i):
            eng.multiexp_batch(handle, scb, 32, n, cnt, cid, out=outb)
        sms = T.run(step_stream, max(2, a.steps // 4), 2)
        eng.multiexp_resident(handle, scal[1], 32, n, cid, out=out_dev); T.sync_all()
        same = eng.normalize(cid, outb[3 * n8: 6 * n8].clone()) == eng.normalize(cid, out_dev)
        assert same, "an MSM of the stream differs from the single-MSM result"
        stream_blk = {"msms_per_step": cnt, "ms_per_msm": sms / cnt, "value": total_points * cnt / (sms * 1e-3), "unit": "points/s", "results_equal_single_msm": bool(same),
                      "api": "b200msm_g1_multiexp_batch (count = %d MSMs of 2^%d points, scalars resident; 8 worker contexts x 1 lane)" % (cnt, a.log2n)}
        del scb, outb

    # ---- e2e: host buffers through the reference-facing entry point (H2D of bases + scalars, D2H of the result, every step)
    hb = torch.empty(n * 2 * n8, dtype=torch.uint8).pin_memory(); hb.copy_(bases)
    hs = [torch.empty(n * 32, dtype=torch.uint8).pin_memory() for _ in range(NSETS)]
    for k in range(NSETS): hs[k].copy_(scal[k])
    hout = torch.zeros(3 * n8, dtype=torch.uint8).pin_memory()

    def step_host(i):
        eng.multiexp_affine(cid, hb, hs[i % NSETS], 32, n, out=hout)          # synchronous: result is in host memory on return
        if world > 1:
            out_dev.copy_(hout, non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_dev)
            rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, hout.data_ptr())
            assert rc == 0, rc

    e2e_ms = T.run(step_host, max(3, min(a.steps, 10)), 2, wall=True)
    e2e_ok = None
    step_host(0)                      # every rank: the step holds a collective when N > 1
    if rank == 0:
        e2e_ok = eng.normalize(cid, bytes(hout.numpy())) == expect
        assert e2e_ok, "e2e (host buffers) result fails the known-answer identity"
    T.sync_all()
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1 only: (1) strong scaling on BASELINE config 4 (one MSM of 2^24 points on ONE GPU, then sharded over the N ranks), both results checked against
    # the known answer; (2) the weak workload through ONE multi-device context behind the C ABI (b200msm_create_multi) driven by rank 0 alone.
    strong_blk = None; abi_blk = None
    if world > 1 and not strong and cid < 2:
        if not a.no_strong: strong_blk = strong_scaling_block(a, torch, dist, b200msm, eng, T, cpu_group, rank, world, dev, cid, n8)
        if not a.no_abi_multi: abi_blk = abi_multi_block(a, torch, dist, b200msm, cpu_group, rank, world, cid, n8, hb, hs, n, seed, sums, k0, eng)

    # ---- kernel-level timings (per-phase CUDA events inside the engine) for the roofline, same inputs, rank 0 only
    line = None
    if rank == 0:
        roof, st = kernel_roofline(eng, handle, scal, n, cid, cname, a, out_dev)
        adds = st["affine_adds"]
        cfg = workload_config(cname, a.log2n, world, a.log2n_total)
        line = {"metric": metric_name(cname), "value": total_points / (ms * 1e-3), "unit": "points/s",
                "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq, 32x32->64 IMAD)", "data": "synthetic",
                "config": cfg,
                "plan": {"parallelism": "point-range shards x%d + all_gather of partials" % world if world > 1 else "single GPU", "inputs": "bases and scalars resident in HBM (device pointers through the C ABI), result left on the device",
                         "window_bits": int(st["window_bits"]), "windows": int(st["windows"]), "tree_rounds": int(round(st["tree_rounds"]))},
                "result_checked": bool(checked), "multi_gpu_result_checked": bool(checked) if world > 1 else None,
                "result_check": "one full step of the timed path (%d points over %d GPU(s), all-gathered and summed) equals (sum_i s_i*k_i mod r)*G from host integers" % (total_points, world),
                "clocks": clocks,
                "e2e": {"value": total_points / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": n * (2 * n8 + 32), "d2h_bytes_per_step": 3 * n8, "api": "b200msm_g1_multiexp_affine with pinned host buffers", "result_checked": bool(e2e_ok)},
                "gpu_launches": int(launches),
                "roofline": roof,
                "phases_ms": {k: round(v, 4) for k, v in st.items() if k.startswith("ms_")},
                "pairs": st["pairs"], "affine_adds": adds, "resident_window_table": win, "stream_of_msms": stream_blk}
        if strong_blk is not None: line["strong_2p%d" % a.strong_log2n] = strong_blk
        if abi_blk is not None: line["abi_multi"] = abi_blk
    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own code on the host cores, the SAME workload for one step, plus the single-instance row
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            r = time_reference(cname, min(total_points, 1 << 20), 1, 0, 60.0)
            line["cpu_baseline"] = {"value": r["pps"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"] + "; 1 step", "ms_per_sample": r["ms"]}
            if cid < 2: line["cpu_baseline"]["single_thread"] = time_reference_single(cname, 16)
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "skipped", "sample": "measured at N=1 only"}
    if rank == 0: _emit(line)
    eng.free_bases(handle)
    if world > 1: dist.barrier(); dist.destroy_process_group()


def strong_scaling_block(a, torch, dist, b200msm, eng, T, cpu_group, rank, world, dev, cid, n8):
    """BASELINE config 4 inside the driver's own --gpus N run: one MSM of 2^K points (K = --strong-log2n, default 24) timed on rank 0's GPU alone
    and then sharded by point range over all N ranks (NCCL all_gather + sum in the step); both results compared with the known answer."""
    K = a.strong_log2n; ntot = 1 << K; sseed = SEED + K
    from b200msm.sharded import shard_range
    lo, hi = shard_range(ntot, rank, world); m = hi - lo
    g = torch.Generator(device=dev); g.manual_seed(777)
    gathered = torch.zeros(world * 3 * n8, dtype=torch.uint8, device=dev); tot = torch.zeros(3 * n8, dtype=torch.uint8, device=dev); part = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    res = {"log2n_total": K}
    # (1) one GPU: rank 0 holds the whole problem; the other ranks wait on a host-side barrier
    full_sc = None; expect = None
    if rank == 0:
        fb = torch.empty(ntot * 2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, sseed, 0, ntot, fb)
        full_sc = torch.randint(0, 256, (ntot * 32,), dtype=torch.uint8, device=dev, generator=g)
        h = eng.upload_bases(cid, fb, ntot); del fb
        for i in range(2): eng.multiexp_resident(h, full_sc, 32, ntot, cid, out=part)
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); k1 = 5
        e0.record(T.stream)
        for i in range(k1): eng.multiexp_resident(h, full_sc, 32, ntot, cid, out=part)
        e1.record(T.stream); torch.cuda.synchronize(dev)
        res["ms_1gpu"] = e0.elapsed_time(e1) / k1
        g0 = torch.empty(2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, sseed, 0, 1, g0)
        expect = known_answer_point(eng, cid, g0, _splitmix64(sseed) or 1, scalar_dot(sseed, 0, full_sc, ntot, cid))
        res["result_1gpu_checked"] = eng.normalize(cid, part) == expect
        eng.free_bases(h)
    # every rank needs ITS slice of the same scalars: rank 0 broadcasts them over NCCL (outside any timed region)
    dist.barrier(group=cpu_group)
    if rank != 0: full_sc = torch.empty(ntot * 32, dtype=torch.uint8, device=dev)
    dist.broadcast(full_sc, src=0)
    my_sc = full_sc[lo * 32: hi * 32].clone(); del full_sc
    mb = torch.empty(m * 2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, sseed, lo, m, mb)
    h = eng.upload_bases(cid, mb, m); del mb

    def step(i):
        eng.multiexp_resident(h, my_sc, 32, m, cid, out=part)
        dist.all_gather_into_tensor(gathered, part)
        rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, tot.data_ptr())
        assert rc == 0, rc
    res["ms_Ngpu"] = T.run(step, 10, 3)
    eng.free_bases(h)
    if rank == 0:
        res["n_gpus"] = world; res["speedup"] = res["ms_1gpu"] / res["ms_Ngpu"]
        res["result_Ngpu_checked"] = eng.normalize(cid, tot) == expect
        res["points_per_s_Ngpu"] = ntot / (res["ms_Ngpu"] * 1e-3)
        assert res["result_1gpu_checked"] and res["result_Ngpu_checked"], "strong-scaling MSM fails the known-answer identity"
        return res
    return None


def abi_multi_block(a, torch, dist, b200msm, cpu_group, rank, world, cid, n8, hb, hs, n, seed, sums, k0, eng0):
    """The weak workload (N * 2^log2n points) through ONE context over all N GPUs (b200msm_create_multi): what a JS / C / Rust host that binds
    include/b200msm.h gets.  Rank 0 drives it alone (the other ranks idle on a host-side barrier; their GPUs keep their memory but run nothing):
    host buffers in, result to the host; then resident shards.  Checked against the same known answer as the torchrun path."""
    # rank 0 needs every rank's host inputs: gather them through gloo (outside any timed region)
    nb = n * 2 * n8
    allb = [torch.empty(nb, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    alls = [torch.empty(n * 32, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(hb.clone(), allb, dst=0, group=cpu_group); dist.gather(hs[0].clone(), alls, dst=0, group=cpu_group)
    res = None
    if rank == 0:
        tb = torch.cat(allb).pin_memory(); ts = torch.cat(alls).pin_memory(); del allb, alls
        tot = n * world
        me = b200msm.Engine(devices=list(range(world)))
        out = torch.zeros(3 * n8, dtype=torch.uint8).pin_memory()
        for i in range(3): me.multiexp_affine(cid, tb, ts, 32, tot, out=out)
        k = 10; t0 = time.perf_counter()
        for i in range(k): me.multiexp_affine(cid, tb, ts, 32, tot, out=out)
        e2e = (time.perf_counter() - t0) * 1e3 / k
        g0 = torch.empty(2 * n8, dtype=torch.uint8, device=torch.device("cuda", 0)); eng0.generate_bases(cid, seed, 0, 1, g0)
        expect = known_answer_point(eng0, cid, g0, k0, sum(sums) % R_ORDER[cid & 1])
        ok1 = eng0.normalize(cid, bytes(out.numpy())) == expect
        h = me.upload_bases(cid, tb, tot)
        for i in range(3): me.multiexp_resident(h, ts, 32, tot, cid, out=out)
        t0 = time.perf_counter()
        for i in range(k): me.multiexp_resident(h, ts, 32, tot, cid, out=out)
        resid = (time.perf_counter() - t0) * 1e3 / k
        ok2 = eng0.normalize(cid, bytes(out.numpy())) == expect
        me.free_bases(h); me.close()
        assert ok1 and ok2, "multi-device context result fails the known-answer identity"
        res = {"api": "b200msm_create_multi over %d devices, one host process; b200msm_g1_multiexp_affine (pinned host buffers) / b200msm_upload_bases + b200msm_g1_multiexp_resident (host scalars)" % world,
               "points_per_step": tot, "e2e_ms_per_step": e2e, "e2e_points_per_s": tot / (e2e * 1e-3), "h2d_bytes_per_step": tot * (2 * n8 + 32),
               "resident_bases_ms_per_step": resid, "resident_bases_points_per_s": tot / (resid * 1e-3), "h2d_bytes_per_step_resident": tot * 32,
               "result_checked": True, "timing": "wall clock around synchronous calls (results are in host memory on return), 10 steps after 3 warm-up"}
    dist.barrier(group=cpu_group)
    return res


BATCH_WORKERS = 8          # the engine's default for b200msm_g1_multiexp_batch (option "batch_workers"): one lane per worker


def run_batched(a):
    """BASELINE config 5: a.batch independent MSMs of 2^log2n points over the same bases, MSM j on rank j % world (replicas, no collective
    on the data path; one barrier brackets the step).  value = points of the whole batch / second, scalars resident in HBM;
    e2e = the same with every MSM's scalars in pinned host memory and results returned to the host."""
    import torch
    import torch.distributed as dist
    import b200msm
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available(): raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1: dist.init_process_group("nccl", device_id=dev)
    cname = a.curve; cid = CURVE_ID[cname]; n8 = b200msm.N8[cid]; n = 1 << a.log2n
    mine = len(range(rank, a.batch, world))                       # MSMs of this rank per step
    eng = b200msm.Engine(local); stream = torch.cuda.current_stream(dev); eng.set_stream(stream.cuda_stream)
    if world * BATCH_WORKERS > (os.cpu_count() or 1): eng.set_option("batch_blocking", 1)      # the ranks' worker threads outnumber the host cores: sleep in host waits instead of spinning
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, SEED + a.log2n, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(99 + rank)
    scal = torch.randint(0, 256, (max(1, mine) * n * 32,), dtype=torch.uint8, device=dev, generator=g)
    out_dev = torch.zeros(max(1, mine) * 3 * n8, dtype=torch.uint8, device=dev)
    handle = eng.upload_bases(cid, bases, n)
    hwin = None if a.no_window_table else eng.upload_bases_windowed(cid, bases, n, 32, a.table_window_bits)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warm):
        for _ in range(warm): fn()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps): fn()
        sync_all()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_dev(h=None):
        if mine: eng.multiexp_batch(h or handle, scal, 32, n, mine, cid, out=out_dev)      # returns when all results are in out_dev
    if rank == 0:       # correctness guard: first MSM of the batch vs a single call
        one = eng.multiexp_resident(handle, scal[: n * 32], 32, n, cid)
        step_dev(); torch.cuda.synchronize(dev)
        assert eng.normalize(cid, out_dev[: 3 * n8]) == eng.normalize(cid, one), "batched result differs from the single-MSM path"
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    ms = timed(step_dev, a.steps, max(3, a.warmup))
    wms = timed(lambda: step_dev(hwin), a.steps, max(3, a.warmup)) if hwin else None
    hs = torch.empty(max(1, mine) * n * 32, dtype=torch.uint8).pin_memory(); hs.copy_(scal)
    hout = torch.zeros(max(1, mine) * 3 * n8, dtype=torch.uint8).pin_memory()

    def step_host():
        if mine: eng.multiexp_batch(handle, hs, 32, n, mine, cid, out=hout)
    e2e_ms = timed(step_host, max(3, min(a.steps, 10)), 2)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        roof, st = kernel_roofline(eng, handle, [scal[: n * 32]], n, cid, cname, a, out_dev[: 3 * n8])
        total = a.batch * n
        line = {"metric": metric_name(cname), "value": total / (ms * 1e-3), "unit": "points/s",
                "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq, 32x32->64 IMAD)", "data": "synthetic",
                "config": {"workload": "batched: %d independent %s G1 MSMs of 2^%d points per step over one resident base set, MSM j on GPU j %% %d, %d worker contexts per GPU; uniform 256-bit scalars"
                                       % (a.batch, CURVE_LABEL[cname].replace(" G1", ""), a.log2n, world, BATCH_WORKERS),
                           "curve": cname, "log2n_per_msm": a.log2n, "batch": a.batch, "parallelism": "replicas (independent MSMs), no collective",
                           "window_bits": int(st["window_bits"]), "windows": int(st["windows"]),
                           "cache": "no L2 flush: each MSM's working set (bases %d MiB + sort/tree scratch) exceeds the 126 MB L2 and %d MSMs run concurrently" % (n * 2 * n8 >> 20, BATCH_WORKERS)},
                "msm_per_s": a.batch / (ms * 1e-3), "ms_per_msm": ms / a.batch,
                "clocks": clocks,
                "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": a.batch * n * 32, "d2h_bytes_per_step": a.batch * 3 * n8,
                        "api": "b200msm_g1_multiexp_batch with pinned host scalars / results, bases resident"},
                "gpu_launches": int(st["launches"]) * a.batch,
                "roofline": roof, "phases_ms_single_msm": {k: round(v, 4) for k, v in st.items() if k.startswith("ms_")},
                "resident_window_table": ({"ms_per_step": wms, "value": total / (wms * 1e-3), "unit": "points/s", "msm_per_s": a.batch / (wms * 1e-3)} if wms else None),
                "cpu_baseline": {"value": None, "unit": "points/s", "cores": 0, "kind": "skipped", "sample": "see the default (single) workload"}}
        if world == 1 and not a.no_cpu_baseline:
            try:
                r = time_reference(cname, 1 << min(a.log2n, 20), 1, 0, 60.0)
                line["cpu_baseline"] = {"value": r["pps"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"] + "; 1 step (one MSM of the batch)", "ms_per_sample": r["ms"]}
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        _emit(line)
    eng.free_bases(handle)
    if hwin: eng.free_bases(hwin)
    if world > 1: dist.barrier(); dist.destroy_process_group()


def run_ntt(a):
    """one Fr NTT (frm_fft) of 2^log2n elements per GPU per step; N > 1 runs independent replicas (the transform does
    not shard in this design: no collective).  value = elements transformed per second, data resident in HBM; e2e = pinned host in/out."""
    import torch
    import torch.distributed as dist
    import b200msm
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available(): raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1: dist.init_process_group("nccl", device_id=dev)
    cname = a.curve if a.curve in ("bls12381", "bn128") else "bls12381"; cid = CURVE_ID[cname]; lg = a.log2n; n = 1 << lg
    eng = b200msm.Engine(local); stream = torch.cuda.current_stream(dev); eng.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev); g.manual_seed(5 + rank)
    xs = []
    for _ in range(2):
        x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g); x[:, 31] &= 0x0F      # < 2^252 < r: reduced Montgomery elements
        xs.append(x.reshape(-1).contiguous())
    out = torch.empty_like(xs[0])

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)
    if rank == 0:   # correctness guard (not timed, oracle-free): ifft(fft(x)) == x at full size; byte parity with frm_fft is tests/test_gpu_parity.py
        eng.fr_fft(cid, xs[0], lg, out=out); eng.fr_fft(cid, out, lg, inverse=True, out=out); torch.cuda.synchronize(dev)
        assert torch.equal(out, xs[0]), "ifft(fft(x)) != x"
    for i in range(max(3, a.warmup)): eng.fr_fft(cid, xs[i % 2], lg, out=out)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    l0 = eng.counter("launches")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(a.steps): eng.fr_fft(cid, xs[i % 2], lg, out=out)
    e1.record(stream); sync_all()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (eng.counter("launches") - l0) // max(1, a.steps)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ph = {}
    for i in range(5):
        eng.fr_fft(cid, xs[i % 2], lg, out=out)
        for k, v in eng.fr_fft_last_phases().items(): ph[k] = ph.get(k, 0) + v / 5
    hin = torch.empty(n * 32, dtype=torch.uint8).pin_memory(); hin.copy_(xs[0]); hout = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    for _ in range(2): eng.fr_fft(cid, hin, lg, out=hout)
    sync_all(); t0 = time.perf_counter()
    k = max(3, min(a.steps, 10))
    for _ in range(k): eng.fr_fft(cid, hin, lg, out=hout)
    sync_all(); e2e_ms = (time.perf_counter() - t0) * 1e3 / k
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        passes = int(round(ph["radix4_passes"] + ph["radix2_passes"]))
        alg_bytes = n * 32 * 2                                     # one global pass reads and writes every element once
        per_pass_ms = ph["ms_passes"] / max(1, passes)
        roof = {"bound": "hbm", "kernel": "k_ntt_stage4 (two radix-2 stages per pass over the array; %d radix-4 + %d radix-2 passes per transform)" % (ph["radix4_passes"], ph["radix2_passes"]),
                "achieved": alg_bytes / (per_pass_ms * 1e-3) / 1e9 if per_pass_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                "frac": (alg_bytes / (per_pass_ms * 1e-3) / 1e9 / hbm_peak) if per_pass_ms > 0 else None,
                "traffic": profile_traffic("k_ntt_stage4|%s|2^%d" % ("bls12381_fr" if cid == 0 else "bn128_fr", lg)),      # from the committed ncu capture (profiles/r2_ncu_traffic.json), None when there is none for this size
                "avg_launch_ms": per_pass_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_src,
                "note": "each pass also performs n/2 (radix-2) or n (radix-4) Fr multiplications: %.1f G multiplications/s in the passes"
                        % ((ph["radix4_passes"] * n + ph["radix2_passes"] * n / 2) / (ph["ms_passes"] * 1e-3) / 1e9 if ph["ms_passes"] > 0 else 0.0)}
        line = {"metric": ("bls12381" if cid == 0 else "bn254") + "_fr_ntt_elements_per_s", "value": n * world / (ms * 1e-3), "unit": "elements/s",
                "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fr, 32x32->64 IMAD)", "data": "synthetic",
                "config": {"workload": "%s Fr NTT (frm_fft), 2^%d elements per GPU per step, forward transform, data resident in HBM" % ("BLS12-381" if cid == 0 else "BN254", lg),
                           "curve": cname, "log2n": lg, "parallelism": "replicas only (independent transforms)" if world > 1 else "single GPU",
                           "cache": "no L2 flush: the array (%d MiB) and its twiddle table (%d MiB) exceed the 126 MB L2; two input arrays alternate" % (n * 32 >> 20, n * 16 >> 20)},
                "clocks": clocks,
                "e2e": {"value": n * world / (e2e_ms * 1e-3), "unit": "elements/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
                        "api": "b200msm_fr_fft with pinned host buffers"},
                "gpu_launches": int(launches), "roofline": roof, "phases_ms": {k: round(v, 4) for k, v in ph.items()}}
        if world == 1 and not a.no_cpu_baseline:
            try:
                import refwasm, random
                pb = refwasm.RefModule(cname); m = 1 << 16
                hb = random.Random(1).getrandbits(252 * 1).to_bytes(32, "little") * m
                p = pb.alloc(len(hb) + 64); pb.write(p, hb)
                t0 = time.perf_counter(); reps = 3
                for _ in range(reps): pb.frm_fft(p, m)
                dt = (time.perf_counter() - t0) / reps
                line["cpu_baseline"] = {"value": m / dt, "unit": "elements/s", "cores": 1, "kind": "reference", "ms_per_sample": dt * 1e3,
                                        "sample": "reference WASM (wasmcurves frm_fft) AOT-compiled via C, one transform of 2^16 elements per step on one host core"}
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "skipped", "sample": "measured at N=1 only"}
        _emit(line)
    if world > 1: dist.barrier(); dist.destroy_process_group()


def _emit(line):
    """the ONE JSON line goes to the real stdout; everything else printed while running (NCCL's version banner, warnings) went to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1); os.dup2(2, 1)           # library chatter on fd 1 must not break the one-line contract
    args = parse()
    if args.impl == "reference": run_reference(args)
    elif args.workload == "batched": run_batched(args)
    elif args.workload == "ntt": run_ntt(args)
    else: run_ours(args)
