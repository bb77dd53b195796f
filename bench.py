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
`--impl reference` times the reference's own WASM MSM (compiled natively, oracle/_ref) on the host cores.
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
    ap.add_argument("--ref-log2n", type=int, default=16, help="points per reference step (bounded sample)")
    ap.add_argument("--workload", default="single", choices=["single", "batched", "ntt"],
                    help="single: one MSM of 2^log2n points per GPU per step (default, the headline); batched: BASELINE config 5, --batch independent MSMs of 2^log2n points (default 2^18) spread over the GPUs per step")
    ap.add_argument("--batch", type=int, default=64)
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


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    cname, bases, scalars, n = args
    import refwasm
    pb = refwasm.RefModule(cname)
    return pb.msm_affine_raw(bases, scalars, 32, n)


def reference_msm_parallel(pool, cname, bases, scalars, n, nproc, n8):
    """One MSM of n points on nproc host processes: point-range slices, each through the reference's own
    g1m_multiexpAffine (one WASM instance per worker, as ffjavascript's worker pool does), partials summed with g1m_add."""
    import refwasm
    per = (n + nproc - 1) // nproc
    jobs = []
    for k in range(nproc):
        lo, hi = k * per, min(n, (k + 1) * per)
        if lo >= hi: break
        jobs.append((cname, bases[lo * 2 * n8: hi * 2 * n8], scalars[lo * 32: hi * 32], hi - lo))
    parts = pool.map(_ref_worker, jobs)
    pb = reference_msm_parallel._pb.get(cname)
    if pb is None:
        pb = reference_msm_parallel._pb[cname] = refwasm.RefModule(cname)
    mark = pb.heap_mark()
    pacc = pb.alloc(3 * n8); pt = pb.alloc(3 * n8)
    pb.write(pacc, parts[0])
    for part in parts[1:]:
        pb.write(pt, part); pb.g1m_add(pacc, pt, pacc)
    out = pb.normalize_read(pacc)
    pb.heap_release(mark)
    return out


reference_msm_parallel._pb = {}


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
    return n / dt, dt * 1e3, 1, "reference", "reference WASM (upstream wasmcurves g2m_multiexpAffine) AOT-compiled via C; one MSM of 2^%d points per step on one host core" % log2n


def time_reference(cname, log2n, steps, warmup):
    """returns (points_per_s, ms_per_step, cores, kind, sample description)"""
    import multiprocessing as mp
    import pyref, coracle, refwasm
    if cname.endswith("_g2"): return time_reference_g2(cname, min(log2n, 12), steps, warmup)
    cv = pyref.CURVES[cname]
    n = 1 << log2n
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    bases = coracle.generate_bases(cv.cid, pyref.affine_to_bytes(cv, cv.G), SEED + log2n, 0, n)
    import random
    rnd = random.Random(log2n)
    scalars = rnd.getrandbits(256 * n).to_bytes(32 * n, "little")
    if refwasm.available(cname):
        kind = "reference"
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            for _ in range(warmup): reference_msm_parallel(pool, cname, bases, scalars, n, cores, cv.n8)
            t0 = time.perf_counter()
            for _ in range(steps): res = reference_msm_parallel(pool, cname, bases, scalars, n, cores, cv.n8)
            dt = (time.perf_counter() - t0) / steps
        check = coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, bases, scalars, 32, n))
        assert pyref.canonical_bytes(cv, res) == check, "reference arm result differs from the C oracle"
        what = "reference WASM (upstream wasmcurves g1m_multiexpAffine) AOT-compiled via C"
    else:
        kind = "port"; cores = 1
        for _ in range(warmup): coracle.multiexp_affine(cv.cid, bases, scalars, 32, n)
        t0 = time.perf_counter()
        for _ in range(steps): coracle.multiexp_affine(cv.cid, bases, scalars, 32, n)
        dt = (time.perf_counter() - t0) / steps
        what = "C port of the upstream algorithm (oracle/msm_oracle.c)"
    sample = "%s; one MSM of 2^%d points per step, point-range slices over %d host processes, partials summed with g1m_add" % (what, log2n, cores)
    return n / dt, dt * 1e3, cores, kind, sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0: return
    cname = a.curve
    steps = max(1, min(a.steps, 5)); warm = max(1, min(a.warmup, 1))
    pps, ms, cores, kind, sample = time_reference(cname, a.ref_log2n, steps, warm)
    line = {"impl": "reference", "metric": metric_name(cname), "value": pps, "unit": "points/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq)", "data": "synthetic",
            "config": {"workload": "%s MSM, bounded sample 2^%d points per step (of the 2^%d-point workload), uniform 256-bit scalars" % (cname, a.ref_log2n, a.log2n),
                       "curve": cname, "log2n_per_step": a.ref_log2n},
            "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def kernel_roofline(eng, handle, scal, n, cid, cname, a, out_dev):
    """per-phase CUDA-event timings inside the engine (single lane, same inputs) -> (roofline object, averaged stats)"""
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
    # dominant kernel: the first (largest) k_tree_bwd launch = the backward pass of tree round 0, which does 5 of the 6
    # field multiplications of every batch-affine addition of that round
    bwd0_ms = st["ms_k_tree_bwd_round0"]; adds0 = st["affine_adds_round0"]
    alg_lp = adds0 * 5 * lp                                   # algorithmic limb products of that launch
    achieved = alg_lp / (bwd0_ms * 1e-3) if bwd0_ms > 0 else 0.0
    # algorithmic HBM bytes of that launch per addition: 2 input points + prefix product + share of the thread inverse + output point
    bytes_per_add = 2 * 2 * n8 + n8 + n8 / 8 + 2 * n8
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture (profiles/), 2^20 BLS12-381 only
    traffic = 3.04e9 if (cname == "bls12381" and a.log2n == 20) else None      # profiles/r1b_ncu_full_tree_kernels_2p20_bls.csv: 2.236 GB read + 0.804 GB written
    roof = {"bound": "imad", "kernel": "k_tree_bwd<FIRST=1> (batch-affine backward pass, tree round 0: one launch per step)",
            "achieved": achieved / 1e12, "peak": imad / 1e12, "unit": "T limb-products/s (32x32+64 IMAD.WIDE.U32)",
            "frac": (achieved / imad) if imad else None, "traffic": traffic,
            "avg_launch_ms": bwd0_ms, "algorithmic_units_per_launch": alg_lp, "additions_per_launch": adds0,
            "peak_source": "measured in this run by b200msm_probe_imad: register-resident IMAD.WIDE.U32 carry chains on all SMs (the instruction the field multiplier is made of); plain 32-bit IMAD runs at twice this rate",
            "hbm": {"achieved": adds0 * bytes_per_add / (bwd0_ms * 1e-3) / 1e9 if bwd0_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (adds0 * bytes_per_add / (bwd0_ms * 1e-3) / 1e9 / hbm_peak) if bwd0_ms > 0 else None,
                    "algorithmic_bytes_per_launch": adds0 * bytes_per_add, "peak_source": hbm_src},
            "all_rounds": {"kernel_group": "k_tree_bwd, all rounds", "limb_products": adds * 5 * lp, "ms": st["ms_k_tree_bwd"],
                           "frac_of_imad_peak": (adds * 5 * lp / (st["ms_k_tree_bwd"] * 1e-3) / imad) if imad and st["ms_k_tree_bwd"] > 0 else None},
            "whole_accumulate": {"limb_products": adds * FQMUL_PER_AFFINE_ADD * lp, "ms": st["ms_accumulate"],
                                 "frac_of_imad_peak": (adds * FQMUL_PER_AFFINE_ADD * lp / (st["ms_accumulate"] * 1e-3) / imad) if imad and st["ms_accumulate"] > 0 else None},
            "fqmul_per_s_measured": fq, "fqmul_frac_of_imad_peak": fq * lp / imad if imad else None}
    return roof, st


def run_ours(a):
    import torch
    import torch.distributed as dist
    import b200msm
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cname = a.curve; cid = CURVE_ID[cname]; n8 = b200msm.N8[cid]
    strong = a.log2n_total > 0
    if strong:
        from b200msm.sharded import shard_range
        lo_, hi_ = shard_range(1 << a.log2n_total, rank, world); n = hi_ - lo_; first_pt = lo_
    else:
        n = 1 << a.log2n; first_pt = rank * n
    eng = b200msm.Engine(local)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)

    # ---- synthetic inputs, generated on the device: bases P_i = k_i * G (global index range of this rank), uniform 256-bit scalars
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
    eng.generate_bases(cid, SEED + a.log2n, first_pt, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    NSETS = 2
    scal = [torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g) for _ in range(NSETS)]
    handle = eng.upload_bases(cid, bases, n)        # device -> device copy: the engine's own resident copy
    out_dev = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world * 3 * n8, dtype=torch.uint8, device=dev) if world > 1 else None
    total_dev = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)

    def step_device(i):
        eng.multiexp_resident(handle, scal[i % NSETS], 32, n, cid, out=out_dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_dev)
            rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, total_dev.data_ptr())
            assert rc == 0, rc

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- correctness guard, oracle-free (the oracle is test infrastructure; in this file only the cpu_baseline leg touches it): the bases
    # are P_i = k_i*G with a known splitmix64 stream, so  sum_i s_i*P_i = (sum_i s_i*k_i mod r)*G = t*P_0  with  t = (sum_i s_i*k_i)*k_0^-1 mod r.
    # Left side: the full pipeline on a 2^12-point prefix; right side: a one-point MSM.  Not timed.
    if rank == 0:
        R_ORDER = {0: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 1: 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001}[cid & 1]
        m = 1 << 12
        hs = bytes(scal[0][: m * 32].cpu().numpy())
        ks = [(_splitmix64(SEED + a.log2n + first_pt + i) or 1) for i in range(m)]
        tot = sum(int.from_bytes(hs[32 * i: 32 * i + 32], "little") * ks[i] for i in range(m)) % R_ORDER
        t = tot * pow(ks[0], -1, R_ORDER) % R_ORDER
        lhs = eng.normalize(cid, eng.multiexp_affine(cid, bases[: m * 2 * n8], scal[0][: m * 32], 32, m))
        rhs = eng.normalize(cid, eng.multiexp_affine(cid, bases[: 2 * n8], t.to_bytes(32, "little"), 32, 1))
        assert lhs == rhs and any(lhs), "GPU MSM fails the known-answer identity sum_i s_i*k_i*G"

    for i in range(max(3, a.warmup)): step_device(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    launches0 = eng.counter("launches")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for i in range(a.steps): step_device(i)
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (eng.counter("launches") - launches0) // max(1, a.steps)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

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
            if world > 1:
                dist.all_gather_into_tensor(gathered, outw)
                rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, total_dev.data_ptr())
                assert rc == 0, rc
        for i in range(max(3, a.warmup)): step_win(i)
        eng.multiexp_resident(handle, scal[(max(3, a.warmup) - 1) % NSETS], 32, n, cid, out=out_dev)
        sync_all()
        same = eng.normalize(cid, outw) == eng.normalize(cid, out_dev)
        assert same, "window-table result differs from the ordinary pipeline"
        l0 = eng.counter("launches")
        e0.record(stream)
        for i in range(a.steps): step_win(i)
        e1.record(stream)
        sync_all()
        wms = e0.elapsed_time(e1) / a.steps
        t = torch.tensor([wms], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wms = float(t.item())
        _, wst = eng.multiexp_resident(hwin, scal[0], 32, n, cid, out=outw, want_stats=True)
        win = {"ms_per_step": wms, "value": ((1 << a.log2n_total) if strong else n * world) / (wms * 1e-3), "unit": "points/s",
               "window_bits": int(wst["window_bits"]), "windows": int(wst["windows"]), "affine_adds": int(wst["affine_adds"]),
               "table_bytes_per_gpu": int(wst["windows"]) * n * 2 * n8, "table_build_ms": build_ms,
               "gpu_launches": int((eng.counter("launches") - l0) // max(1, a.steps + 1)), "result_equals_ordinary_path": bool(same),
               "api": "b200msm_upload_bases_windowed + b200msm_g1_multiexp_resident (rows 2^(window offset) * P_i precomputed once per base set; all windows share one bucket array)"}
        eng.free_bases(hwin)

    # ---- e2e: host buffers through the reference-facing entry point (H2D of bases + scalars, D2H of the result, every step)
    hb = torch.empty(n * 2 * n8, dtype=torch.uint8).pin_memory(); hb.copy_(bases)
    hs = [torch.empty(n * 32, dtype=torch.uint8).pin_memory() for _ in range(NSETS)]
    for k in range(NSETS): hs[k].copy_(scal[k])
    hout = torch.zeros(3 * n8, dtype=torch.uint8).pin_memory()
    gout = [torch.zeros(3 * n8, dtype=torch.uint8) for _ in range(world)] if world > 1 else None

    def step_host(i):
        eng.multiexp_affine(cid, hb, hs[i % NSETS], 32, n, out=hout)          # synchronous: result is in host memory on return
        if world > 1:
            out_dev.copy_(hout, non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_dev)
            rc = b200msm.lib.b200msm_g1_sum(eng._ctx, cid, gathered.data_ptr(), world, hout.data_ptr())
            assert rc == 0, rc

    for i in range(2): step_host(i)
    sync_all()
    e2e_steps = max(3, min(a.steps, 10))
    t0 = time.perf_counter(); e0.record(stream)
    for i in range(e2e_steps): step_host(i)
    e1.record(stream)
    sync_all()
    e2e_ms = max((time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernel-level timings (per-phase CUDA events inside the engine) for the roofline, same inputs, rank 0 only
    line = None
    if rank == 0:
        roof, st = kernel_roofline(eng, handle, scal, n, cid, cname, a, out_dev)
        adds = st["affine_adds"]
        total_points = (1 << a.log2n_total) if strong else n * world
        line = {"metric": metric_name(cname), "value": total_points / (ms * 1e-3), "unit": "points/s",
                "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq, 32x32->64 IMAD)", "data": "synthetic",
                "config": {"workload": ("%s MSM, %d points per GPU (%d points per step), uniform 256-bit scalars, bases P_i = k_i*G resident in HBM"
                                        % (CURVE_LABEL[cname], n, total_points)),
                           "curve": cname, "log2n_per_gpu": a.log2n, "parallelism": "point-range shards x%d + all_gather of partials" % world if world > 1 else "single GPU",
                           "window_bits": int(st["window_bits"]), "windows": int(st["windows"]), "tree_rounds": int(round(st["tree_rounds"])),
                           "cache": "no L2 flush: per-step working set (bases %d MiB + scalars %d MiB + sort/tree scratch > 1 GiB) exceeds the 126 MB L2; %d scalar sets alternate"
                                    % (n * 2 * n8 >> 20, n * 32 >> 20, NSETS)},
                "clocks": clocks,
                "e2e": {"value": total_points / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": n * (2 * n8 + 32), "d2h_bytes_per_step": 3 * n8, "api": "b200msm_g1_multiexp_affine with pinned host buffers"},
                "gpu_launches": int(launches),
                "roofline": roof,
                "phases_ms": {k: round(v, 4) for k, v in st.items() if k.startswith("ms_")},
                "pairs": st["pairs"], "affine_adds": adds, "resident_window_table": win}
    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own code on the host cores, bounded sample
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            pps, rms, cores, kind, sample = time_reference(cname, a.ref_log2n, 2, 1)
            line["cpu_baseline"] = {"value": pps, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample, "ms_per_sample": rms}
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "skipped", "sample": "measured at N=1 only"}
    if rank == 0: _emit(line)
    eng.free_bases(handle)
    if world > 1: dist.barrier(); dist.destroy_process_group()


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
                                       % (a.batch, CURVE_LABEL[cname].replace(" G1", ""), a.log2n, world, 4),
                           "curve": cname, "log2n_per_msm": a.log2n, "batch": a.batch, "parallelism": "replicas (independent MSMs), no collective",
                           "window_bits": int(st["window_bits"]), "windows": int(st["windows"]),
                           "cache": "no L2 flush: each MSM's working set (bases %d MiB + sort/tree scratch) exceeds the 126 MB L2 and %d MSMs run concurrently" % (n * 2 * n8 >> 20, 4)},
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
                pps, rms, cores, kind, sample = time_reference(cname, a.ref_log2n, 2, 1)
                line["cpu_baseline"] = {"value": pps, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample, "ms_per_sample": rms}
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
                "traffic": 1.02e9 if (cid == 0 and lg == 24) else None,      # profiles/r1b_ncu_full_ntt_2p24_bls_fr.csv: 0.54 GB read + 0.48 GB written per pass
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
