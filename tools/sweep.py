#!/usr/bin/env python3
"""Phase-time sweep over problem sizes (development tool): python tools/sweep.py [--curve bls12381] [--sizes 14,16,18,20] [--opt key=val ...]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm

ap = argparse.ArgumentParser()
ap.add_argument("--curve", default="bls12381"); ap.add_argument("--sizes", default="14,16,18,20"); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--opt", action="append", default=[]); ap.add_argument("--probe", action="store_true"); ap.add_argument("--roofline", action="store_true", help="compact rows with the whole-MSM fraction of the IMAD roofline"); ap.add_argument("--windowed", type=int, default=-1, help="upload with a precomputed window table of this width (0 = auto)")
a = ap.parse_args()
cid = {"bls12381": 0, "bn128": 1, "bls12381_g2": 2, "bn128_g2": 3}[a.curve]; n8 = b200msm.N8[cid]
eng = b200msm.Engine(0); dev = torch.device("cuda", 0)
eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
for kv in a.opt:
    k, v = kv.split("="); eng.set_option(k, int(v))
if a.probe:
    imad = eng.probe_imad(); fq = eng.probe_fqmul(cid); im32 = eng.probe_imad32()
    try: eng.set_option("probe29", 1); fq29 = eng.probe_fqmul(cid); eng.set_option("probe29", 0)      # -DB200_EXPERIMENTS builds only
    except Exception: fq29 = None
    eng.set_option("probe_sqr", 1); fsq = eng.probe_fqmul(cid); eng.set_option("probe_sqr", 0)
    print(json.dumps({"imad_wide_per_s": imad, "imad32_per_s": im32, "fqmul_per_s": fq, "fqmul29_per_s": fq29, "fqsqr_per_s": fsq, "sqr_over_mul": fsq / fq,
                      "fqmul_frac_of_imad_wide_peak": fq * (300 if cid == 0 else 136) / imad}))
for lg in [int(x) for x in a.sizes.split(",")]:
    n = 1 << lg
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
    eng.generate_bases(cid, 0xB2000000 + lg, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g)
    h = eng.upload_bases(cid, bases, n) if a.windowed < 0 else eng.upload_bases_windowed(cid, bases, n, 32, a.windowed); out = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    for _ in range(2): eng.multiexp_resident(h, sc, 32, n, cid, out=out)
    eng.multiexp_resident(h, sc, 32, n, cid, out=out, want_stats=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): eng.multiexp_resident(h, sc, 32, n, cid, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    agg = {}
    for _ in range(a.reps):
        _, st = eng.multiexp_resident(h, sc, 32, n, cid, out=out, want_stats=True)
        for k, v in st.items(): agg[k] = agg.get(k, 0) + v / a.reps
    row = {"log2n": lg, "ms": round(ms, 3), "c": int(agg["window_bits"]), "W": int(agg["windows"]), "rounds": int(agg["tree_rounds"]),
           "pairs": int(agg["pairs"]), "adds": int(agg["affine_adds"]), "launches": int(agg["launches"])}
    row.update({k[3:]: round(v, 3) for k, v in agg.items() if k.startswith("ms_")})
    if a.roofline:      # whole-MSM fraction of the integer-multiply roofline: 6 field multiplications per batch-affine addition
        lp = {0: 300, 1: 136, 2: 900, 3: 408}[cid]
        if "imad" not in globals(): imad = eng.probe_imad()
        row = {"curve": a.curve, "log2n": lg, "ms": row["ms"], "Mpoints_per_s": round(n / row["ms"] / 1e3, 1), "c": row["c"], "windows": row["W"], "adds": row["adds"],
               "frac_of_imad_roofline": round(row["adds"] * 6 * lp / (row["ms"] * 1e-3) / imad, 3), "table": a.windowed >= 0}
    print(json.dumps(row), flush=True)
    eng.free_bases(h); del bases, sc
