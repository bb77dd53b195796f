#!/usr/bin/env python3
"""Round-2 A/B harness (development tool): one process, several engine configurations, JSON lines out.

    python tools/exp_r2.py --sizes 16,18,20 --configs "base;lanes=1;l2_persist=100" [--lib path/to/variant.so] [--phases]

Every configuration is `key=val,key=val` of b200msm_set_option keys ("base" = defaults).  For each (size, configuration) the MSM over
resident bases is timed with CUDA events (best of --reps blocks of --iters MSMs) and, with --phases, the engine's single-lane phase
profile is printed.  Results are checked against the first configuration's normalized point (bit-exact) before they are reported.
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--curve", default="bls12381"); ap.add_argument("--sizes", default="20"); ap.add_argument("--configs", default="base")
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--lib", default="")
ap.add_argument("--phases", action="store_true"); ap.add_argument("--windowed", type=int, default=-1); ap.add_argument("--tag", default="")
a = ap.parse_args()
if a.lib: os.environ["B200MSM_LIB"] = os.path.abspath(a.lib)
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm

cid = {"bls12381": 0, "bn128": 1, "bls12381_g2": 2, "bn128_g2": 3}[a.curve]; n8 = b200msm.N8[cid]
dev = torch.device("cuda", 0)
DEFAULTS = {"lanes": 4, "sort_groups": 1, "window_bits": 0, "tree_rounds": -1, "ba_k": 0, "pt_k": 8, "persist": 592, "combine": 0, "accumulate": 0, "subslots": 0, "group_pairs": 0, "xonly": 1, "fold_cluster": 1, "fused_round": 0, "fused_grid": 444, "block_tree": 0, "meta_upfront": 0, "fused_tiles": 592, "fused_kmax": 16}


def parse_cfg(c):
    if c in ("", "base"): return {}
    return {kv.split("=")[0]: int(kv.split("=")[1]) for kv in c.split(",")}


for lg in [int(x) for x in a.sizes.split(",")]:
    n = 1 << lg
    eng = b200msm.Engine(0); eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
    eng.generate_bases(cid, 0xB2000000 + lg, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    sc = [torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    h = eng.upload_bases(cid, bases, n) if a.windowed < 0 else eng.upload_bases_windowed(cid, bases, n, 32, a.windowed)
    out = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    ref = None
    for cfg in a.configs.split(";"):
        opts = parse_cfg(cfg)
        for k, v in DEFAULTS.items():
            try: eng.set_option(k, v)
            except Exception: pass      # an older build of the library may not know a newer key
        try:
            for k, v in opts.items(): eng.set_option(k, v)
        except Exception as ex:
            print(json.dumps({"log2n": lg, "config": cfg, "error": repr(ex)}), flush=True); continue
        for i in range(3): eng.multiexp_resident(h, sc[i % 2], 32, n, cid, out=out)
        eng.multiexp_resident(h, sc[0], 32, n, cid, out=out); torch.cuda.synchronize()
        res = eng.normalize(cid, out)
        if ref is None: ref = res
        best = 1e9
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = eng.counter("launches")
        for _ in range(a.reps):
            torch.cuda.synchronize(); e0.record()
            for i in range(a.iters): eng.multiexp_resident(h, sc[i % 2], 32, n, cid, out=out)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / a.iters)
        row = {"tag": a.tag, "curve": a.curve, "log2n": lg, "config": cfg or "base", "ms": round(best, 4), "same_point": res == ref,
               "launches": (eng.counter("launches") - l0) // (a.reps * a.iters)}
        if a.phases:
            agg = {}
            eng.multiexp_resident(h, sc[0], 32, n, cid, out=out, want_stats=True)
            for i in range(5):
                _, st = eng.multiexp_resident(h, sc[i % 2], 32, n, cid, out=out, want_stats=True)
                for k, v in st.items(): agg[k] = agg.get(k, 0) + v / 5
            row["phases"] = {k[3:]: round(v, 3) for k, v in agg.items() if k.startswith("ms_")}
            row.update({"c": int(agg["window_bits"]), "W": int(agg["windows"]), "rounds": int(agg["tree_rounds"]), "adds": int(agg["affine_adds"])})
        print(json.dumps(row), flush=True)
    eng.free_bases(h); eng.close(); del bases, sc
