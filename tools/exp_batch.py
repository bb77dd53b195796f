#!/usr/bin/env python3
"""Development A/B for the batched entry point (b200msm_g1_multiexp_batch): python tools/exp_batch.py --configs "base;fold_cluster=0" """
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser(); ap.add_argument("--log2n", type=int, default=18); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--configs", default="base")
a = ap.parse_args()
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm
dev = torch.device("cuda", 0); cid = 0; n8 = 48; n = 1 << a.log2n
DEFAULTS = {"lanes": 4, "sort_groups": 1, "fold_cluster": 1, "batch_workers": 4, "issue_threads": 0}
for cfg in a.configs.split(";"):
    eng = b200msm.Engine(0); eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    for k, v in DEFAULTS.items(): eng.set_option(k, v)
    if cfg not in ("", "base"):
        for kv in cfg.split(","): eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, 0xB2000000 + a.log2n, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(5)
    sc = torch.randint(0, 256, (a.batch * n * 32,), dtype=torch.uint8, device=dev, generator=g)
    h = eng.upload_bases(cid, bases, n); out = torch.zeros(a.batch * 3 * n8, dtype=torch.uint8, device=dev)
    for _ in range(2): eng.multiexp_batch(h, sc, 32, n, a.batch, cid, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); best = 1e9
    for _ in range(3):
        e0.record(); eng.multiexp_batch(h, sc, 32, n, a.batch, cid, out=out); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"log2n": a.log2n, "batch": a.batch, "config": cfg, "ms_batch": round(best, 2), "ms_per_msm": round(best / a.batch, 4)}), flush=True)
    eng.free_bases(h); eng.close()
