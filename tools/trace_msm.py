#!/usr/bin/env python3
"""Kernel timeline of one MSM (development tool): CUPTI activity records through torch.profiler -- exact start / end of every kernel on
every lane's stream, no ncu serialisation.

    python tools/trace_msm.py --log2n 20 [--configs "base;lanes=1"] [--windowed 0] --out gpurun_out/trace_2p20.json

Prints, per configuration: the MSM's wall time on the device, the time during which 1 / 2 / 3+ kernels were in flight, idle gaps, and per
kernel name the summed duration and the EXCLUSIVE time (instants at which it was the only kernel running); writes the raw records as JSON.
"""
import argparse, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--curve", default="bls12381"); ap.add_argument("--log2n", type=int, default=20); ap.add_argument("--configs", default="base")
ap.add_argument("--windowed", type=int, default=-1); ap.add_argument("--out", default=""); ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--dump", action="store_true", help="print every kernel of the last MSM")
a = ap.parse_args()
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm
from torch.profiler import profile, ProfilerActivity

cid = {"bls12381": 0, "bn128": 1, "bls12381_g2": 2, "bn128_g2": 3}[a.curve]; n8 = b200msm.N8[cid]
dev = torch.device("cuda", 0); n = 1 << a.log2n
eng = b200msm.Engine(0); eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
eng.generate_bases(cid, 0xB2000000 + a.log2n, 0, n, bases)
g = torch.Generator(device=dev); g.manual_seed(a.log2n)
sc = [torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
h = eng.upload_bases(cid, bases, n) if a.windowed < 0 else eng.upload_bases_windowed(cid, bases, n, 32, a.windowed)
out = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
result = {}
for cfg in a.configs.split(";"):
    if cfg not in ("", "base"):
        for kv in cfg.split(","): eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    for i in range(4): eng.multiexp_resident(h, sc[i % 2], 32, n, cid, out=out)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.iters):
            eng.multiexp_resident(h, sc[i % 2], 32, n, cid, out=out); torch.cuda.synchronize()
    path = "/tmp/trace_%d.json" % os.getpid(); prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    # split into MSMs at the largest gaps (each iteration ends in a synchronize)
    gaps = sorted(range(1, len(ev)), key=lambda i: ev[i]["ts"] - max(x["ts"] + x["dur"] for x in ev[max(0, i - 40):i]), reverse=True)[:a.iters - 1]
    cut = [0] + sorted(gaps) + [len(ev)]
    last = ev[cut[-2]:cut[-1]]
    t0 = last[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in last)
    recs = [{"name": e["name"].split("<")[0].split("(")[0].replace("void ", "").replace("b200::", ""), "stream": e["args"].get("stream"), "t": round(e["ts"] - t0, 2), "d": round(e["dur"], 2)} for e in last]
    # sweep: concurrency histogram and exclusive time per kernel name
    pts = []
    for i, r in enumerate(recs): pts.append((r["t"], 1, i)); pts.append((r["t"] + r["d"], -1, i))
    pts.sort()
    live = set(); prev = 0.0; conc = collections.Counter(); excl = collections.Counter(); tot = collections.Counter(); cnt = collections.Counter()
    for t, kind, i in pts:
        dt = t - prev
        if dt > 0:
            conc[min(len(live), 3)] += dt
            if len(live) == 1: excl[recs[next(iter(live))]["name"]] += dt
        prev = t
        if kind == 1: live.add(i)
        else: live.discard(i)
    for r in recs: tot[r["name"]] += r["d"]; cnt[r["name"]] += 1
    streams = sorted({r["stream"] for r in recs})
    print("== %s  2^%d  config=%s  wall %.3f ms  kernels %d  streams %d" % (a.curve, a.log2n, cfg or "base", (t1 - t0) / 1000, len(recs), len(streams)))
    print("   in flight: idle %.3f  one %.3f  two %.3f  three+ %.3f ms" % tuple(conc[k] / 1000 for k in range(4)))
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("   %-28s n=%3d  sum %.3f ms  exclusive %.3f ms" % (name[:28], cnt[name], v / 1000, excl[name] / 1000))
    if a.dump:
        for r in recs: print("   %9.1f +%8.1f  s%-3s %s" % (r["t"], r["d"], streams.index(r["stream"]), r["name"]))
    result[cfg or "base"] = recs
if a.out: json.dump(result, open(a.out, "w"))
eng.free_bases(h); eng.close()
