#!/usr/bin/env python3
"""Batched throughput (BASELINE config 5): B independent MSMs of 2^log2n points streamed over C engine contexts per GPU.
Each context = own stream + scratch; host threads issue the calls (ctypes releases the GIL), so one MSM's latency-bound
tails overlap another MSM's throughput kernels.   python tools/batched.py --log2n 18 --batch 64 --contexts 1,2,4"""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=18); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--contexts", default="1,2,4")
ap.add_argument("--curve", default="bls12381"); ap.add_argument("--gpu", type=int, default=0)
a = ap.parse_args()
cid = 0 if a.curve == "bls12381" else 1; n8 = b200msm.N8[cid]; n = 1 << a.log2n
dev = torch.device("cuda", a.gpu); torch.cuda.set_device(dev)
gen = b200msm.Engine(a.gpu)
bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev); gen.generate_bases(cid, 0xB2000000 + a.log2n, 0, n, bases)
g = torch.Generator(device=dev); g.manual_seed(7)
scal = [torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g) for _ in range(8)]
torch.cuda.synchronize()
ref = None
for C in [int(x) for x in a.contexts.split(",")]:
    engs = [b200msm.Engine(a.gpu) for _ in range(C)]
    hs = [e.upload_bases(cid, bases, n) for e in engs]
    outs = [[None] * a.batch for _ in range(1)]
    res = [None] * a.batch

    def worker(k):
        e, h = engs[k], hs[k]
        for j in range(k, a.batch, C):
            res[j] = e.multiexp_resident(h, scal[j % 8], 32, n, cid)      # host result: call returns when this MSM is done
    for k in range(C): worker(k) if False else None
    # warm-up
    ts = [threading.Thread(target=worker, args=(k,)) for k in range(C)]
    [t.start() for t in ts]; [t.join() for t in ts]
    t0 = time.perf_counter()
    ts = [threading.Thread(target=worker, args=(k,)) for k in range(C)]
    [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.perf_counter() - t0
    if ref is None: ref = list(res)
    same = all(gen.normalize(cid, res[j]) == gen.normalize(cid, ref[j]) for j in range(0, a.batch, 7))
    print(json.dumps({"log2n": a.log2n, "batch": a.batch, "contexts": C, "ms_per_msm": round(dt * 1e3 / a.batch, 3),
                      "msm_per_s": round(a.batch / dt, 1), "Mpoints_per_s": round(a.batch * n / dt / 1e6, 1), "results_match": same}), flush=True)
    for e, h in zip(engs, hs): e.free_bases(h); e.close()
