#!/usr/bin/env python3
"""Full-size correctness + timing check for large MSMs (BASELINE config 4 sizes) on one GPU:
python tools/bigcheck.py --sizes 22,24 [--curve bls12381].  Exact known answer: bases are P_i = k_i*G with a known
splitmix64 stream, so sum_i s_i*P_i = (sum_i s_i*k_i mod r)*G (host big integers + one oracle scalar multiplication)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import numpy as np, torch
import b200msm, pyref, coracle

ap = argparse.ArgumentParser(); ap.add_argument("--sizes", default="22,24"); ap.add_argument("--curve", default="bls12381"); ap.add_argument("--pattern", default="uniform", choices=["uniform", "equal", "small", "sparse"])
a = ap.parse_args()
cv = pyref.CURVES[a.curve]; cid = cv.cid; n8 = cv.n8
eng = b200msm.Engine(0); dev = torch.device("cuda", 0)
eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)


def splitmix(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


for lg in [int(x) for x in a.sizes.split(",")]:
    n = 1 << lg; seed = 0xB2000000 + lg
    t0 = time.time()
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev)
    eng.generate_bases(cid, seed, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g)
    if a.pattern == "equal": sc = sc[:32].repeat(n)                                   # every scalar identical: one bucket per window holds all points
    elif a.pattern == "small": sc = (sc.view(n, 32) * (torch.arange(32, device=dev) < 1)).reshape(-1).contiguous()   # scalars < 256
    elif a.pattern == "sparse": sc = (sc.view(n, 32) * (torch.rand(n, 1, device=dev, generator=g) < 0.05)).to(torch.uint8).reshape(-1).contiguous()   # 95 % zero scalars
    tgen = time.time() - t0
    out = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    h = eng.upload_bases(cid, bases, n)
    eng.multiexp_resident(h, sc, 32, n, cid, out=out); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.multiexp_resident(h, sc, 32, n, cid, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    got = eng.normalize(cid, bytes(out.cpu().numpy()))
    # exact answer, chunked to bound host memory
    t0 = time.time(); total = 0; CH = 1 << 21
    with np.errstate(over="ignore"):
        for lo in range(0, n, CH):
            m = min(CH, n - lo)
            k = splitmix(np.uint64(seed) + np.arange(lo, lo + m, dtype=np.uint64)); k[k == 0] = 1
            w = sc[lo * 32:(lo + m) * 32].cpu().numpy().view("<u8").reshape(m, 4).astype(object)
            s = w[:, 0] + (w[:, 1] << 64) + (w[:, 2] << 128) + (w[:, 3] << 192)
            total = (total + int((s * k.astype(object)).sum())) % cv.r
    exp = coracle.normalize(cid, coracle.times_scalar_affine(cid, pyref.affine_to_bytes(cv, cv.G), total.to_bytes(32, "little")))
    print(json.dumps({"curve": a.curve, "pattern": a.pattern, "log2n": lg, "ms": round(ms, 3), "Mpoints_per_s": round(n / ms / 1e3, 1), "exact_match": got == exp,
                      "gen_s": round(tgen, 2), "host_check_s": round(time.time() - t0, 1), "gpu_mem_GiB": round(torch.cuda.mem_get_info()[0] / 2**30, 1)}), flush=True)
    eng.free_bases(h); del bases, sc
