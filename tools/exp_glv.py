#!/usr/bin/env python3
"""Development probe: the MSM over GLV-split inputs (reference flow: g1m_glv_preprocessEndomorphism then multiExp on 2N points, test/glv.js:103-192)
against the plain MSM, same points and scalars, device-resident inputs.  python tools/exp_glv.py --sizes 16,18,20 [--wb 0,15,16]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser(); ap.add_argument("--sizes", default="18,20"); ap.add_argument("--wb", default="0"); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm
dev = torch.device("cuda", 0); cid = 0; n8 = 48
eng = b200msm.Engine(0); eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
def timed(fn, iters):
    for _ in range(3): fn()
    best = 1e9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / iters)
    return best
for lg in [int(x) for x in a.sizes.split(",")]:
    n = 1 << lg
    bases = torch.empty(n * 2 * n8, dtype=torch.uint8, device=dev); eng.generate_bases(cid, 0xB2000000 + lg, 0, n, bases)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device=dev, generator=g)
    sc.view(n, 32)[:, 31] &= 0x3f          # < r, like a prover's scalars
    out = torch.zeros(3 * n8, dtype=torch.uint8, device=dev); out2 = torch.zeros(3 * n8, dtype=torch.uint8, device=dev)
    p2 = torch.empty(2 * n * 2 * n8, dtype=torch.uint8, device=dev); s2 = torch.empty(2 * n * 32, dtype=torch.uint8, device=dev)
    for wb in [int(x) for x in a.wb.split(",")]:
        eng.set_option("window_bits", wb)
        t_plain = timed(lambda: eng.multiexp_affine(cid, bases, sc, 32, n, out=out), a.iters)
        def glv():
            eng.glv_preprocess(cid, bases, sc, n, out_points=p2, out_scalars=s2)
            eng.multiexp_affine(cid, p2, s2, 32, 2 * n, out=out2)
        t_glv = timed(glv, a.iters)
        s16 = torch.empty(2 * n * 16, dtype=torch.uint8, device=dev)
        def glv16():
            eng.glv_preprocess(cid, bases, sc, n, out_points=p2, out_scalars=s2)
            s16.view(2 * n, 16).copy_(s2.view(2 * n, 32)[:, :16])
            eng.multiexp_affine(cid, p2, s16, 16, 2 * n, out=out2)
        t_glv16 = timed(glv16, a.iters)
        same16 = eng.normalize(cid, out) == eng.normalize(cid, out2)
        t_msm16 = timed(lambda: eng.multiexp_affine(cid, p2, s16, 16, 2 * n, out=out2), a.iters)
        t_pre = timed(lambda: eng.glv_preprocess(cid, bases, sc, n, out_points=p2, out_scalars=s2), a.iters)
        same = eng.normalize(cid, out) == eng.normalize(cid, out2)
        print(json.dumps({"log2n": lg, "window_bits": wb, "ms_plain": round(t_plain, 4), "ms_glv_total": round(t_glv, 4), "ms_glv_preprocess": round(t_pre, 4), "same_point": same, "ms_glv16_total": round(t_glv16, 4), "ms_msm16_only": round(t_msm16, 4), "same16": same16}), flush=True)
eng.close()
