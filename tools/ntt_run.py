"""one NTT per size (development / profiling tool): python tools/ntt_run.py --log2n 24 [--curve bn128] [--reps 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import torch, b200msm
ap = argparse.ArgumentParser(); ap.add_argument("--log2n", default="24"); ap.add_argument("--curve", default="bls12381"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args(); cid = 0 if a.curve == "bls12381" else 1
eng = b200msm.Engine(0); dev = torch.device("cuda", 0); eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
for lg in [int(x) for x in a.log2n.split(",")]:
    n = 1 << lg
    x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev); x[:, 31] &= 0x0F; x = x.reshape(-1).contiguous(); out = torch.empty_like(x)
    eng.fr_fft(cid, x, lg, out=out); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(a.reps): eng.fr_fft(cid, x, lg, out=out)
    e1.record(); torch.cuda.synchronize()
    ph = eng.fr_fft_last_phases()
    print(json.dumps({"log2n": lg, "ms": round(e0.elapsed_time(e1) / a.reps, 4), **{k: round(v, 4) for k, v in ph.items()}}), flush=True)
