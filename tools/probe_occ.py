"""occupancy sensitivity of the field multiplier and the FP64 pipe rate (development tool)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "zprize-wasm-msm_b200")): sys.path.insert(0, p)
import b200msm
eng = b200msm.Engine(0)
imad = eng.probe_imad()
print(json.dumps({"imad_wide_per_s": imad, "imad32_per_s": eng.probe_imad32(), "dfma_per_s": eng.probe_dfma()}))
print(json.dumps(eng.probe_dualpipe()))
for smem, warps in ((0, 32), (56 * 1024, 24), (100 * 1024, 16), (200 * 1024, 8)):
    eng.set_option("probe_smem", smem)
    for sq in (0, 1):
        eng.set_option("probe_sqr", sq)
        fq = eng.probe_fqmul(0)
        print(json.dumps({"warps_per_sm": warps, "sqr": sq, "fq_per_s": fq, "frac_of_imad_peak": fq * (234 if sq else 300) / imad}))
