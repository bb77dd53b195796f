"""GPU parity tests added in round 2 (run with -m gpu on a B200), all through the C ABI:
scalars wider than 32 bytes, the optimised-binary-GCD inversion on the device, BASELINE config 4 sizes (2^22, 2^24) with exact known
answers, and the multi-device context (b200msm_create_multi) -- on one GPU with repeated ordinals (two or three shards on the same
device), on every visible GPU when there are several, and the one-process-per-GPU NCCL form under torch.multiprocessing."""
import os, random
import pytest
import pyref, coracle
from util import curve, make_bases, make_scalars, oracle_msm, gen_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import b200msm
    e = b200msm.Engine()
    yield e
    e.close()


def _norm(eng, cv, jac): return eng.normalize(cv.cid, jac)


# ---------------------------------------------------------------- f1m_inverse on the device (bingcd.h compiled by nvcc)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
def test_device_inverse_edges(eng, cname):
    cv = curve(cname); rnd = random.Random(5); L = cv.q.bit_length()
    xs = [0, 1, 2, cv.q - 1, cv.q - 2, (cv.q - 1) // 2, cv.R % cv.q, (1 << (L - 1)) % cv.q] + [(1 << k) % cv.q for k in range(0, L, 5)]
    xs += [cv.q - (1 << k) for k in range(0, L - 1, 7)] + [rnd.randrange(cv.q) for _ in range(2000)] + [rnd.randrange(1 << k) for k in range(1, L - 1, 3)]
    a = b"".join(pyref.fe_bytes(cv, x) for x in xs)
    got = eng.fq_op(cv.cid, 4, a)
    # Montgomery semantics: in = x (representing x/R), out = R^2 / x
    exp = b"".join(pyref.fe_bytes(cv, (cv.R * cv.R * pow(x, -1, cv.q)) % cv.q if x else 0) for x in xs)
    assert got == exp
    assert eng.fq_op(cv.cid, 8, a) == exp          # Fermat agrees


# ---------------------------------------------------------------- scalars wider than 32 bytes (build_multiexp.js:251-371 takes any scalarSize)
@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("ssz", [33, 40, 64, 65])
def test_msm_wide_scalars(eng, cname, ssz):
    cv = curve(cname); n = 300
    bases = make_bases(cv, n, 61); rnd = random.Random(ssz)
    sc = bytes(rnd.getrandbits(8) for _ in range(n * ssz))
    exp = oracle_msm(cv, bases, sc, ssz, n)
    assert _norm(eng, cv, eng.multiexp_affine(cv.cid, bases, sc, ssz, n)) == exp
    # the same through resident bases, and with the high bytes all zero (== the 32-byte MSM)
    h = eng.upload_bases(cv.cid, bases, n)
    try:
        assert _norm(eng, cv, eng.multiexp_resident(h, sc, ssz, n, cv.cid)) == exp
    finally:
        eng.free_bases(h)
    sc32 = make_scalars(n, 9, "u256")
    padded = b"".join(sc32[32 * i: 32 * i + 32] + bytes(ssz - 32) for i in range(n))
    assert _norm(eng, cv, eng.multiexp_affine(cv.cid, bases, padded, ssz, n)) == oracle_msm(cv, bases, sc32, 32, n)


def test_two_torsion_like_input_does_not_poison_the_batch(eng):
    """an (unchecked, off-curve) point with y == 0 that meets itself in a bucket doubles to infinity (2y = 0); in the batch-affine
    tree its denominator must stay out of the shared product, or the single inversion of the round fails for EVERY addition.
    The bad pair is alone in its bucket (only they have bit 250 set), so the expected sum is exactly the MSM of the good points."""
    cv = curve("bls12381"); n = 1000
    bad = pyref.fe_bytes(cv, 5 * cv.R % cv.q) + bytes(cv.n8)             # x = 5, y = 0 (Montgomery form)
    rnd = random.Random(72)
    good_b = make_bases(cv, n, 71); good_s = b"".join(rnd.getrandbits(200).to_bytes(32, "little") for _ in range(n))
    bases = good_b[: 500 * 2 * cv.n8] + bad + bad + good_b[500 * 2 * cv.n8:]
    sc = good_s[: 500 * 32] + (1 << 250).to_bytes(32, "little") * 2 + good_s[500 * 32:]
    exp = oracle_msm(cv, good_b, good_s, 32, n)
    for mode, rounds in ((2, 2), (2, -1), (1, -1)):
        eng.set_option("accumulate", mode); eng.set_option("tree_rounds", rounds)
        try:
            got = _norm(eng, cv, eng.multiexp_affine(cv.cid, bases, sc, 32, n + 2))
        finally:
            eng.set_option("accumulate", 0); eng.set_option("tree_rounds", -1)
        assert got == exp, (mode, rounds)
    # and alone: P + P = infinity
    eng.set_option("tree_rounds", 1)
    try:
        assert _norm(eng, cv, eng.multiexp_affine(cv.cid, bad + bad, (7).to_bytes(32, "little") * 2, 32, 2)) == bytes(2 * cv.n8)
    finally:
        eng.set_option("tree_rounds", -1)


# ---------------------------------------------------------------- BASELINE config 4 sizes on one GPU, exact known answer
def _splitmix64_np(x):
    import numpy as np
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def known_answer(cv, seed, first, n, sc_u8_tensor):
    """(sum_i s_i * k_i mod r) * G for bases P_i = splitmix64(seed + first + i) * G: host big integers + one oracle scalar multiplication"""
    import numpy as np
    total = 0; CH = 1 << 20
    with np.errstate(over="ignore"):
        for lo in range(0, n, CH):
            m = min(CH, n - lo)
            k = _splitmix64_np(np.uint64(seed) + np.arange(first + lo, first + lo + m, dtype=np.uint64)); k[k == 0] = 1
            w = sc_u8_tensor[lo * 32:(lo + m) * 32].cpu().numpy().view("<u8").reshape(m, 4).astype(object)
            s = w[:, 0] + (w[:, 1] << 64) + (w[:, 2] << 128) + (w[:, 3] << 192)
            total = (total + int((s * k.astype(object)).sum())) % cv.r
    return coracle.normalize(cv.cid, coracle.times_scalar_affine(cv.cid, gen_bytes(cv), total.to_bytes(32, "little")))


@pytest.mark.parametrize("cname,lg", [("bls12381", 22), ("bls12381", 24), ("bn128", 22)])
def test_big_sizes_known_answer(eng, cname, lg):
    import torch
    cv = curve(cname); n = 1 << lg; seed = 0xB2000000 + lg
    d = torch.empty(n * 2 * cv.n8, dtype=torch.uint8, device="cuda")
    eng.generate_bases(cv.cid, seed, 0, n, d)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
    got = _norm(eng, cv, eng.multiexp_affine(cv.cid, d, sc, 32, n))
    assert got == known_answer(cv, seed, 0, n, sc)
    del d, sc; torch.cuda.empty_cache()


# ---------------------------------------------------------------- multi-device context behind the C ABI
def _multi_cases(me, one, cv, n, seed):
    """every sharded entry point of a multi context `me` against the single-device engine `one` (itself pinned against the oracle above)"""
    bases = make_bases(cv, n, seed); sc = make_scalars(n, seed + 1, "u256")
    exp = oracle_msm(cv, bases, sc, 32, n) if n <= 5000 else _norm(one, cv, one.multiexp_affine(cv.cid, bases, sc, 32, n))
    assert _norm(one, cv, me.multiexp_affine(cv.cid, bases, sc, 32, n)) == exp
    # chunk API: one window, sharded
    assert _norm(one, cv, me.multiexp_affine_chunk(cv.cid, bases, sc, 32, n, 37, 13)) == _norm(one, cv, one.multiexp_affine_chunk(cv.cid, bases, sc, 32, n, 37, 13))
    # Jacobian bases
    jac = one.batch_convert(cv.cid, "toJacobian", bases, n)
    assert _norm(one, cv, me.multiexp_jacobian(cv.cid, jac, sc, 32, n)) == exp
    # resident: sharded, replicated, windowed; prefixes of the uploaded set (ragged against the shard boundaries)
    for repl in (0, 1):
        me.set_option("multi_replicate", repl)
        for windowed in (False, True):
            h = me.upload_bases_windowed(cv.cid, bases, n, 32, 0) if windowed else me.upload_bases(cv.cid, bases, n)
            try:
                assert _norm(one, cv, me.multiexp_resident(h, sc, 32, n, cv.cid)) == exp, (repl, windowed)
                for m in sorted({1, n // 3, n - 1} - {0}):
                    e2 = oracle_msm(cv, bases[: m * 2 * cv.n8], sc[: m * 32], 32, m) if m <= 5000 else _norm(one, cv, one.multiexp_affine(cv.cid, bases[: m * 2 * cv.n8], sc[: m * 32], 32, m))
                    assert _norm(one, cv, me.multiexp_resident(h, sc[: m * 32], 32, m, cv.cid)) == e2, (repl, windowed, m)
                # batch of 5 MSMs over the same bases
                cnt = 5; m = max(1, n // 2)
                scs = b"".join(make_scalars(m, 900 + j, "u256") for j in range(cnt))
                outs = me.multiexp_batch(h, scs, 32, m, cnt, cv.cid)
                for j in range(cnt):
                    ej = _norm(one, cv, one.multiexp_affine(cv.cid, bases[: m * 2 * cv.n8], scs[j * m * 32:(j + 1) * m * 32], 32, m))
                    assert _norm(one, cv, outs[j * 3 * cv.n8:(j + 1) * 3 * cv.n8]) == ej, (repl, windowed, j)
            finally:
                me.free_bases(h)
    me.set_option("multi_replicate", 0)


@pytest.mark.parametrize("cname", ["bls12381", "bn128"])
@pytest.mark.parametrize("shards", [2, 3])
def test_multi_context_shards_on_one_gpu(eng, cname, shards):
    """b200msm_create_multi with a repeated ordinal: the sharding, per-device threads and the host combination of the partials, on one GPU"""
    import b200msm
    cv = curve(cname)
    me = b200msm.Engine(devices=[0] * shards)
    try:
        assert me.device_count == shards
        me.set_option("multi_min_points", 1)                 # shard even tiny problems
        for n, seed in ((1, 3), (2, 4), (7, 5), (1000, 6), (20000, 7)):
            _multi_cases(me, eng, cv, n, seed)
        me.set_option("multi_min_points", 1 << 15)           # default policy: a small MSM stays on one device
        bases = make_bases(cv, 500, 8); sc = make_scalars(500, 9, "u256")
        assert _norm(eng, cv, me.multiexp_affine(cv.cid, bases, sc, 32, 500)) == oracle_msm(cv, bases, sc, 32, 500)
        # n = 0 and infinity inputs
        assert _norm(eng, cv, me.multiexp_affine(cv.cid, b"", b"", 32, 0)) == bytes(2 * cv.n8)
    finally:
        me.close()


def test_multi_context_g2_and_device_inputs(eng):
    import torch, b200msm
    me = b200msm.Engine(devices=[0, 0]); me.set_option("multi_min_points", 1)
    try:
        cid = 2; n = 3000; n8 = 96
        d = torch.empty(n * 2 * n8, dtype=torch.uint8, device="cuda"); eng.generate_bases(cid, 0xB2000000, 0, n, d)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
        out = torch.zeros(3 * n8, dtype=torch.uint8, device="cuda")
        me.multiexp_affine(cid, d, sc, 32, n, out=out)       # device inputs and a device output
        assert eng.normalize(cid, out) == eng.normalize(cid, eng.multiexp_affine(cid, d, sc, 32, n))
    finally:
        me.close()


def test_multi_context_all_visible_gpus(eng):
    """the real thing when the box has several GPUs: one context over all of them, full-size known answer"""
    import torch, b200msm
    G = torch.cuda.device_count()
    if G < 2: pytest.skip("needs >= 2 GPUs")
    cv = curve("bls12381"); lg = 20; n = G << lg; seed = 0xB2000000 + lg
    me = b200msm.Engine(devices=list(range(G)))
    try:
        d = torch.empty(n * 2 * cv.n8, dtype=torch.uint8, device="cuda:0"); eng.generate_bases(cv.cid, seed, 0, n, d)
        g = torch.Generator(device="cuda:0"); g.manual_seed(lg)
        sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda:0", generator=g)
        exp = known_answer(cv, seed, 0, n, sc)
        hb = d.cpu(); hs = sc.cpu()                          # host buffers: every GPU pulls its own slice
        assert _norm(eng, cv, me.multiexp_affine(cv.cid, hb.numpy(), hs.numpy(), 32, n)) == exp
        h = me.upload_bases(cv.cid, hb.numpy(), n)
        assert _norm(eng, cv, me.multiexp_resident(h, hs.numpy(), 32, n, cv.cid)) == exp
        me.free_bases(h)
        _multi_cases(me, eng, cv, 20000, 11)
    finally:
        me.close()


# ---------------------------------------------------------------- one process per GPU + NCCL (the torchrun form bench.py uses)
def _nccl_worker(rank, world, port, lg, q):
    import sys, torch, torch.distributed as dist
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zprize-wasm-msm_b200")):
        if p not in sys.path: sys.path.insert(0, p)
    import b200msm
    from b200msm.sharded import engine_sharded_msm
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cv = pyref.BLS12_381; n = 1 << lg; seed = 0xB2000000 + lg
        e = b200msm.Engine(rank)
        d = torch.empty(n * 2 * cv.n8, dtype=torch.uint8, device="cuda"); e.generate_bases(cv.cid, seed, 0, n, d)      # every rank holds the full inputs, uses its slice
        g = torch.Generator(device="cuda"); g.manual_seed(lg)
        sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
        res = engine_sharded_msm(e, cv.cid, bytes(d.cpu().numpy()), bytes(sc.cpu().numpy()), 32, n, device=torch.device("cuda", rank))
        q.put((rank, e.normalize(cv.cid, res)))
        e.close()
    finally:
        dist.destroy_process_group()


def test_sharded_msm_two_ranks_nccl(eng):
    import torch, torch.multiprocessing as mp
    if torch.cuda.device_count() < 2: pytest.skip("needs >= 2 GPUs")
    lg = 16; cv = curve("bls12381"); n = 1 << lg; seed = 0xB2000000 + lg
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, lg, q)) for r in range(2)]
    for p in procs: p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs: p.join(timeout=120)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
    exp = known_answer(cv, seed, 0, n, sc)
    assert got[0] == exp and got[1] == exp
