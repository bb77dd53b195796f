"""CPU test of the multi-rank plumbing (world_size 2, gloo): point-range partition, all_gather of the partial points,
combination.  The per-rank MSM and the point sum are played by the oracle here; on GPUs the engine is injected instead
(b200msm.sharded.engine_sharded_msm, exercised by bench.py --gpus N and tests/test_gpu_parity.py)."""
import os, socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyref, coracle
from util import make_bases, make_scalars, oracle_msm


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cname, n, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "zprize-wasm-msm_b200"), here):
        if p not in sys.path: sys.path.insert(0, p)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharded", os.path.join(root, "zprize-wasm-msm_b200", "b200msm", "sharded.py"))
    sharded = importlib.util.module_from_spec(spec); spec.loader.exec_module(sharded)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cv = pyref.CURVES[cname]
    bases = make_bases(cv, n, 99); sc = make_scalars(n, 7, "u256")

    def local(b, s, ss, m): return coracle.multiexp_affine(cv.cid, b, s, ss, m)

    def combine(parts):
        acc = parts[0]
        for p in parts[1:]: acc = coracle.add(cv.cid, acc, p)
        return acc
    res = sharded.sharded_msm(local, combine, bases, sc, 32, n, 2 * cv.n8)
    q.put((rank, coracle.normalize(cv.cid, res)))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("cname,n", [("bls12381", 301), ("bn128", 64), ("bls12381", 1)])
def test_sharded_msm_world2_gloo(cname, n):
    world = 2; port = _free_port()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cname, n, q)) for r in range(world)]
    for p in procs: p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    cv = pyref.CURVES[cname]
    exp = oracle_msm(cv, make_bases(cv, n, 99), make_scalars(n, 7, "u256"), 32, n)
    assert all(o[1] == exp for o in outs)


def test_shard_ranges_cover_exactly():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("sharded", os.path.join(root, "zprize-wasm-msm_b200", "b200msm", "sharded.py"))
    sharded = importlib.util.module_from_spec(spec); spec.loader.exec_module(sharded)
    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            rs = [sharded.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
