"""CPU: host-side logic of the product -- the C-ABI library loads and exports every symbol of include/tnad.h,
the einsum -> GEMM planner's stride algebra, the host mirror of the reference API, and the multi-process
(gloo, world_size 2) plumbing of the replicated / sweep paths.  No compute call touches a GPU here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import tnad_b200 as T
import tnad_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI -------------------------------------------------------------------------------------------
def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tnad.h")).read()
    return sorted(set(re.findall(r"TNAD_API\s+[\w\s\*]+?\b(tnad_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = T.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tnad.h but not exported"
    assert set(declared) == set(T.SIGNATURES), "ctypes binding table and header disagree"
    assert lib.tnad_version() >= 100


def test_no_cpu_fallback_without_device():
    lib = T.load_library()
    import ctypes as C
    h = C.c_void_p()
    rc = lib.tnad_create(99, C.byref(h))       # no such device anywhere
    assert rc != 0 and not h.value
    msg = lib.tnad_last_error(None).decode()
    assert "device" in msg.lower()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensornetworkad.jl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "tnad_oracle" not in src and "oracle/" not in src, fn
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        for line in open(os.path.join(pkg, "csrc", fn)):
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line, (fn, line)


# ---- einsum -> GEMM planner ------------------------------------------------------------------------------
DRIVER_SPECS = """iba,ad->ibd ibd,dcl->ibcl ibcl,jkcb->ijlk pq,qj->pj pi,pj->ij abi,aed->ibed ibed,bjce->ijcd ijcd,dck->ijk
pi,ij->pj pj,qj->pq qp,qj->pj ijk,dck->ijcd ijcd,ijk->dck ijcd,bjce->ibed ibed,ijcd->bjce ibed,aed->abi abi,ibed->aed
ijlk,jkcb->ibcl ibcl,ijlk->jkcb ibcl,dcl->ibd ibd,ibcl->dcl ibd,ad->iba iba,ibd->ad
ica,ab->icb eg,gfk->efk icb,bde->icde icde,cjfdlm->iejflm iejflm,efk->ijklm abckl,ijkl->abcij abcij,ij->abc abc,ij->abcij
ijklm,efk->iejflm iejflm,ijklm->efk iejflm,cjfdlm->icde icde,iejflm->cjfdlm icde,bde->icb icb,icde->bde efk,gfk->eg
eg,efk->gfk icb,ab->ica ica,icb->ab npu,por->nour dom,lmn->dlno nour,dlno->urdl urdl,dlno->nour nour,urdl->dlno
nour,por->npu npu,nour->por dlno,lmn->dom dom,dlno->lmn mi,mj->ij ij,nj->in mi,in->mn mi,ij->mj mj,nj->mn mk,m->k mk,k->m
ia,ajb->ijb ijb,bk->ijk alc,ckd->alkd bjd,bia->jdia alkd,jdia->ijkl mr,mq->rq mr,rq->mq ki,kj->ij ik,kj->ij""".split()


def _off(levels, i):
    o = 0
    for l, (n, s) in enumerate(levels):
        if l == len(levels) - 1:
            o += i * s
        else:
            o += (i % n) * s
            i //= n
    return o


def _run_plan(spec, A, B):
    p = T.contract_plan(spec, A.shape, B.shape)
    lhs, out = spec.split("->")
    la, lb = lhs.split(",")
    ext = dict(zip(la, A.shape)); ext.update(zip(lb, B.shape))
    cshape = tuple(ext[l] for l in out)
    Af, Bf = A.reshape(-1, order="F"), B.reshape(-1, order="F")
    Cf = np.zeros(int(np.prod(cshape)))
    am = np.array([_off(p["am"], i) for i in range(p["M"])]); cm = np.array([_off(p["cm"], i) for i in range(p["M"])])
    ak = np.array([_off(p["ak"], i) for i in range(p["K"])]); bk = np.array([_off(p["bk"], i) for i in range(p["K"])])
    bn = np.array([_off(p["bn"], i) for i in range(p["N"])]); cn = np.array([_off(p["cn"], i) for i in range(p["N"])])
    for b in range(p["batch"]):
        ab, bb, cb = _off(p["ab"], b), _off(p["bb"], b), _off(p["cb"], b)
        Cf[cb + cm[:, None] + cn[None, :]] = Af[ab + am[:, None] + ak[None, :]] @ Bf[bb + bk[:, None] + bn[None, :]]
    for vec, kfast, r, k in ((p["a_vec"], p["a_kfast"], p["am"], p["ak"]), (p["b_vec"], p["b_kfast"], p["bn"], p["bk"])):
        if vec:
            fast = k if kfast else r
            assert fast[0][1] == 1 and fast[0][0] % 2 == 0
    return Cf.reshape(cshape, order="F"), p


@pytest.mark.parametrize("spec", DRIVER_SPECS)
def test_contract_plan_matches_einsum(spec):
    rng = np.random.default_rng(abs(hash(spec)) % 2 ** 32)
    labs = sorted(set(re.sub("[^a-z]", "", spec)))
    for trial in range(3):
        ext = {l: int(rng.integers(1 if trial == 2 else 2, 7)) for l in labs}
        lhs, _ = spec.split("->")
        la, lb = lhs.split(",")
        A = rng.standard_normal([ext[l] for l in la]); B = rng.standard_normal([ext[l] for l in lb])
        C, p = _run_plan(spec, A, B)
        assert np.allclose(C, np.einsum(spec, A, B), atol=1e-12)
        for s in ("am", "ak", "bk", "bn", "cm", "cn"):
            assert 1 <= len(p[s]) <= 4


def test_contract_plan_headline_shapes_are_vectorised():
    # d=4, chi=128: the big contractions must take the 16-byte cp.async path on both operands
    chi, D = 128, 16
    for spec, sa, sb in [("ibcl,jkcb->ijlk", (chi, D, D, chi), (D, D, D, D)), ("ibd,dcl->ibcl", (chi, D, chi), (chi, D, chi)),
                         ("ibed,bjce->ijcd", (chi, D, D, chi), (D, D, D, D)), ("pq,qj->pj", (chi * D, chi * D), (chi * D, chi))]:
        p = T.contract_plan(spec, sa, sb)
        assert p["a_vec"] == 1 and p["b_vec"] == 1, spec


def test_contract_plan_rejects_bad_specs():
    with pytest.raises(T.TnadError):
        T.contract_plan("ab,bc->ad", (2, 3), (3, 4))
    with pytest.raises(T.TnadError):
        T.contract_plan("ab,bc->ac", (2, 3), (4, 4))


# ---- host mirror of the reference API ---------------------------------------------------------------------
def test_model_builders_match_oracle():
    for b in (0.2, 0.44, 0.9):
        assert np.array_equal(T.model_tensor(T.Ising(), b), O.model_tensor_ising(b))
        assert np.array_equal(T.mag_tensor(T.Ising(), b), O.mag_tensor_ising(b))
        assert np.allclose(T.dmodel_tensor(T.Ising(), b), O.dmodel_tensor_ising(b), atol=1e-15)
        assert T.magofbeta(T.Ising(), b) == O.magofbeta(b)
    assert np.array_equal(T.hamiltonian(T.Heisenberg()), O.hamiltonian_heisenberg())
    assert np.array_equal(T.hamiltonian(T.Heisenberg(2.0, 0.5, 0.5)), O.hamiltonian_heisenberg(2.0, 0.5, 0.5))
    assert np.array_equal(T.hamiltonian(T.TFIsing(0.7)), O.hamiltonian_tfising(0.7))
    assert np.array_equal(T.diaglocalhamiltonian([0.3, 0.1, -0.43]), O.diaglocalhamiltonian([0.3, 0.1, -0.43]))
    assert np.allclose(T.tensorfromclassical([[0.3, -0.3], [-0.3, 0.3]]), T.model_tensor(T.Ising(), 0.3), atol=1e-14)


def test_ipeps_types():
    # test/ctmrg.jl:6-14
    assert isinstance(T.SquareLattice(), T.AbstractLattice)
    assert isinstance(T.IPEPS(np.random.randn(2, 3, 3, 3, 3)), T.IPEPS)
    sq = T.SquareIPEPS(np.random.randn(3, 3, 3, 3, 2))
    assert T.getd(sq) == 3 and T.gets(sq) == 2
    with pytest.raises(T.DimensionMismatch):
        T.SquareIPEPS(np.random.randn(3, 3, 4, 3, 2))
    x = T.indexperm_symmetrize(sq).bulk
    assert np.allclose(x, O.indexperm_symmetrize(sq.bulk)) and np.linalg.norm(x) == pytest.approx(1.0)
    for p in [(0, 3, 2, 1, 4), (2, 1, 0, 3, 4), (1, 0, 3, 2, 4), (3, 2, 1, 0, 4)]:
        assert np.allclose(x, np.transpose(x, p))


def test_runtime_constructor_random_env():
    rt = T.SquareCTMRGRuntime(np.random.randn(2, 2, 2, 2), "random", 10, rng=np.random.default_rng(0))
    assert isinstance(rt, T.CTMRGRuntime) and T.getchi(rt) == 10 and T.getD(rt) == 2
    assert np.allclose(rt.corner, rt.corner.T) and np.allclose(rt.edge, np.transpose(rt.edge, (2, 1, 0)))


def test_fixedpoint_semantics():
    # test/fixedpoint.jl:13-26
    nxt = lambda g: 0.5 * (g + 9 / g)  # noqa: E731
    assert T.fixedpoint(nxt, 9, lambda x: True) == 9

    class Stop2:
        def __init__(self):
            self.c = 0

        def __call__(self, v):
            self.c += 1
            return self.c == 2
    assert T.fixedpoint(nxt, 9, Stop2()) == pytest.approx(0.5 * (9 + 1))
    st = T.StopFunction(np.full(3, np.inf), -1, 0.0, 2)
    n = 0
    state = (None, np.full(3, np.inf))
    while not st(state):
        n += 1
        state = (None, np.arange(3.0) + n)
    assert n == 3                                     # counter from -1 => maxit + 1 applications


def test_num_grad():
    assert T.num_grad(lambda x: x * x, 3.0) == pytest.approx(6.0, rel=1e-6)
    assert np.allclose(T.num_grad(np.trace, np.random.rand(2, 2)), np.eye(2), atol=1e-8)


def test_argument_checks_raise_before_any_gpu_work():
    c = T.Context.__new__(T.Context)            # no device needed: checks run in the Python layer first
    with pytest.raises(T.DimensionMismatch):
        T.Context.energy(c, np.zeros((2, 2, 2, 2)), np.zeros((2, 2, 3, 2, 2)), 4, 0.0, 1)
    with pytest.raises(T.DimensionMismatch):
        T.Context.trg_forward(c, np.zeros((2, 3, 3, 2)), 4, 1)


# ---- N > 1 plumbing (gloo, world_size 2) ----------------------------------------------------------------------
def test_gloo_world2_replica_plumbing(tmp_path):
    script = tmp_path / "w2.py"
    script.write_text(
        "import os, sys, json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "from tnad_b200.sweep import partition\n"
        "d = bench.Dist(backend='gloo')\n"
        "assert d.on and d.world == 2\n"
        "mx = d.max(10.0 + d.rank); sm = d.sum(1.0 + d.rank)\n"
        "mine = partition(7, d.rank, d.world)\n"
        "d.barrier()\n"
        f"open(os.path.join({str(tmp_path)!r}, 'r%d.json' % d.rank), 'w').write(json.dumps({{'rank': d.rank, 'max': mx, 'sum': sm, 'mine': mine}}))\n"
        "d.close()\n")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"r{k}.json")) for k in range(2)]
    assert all(x["max"] == 11.0 and x["sum"] == 3.0 for x in rows)
    got = sorted(i for x in rows for i in x["mine"])
    assert got == list(range(7))                      # every instance exactly once, no overlap


# ---- chi-sharded ctmrgstep: the schedule (data) interpreted with NumPy under gloo, against the oracle --------------
_SHARD_SCRIPT = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, torch.distributed as dist
import tnad_oracle as O
from tnad_b200.sharded import shard_plan, buffer_sizes, step_schedule
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
chi, d = 8, 2
D = d * d
rng = np.random.default_rng(5)
a = rng.standard_normal((d, d, d, d, 2)); a = O.indexperm_symmetrize(a)
bulk = O.double_layer(a)[1]
corner = rng.standard_normal((chi, chi)); corner = corner + corner.T
edge = rng.standard_normal((chi, D, chi)); edge = edge + edge.transpose(2, 1, 0)
p = shard_plan(chi, D, world, rank)
buf = {k: np.zeros(v) for k, v in buffer_sizes(p).items()}
for k, v in (('bulk', bulk), ('corner', corner), ('edge', edge)):
    buf[k][:] = v.ravel(order='F')
view = lambda name, off, dims: buf[name][off:off + int(np.prod(dims))].reshape(dims, order='F')
for op in step_schedule(p):
    kind = op[0]
    if kind == 'contract':
        _, spec, (A, oa, da), (B, ob, db), out = op
        r = np.einsum(spec, view(A, oa, da), view(B, ob, db))
        buf[out][:r.size] = r.ravel(order='F')
    elif kind == 'gather':
        full, part = torch.from_numpy(buf[op[1]]), torch.from_numpy(buf[op[2]])
        dist.all_gather_into_tensor(full, part)
    elif kind == 'permute':
        buf[op[4]][:] = np.transpose(view(op[1], 0, op[2]), op[3]).ravel(order='F')
    elif kind == 'svd_symmetrized':
        n = op[2]
        M = view(op[1], 0, (n, n)); U, S, V = O.svd(M + M.T)
        buf[op[3]][:] = U.ravel(order='F'); buf[op[4]][:] = S; buf[op[5]][:] = V.ravel(order='F')
    elif kind == 'finish':
        c1 = view(op[1], 0, (chi, chi)); e1 = view(op[2], 0, (chi, D, chi))
        c2 = c1 + c1.T; e2 = e1 + e1.transpose(2, 1, 0)
        buf[op[3]][:] = (c2 / np.linalg.norm(c2)).ravel(order='F'); buf[op[4]][:] = (e2 / np.linalg.norm(e2)).ravel(order='F')
    else:
        raise SystemExit('unknown op ' + kind)
c_ref, e_ref, vals = O.ctmrgstep(bulk, corner, edge)
c = view('corner_out', 0, (chi, chi)); e = view('edge_out', 0, (chi, D, chi))
err = max(np.abs(c - c_ref).max(), np.abs(e - e_ref).max(), np.abs(buf['S'] / buf['S'][0] - vals).max())
json.dump({'rank': rank, 'err': float(err), 'width': p.width, 'start': p.start}, open(os.path.join(OUT, 'r%d.json' % rank), 'w'))
dist.barrier(); dist.destroy_process_group()
"""


def test_shard_plan_rejects_uneven_split():
    from tnad_b200.sharded import shard_plan, step_schedule, buffer_sizes
    with pytest.raises(ValueError):
        shard_plan(10, 4, 4, 0)
    with pytest.raises(ValueError):
        shard_plan(8, 4, 2, 2)
    p = shard_plan(256, 25, 8, 3)
    assert (p.width, p.start, p.n) == (32, 96, 6400)
    names = set(buffer_sizes(p))
    for op in step_schedule(p):          # every buffer the schedule touches exists, slices stay inside it
        if op[0] == "contract":
            for nm, off, dims in (op[2], op[3]):
                assert nm in names and off + int(np.prod(dims)) <= buffer_sizes(p)[nm]
            assert op[4] in names


def test_gloo_world2_sharded_ctmrgstep_schedule(tmp_path):
    script = tmp_path / "shard.py"
    script.write_text(f"ROOT = {ROOT!r}\nOUT = {str(tmp_path)!r}\n" + _SHARD_SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"r{k}.json")) for k in range(2)]
    assert [x["start"] for x in rows] == [0, 4] and all(x["width"] == 4 for x in rows)
    assert all(x["err"] < 1e-12 for x in rows), rows


# ---- NumPy statements of the two eigensolver algorithms the kernels implement (tools/*_proto.py) ------------------------
def _load_tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind", ["random", "graded", "wilkinson"])
def test_divide_and_conquer_prototype(kind):
    P = _load_tool("stedc_proto")
    rng = np.random.default_rng(2)
    n = 90
    if kind == "random":
        d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    elif kind == "graded":
        d, e = 10.0 ** (-rng.uniform(0, 16, n)), 10.0 ** (-rng.uniform(0, 16, n - 1))
    else:
        d, e = np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
    lam, Z = P.stedc(d, e, smax=8)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    o = np.argsort(lam)
    ref = np.linalg.eigvalsh(Tm)
    nrm = np.abs(ref).max()
    assert np.abs(lam[o] - ref).max() < 1e-13 * nrm
    assert np.abs(Tm @ Z - Z * lam).max() < 1e-13 * nrm and np.abs(Z.T @ Z - np.eye(n)).max() < 1e-13
    assert P.leaf_size(6400, 64) == (50, 7) and P.leaf_size(2048, 16) == (16, 7)


@pytest.mark.parametrize("kind", ["random", "low_rank", "decaying"])
def test_one_barrier_tridiagonalisation_prototype(kind):
    P = _load_tool("sytrd1b_proto")
    rng = np.random.default_rng(4)
    n = 70
    a = rng.standard_normal((n, n))
    if kind == "random":
        a = a + a.T
    elif kind == "low_rank":
        b = rng.standard_normal((n, 6))
        a = b @ b.T
    else:
        q, _ = np.linalg.qr(a)
        a = (q * 10.0 ** (-np.arange(n) / 4.0)) @ q.T
        a = a + a.T
    P.THETA = 0.1
    P.REDO[0] = 0
    d, e, Vh, tau = P.sytrd_one_barrier(a, nb=16)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    Q = P.form_q(Vh, tau)
    nrm = np.linalg.norm(a, 2)
    assert np.abs(Q.T @ Q - np.eye(n)).max() < 1e-13
    assert np.abs(Q @ Tm @ Q.T - a).max() < 1e-13 * nrm
    if kind == "low_rank":
        assert P.REDO[0] >= 1          # the cancellation guard must fire at the rank boundary


# ---- bench.py contract: the committed result line carries every key the driver reads -------------------------------
@pytest.mark.parametrize("name", ["r1_bench_n1.json", "r2_bench_n1.json"])
def test_recorded_bench_line_has_contract_keys(name):
    import json
    path = os.path.join(ROOT, "profiles", name)
    line = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["dtype"] == "f64" and line["higher_is_better"] is False and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    r = line["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = line["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    assert line["gpu_launches"] > 0 and line["clocks"]["reasons"] == []
    if name.startswith("r2"):     # round 2: same-config CPU sample checked against the GPU in the run, whole metric in the line
        assert line["parity_check"]["ok"] is True and line["contractions"]["frac_of_dmma_peak"] > 0.5
        assert line["trg_chi64_iters_per_s"] > 1.8 and line["trg_chi64"]["parity_vs_oracle_fixture"]["lnZ_rel_err"] < 1e-10


@pytest.mark.parametrize("n", [2, 4, 8])
def test_recorded_multi_gpu_lines(n):
    import json
    line = json.loads(open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")).read().strip().splitlines()[-1])
    assert line["n_gpus"] == n and line["scaling"] == "weak" and line["value"] > 0
    assert line["sweep_instances_per_s"] > 0 and line["sharded_s_per_step"] > 0
    for k in ("sharded_ms_contract", "sharded_ms_gather", "sharded_ms_svd"):
        assert line[k] >= 0


def test_tma_shared_memory_layout_prototype():
    """The fragment addressing of gemm_tma.cu (128-byte swizzle, k permutation, lane -> row assignment): every lane gets
    exactly A[row(blk, gq), k(kk, t)] from both tile layouts, and every 128-bit load is conflict-free per quarter warp."""
    P = _load_tool("tma_layout_proto")
    rng = np.random.default_rng(5)
    tile = rng.standard_normal((64, 16))
    assert sorted(P.kmap(kk, t) for kk in range(4) for t in range(4)) == list(range(16))      # a bijection onto the k-tile
    for kfast in (True, False):
        img = P.tile_image(tile, kfast)
        assert not np.isnan(img).any()                                                        # the layout fills the stage exactly
        for w0 in (0, 32):
            frag, loads = P.warp_fragments(img, kfast, w0)
            for blk in range(4):
                for kk in range(4):
                    for lane in range(32):
                        gq, t = lane >> 2, lane & 3
                        assert frag[blk, kk, lane] == tile[P.rowmap(kfast, w0, blk, gq), P.kmap(kk, t)]
            assert len(loads) == 8 and all(P.wavefronts_128(a) == 4 for a in loads)          # 4 = the minimum for 512 bytes
            rows = sorted(P.rowmap(kfast, w0, blk, gq) for blk in range(4) for gq in range(8))
            assert rows == list(range(w0, w0 + 32))                                           # the warp tile is covered once


def test_golub_kahan_prototype():
    P = _load_tool("gebrd_proto")
    rng = np.random.default_rng(3)
    for m, n, rank in [(40, 40, None), (55, 30, None), (48, 48, 7)]:
        a = rng.standard_normal((m, n)) if rank is None else rng.standard_normal((m, rank)) @ rng.standard_normal((rank, n))
        U, s, V = P.svd_via_gk(a, smax=8)
        ref = np.linalg.svd(a, compute_uv=False)
        r = int(np.sum(ref > 1e-12 * ref[0]))
        assert np.abs(s - ref).max() < 1e-13 * ref[0]
        assert np.abs((U * s) @ V.T - a).max() < 1e-13 * ref[0]
        assert np.abs(U[:, :r].T @ U[:, :r] - np.eye(r)).max() < 1e-10


# ---- two-stage tridiagonalisation: the NumPy statement of the kernels' data flow (tools/twostage_proto.py) ----------
def test_twostage_proto_direct_and_systolic():
    import importlib.util
    spec = importlib.util.spec_from_file_location("twostage_proto", os.path.join(ROOT, "tools", "twostage_proto.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    rng = np.random.default_rng(0)
    # dense -> band -> tridiagonal -> back-transformation, textbook order
    for n, b in [(37, 4), (50, 16), (9, 4), (6, 4)]:
        A = rng.standard_normal((n, n)); A = A + A.T
        lam, U, (Bd, d, e) = P.eigh_twostage(A, b)
        assert np.abs(np.tril(Bd, -b - 1)).max() == 0.0
        assert np.abs(lam - np.linalg.eigvalsh(A)).max() < 1e-12 * np.abs(lam).max()
        assert np.abs(A @ U - U * lam).max() < 1e-12 * np.abs(lam).max() and np.abs(U.T @ U - np.eye(n)).max() < 1e-13
    # the systolic formulation the kernels use (windows, mailboxes, zero padding, staged Q2 pass), b = 32
    for n in (3, 5, 33, 34, 65, 97, 130):
        err_t, res, orth = P.check_systolic(n, rng)
        assert res < 1e-12 and orth < 1e-13, (n, res, orth)
