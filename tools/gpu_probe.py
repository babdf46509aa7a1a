"""Bottom-up GPU probe (not a pytest file): runs every layer against the oracle, never stops at the
first failure, and writes a log to gpurun_out/probe.log.  Used during bring-up; the pytest files
under tests/ are the parity tests proper.

    python tests/gpu_probe.py [quick|full]
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T          # noqa: E402
import tnad_oracle as O        # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "w")
RESULTS = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def check(name, fn):
    t = time.time()
    try:
        ok, info = fn()
    except Exception as e:  # noqa: BLE001
        ok, info = False, "EXC " + repr(e) + "\n" + traceback.format_exc()
    RESULTS.append((name, ok))
    log(("PASS" if ok else "FAIL"), name, f"[{time.time() - t:.2f}s]", info)


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def main(mode):
    ctx = T.Context(0)
    rng = np.random.default_rng(0)

    # ---- 1. contractions -----------------------------------------------------------------
    cases = [
        ("ab,bc->ac", (5, 7), (7, 3)),
        ("ab,bc->ac", (128, 64), (64, 128)),
        ("ab,bc->ac", (257, 130), (130, 191)),
        ("ab,cb->ac", (200, 96), (150, 96)),
        ("ba,bc->ac", (96, 200), (96, 150)),
        ("ba,cb->ac", (96, 200), (150, 96)),
        ("ab,bc->ac", (512, 512), (512, 512)),
        ("iba,ad->ibd", (6, 4, 6), (6, 6)),
        ("ibcl,jkcb->ijlk", (20, 4, 4, 20), (4, 4, 4, 4)),
        ("ibcl,jkcb->ijlk", (7, 3, 3, 7), (3, 3, 3, 3)),
        ("abi,aed->ibed", (20, 4, 20), (20, 4, 20)),
        ("ibed,bjce->ijcd", (20, 4, 4, 20), (4, 4, 4, 4)),
        ("ijcd,dck->ijk", (20, 4, 4, 20), (20, 4, 20)),
        ("icde,cjfdlm->iejflm", (10, 4, 4, 10), (4, 4, 4, 4, 2, 2)),
        ("iejflm,efk->ijklm", (10, 10, 4, 4, 2, 2), (10, 4, 10)),
        ("abcij,ij->abc", (10, 4, 10, 2, 2), (2, 2)),
        ("abc,ij->abcij", (10, 4, 10), (2, 2)),
        ("npu,por->nour", (9, 7, 5), (7, 9, 6)),
        ("nour,dlno->urdl", (9, 9, 5, 6), (5, 6, 9, 9)),
        ("mk,m->k", (37, 11), (37,)),
        ("mk,k->m", (37, 11), (11,)),
        ("ibcl,jkcb->ijlk", (128, 16, 16, 128), (16, 16, 16, 16)),
    ]
    for spec, sa, sb in cases:
        def f(spec=spec, sa=sa, sb=sb):
            A, B = rng.standard_normal(sa), rng.standard_normal(sb)
            ref = np.einsum(spec, A, B, optimize=True)
            got = ctx.contract(spec, A, B)
            e = relerr(got, ref)
            C0 = rng.standard_normal(ref.shape)
            got2 = ctx.contract(spec, A, B, alpha=-0.5, beta=2.0, Cin=C0)
            e2 = relerr(got2, -0.5 * ref + 2.0 * C0)
            return e < 1e-13 and e2 < 1e-13, f"rel {e:.2e} {e2:.2e}"
        check(f"contract {spec} {sa} {sb}", f)

    # ---- 2. SVD ---------------------------------------------------------------------------
    def svd_case(A, sym=False):
        U, S, V = ctx.svd(A)
        k = min(A.shape)
        Sref = np.linalg.svd(A, compute_uv=False)
        rec = relerr((U * S) @ V.T, A)
        ou = np.abs(U.T @ U - np.eye(k)).max()
        ov = np.abs(V.T @ V - np.eye(k)).max()
        es = np.abs(S - Sref).max() / Sref[0]
        ok = rec < 2e-13 and ou < 1e-12 and ov < 1e-12 and es < 1e-13 and np.all(np.diff(S) <= 0)
        return ok, f"rec {rec:.1e} orthU {ou:.1e} orthV {ov:.1e} dS {es:.1e} sweeps {ctx.last_sweeps}"

    for shp in [(4, 4), (9, 9), (30, 30), (64, 64), (80, 80), (100, 60), (60, 100), (128, 128), (200, 200), (400, 400)]:
        check(f"svd random {shp}", lambda shp=shp: svd_case(rng.standard_normal(shp)))

    def rankdef():
        B = rng.standard_normal((50, 7))
        return svd_case(B @ rng.standard_normal((7, 50)))
    check("svd rank-deficient 50x50 rank 7", rankdef)

    def ising_mat():
        a = O.model_tensor_ising(0.5)
        m = np.reshape(np.transpose(a, (2, 1, 0, 3)), (4, 4), order="F")
        return svd_case(m)
    check("svd ising 4x4 (rank 2, exact zeros)", ising_mat)

    def symm_pm():
        Q, _ = np.linalg.qr(rng.standard_normal((96, 96)))
        lam = np.concatenate([np.linspace(1, 2, 40), -np.linspace(1, 2, 40), np.zeros(16)])
        return svd_case((Q * lam) @ Q.T)
    check("svd symmetric with +-pairs and null space", symm_pm)

    # ---- 3. trg_svd / svd_back ------------------------------------------------------------
    def trgsvd():
        t = rng.standard_normal((6, 5, 4, 7))
        u, v = ctx.trg_svd(t, 100, 0.0)
        rec = relerr(np.einsum("ija,akl->ijkl", u, v), t)
        u2, v2 = ctx.trg_svd(t, 8, 1e-16)
        uo, vo, _ = O.trg_svd(t, 8, 1e-16)
        rec2 = relerr(np.einsum("ija,akl->ijkl", u2, v2), np.einsum("ija,akl->ijkl", uo, vo))
        return rec < 1e-13 and rec2 < 1e-12 and u2.shape == uo.shape, f"rec {rec:.1e} trunc {rec2:.1e} {u2.shape}"
    check("trg_svd", trgsvd)

    for (m, n) in [(6, 3), (3, 6), (5, 5), (40, 40)]:
        def sb(m=m, n=n):
            A = rng.standard_normal((m, n))
            U, S, V = O.svd(A)
            k = min(m, n)
            dU, dS, dV = rng.standard_normal((m, k)), rng.standard_normal(k), rng.standard_normal((n, k))
            errs = []
            for mask in [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1)]:
                a = [x if f else None for x, f in zip((dU, dS, dV), mask)]
                ref = O.svd_back(U, S, V, *a)
                got = ctx.svd_back(U, S, V, *a)
                errs.append(relerr(got, ref))
            return max(errs) < 1e-12, "rel " + " ".join(f"{e:.1e}" for e in errs)
        check(f"svd_back {m}x{n}", sb)

    # ---- 4. TRG ---------------------------------------------------------------------------
    def trg_case(beta, chi, niter, golden=None, gold_g=None):
        a = O.model_tensor_ising(beta)
        ref, gref = O.trg_value_and_grad(a, chi, niter)
        t0 = time.time()
        lnz, g = T.trg_value_and_grad(a, chi, niter, ctx=ctx)
        dt = time.time() - t0
        db = float(np.sum(g * O.dmodel_tensor_ising(beta)))
        dbref = float(np.sum(gref * O.dmodel_tensor_ising(beta)))
        e1 = abs(lnz - ref) / abs(ref)
        e2 = abs(db - dbref) / abs(dbref)
        info = f"lnZ {lnz!r} ref {ref!r} rel {e1:.1e}; dbeta {db!r} ref {dbref!r} rel {e2:.1e}; {dt:.2f}s"
        if golden is not None:
            info += f"; golden rel {abs(lnz - golden) / abs(golden):.1e}"
        if gold_g is not None:
            info += f"; golden grad rel {abs(db - gold_g) / abs(gold_g):.1e}"
        return e1 < 1e-10 and e2 < 1e-8, info
    check("trg ising 0.4 chi5 n5", lambda: trg_case(0.4, 5, 5, 0.8919788686747141))
    check("trg ising 0.5 chi5 n5", lambda: trg_case(0.5, 5, 5, None, 1.7502426939979507))
    check("trg ising 0.5 chi20 n20 (C1)", lambda: trg_case(0.5, 20, 20, 1.0257933734351765, 1.7455677143228514))

    # ---- 5. CTMRG ---------------------------------------------------------------------------
    def step_case(D, chi):
        bulk = rng.standard_normal((D, D, D, D))
        bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
        c, e = O.init_random(bulk, chi, rng)
        cr, er, vr = O.ctmrgstep(bulk, c, e)
        cg, eg, vg = ctx.ctmrgstep(bulk, c, e)
        ev = np.abs(vg - vr).max()
        ec = np.abs(np.abs(cg) - np.abs(cr)).max()
        ee = np.abs(np.abs(eg) - np.abs(er)).max()
        return ev < 1e-12 and ec < 1e-11 and ee < 1e-11, f"vals {ev:.1e} |corner| {ec:.1e} |edge| {ee:.1e}"
    check("ctmrgstep D=2 chi=5", lambda: step_case(2, 5))
    check("ctmrgstep D=3 chi=10", lambda: step_case(3, 10))
    check("ctmrgstep D=4 chi=20", lambda: step_case(4, 20))

    def mag_case(beta, chi, seed):
        a, m = O.model_tensor_ising(beta), O.mag_tensor_ising(beta)
        c0, e0 = O.init_raw(a, chi)
        cg, eg = ctx.ctmrg_init_raw(a, chi)
        e_init = max(np.abs(cg - c0).max(), np.abs(eg - e0).max())
        c0, e0 = O.init_random(a, chi, np.random.default_rng(seed))      # :random breaks the Z2 symmetry
        co, eo, vo, no = O.ctmrg(a, c0, e0, 1e-10, 300)
        cg, eg, vg, ng = ctx.ctmrg(a, c0, e0, 1e-10, 300)
        mo = O.magnetisation_readout(a, m, co, eo)
        mg = ctx.magnetisation_readout(a, m, cg, eg)
        return (e_init < 1e-14 and abs(mo - mg) < 1e-8 and abs(mg - O.magofbeta(beta)) < 1e-4), \
            f"init {e_init:.1e} mag {mg!r} oracle {mo!r} onsager {O.magofbeta(beta)!r} steps {ng}/{no}"
    check("ctmrg ising beta=0.6 chi=8 magnetisation", lambda: mag_case(0.6, 8, 5))
    check("ctmrg ising beta=0.3 chi=16 magnetisation", lambda: mag_case(0.3, 16, 5))

    # ---- 6. energy + gradient ------------------------------------------------------------------
    h = O.hamiltonian_heisenberg()

    def energy_case(d, chi, tol, maxit, s=2, seed=0):
        A = O.indexperm_symmetrize(np.random.default_rng(seed).standard_normal((d, d, d, d, s)))
        info = {}
        t0 = time.time()
        yo, go = O.energy_value_and_grad(h, A, chi, tol, maxit, info=info)
        t1 = time.time()
        yg, gg = T.energy_and_gradient(h, A, chi, tol, maxit, ctx=ctx)
        t2 = time.time()
        e1 = abs(yg - yo) / abs(yo)
        e2 = relerr(gg, go)
        tm = ctx.last_timing()
        return (e1 < 1e-10 and e2 < 1e-8 and ctx.last_steps == info["nsteps"],
                f"E {yg!r} oracle {yo!r} rel {e1:.1e}; grad rel {e2:.1e}; steps {ctx.last_steps}/{info['nsteps']}; "
                f"cpu {t1 - t0:.2f}s gpu {t2 - t1:.2f}s timing {tm}")
    check("energy d=2 chi=4 tol=0 maxit=10", lambda: energy_case(2, 4, 0.0, 10))
    check("energy d=2 chi=20 tol=1e-6 maxit=100 (C3)", lambda: energy_case(2, 20, 1e-6, 100))
    check("energy d=3 chi=12 tol=0 maxit=6", lambda: energy_case(3, 12, 0.0, 6))

    # ---- 7. sizes of the headline config -----------------------------------------------------------
    if mode == "full":
        def big_svd(n):
            A = rng.standard_normal((n, n))
            A = A + A.T
            t0 = time.time()
            U, S, V = ctx.svd(A)
            dt = time.time() - t0
            rec = relerr((U * S) @ V.T, A)
            ou = np.abs(U.T @ U - np.eye(n)).max()
            return rec < 5e-13 and ou < 1e-12, f"n={n} {dt:.3f}s (incl. PCIe) sweeps {ctx.last_sweeps} rec {rec:.1e} orth {ou:.1e}"
        check("svd symmetric 1024", lambda: big_svd(1024))
        check("svd symmetric 2048", lambda: big_svd(2048))

        def c4(maxit):
            A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))
            t0 = time.time()
            yg, gg = T.energy_and_gradient(h, A, 128, 0.0, maxit, ctx=ctx)
            dt = time.time() - t0
            return True, f"E {yg!r} |g| {np.linalg.norm(gg):.6e} steps {ctx.last_steps} wall {dt:.2f}s timing {ctx.last_timing()} launches {ctx.launch_count()}"
        check("C4 energy+grad d=4 chi=128 maxit=1", lambda: c4(1))
        check("C4 energy+grad d=4 chi=128 maxit=1 (again, warm)", lambda: c4(1))

    nfail = sum(1 for _, ok in RESULTS if not ok)
    log(f"SUMMARY: {len(RESULTS) - nfail} passed, {nfail} failed")
    for name, ok in RESULTS:
        if not ok:
            log("  failed:", name)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "quick"))
