"""End-to-end example workloads (gmg needs transpose, diagonal, the COO
constructor, scalar multiply and chained SpGEMM; pde = CG on Poisson)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
EX = os.path.join(ROOT, "examples")


def test_gmg_preconditioned_cg_converges_fast():
    sys.path.insert(0, EX)
    import gmg

    iters, rel = gmg.solve(128, 4, rtol=1e-10, maxiter=100, verbose=False)
    assert rel < 1e-9
    assert iters < 40          # plain CG needs several hundred iterations on this grid
    # Galerkin operator check against scipy on a small grid
    import scipy.sparse as sp
    import legate_sparse as sparse
    from _common import poisson2d

    A = poisson2d(sparse, 16)
    R, nc = gmg.full_weighting(16)
    Ac = R @ A @ (R.T * 4.0)
    Rs, As = R.toscipy(), A.toscipy()
    E = (Rs @ As @ (Rs.T * 4.0)).tocsr()
    assert np.allclose(Ac.todense(), np.asarray(E.todense()), rtol=1e-12, atol=1e-13)


def test_example_clis_run():
    env = dict(os.environ)
    for cmd in (["spmv_microbenchmark.py", "--nmin", "64k", "--nmax", "128k", "-i", "5"],
                ["spgemm_microbenchmark.py", "-n", "64k", "-i", "2"],
                ["pde.py", "-n", "64"], ["pde.py", "-n", "64", "--throughput", "-m", "50"]):
        out = subprocess.run([sys.executable, os.path.join(EX, cmd[0])] + cmd[1:], capture_output=True, text=True,
                             timeout=600, cwd=EX, env=env)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "relative residual" in out.stdout
