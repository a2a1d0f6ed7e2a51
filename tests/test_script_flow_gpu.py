"""GPU: the call pattern of the reference's scripts, reproduced without the reference (which is
not on the GPU box): module globals assigned in the parent, a FORK pool created after the data is
loaded, one HEALPix pixel per task (picca_cf.py:449-473), dmat chunks round-robin with
np.random.seed(healpixs[0]) per chunk (picca_dmat.py:21-38, 471-501).  CUDA must not be touched in
the parent before the fork; every worker binds a device lazily."""
import multiprocessing

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases

pytestmark = pytest.mark.gpu


def corr_func(healpixs):
    """verbatim shape of picca_cf.py:21-37"""
    from picca_b200 import cf
    cf.fill_neighs(healpixs)
    return cf.compute_xi(healpixs)


def calc_dmat(healpixs):
    """verbatim shape of picca_dmat.py:21-38"""
    from picca_b200 import cf
    cf.fill_neighs(healpixs)
    np.random.seed(healpixs[0])
    return cf.compute_dmat(healpixs)


def _in_subprocess(target, queue):
    queue.put(target())


def _cf_flow():
    from multiprocessing import Lock, Value
    from picca_b200 import cf
    cfg = cases.CF_CASES["default"]
    data, num, z_min, cosmo = cases.forests()
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    cf.counter = Value("i", 0)
    cf.lock = Lock()
    cpu_data = {hp: [hp] for hp in data}
    context = multiprocessing.get_context("fork")
    pool = context.Pool(processes=2)
    out = pool.map(corr_func, sorted(cpu_data.values()))
    pool.close()
    pool.join()
    stacked = np.array(out)                      # picca_cf.py:466
    return stacked, cf.counter.value, num


def test_cf_fork_pool_matches_reference_golden(tmp_path):
    # run the whole script-like flow in a fresh process so that this pytest process (which has
    # already initialised CUDA in other tests) is not the forking parent
    ctx = multiprocessing.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_in_subprocess, args=(_cf_flow, q))
    p.start()
    stacked, counted, num = q.get(timeout=600)
    p.join()
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cf.npz"))["cf_default"]
    assert stacked.shape == gold.shape                      # (n_healpix, 6, nb)
    assert counted == num                                   # shared progress counter advanced
    assert np.array_equal(stacked[:, 5, :].astype(np.int64), gold[:, 5, :].view(np.int64))
    for k in range(5):
        np.testing.assert_allclose(stacked[:, k, :], gold[:, k, :], rtol=1e-9, atol=1e-300)


def _dmat_flow():
    from multiprocessing import Lock, Value
    from picca_b200 import cf
    cfg = cases.DMAT_CASES["default"]
    data, num, z_min, cosmo = cases.dmat_forests()
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    cf.counter = Value("i", 0)
    cf.lock = Lock()
    nproc = 2
    cpu_data = {}
    for index, healpix in enumerate(sorted(data)):     # picca_dmat.py:471-476
        cpu_data.setdefault(index % nproc, []).append(healpix)
    context = multiprocessing.get_context("fork")
    pool = context.Pool(processes=nproc)
    dmat_data = pool.map(calc_dmat, sorted(cpu_data.values()))
    pool.close()
    pool.join()
    merged = [np.array([item[k] for item in dmat_data]).sum(axis=0) for k in range(8)]  # :494-501
    chunks = sorted(cpu_data.values())
    return merged, chunks


def test_dmat_fork_pool_chunks_match_oracle():
    """--nproc 2: each chunk reseeds with its first HEALPix id, so the result differs from the
    single-chunk golden; compare with the oracle run chunk by chunk in this process."""
    ctx = multiprocessing.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_in_subprocess, args=(_dmat_flow, q))
    p.start()
    merged, chunks = q.get(timeout=600)
    p.join()
    from oracle import cf as ocf
    cfg = cases.DMAT_CASES["default"]
    data, num, z_min, cosmo = cases.dmat_forests()
    helpers.configure(ocf, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    want = None
    for chunk in chunks:
        ocf.fill_neighs(chunk)
        np.random.seed(chunk[0])
        res = ocf.compute_dmat(chunk)
        want = list(res) if want is None else [a + b for a, b in zip(want, res)]
    assert int(merged[6]) == int(want[6]) and int(merged[7]) == int(want[7])
    for k in range(6):
        scale = np.abs(want[k]).max()
        np.testing.assert_allclose(merged[k], want[k], rtol=1e-9, atol=1e-12 * scale)
