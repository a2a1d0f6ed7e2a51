"""Generate the golden vectors by running the LIVE reference (its Numba path, imported unmodified
from /root/reference through tests/refharness) on the seeded cases of ``cases.py``.

    python -m tests.golden.make_golden          # writes tests/golden/golden_*.npz

Only works where /root/reference exists; the resulting small .npz files are committed and travel
to the GPU box.  Host of record: see ``host`` inside each file (NumPy's arccos differs in the last
ulp between SIMD back-ends).
"""
import os
import platform
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.refharness import load  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def checksum(data):
    tot = 0.
    for hp in sorted(data):
        for d in data[hp]:
            for name in ("weights", "delta", "z", "r_comov", "dist_m", "log_lambda"):
                v = getattr(d, name, None)
                if v is not None:
                    tot += float(np.sum(np.asarray(v, dtype=np.float64)))
    return tot


def neighbour_ids(data, hps):
    counts, ids = [], []
    for hp in hps:
        for d in data[hp]:
            counts.append(len(d.neighbours))
            ids.extend(int(o.thingid) for o in d.neighbours)
    return np.array(counts, dtype=np.int64), np.array(ids, dtype=np.int64)


def run_cf(out):
    for name, cfg in cases.CF_CASES.items():
        cf, _, _, _, _ = load.reference_modules()
        cf.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        second = cfg.pop("second", False)
        plates = cfg.pop("plates", False)
        data, num, z_min, cosmo = cases.forests(plates=plates)
        rdata = load.to_reference_deltas(data)
        over = dict(cfg)
        z_min2 = None
        if second:
            data2, num2, z_min2, _ = cases.forests(second=True, plates=plates)
            over["data2"] = load.to_reference_deltas(data2)
            over["num_data2"] = num2
        helpers.configure(cf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
        hps = sorted(rdata)
        rows, counts, ids = [], [], []
        for hp in hps:
            cf.fill_neighs([hp])
            c, i = neighbour_ids(rdata, [hp])
            counts.append(c)
            ids.append(i)
            res = cf.compute_xi([hp])
            rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:5]] +
                                 [np.asarray(res[5], dtype=np.int64).view(np.float64)]))
        out["cf_%s" % name] = np.stack(rows)
        out["cf_%s_nbcount" % name] = np.concatenate(counts)
        out["cf_%s_nbid" % name] = np.concatenate(ids)
        out["cf_%s_checksum" % name] = np.array([checksum(data)])
        print("cf", name, "pairs", int(np.stack(rows)[:, 5].view(np.int64).sum()))


def pack8(res):
    return dict(weights_dmat=res[0], dmat=res[1], r_par_eff=res[2], r_trans_eff=res[3],
                z_eff=res[4], weight_eff=res[5], counts=np.array([res[6], res[7]], dtype=np.int64))


def run_dmat(out):
    for name, cfg in cases.DMAT_CASES.items():
        cf, _, _, _, _ = load.reference_modules()
        cf.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        second = cfg.pop("second", False)
        data, num, z_min, cosmo = cases.dmat_forests()
        rdata = load.to_reference_deltas(data)
        over = dict(cfg)
        z_min2 = None
        if second:
            data2, num2, z_min2, _ = cases.dmat_forests(second=True)
            over["data2"] = load.to_reference_deltas(data2)
            over["num_data2"] = num2
        helpers.configure(cf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
        hps = sorted(rdata)
        cf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_dmat.py:36
        res = cf.compute_dmat(hps)
        for key, val in pack8(res).items():
            out["dmat_%s_%s" % (name, key)] = np.asarray(val)
        print("dmat", name, "pairs", res[6], "used", res[7])


def run_metal(out):
    for name, cfg in cases.METAL_CASES.items():
        cf, _, _, _, _ = load.reference_modules()
        cf.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        second = cfg.pop("second", False)
        pair = cfg.pop("pair")
        data, num, z_min, cosmo = cases.forests()
        rdata = load.to_reference_deltas(data)
        over = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo)
        z_min2 = None
        if second:
            data2, num2, z_min2, _ = cases.forests(second=True)
            over["data2"] = load.to_reference_deltas(data2)
            over["num_data2"] = num2
        helpers.configure(cf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
        for k, v in over.items():
            setattr(cf, k, v)
        hps = sorted(rdata)
        cf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_metal_dmat.py:48
        res = cf.compute_metal_dmat(hps, abs_igm1=pair[0], abs_igm2=pair[1])
        for key, val in pack8(res).items():
            out["metal_%s_%s" % (name, key)] = np.asarray(val)
        print("metal", name, "pairs", res[6], "used", res[7], "sum", res[1].sum())


def run_xcf(out):
    for name, cfg in cases.XCF_CASES.items():
        _, xcf, _, _, _ = load.reference_modules()
        xcf.userprint = lambda *a, **k: None
        data, num, z_min, cosmo = cases.forests()
        objs, z_min2 = cases.quasars(cosmo)
        rdata, robjs = load.to_reference_deltas(data), load.to_reference_qsos(objs)
        helpers.configure(xcf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2),
                          objs=robjs, **cfg)
        hps = sorted(rdata)
        rows, counts, ids = [], [], []
        for hp in hps:
            xcf.fill_neighs([hp])
            c, i = neighbour_ids(rdata, [hp])
            counts.append(c)
            ids.append(i)
            res = xcf.compute_xi([hp])
            rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:5]] +
                                 [np.asarray(res[5], dtype=np.int64).view(np.float64)]))
        out["xcf_%s" % name] = np.stack(rows)
        out["xcf_%s_nbcount" % name] = np.concatenate(counts)
        out["xcf_%s_nbid" % name] = np.concatenate(ids)
        print("xcf", name, "pairs", int(np.stack(rows)[:, 5].view(np.int64).sum()))


def run_xmetal(out):
    for name, cfg in cases.XMETAL_CASES.items():
        _, xcf, _, constants, _ = load.reference_modules()
        constants.ABSORBER_IGM.update(cases.EXTRA_ABSORBERS)
        xcf.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        abs_igm = cfg.pop("abs_igm")
        data, num, z_min, cosmo = cases.forests()
        objs, z_min2 = cases.quasars(cosmo)
        rdata, robjs = load.to_reference_deltas(data), load.to_reference_qsos(objs)
        over = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo)
        helpers.configure(xcf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2),
                          objs=robjs, **over)
        for k, v in over.items():
            setattr(xcf, k, v)
        hps = sorted(rdata)
        xcf.fill_neighs(hps)
        np.random.seed(hps[0])
        res = xcf.compute_metal_dmat(hps, abs_igm=abs_igm)
        for key, val in pack8(res).items():
            out["xmetal_%s_%s" % (name, key)] = np.asarray(val)
        kept = sum(d.neighbours is not None for hp in hps for d in rdata[hp])
        out["xmetal_%s_skipped" % name] = np.array([kept])
        print("xmetal", name, "pairs", res[6], "used", res[7], "sum", res[1].sum(), "skipped", kept)


def run_co(out):
    from picca_b200 import synth
    for name, cfg in cases.CO_CASES.items():
        assert shims_install()
        import picca.co
        import importlib
        co = importlib.reload(picca.co)
        co.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        second = cfg.pop("second", False)
        cosmo = synth.FlatLCDM()
        objs, z_min = cases.quasars(cosmo)
        robjs = load.to_reference_qsos(objs)
        for k, v in cfg.items():
            setattr(co, k, v)
        co.objs, co.objs2, z_min2 = robjs, None, None
        if second:
            objs2, z_min2 = cases.quasars2(cosmo)
            co.objs2 = load.to_reference_qsos(objs2)
        co.ang_max = cases.ang_max_for(cosmo, cfg, z_min, z_min2)
        co.nside = 16
        co.num_data = sum(len(v) for v in robjs.values())
        co.lock, co.counter = load.DummyLock(), load.DummyCounter()
        rows = []
        for hp in sorted(robjs):
            co.fill_neighs([hp])
            res = co.compute_xi([hp])
            rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:4]] +
                                 [np.asarray(res[4], dtype=np.int64).view(np.float64)]))
        out["co_%s" % name] = np.stack(rows)
        print("co", name, "pairs", int(np.stack(rows)[:, 4].view(np.int64).sum()))


def shims_install():
    from tests.refharness import shims
    return shims.install()


def run_xdmat(out):
    for name, cfg in cases.XDMAT_CASES.items():
        _, xcf, _, _, _ = load.reference_modules()
        xcf.userprint = lambda *a, **k: None
        data, num, z_min, cosmo = cases.dmat_forests()
        objs, z_min2 = cases.quasars(cosmo)
        rdata, robjs = load.to_reference_deltas(data), load.to_reference_qsos(objs)
        helpers.configure(xcf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2),
                          objs=robjs, **cfg)
        hps = sorted(rdata)
        xcf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_xdmat.py:37
        res = xcf.compute_dmat(hps)
        for key, val in pack8(res).items():
            out["xdmat_%s_%s" % (name, key)] = np.asarray(val)
        print("xdmat", name, "pairs", res[6], "used", res[7])


def pack_wick(res, n_t):
    out = dict(weights_wick=res[0], num_pairs_wick=np.asarray(res[1], dtype=np.int64),
               counts=np.array([res[2], res[3]], dtype=np.int64))
    for k in range(n_t):
        out["t%d" % (k + 1)] = res[4 + k]
    return out


def run_wick(out):
    for name, cfg in cases.WICK_CASES.items():
        cf, _, _, _, _ = load.reference_modules()
        cf.userprint = lambda *a, **k: None
        cfg = dict(cfg)
        second = cfg.pop("second", False)
        data, num, z_min, cosmo = cases.dmat_forests()
        rdata = cases.set_fname(load.to_reference_deltas(data), "D1")
        var1, xi1 = cases.wick_1d("D1")
        over = dict(cfg, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1})
        z_min2 = None
        if second:
            data2, num2, z_min2, _ = cases.dmat_forests(second=True)
            over["data2"] = cases.set_fname(load.to_reference_deltas(data2), "D2")
            over["num_data2"] = num2
            var2, xi2 = cases.wick_1d("D2")
            over["get_variance_1d"]["D2"], over["xi_1d"]["D2"] = var2, xi2
        helpers.configure(cf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
        for k, v in over.items():
            setattr(cf, k, v)
        hps = sorted(rdata)
        cf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_wick.py:35
        res = cf.compute_wick_terms(hps)
        for key, val in pack_wick(res, 3).items():
            out["wick_%s_%s" % (name, key)] = np.asarray(val)
        print("wick", name, "forests", res[2], "used", res[3], "pairs", int(res[1].sum()),
              "t3", np.abs(res[6]).sum())


def run_xwick(out):
    for name, cfg in cases.XWICK_CASES.items():
        _, xcf, _, _, _ = load.reference_modules()
        xcf.userprint = lambda *a, **k: None
        data, num, z_min, cosmo = cases.xwick_forests()
        objs, z_min2 = cases.quasars(cosmo)
        rdata = cases.set_fname(load.to_reference_deltas(data), "D1")
        robjs = load.to_reference_qsos(objs)
        var1, xi1 = cases.wick_1d("D1")
        over = dict(cfg, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1}, xi_wick=None)
        helpers.configure(xcf, rdata, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2),
                          objs=robjs, **over)
        for k, v in over.items():
            setattr(xcf, k, v)
        hps = sorted(rdata)
        xcf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_xwick.py:36
        res = xcf.compute_wick_terms(hps)
        for key, val in pack_wick(res, 4).items():
            out["xwick_%s_%s" % (name, key)] = np.asarray(val)
        print("xwick", name, "forests", res[2], "used", res[3], "pairs", int(res[1].sum()),
              "t4", np.abs(res[7]).sum())


def main():
    todo = (("cf", run_cf), ("dmat", run_dmat), ("xcf", run_xcf), ("xdmat", run_xdmat),
            ("metal", run_metal), ("xmetal", run_xmetal), ("co", run_co), ("wick", run_wick),
            ("xwick", run_xwick))
    only = sys.argv[1:]
    for tag, fn in todo:
        if only and tag not in only:
            continue
        out = {"host": np.array([platform.processor() + " numpy " + np.__version__])}
        fn(out)
        np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % tag), **out)


if __name__ == "__main__":
    main()
