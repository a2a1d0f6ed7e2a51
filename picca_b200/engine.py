"""Device-side operations of the pair-counting path, one process per GPU.

Thin host code over the C ABI (``include/picca_b200.h``): torch is used only for device memory,
streams and host<->device copies.  Every function here launches hand-written sm_100a kernels from
``libpicca_b200.so``; nothing falls back to the CPU.
"""
import ctypes
import multiprocessing
import os

import numpy as np

from . import _lib, catalog as _catalog

MODE_AUTO, MODE_CROSS, MODE_XCF = 0, 1, 2

_ENGINE = None


def _pick_device():
    """Device for this process: explicit env, torchrun's LOCAL_RANK, else the fork-pool worker
    index (the unchanged scripts fork ``--nproc`` workers, picca_cf.py:454-457)."""
    import torch
    n = torch.cuda.device_count()
    if n == 0:
        raise RuntimeError("picca_b200: no CUDA device visible; there is no CPU fallback")
    if "PICCA_B200_DEVICE" in os.environ:
        return int(os.environ["PICCA_B200_DEVICE"]) % n
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"]) % n
    ident = multiprocessing.current_process()._identity
    if ident:
        return (ident[0] - 1) % n
    return 0


class PairList:
    """CSR forest-pair list on the device (``pb2_pairs``) for the lines of sight ``f1_index``."""

    def __init__(self, engine, f1_index, nb_offset, nb_f1, nb_f2, nb_ang, nb_cos, nb_sin):
        self.engine = engine
        self.f1_index = f1_index
        self.nb_offset = nb_offset
        self.nb_f1, self.nb_f2 = nb_f1, nb_f2
        self.nb_ang, self.nb_cos, self.nb_sin = nb_ang, nb_cos, nb_sin
        self.nb_keep = None
        self.n_f1 = int(f1_index.numel())
        self.n_pairs = int(nb_f2.numel())
        self._host_offset = None
        self._host_f2 = None

    def struct(self):
        s = _lib.Pairs()
        s.n_f1, s.n_pairs = self.n_f1, self.n_pairs
        s.f1_index = self.f1_index.data_ptr()
        s.nb_offset = self.nb_offset.data_ptr()
        s.nb_f1 = self.nb_f1.data_ptr()
        s.nb_f2 = self.nb_f2.data_ptr()
        s.nb_ang = self.nb_ang.data_ptr()
        s.nb_cos = self.nb_cos.data_ptr()
        s.nb_sin = self.nb_sin.data_ptr()
        s.nb_keep = self.nb_keep.data_ptr() if self.nb_keep is not None else None
        return s

    def subset(self, keep):
        """The sub-list of the forest pairs flagged in ``keep`` (bool array / tensor over the
        pairs), same lines of sight: what ``np.array(neighbours)[w]`` selects in cf.py:444-447."""
        import torch
        dev = self.nb_f2.device
        keep = torch.as_tensor(np.asarray(keep, dtype=bool) if not torch.is_tensor(keep) else keep,
                               device=dev).to(torch.bool)
        counts = torch.zeros(self.n_f1, dtype=torch.int64, device=dev)
        counts.index_add_(0, self.nb_f1[keep].to(torch.int64),
                          torch.ones(int(keep.sum().item()), dtype=torch.int64, device=dev))
        offset = torch.zeros(self.n_f1 + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, dim=0, out=offset[1:])
        return PairList(self.engine, self.f1_index, offset, self.nb_f1[keep].contiguous(),
                        self.nb_f2[keep].contiguous(), self.nb_ang[keep].contiguous(),
                        self.nb_cos[keep].contiguous(), self.nb_sin[keep].contiguous())

    def host_offset(self):
        if self._host_offset is None:
            self._host_offset = self.nb_offset.cpu().numpy()
        return self._host_offset

    def host_f2(self):
        if self._host_f2 is None:
            self._host_f2 = self.nb_f2.cpu().numpy()
        return self._host_f2

    def set_host_angles(self, ang):
        """Parity mode: replace the device-computed angles by host (NumPy) ones, including
        cos(ang/2) and sin(ang/2) as the reference's libm evaluates them (cf.py:356-357)."""
        import torch
        ang = np.ascontiguousarray(ang, dtype=np.float64)
        dev = self.nb_ang.device
        self.nb_ang = torch.from_numpy(ang).to(dev)
        self.nb_cos = torch.from_numpy(np.cos(ang / 2)).to(dev)
        self.nb_sin = torch.from_numpy(np.sin(ang / 2)).to(dev)


class Engine:
    def __init__(self, device=None):
        import torch
        self.torch = torch
        self.lib = _lib.lib()
        self.device_index = _pick_device() if device is None else int(device)
        torch.cuda.set_device(self.device_index)
        self.device = torch.device("cuda", self.device_index)
        self._cats = {}

    # ------------------------------------------------------------------ memory
    collect_dmat_stats = False   # bench.py: read pb2_dmat_stats after every dmat launch
    last_dmat_stats = None
    sum_dmat_stats = None        # ... summed over the launches since it was last reset
    dmat_kernel_ms_log = None    # bench.py: list that receives the kernel time of every launch

    def stream_ptr(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def device_catalog(self, host_cat, pin=False, cache=True):
        key = id(host_cat)
        if cache and key in self._cats and self._cats[key].host is host_cat:
            return self._cats[key]
        dev = _catalog.DeviceCatalog(host_cat, self.device, pin=pin)
        if cache:
            self._cats[key] = dev
        return dev

    def drop_catalogs(self):
        self._cats.clear()

    # ------------------------------------------------------------------ neighbours
    def neighbour_counts(self, cat1, cat2, params, mode, f1_index):
        """``len(delta.neighbours)`` of every listed line of sight (int32 device tensor) without
        building the lists: what the --rej draw of a chunk needs from the forests of other
        ranks."""
        torch = self.torch
        if not torch.is_tensor(f1_index):
            f1_index = torch.as_tensor(np.ascontiguousarray(f1_index, dtype=np.int32),
                                       device=self.device)
        n_f1 = int(f1_index.numel())
        count = torch.empty(n_f1, dtype=torch.int32, device=self.device)
        if n_f1:
            _lib.check(self.lib.pb2_neigh_count(
                ctypes.byref(cat1.struct), ctypes.byref(cat2.struct), ctypes.byref(params),
                ctypes.c_int32(mode), ctypes.c_int64(n_f1), ctypes.c_void_p(f1_index.data_ptr()),
                ctypes.c_void_p(count.data_ptr()), self.stream_ptr()), "pb2_neigh_count")
        return count

    def neighbours(self, cat1, cat2, params, mode, f1_index):
        """cf.fill_neighs / xcf.fill_neighs on the device for the lines of sight ``f1_index``
        (int32 tensor on the device or array-like)."""
        torch = self.torch
        if not torch.is_tensor(f1_index):
            f1_index = torch.as_tensor(np.ascontiguousarray(f1_index, dtype=np.int32),
                                       device=self.device)
        n_f1 = int(f1_index.numel())
        count = self.neighbour_counts(cat1, cat2, params, mode, f1_index)
        nb_offset = torch.zeros(n_f1 + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(count, dim=0, out=nb_offset[1:])
        n_pairs = int(nb_offset[-1].item()) if n_f1 else 0
        nb_f1 = torch.empty(n_pairs, dtype=torch.int32, device=self.device)
        nb_f2 = torch.empty(n_pairs, dtype=torch.int32, device=self.device)
        nb_ang = torch.empty(n_pairs, dtype=torch.float64, device=self.device)
        nb_cos = torch.empty(n_pairs, dtype=torch.float64, device=self.device)
        nb_sin = torch.empty(n_pairs, dtype=torch.float64, device=self.device)
        if n_pairs:
            _lib.check(self.lib.pb2_neigh_fill(
                ctypes.byref(cat1.struct), ctypes.byref(cat2.struct), ctypes.byref(params),
                ctypes.c_int32(mode), ctypes.c_int64(n_f1), ctypes.c_void_p(f1_index.data_ptr()),
                ctypes.c_void_p(nb_offset.data_ptr()), ctypes.c_void_p(nb_f1.data_ptr()),
                ctypes.c_void_p(nb_f2.data_ptr()), ctypes.c_void_p(nb_ang.data_ptr()),
                ctypes.c_void_p(nb_cos.data_ptr()), ctypes.c_void_p(nb_sin.data_ptr()),
                self.stream_ptr()), "pb2_neigh_fill")
        return PairList(self, f1_index, nb_offset, nb_f1, nb_f2, nb_ang, nb_cos, nb_sin)

    # ------------------------------------------------------------------ correlation
    def xi(self, cat1, cat2, params, pairs, out_row, n_rows, cross_obj=False, variant=0,
           normalise=False, out=None):
        """Accumulate the pair histograms.  Returns a device tensor [n_rows, 6, nb] (float64 slots;
        slot 5 holds int64 counts)."""
        torch = self.torch
        nb = params.num_bins_r_par * params.num_bins_r_trans
        if out is None:
            out = torch.zeros((n_rows, 6, nb), dtype=torch.float64, device=self.device)
        if not torch.is_tensor(out_row):
            out_row = torch.as_tensor(np.ascontiguousarray(out_row, dtype=np.int32),
                                      device=self.device)
        ps = pairs.struct()
        fn = self.lib.pb2_xi_cross if cross_obj else self.lib.pb2_xi_auto
        _lib.check(fn(ctypes.byref(cat1.struct), ctypes.byref(cat2.struct), ctypes.byref(params),
                      ctypes.byref(ps), ctypes.c_void_p(out_row.data_ptr()),
                      ctypes.c_int64(n_rows), ctypes.c_void_p(out.data_ptr()),
                      ctypes.c_int32(variant), self.stream_ptr()),
                   "pb2_xi_cross" if cross_obj else "pb2_xi_auto")
        if normalise:
            _lib.check(self.lib.pb2_xi_normalise(ctypes.c_int64(n_rows), ctypes.c_int32(nb),
                                                 ctypes.c_void_p(out.data_ptr()),
                                                 self.stream_ptr()), "pb2_xi_normalise")
        return out

    # ------------------------------------------------------------------ distortion matrix
    def dmat_outputs(self, params):
        """Zeroed accumulators (weights_dmat[nb], dmat[nb, nbm], r_par_eff, r_trans_eff, z_eff,
        weight_eff [nbm]) on the device."""
        torch = self.torch
        nb = params.num_bins_r_par * params.num_bins_r_trans
        nbm = params.num_model_bins_r_par * params.num_model_bins_r_trans
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.device)
        return z(nb), z(nb, nbm), z(nbm), z(nbm), z(nbm), z(nbm)

    def dmat(self, cat1, cat2, params, pairs, cross_obj=False, out=None):
        """Accumulate the distortion matrix over the kept pairs into ``out`` (default: fresh
        zeroed accumulators).  Returns device tensors (weights_dmat[nb], dmat[nb, nbm],
        r_par_eff, r_trans_eff, z_eff, weight_eff [nbm])."""
        torch = self.torch
        weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff = \
            out if out is not None else self.dmat_outputs(params)
        nbytes = int(self.lib.pb2_dmat_scratch_bytes(
            ctypes.byref(cat1.struct), ctypes.byref(cat2.struct), ctypes.byref(params),
            ctypes.c_int32(int(cross_obj)))) + 8 * pairs.n_pairs + 256
        scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)
        ps = pairs.struct()
        fn = self.lib.pb2_dmat_cross if cross_obj else self.lib.pb2_dmat_auto
        _lib.check(fn(ctypes.byref(cat1.struct), ctypes.byref(cat2.struct), ctypes.byref(params),
                      ctypes.byref(ps), ctypes.c_void_p(weights_dmat.data_ptr()),
                      ctypes.c_void_p(dmat.data_ptr()), ctypes.c_void_p(r_par_eff.data_ptr()),
                      ctypes.c_void_p(r_trans_eff.data_ptr()), ctypes.c_void_p(z_eff.data_ptr()),
                      ctypes.c_void_p(weight_eff.data_ptr()), ctypes.c_void_p(scratch.data_ptr()),
                      ctypes.c_int64(nbytes), self.stream_ptr()),
                   "pb2_dmat_cross" if cross_obj else "pb2_dmat_auto")
        if not cross_obj and self.collect_dmat_stats:
            out3 = (ctypes.c_double * 3)()
            _lib.check(self.lib.pb2_dmat_stats(ctypes.c_void_p(scratch.data_ptr()), out3,
                                               self.stream_ptr()), "pb2_dmat_stats")
            self.last_dmat_stats = {"as_written_ops": out3[0], "sum_unique_model_bins": out3[1],
                                    "in_range_pixel_pairs": out3[2]}
            tot = self.sum_dmat_stats or {}
            self.sum_dmat_stats = {k: tot.get(k, 0.) + v for k, v in self.last_dmat_stats.items()}
        if self.dmat_kernel_ms_log is not None:
            self.dmat_kernel_ms_log.append(float(self.lib.pb2_last_kernel_ms()))
        return weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff

    # ------------------------------------------------------------------ measurement
    def fp64_peak(self, iters=4096):
        ops = ctypes.c_double(0.)
        ms = ctypes.c_double(0.)
        _lib.check(self.lib.pb2_fp64_peak(ctypes.c_int32(iters), ctypes.byref(ops),
                                          ctypes.byref(ms)), "pb2_fp64_peak")
        return ops.value, ms.value

    def launch_count(self):
        return int(self.lib.pb2_launch_count())


def get_engine():
    """Process-wide engine, created lazily (never before a fork: CUDA must not be initialised in
    the parent of the scripts' fork pools)."""
    global _ENGINE
    if _ENGINE is None or _ENGINE._pid != os.getpid():
        _ENGINE = Engine()
        _ENGINE._pid = os.getpid()
    return _ENGINE
