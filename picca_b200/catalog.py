"""Delta / object loader: packs ``dict[healpix] -> list[Delta|QSO]`` (what ``picca.io.read_deltas``
and ``read_objects`` return, reference py/picca/io.py:383-512, :515-614) into SoA + CSR buffers and
places them in HBM.

Layout: lines of sight in ascending-HEALPix, list order -- the iteration order of
the reference's fill_neighs (cf.py:91-122).  Per pixel: r_comov, dist_m, z, weights, delta*weights,
log_lambda (fp64, one contiguous array each, 8-byte elements, forests back to back).  Per line of
sight: CSR offset, unit vector, ra, dec, cos_dec, z_qso, thingid, plate, fiberid, order, HEALPix
row.  Per HEALPix pixel: member range and the bounding cap of its members (used by the device
neighbour search instead of healpy.query_disc).
"""
import operator

import numpy as np

from . import _lib

_PIXEL_FIELDS = ("r_comov", "dist_m", "z", "weights", "delta_w", "z_w", "log_lambda")
_DIAG_FIELDS = ("dg_offset", "dg_count", "dg_rec", "il_offset", "il_rec")  # dg_rec / il_rec: device
DIAG_DUMMY_COL = 1e300   # distance of the dummy pixels around a line of sight (interleaved copy)
DIAG_DUMMY_ROW = 1e299   # ... and after it in the natural-order copy
_LOS_F64 = ("x_cart", "y_cart", "z_cart", "ra", "dec", "cos_dec", "z_qso")
_LOS_I64 = ("thingid", "plate", "fiberid")


class HostCatalog:
    """Packed catalogue in host memory (NumPy)."""

    def __init__(self):
        self.healpixs = None      # sorted list of HEALPix ids
        self.objs = None          # flat list of the original objects, catalogue order
        self.n_los = 0
        self.n_pix = 0
        self.arrays = {}          # name -> ndarray
        self.sorted = 1
        self.max_pix = 0
        self.ids_are_int = True
        self.thingid_remapped = False
        self.meta_deferred = False   # diag metadata / sorted flag still to come from the device
        self.from_soa = False     # packed straight from a registered SoA (forest.register_soa)
        self.is_object = False
        self.il_total = 0         # diagonal-lane copies (see _diag_metadata)
        self.dg_total = 0
        self.dg_lanes = 0
        self.dg_max_pix = 0
        self.dg_ok = 0
        self.dg_reach = 0.0

    def first_of(self, healpix):
        k = self.hp_index[healpix]
        return int(self.arrays["hp_first"][k]), int(self.arrays["hp_first"][k + 1])

    def nbytes(self):
        return int(sum(a.nbytes for a in self.arrays.values()))


def _as_int64(values):
    """plate / fiberid / thingid columns; string ids (combined re-observations) are flagged."""
    try:
        arr = np.array(values)
        if arr.dtype.kind in "iu":
            return arr.astype(np.int64), True
        if arr.dtype.kind == "f" and np.all(arr == np.floor(arr)):
            return arr.astype(np.int64), True
    except (TypeError, ValueError):
        pass
    return np.zeros(len(values), dtype=np.int64), False


def diag_layout():
    """(lanes, pad, row_pad, chunk_rows) of the packed copies the diagonal-lane xi kernel reads
    (PB2_DIAG_LANES, PB2_DIAG_PAD, PB2_DIAG_ROW_PAD, PB2_DIAG_CHUNK_ROWS of
    include/picca_b200.h)."""
    lanes = int(_lib.lib().pb2_diag_lanes())
    return lanes, 34 * lanes, 8, 32


def _diag_metadata(cat, offset):
    """Geometry of the packed copies the diagonal-lane xi kernel reads (layout:
    include/picca_b200.h, pb2_catalog): per-forest counts of non-zero-weight pixels, record
    offsets of the natural-order and of the interleaved copy, totals, and the flags the launcher
    checks.  The records themselves are written on the device (``pb2_pack_diag``).  Host version
    (``pack`` without ``defer_products``); the product path takes the counts and flags from one
    pass over the SoA in HBM instead (``pb2_catalog_stats``, DeviceCatalog._finish)."""
    A = cat.arrays
    n = cat.n_los
    keep = A["weights"] != 0
    if n and keep.size:
        # non-zero-weight pixels per forest (reduceat repeats an element for an empty segment)
        first_pix = np.minimum(offset[:-1], keep.size - 1)
        count = np.add.reduceat(keep, first_pix, dtype=np.int64)
        count[np.diff(offset) == 0] = 0
    else:
        count = np.zeros(n, np.int64)
    _diag_layout(cat, count)
    fields = ("r_comov", "dist_m", "weights", "delta_w" if "delta_w" in A else "delta", "z")
    # a sum is finite iff every term is (no overflow at these magnitudes): one pass per field,
    # the masked test only when a zero-weight pixel carries the non-finite value
    # (a deferred delta * weights is finite when both factors are, far from overflow)
    finite = all(bool(np.isfinite(A[name].sum())) or bool(np.all(np.isfinite(A[name][keep])))
                 for name in fields)
    cat.dg_ok = int(finite)
    if keep.any() and finite:
        cat.dg_reach = float(max(np.abs(A["r_comov"][keep]).max(),
                                 np.abs(A["dist_m"][keep]).max()))
    else:
        cat.dg_reach = 0.0


def _diag_layout(cat, count):
    """Record offsets and totals of the two packed copies from the per-forest counts."""
    A = cat.arrays
    lanes, pad, row_pad, chunk = diag_layout()
    n = cat.n_los
    count = np.asarray(count, dtype=np.int64)
    first = np.zeros(n + 1, dtype=np.int64)
    first[1:] = np.cumsum(count)
    A["dg_offset"] = np.ascontiguousarray(first[:-1] + row_pad * np.arange(n, dtype=np.int64))
    A["dg_count"] = count.astype(np.int32)
    cat.dg_total = int(first[-1]) + row_pad * n + chunk
    per_plane = (count + lanes - 1) // lanes + 2 * pad // lanes
    il_offset = np.zeros(n + 1, dtype=np.int64)
    il_offset[1:] = np.cumsum(per_plane)
    A["il_offset"] = np.ascontiguousarray(il_offset[:-1])
    cat.il_total = int(il_offset[-1]) + chunk + 64
    cat.dg_lanes = lanes
    cat.dg_max_pix = int(count.max()) if n else 0


def diag_records_host(cat):
    """NumPy statement of what ``pb2_pack_diag`` writes -- the layout specification the tests
    check the device packer against (never used on the product path).  Returns (dg_rec, il_rec).

    Zero-weight pixels are dropped (the reference never counts them, cf.py:318,331).  A pixel is
    a record of six doubles (r_comov, dist_m, weights, delta*weights, z/2, 0).  Natural order with
    ROW_PAD dummies after every line of sight, and a copy interleaved by LANES with PAD dummies
    either side.  Dummies have weight 0 and a distance of 1e300 / 1e299: they add zeros to every
    sum and fall in no bin."""
    A = cat.arrays
    lanes, pad, row_pad, chunk = diag_layout()
    n = cat.n_los
    offset = A["offset"]
    keep = A["weights"] != 0
    los = np.repeat(np.arange(n, dtype=np.int64), np.diff(offset))[keep]
    rec = np.zeros((len(los), 6), dtype=np.float64)
    rec[:, 0], rec[:, 1] = A["r_comov"][keep], A["dist_m"][keep]
    rec[:, 2], rec[:, 3] = A["weights"][keep], A["delta_w"][keep]
    rec[:, 4] = 0.5 * A["z"][keep]
    first = np.zeros(n + 1, dtype=np.int64)
    first[1:] = np.cumsum(A["dg_count"].astype(np.int64))
    rank = np.arange(len(los), dtype=np.int64) - first[:-1][los]   # pixel index inside its forest
    dg_rec = np.zeros((cat.dg_total, 6), dtype=np.float64)
    dg_rec[:, 0] = dg_rec[:, 1] = DIAG_DUMMY_ROW
    dg_rec[A["dg_offset"][los] + rank] = rec
    jp = rank + pad
    il_rec = np.zeros((lanes * cat.il_total, 6), dtype=np.float64)
    il_rec[:, 0] = il_rec[:, 1] = DIAG_DUMMY_COL
    il_rec[(jp % lanes) * cat.il_total + A["il_offset"][los] + jp // lanes] = rec
    return dg_rec.reshape(-1), il_rec.reshape(-1)


def pack(data, is_object=False, ang_correlation=False, defer_products=False, rows=None):
    """Pack a ``dict[healpix] -> list`` into a HostCatalog.

    ``rows = (r0, r1)``: only the HEALPix pixels ``sorted(data)[r0:r1]`` (a band of a multi-GPU
    shard and its halo, ``dist.BandShard``): with a registered SoA the result is made of views
    into it, at a cost proportional to the band.

    ``defer_products``: leave ``delta_w = delta * weights`` and ``z_w = z * weights`` to the device
    (``pb2_derive_products`` when the catalogue is placed in HBM): the host then hands over
    ``delta`` instead, one pass over the pixels less and 8 B/pixel less to upload.

    ``ang_correlation``: the reference then feeds ``10**log_lambda`` in place of both distances
    (cf.py:186-208, xcf.py:161-182); the packed r_comov/dist_m hold that instead.
    """
    from . import forest as _forest
    all_hps = sorted(data)
    reg = None if is_object else _forest.soa_of(data)
    clean = reg is not None and len(reg["objs"]) > 0 and _forest.registered_clean(data, reg)
    if rows is not None and not clean:   # no registered SoA to slice: pack the band's own dict
        return pack({hp: data[hp] for hp in all_hps[rows[0]:rows[1]]}, is_object=is_object,
                    ang_correlation=ang_correlation, defer_products=defer_products)
    cat = HostCatalog()
    cat.is_object = is_object
    cat.healpixs = all_hps if rows is None else all_hps[rows[0]:rows[1]]
    cat.hp_index = {hp: k for k, hp in enumerate(cat.healpixs)}
    l0 = 0   # first line of sight / first pixel of the band inside the registered SoA
    if rows is not None:
        l0 = sum(len(data[hp]) for hp in all_hps[:rows[0]])
    counts = np.array([len(data[hp]) for hp in cat.healpixs], dtype=np.int64)
    n = int(counts.sum())
    if clean:
        # the producer (B200 loader, synthetic generator) gathered these when it built the forests
        los = reg["los"]
        col = lambda name: los[name][l0:l0 + n]
        objs = reg["objs"][l0:l0 + n]
    else:
        objs = [obj for hp in cat.healpixs for obj in data[hp]]
        # one attrgetter call per object for all positional attributes (a tenth of the getattr
        # calls of a column-by-column walk: 0.5 s -> 0.15 s per 170k quasars)
        fast_cols = {}
        names_f = ("x_cart", "y_cart", "z_cart", "ra", "dec", "cos_dec", "z_qso") + \
            (("r_comov", "dist_m", "weights") if is_object and not ang_correlation else ())
        if n:
            try:
                get = operator.attrgetter(*names_f)
                table = np.array([get(o) for o in objs], dtype=np.float64)
                if table.shape == (n, len(names_f)):
                    fast_cols = {nm: np.ascontiguousarray(table[:, k]) for k, nm in enumerate(names_f)}
            except (TypeError, ValueError, AttributeError):
                fast_cols = {}
            try:
                ids = list(zip(*map(operator.attrgetter("thingid", "plate", "fiberid"), objs)))
                fast_cols.update(thingid=list(ids[0]), plate=list(ids[1]), fiberid=list(ids[2]))
            except AttributeError:
                pass
        col = lambda name: fast_cols[name] if name in fast_cols else [getattr(o, name) for o in objs]
    cat.objs = objs
    cat.n_los = n
    A = cat.arrays
    hp_first = np.zeros(len(cat.healpixs) + 1, dtype=np.int32)
    hp_first[1:] = np.cumsum(counts)
    A["hp_first"] = hp_first
    A["row"] = np.repeat(np.arange(len(cat.healpixs), dtype=np.int32), counts)

    for name in ("x_cart", "y_cart", "z_cart", "ra", "dec", "cos_dec", "z_qso"):
        A[name] = np.array(col(name), dtype=np.float64).reshape(n)
    A["thingid"], ok_t = _as_int64(col("thingid"))
    if not ok_t:
        # non-integer ids: map equal ids to equal integers (only equality is ever used).  The
        # table is process-wide: the neighbour search compares ids ACROSS catalogues (data vs
        # data2, forests vs objects), so equal ids must map to equal integers in all of them
        A["thingid"] = np.array([_ID_TABLE.setdefault(o.thingid, len(_ID_TABLE)) for o in objs],
                                dtype=np.int64)
    cat.thingid_remapped = not ok_t
    A["plate"], ok_p = _as_int64(col("plate"))
    A["fiberid"], ok_f = _as_int64(col("fiberid"))
    cat.ids_are_int = bool(ok_p and ok_f)

    if is_object:
        offset = np.arange(n + 1, dtype=np.int64)
        zq = A["z_qso"]
        if ang_correlation:
            lam = np.array([10.0**o.log_lambda for o in objs], dtype=np.float64).reshape(n)
            A["r_comov"] = lam
            A["dist_m"] = lam.copy()
        else:
            A["r_comov"] = np.array(col("r_comov"), dtype=np.float64).reshape(n)
            A["dist_m"] = np.array(col("dist_m"), dtype=np.float64).reshape(n)
        A["z"] = zq.copy()
        A["weights"] = np.array(col("weights"), dtype=np.float64).reshape(n)
        A["delta_w"] = np.zeros(n, dtype=np.float64)
        A["z_w"] = A["z"] * A["weights"]
        A["log_lambda"] = np.zeros(n, dtype=np.float64)
        A["order"] = np.zeros(n, dtype=np.int32)
    else:
        soa = reg
        # picca_wick.py:369-384 drops `delta` (and others) from every forest to save memory
        has_delta = clean or all(getattr(o, "delta", None) is not None for o in objs)
        fields = ("log_lambda", "weights", "z") + (("delta",) if has_delta else ()) + \
            (() if ang_correlation else ("r_comov", "dist_m"))
        if soa is not None and n and all(k in soa for k in fields) and \
                (clean or _forest.views_intact(objs, soa, fields)):
            # the producer built these forests as views into one array per field in catalogue
            # order: pack without touching the objects' arrays
            cat.from_soa = True
            full_offset = np.asarray(soa["offset"], dtype=np.int64)
            p0, p1 = int(full_offset[l0]), int(full_offset[l0 + n])
            offset = np.ascontiguousarray(full_offset[l0:l0 + n + 1] - p0)
            cat_field = lambda name: soa[name][p0:p1]
        else:
            cat.from_soa = False
            npix = np.array([len(o.weights) for o in objs], dtype=np.int64)
            offset = np.zeros(n + 1, dtype=np.int64)
            offset[1:] = np.cumsum(npix)

            def cat_field(name):
                if n == 0:
                    return np.zeros(0, dtype=np.float64)
                return np.ascontiguousarray(np.concatenate(
                    [np.asarray(getattr(o, name), dtype=np.float64) for o in objs]))

        weights = cat_field("weights")
        delta = cat_field("delta") if has_delta else np.zeros(int(offset[-1]), dtype=np.float64)
        log_lambda = cat_field("log_lambda")
        A["z"] = cat_field("z")
        if ang_correlation:
            lam = 10.0**log_lambda
            A["r_comov"] = lam
            A["dist_m"] = lam.copy()
        else:
            A["r_comov"] = cat_field("r_comov")
            A["dist_m"] = cat_field("dist_m")
        A["weights"] = weights
        if defer_products:
            A["delta"] = np.ascontiguousarray(delta)   # products formed by pb2_derive_products
        else:
            # delta*weights is the product the reference forms first (cf.py:367-368); zero-weight
            # pixels never contribute (cf.py:318, :331) so a NaN delta there must not leak
            A["delta_w"] = np.where(weights != 0, delta * weights, 0.0)
            A["z_w"] = A["z"] * weights
        A["log_lambda"] = log_lambda
        A["order"] = np.array(col("order") if clean else
                              [-1 if getattr(o, "order", None) is None else int(o.order)
                               for o in objs], dtype=np.int32)
    A["offset"] = offset
    cat.n_pix = int(offset[-1])
    lengths = np.diff(offset)
    cat.max_pix = int(lengths.max()) if n else 0

    # deferred: the counts of non-zero weights, the finiteness / sortedness flags and the reach
    # come from ONE pass over the SoA once it is in HBM (pb2_catalog_stats) instead of ~12 NumPy
    # passes here
    cat.meta_deferred = bool(defer_products and not is_object and n > 0)
    if not is_object and not cat.meta_deferred:
        _diag_metadata(cat, offset)

    # sortedness inside each forest (enables the column windows of the pair kernel)
    cat.sorted = 1
    if not is_object and cat.n_pix > 1 and not cat.meta_deferred:
        inner = np.ones(cat.n_pix - 1, dtype=bool)
        inner[offset[1:-1][(offset[1:-1] > 0) & (offset[1:-1] < cat.n_pix)] - 1] = False
        for name in ("r_comov", "dist_m"):
            v = A[name]
            if np.any((v[1:] < v[:-1]) & inner) or not np.isfinite(v.sum()):
                cat.sorted = 0

    # bounding caps per HEALPix pixel
    nhp = len(cat.healpixs)
    cap = np.zeros((4, nhp), dtype=np.float64)
    xyz = np.stack([A["x_cart"], A["y_cart"], A["z_cart"]], axis=1) if n else np.zeros((0, 3))
    for k in range(nhp):
        a, b = hp_first[k], hp_first[k + 1]
        v = xyz[a:b]
        c = v.sum(axis=0)
        norm = np.sqrt((c * c).sum())
        c = c / norm if norm > 0 else v[0]
        dots = np.clip(v @ c, -1.0, 1.0)
        cap[:3, k] = c
        cap[3, k] = float(np.arccos(dots.min())) + 1e-7
    A["cap_x"], A["cap_y"], A["cap_z"], A["cap_rad"] = (np.ascontiguousarray(cap[0]),
                                                        np.ascontiguousarray(cap[1]),
                                                        np.ascontiguousarray(cap[2]),
                                                        np.ascontiguousarray(cap[3]))
    return cat


_REGISTERED = {}   # host pointer -> nbytes of the arrays page-locked by _page_lock


def _page_lock(arr):
    """Page-lock a large, long-lived host array in place (cudaHostRegister) so that its upload is
    a direct DMA instead of a staged copy; released when the array is garbage-collected.  Only
    the per-pixel SoA arrays a producer registered qualify (they outlive the packed catalogue and
    are uploaded again on every re-pack)."""
    import weakref
    import torch
    ptr, nbytes = arr.ctypes.data, arr.nbytes
    if _REGISTERED.get(ptr) == nbytes:
        return True
    rt = torch.cuda.cudart()
    if int(rt.cudaHostRegister(ptr, nbytes, 0)) != 0:
        return False
    _REGISTERED[ptr] = nbytes

    def release(p=ptr):
        _REGISTERED.pop(p, None)
        try:
            rt.cudaHostUnregister(p)
        except Exception:
            pass
    weakref.finalize(arr, release)
    return True


class DeviceCatalog:
    """A HostCatalog resident in HBM (torch tensors own the memory) + its ``pb2_catalog``."""

    def __init__(self, host, device, pin=False):
        import torch
        tensors, self.h2d_bytes = {}, 0
        lock = getattr(host, "from_soa", False)
        for name, arr in host.arrays.items():
            t = torch.from_numpy(arr)
            direct = False
            if pin:
                t = t.pin_memory()
                direct = True
            elif lock and arr.nbytes >= (8 << 20):
                # a band of a registered SoA (pack(rows=...)) is a VIEW: page-lock the array it
                # is cut from, once -- a pageable 1.5 GB upload costs 0.3 s per band and call
                root = arr
                while not root.flags.owndata and isinstance(root.base, np.ndarray):
                    root = root.base
                direct = root.flags.owndata and _page_lock(root)
            tensors[name] = t.to(device, non_blocking=direct)
            self.h2d_bytes += arr.nbytes
        self._finish(host, device, tensors)

    @classmethod
    def from_tensors(cls, host, device, tensors):
        """A catalogue whose SoA arrays are already on the device (e.g. copied from pinned host
        buffers by the caller); the derived copies are built here."""
        self = cls.__new__(cls)
        self.h2d_bytes = 0
        self._finish(host, device, dict(tensors))
        return self

    def _finish(self, host, device, tensors):
        """Derived device-side structures: the packed record copies of the diagonal-lane xi
        kernel (pb2_pack_diag) and the per-forest prefix sums of the forest x object kernel
        (pb2_build_prefix) -- both pure data movement over the SoA, done in HBM."""
        import ctypes
        import torch
        self.host, self.device, self.tensors = host, device, tensors
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        if "delta_w" not in tensors:
            # the host deferred delta * weights and z * weights (pack(defer_products=True))
            delta = tensors.pop("delta")
            tensors["delta_w"] = torch.empty_like(delta)
            tensors["z_w"] = torch.empty_like(delta)
            if host.n_pix:
                ptr = lambda t: ctypes.c_void_p(t.data_ptr())
                _lib.check(_lib.lib().pb2_derive_products(
                    ctypes.c_int64(host.n_pix), ptr(tensors["weights"]), ptr(delta),
                    ptr(tensors["z"]), ptr(tensors["delta_w"]), ptr(tensors["z_w"]), stream),
                    "pb2_derive_products")
        if getattr(host, "meta_deferred", False):
            # one pass over the SoA in HBM: non-zero-weight pixels per forest, finiteness,
            # sortedness, reach; the host turns the counts into the record offsets
            ptr = lambda t: ctypes.c_void_p(t.data_ptr())
            d_count = torch.zeros(host.n_los, dtype=torch.int32, device=device)
            d_flags = torch.zeros(4, dtype=torch.int64, device=device)
            _lib.check(_lib.lib().pb2_catalog_stats(
                ctypes.c_int64(host.n_los), ptr(tensors["offset"]), ptr(tensors["weights"]),
                ptr(tensors["r_comov"]), ptr(tensors["dist_m"]), ptr(tensors["z"]),
                ptr(tensors["delta_w"]), ptr(d_count), ptr(d_flags), stream), "pb2_catalog_stats")
            flags = d_flags.cpu().numpy()
            _diag_layout(host, d_count.cpu().numpy())
            host.dg_ok = int(flags[0] == 0)
            host.sorted = int(flags[1] == 0)
            host.dg_reach = float(flags[2:3].view(np.float64)[0]) if host.dg_ok else 0.0
            host.meta_deferred = False
            for name in ("dg_offset", "dg_count", "il_offset"):
                tensors[name] = torch.from_numpy(host.arrays[name]).to(device)
        else:
            for name in ("dg_offset", "dg_count", "il_offset"):
                if name in host.arrays and name not in tensors:
                    tensors[name] = torch.from_numpy(host.arrays[name]).to(device)
        if not host.is_object and host.n_los:
            tensors["dg_rec"] = torch.empty(6 * host.dg_total, dtype=torch.float64, device=device)
            tensors["il_rec"] = torch.empty(6 * host.dg_lanes * host.il_total, dtype=torch.float64,
                                            device=device)
        self.struct = build_struct(host, tensors)
        if not host.is_object and host.n_los:
            _lib.check(_lib.lib().pb2_pack_diag(
                ctypes.byref(self.struct), ctypes.c_int64(host.dg_total), stream), "pb2_pack_diag")
            # per-forest prefix sums for the forest x object kernel, built on the device
            tensors["px_rec"] = torch.empty(6 * (host.n_pix + host.n_los), dtype=torch.float64,
                                            device=device)
            _lib.check(_lib.lib().pb2_build_prefix(
                ctypes.byref(self.struct), ctypes.c_void_p(tensors["px_rec"].data_ptr()), stream),
                "pb2_build_prefix")
            self.struct.px_rec = tensors["px_rec"].data_ptr()

    def device_bytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.tensors.values()))


def build_struct(host, tensors):
    """``pb2_catalog`` pointing at the device tensors of a packed catalogue."""
    c = _lib.Catalog()
    c.n_los = host.n_los
    c.n_pix = host.n_pix
    for name in ("offset",) + _PIXEL_FIELDS + _LOS_F64 + _LOS_I64 + (
            "order", "row", "hp_first", "cap_x", "cap_y", "cap_z", "cap_rad"):
        setattr(c, name, tensors[name].data_ptr())
    if "dg_rec" in tensors:
        for name in _DIAG_FIELDS:
            setattr(c, name, tensors[name].data_ptr())
        c.il_total = host.il_total
        c.dg_lanes = host.dg_lanes
        c.dg_max_pix = host.dg_max_pix
        c.dg_ok = host.dg_ok
        c.dg_reach = host.dg_reach
    c.n_hp = len(host.healpixs)
    c.sorted = host.sorted
    c.max_pix = host.max_pix
    return c


_HOST_CACHE = {}
_ID_TABLE = {}  # non-integer line-of-sight id -> integer, shared by every catalogue of the process


def _fingerprint(data):
    """Cheap identity of a catalogue dict: which list objects it holds per HEALPix pixel and how
    long they are.  Replacing or resizing a per-pixel list (re-reading, shuffling, trimming)
    changes it; in-place edits of the arrays of a forest do not -- call ``invalidate`` then."""
    return tuple((hp, id(v), len(v)) for hp, v in data.items())


def invalidate(data=None):
    """Forget the packed copy of ``data`` (or of everything): the next cf / xcf call packs and
    uploads again.  Needed after in-place changes of weights / deltas / distances, which the
    fingerprint cannot see."""
    if data is None:
        _HOST_CACHE.clear()
        return
    for key in [k for k in _HOST_CACHE if k[0] == id(data)]:
        del _HOST_CACHE[key]


def cached_pack(data, is_object=False, ang_correlation=False):
    """Pack once per (dict object, flavour, fingerprint); the scripts keep the same dict for a
    whole run.  The mixed-id catalogues of one process share ``_ID_TABLE``."""
    key = (id(data), is_object, bool(ang_correlation))
    mark = _fingerprint(data)
    hit = _HOST_CACHE.get(key)
    if hit is not None and hit[0] is data and hit[2] == mark:
        return hit[1]
    cat = pack(data, is_object=is_object, ang_correlation=ang_correlation,
               defer_products=not is_object)
    _HOST_CACHE[key] = (data, cat, mark)
    return cat
