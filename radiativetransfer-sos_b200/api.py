"""Host-side mirror of the reference interface for the hot path, over the C ABI of libsosgpu.so.

The names follow the reference's subroutines (SOS, SOS_OS, SOS_NOYAUX, SOS_AGGREGATE, SOS_TRPHI_OPTION);
argument meaning and error behaviour (IER = 0 / -1) are the reference's.  There is no CPU fallback: if the
CUDA library cannot be loaded or no CUDA device is present, `Solver()` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsosgpu.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)

SOSGPU_OK = 0
SOSGPU_ERR_NO_DEVICE = -2


class COptics(C.Structure):
    _fields_ = [("nbmu", C.c_int), ("rmu", c_dp), ("ga", c_dp), ("n0", C.c_int), ("tetas", C.c_double),
                ("os_nb", C.c_int), ("alpha", c_dp), ("beta", c_dp), ("gamma", c_dp), ("zeta", c_dp),
                ("a_trunc", C.c_double), ("piz", C.c_double), ("piztr", C.c_double), ("ron", C.c_double),
                ("rho", C.c_double), ("imat_surf", C.c_int), ("ifresnel", C.c_int), ("ind_surf", C.c_double),
                ("surf", c_fp), ("n_surf_rec", C.c_int), ("igmax", C.c_int), ("ipolar", C.c_int),
                ("zout", C.c_double)]


class CTerm(C.Structure):
    _fields_ = [("optics", C.c_int), ("group", C.c_int), ("aik", C.c_double), ("nt", C.c_int),
                ("zprof", c_dp), ("h", c_dp), ("pcaer", c_dp), ("pcmol", c_dp)]


_TERM_DTYPE = np.dtype([("optics", "i4"), ("group", "i4"), ("aik", "f8"), ("nt", "i4"), ("_pad", "i4"),
                        ("zprof", "u8"), ("h", "u8"), ("pcaer", "u8"), ("pcmol", "u8")])
assert _TERM_DTYPE.itemsize == C.sizeof(CTerm)


class CTermOut(C.Structure):
    _fields_ = [("rec", c_dp), ("n_fourier", c_ip), ("n_scatter", c_ip), ("stop_reason", c_ip),
                ("emoins", c_dp), ("eplus", c_dp), ("ttot_tronc", c_dp), ("ttot_vrai", c_dp), ("tauout", c_dp),
                ("ier", c_ip)]


class CGroupOut(C.Structure):
    _fields_ = [("rec", c_dp), ("n_rec", c_ip), ("emoins", c_dp), ("eplus", c_dp), ("ttot_tronc", c_dp),
                ("ttot_vrai", c_dp), ("tauout", c_dp)]


class CDirectModels(C.Structure):
    _fields_ = [("iroujean", C.c_int), ("k0", C.c_double), ("k1", C.c_double), ("k2", C.c_double),
                ("irondeaux", C.c_int), ("ibreon", C.c_int), ("inadal", C.c_int), ("alpha_nadal", C.c_double),
                ("beta_nadal", C.c_double), ("imaignan", C.c_int), ("coef_c_maignan", C.c_double)]


class CCkd(C.Structure):
    _fields_ = [("nb_temp", C.c_int), ("nb_pres", C.c_int), ("nb_conc_h2o", C.c_int), ("tab_temp", c_dp), ("tab_pres", c_dp),
                ("tab_conc_h2o", c_dp), ("nexp", c_ip), ("kdis_ki", c_dp), ("kdis_ki_h2o", c_dp)]


class CGasProfile(C.Structure):
    _fields_ = [("userprofil", c_dp), ("altabs", c_dp), ("ro", c_dp)]


class CProfileTerm(C.Structure):
    _fields_ = [("lamb1", C.c_int), ("ik", C.c_int * 8), ("absprofil", C.c_int), ("iprofil", C.c_int), ("tr", C.c_double),
                ("hr", C.c_double), ("ta", C.c_double), ("ha", C.c_double), ("zmin", C.c_double), ("zmax", C.c_double)]


class CAerComponent(C.Structure):
    _fields_ = [("rn", C.c_double), ("in_", C.c_double), ("alpha0", C.c_double), ("alphaf", C.c_double), ("igranu", C.c_int),
                ("v1", C.c_double), ("v2", C.c_double), ("v3", C.c_double), ("wa", C.c_double)]


class CAerModel(C.Structure):
    _fields_ = [("ncomp", C.c_int), ("comp", C.c_int * 4), ("weight", C.c_double * 4), ("itronc", C.c_int)]


NT_MAX = 600


class CStats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("flops", C.c_double), ("bytes", C.c_double), ("step_ms", C.c_double),
                ("step_launches", C.c_longlong), ("total_ms", C.c_double), ("launches", C.c_longlong),
                ("aggregate_ms", C.c_double), ("useful_flops", C.c_double)]


_lib = None


def load_library():
    """dlopen libsosgpu.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libsosgpu.so is missing (%s): build it with __graft_entry__.build(); "
                               "the SOS hot path has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.sosgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        lib.sosgpu_destroy.argtypes = [C.c_void_p]
        lib.sosgpu_last_error.argtypes = [C.c_void_p]
        lib.sosgpu_last_error.restype = C.c_char_p
        lib.sosgpu_launch_count.argtypes = [C.c_void_p]
        lib.sosgpu_last_kernel_ms.argtypes = [C.c_void_p]
        lib.sosgpu_last_kernel_ms.restype = C.c_double
        lib.sosgpu_launch_count.restype = C.c_longlong
        lib.sosgpu_set_options.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        lib.sosgpu_batch_upload.argtypes = [C.c_void_p, C.POINTER(COptics), C.c_int, C.POINTER(CTerm), C.c_int,
                                            C.c_int, C.POINTER(C.c_void_p)]
        lib.sosgpu_batch_upload_os.argtypes = [C.c_void_p, C.POINTER(COptics), C.c_int, C.POINTER(CTerm), C.c_int,
                                               C.c_int, c_ip, C.POINTER(C.c_void_p)]
        lib.sosgpu_batch_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(CTermOut),
                                         C.POINTER(CGroupOut)]
        lib.sosgpu_batch_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.sosgpu_batch_stats.argtypes = [C.c_void_p, C.POINTER(CStats)]
        lib.sosgpu_batch_group_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.sosgpu_group_finalize.argtypes = [c_dp, c_dp, c_dp, C.c_int]
        lib.sosgpu_set_direct_models.argtypes = [C.c_void_p, C.POINTER(CDirectModels)]
        lib.sosgpu_comm_unique_id.argtypes = [C.c_char_p]
        lib.sosgpu_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        lib.sosgpu_comm_destroy.argtypes = [C.c_void_p]
        lib.sosgpu_comm_barrier.argtypes = [C.c_void_p]
        lib.sosgpu_batch_reduce_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.sosgpu_batch_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(CGroupOut)]
        lib.sosgpu_batch_set_group_optics.argtypes = [C.c_void_p, c_ip]
        lib.sosgpu_batch_set_group_direct.argtypes = [C.c_void_p, c_ip]
        lib.sosgpu_batch_gather_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_ip, C.c_int, C.c_int, c_dp, c_dp]
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def write_updown(fic_up, fic_down, nbmu, itrphi, phios, pas_phi, zout, phi_fin, theta_fin, up, down, fix_sca_index=False):
    """SOS_Up.txt / SOS_Down.txt (SOS_ABS_MAIN.F:2250-2519) from the tables of Solver.trphi_option ([7, nphi, N] each).
    Host-only formatting inside libsosgpu.so; no device needed."""
    lib = load_library()
    up, down = _f64(up), _f64(down)
    assert up.ndim == 3 and up.shape == down.shape and up.shape[0] == 7 and up.shape[2] == nbmu
    lib.sosgpu_write_updown.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, c_dp, c_dp,
                                        c_dp, c_dp, C.c_int, C.c_int]
    pf = _f64(phi_fin) if phi_fin is not None else np.zeros(up.shape[1])
    rc = lib.sosgpu_write_updown(str(fic_up).encode(), str(fic_down).encode(), nbmu, itrphi, float(phios), int(pas_phi), float(zout),
                                 _d(pf), _d(_f64(theta_fin)), _d(up), _d(down), up.shape[1], int(bool(fix_sca_index)))
    if rc != SOSGPU_OK:
        raise RuntimeError("sosgpu_write_updown failed (rc=%d)" % rc)


def write_trans(fictrans, tetas, ttot_tronc, ttot_vrai, tdifmus, rmu_pos, tdifmug):
    """The -SOS.Trans file of SOS_PROC.F:3779-3818 (rmu_pos = RMU(1:N), tdifmug = TDIFMUG(1:N)).  Host-only."""
    lib = load_library()
    rmu_pos, tdifmug = _f64(rmu_pos), _f64(tdifmug)
    lib.sosgpu_write_trans.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, c_dp, c_dp]
    rc = lib.sosgpu_write_trans(str(fictrans).encode(), tetas, ttot_tronc, ttot_vrai, tdifmus, rmu_pos.size, _d(rmu_pos), _d(tdifmug))
    if rc != SOSGPU_OK:
        raise RuntimeError("sosgpu_write_trans failed (rc=%d)" % rc)


def write_flux(ficflux, tetas, ttot_tronc, ttot_vrai, emoins, eplus, tr, hr, ta, ha, zalt, tauabs):
    """The -SOS.Flux file of SOS_PROC.F:3820-3874; returns (TDIR_VRAI, FLUX_DIFF_DOWN, FLUX_DOWN).  Host-only."""
    lib = load_library()
    zalt, tauabs = _f64(zalt), _f64(tauabs)
    out = [C.c_double(0) for _ in range(3)]
    lib.sosgpu_write_flux.argtypes = [C.c_char_p] + [C.c_double] * 9 + [c_dp, c_dp] + [c_dp] * 3
    rc = lib.sosgpu_write_flux(str(ficflux).encode(), tetas, ttot_tronc, ttot_vrai, emoins, eplus, tr, hr, ta, ha, _d(zalt), _d(tauabs),
                               *[C.byref(o) for o in out])
    if rc != SOSGPU_OK:
        raise RuntimeError("sosgpu_write_flux failed (rc=%d)" % rc)
    return tuple(o.value for o in out)


def write_aerosols(path, os_nb, kmat1, kmat2, asym, coef_tronca, piztr, alp, beta11, gamma12, zeta):
    """The aerosol result file of SOS_AEROSOLS (SOS_AEROSOLS.F:2810-2832), read by SOS_PREPA_OS.  Host-side formatting inside
    the library; needs no GPU."""
    lib = load_library()
    lib.sosgpu_write_aerosols.argtypes = [C.c_char_p, C.c_int] + [C.c_double] * 5 + [c_dp] * 4
    rc = lib.sosgpu_write_aerosols(os.fsencode(path), os_nb, kmat1, kmat2, asym, coef_tronca, piztr, _d(_f64(alp)), _d(_f64(beta11)),
                                   _d(_f64(gamma12)), _d(_f64(zeta)))
    if rc != SOSGPU_OK:
        raise RuntimeError("write_aerosols failed (%d)" % rc)


class TermResults:
    """Per-term outputs of SOS / SOS_OS: Fourier records (file order Q,U,I), counts, fluxes, optical depths."""
    pass


class GroupResults:
    """Per-wavelength outputs of the SOS_AGGREGATE chain."""
    pass


class Batch:
    """Inputs resident in HBM (sosgpu_batch)."""

    def __init__(self, solver, handle, nterm, ngroup, rec_stride, wmax, keep):
        self.solver, self.handle = solver, handle
        self.nterm, self.ngroup, self.rec_stride, self.wmax = nterm, ngroup, rec_stride, wmax
        self._keep = keep

    def free(self):
        if self.handle:
            self.solver.lib.sosgpu_batch_free(self.solver.ctx, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Solver:
    """One CUDA device, one stream.  Mirrors SOS (SOS.F:340), SOS_OS (SOS_OS.F:303) and SOS_AGGREGATE."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sosgpu_create(C.byref(h), device)
        if rc != SOSGPU_OK:
            raise RuntimeError("libsosgpu: cannot create a context on CUDA device %d (rc=%d): "
                               "no usable GPU, and the SOS hot path has no CPU fallback" % (device, rc))
        self.ctx = h
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sosgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != SOSGPU_OK:
            raise RuntimeError("libsosgpu %s failed (rc=%d): %s" % (what, rc, self.lib.sosgpu_last_error(self.ctx).decode()))

    @property
    def launches(self):
        return int(self.lib.sosgpu_launch_count(self.ctx))

    def set_direct_models(self, roujean=None, irondeaux=0, ibreon=0, nadal=None, maignan=None):
        """Direct-beam terms of the land-surface models in SOS_TRPHI (SOS_TRPHI.F:1047-1200) for the following synthesis
        calls: roujean=(k0,k1,k2), nadal=(alpha,beta), maignan=coef_c; no argument switches them all off."""
        dm = CDirectModels()
        if roujean is not None:
            dm.iroujean, dm.k0, dm.k1, dm.k2 = 1, roujean[0], roujean[1], roujean[2]
        dm.irondeaux, dm.ibreon = int(irondeaux), int(ibreon)
        if nadal is not None:
            dm.inadal, dm.alpha_nadal, dm.beta_nadal = 1, nadal[0], nadal[1]
        if maignan is not None:
            dm.imaignan, dm.coef_c_maignan = 1, maignan
        self._check(self.lib.sosgpu_set_direct_models(self.ctx, C.byref(dm)), "set_direct_models")

    def set_options(self, field_budget_bytes=0, max_wave_orders=0):
        self.lib.sosgpu_set_options(self.ctx, field_budget_bytes, max_wave_orders)

    # ------------------------------------------------------------------ batch
    def upload(self, workload, term_ids=None, os_level=False, iborm=None, groups=None, ngroup=None):
        """H2D of a Workload (synth.Workload): optics entries are de-duplicated by identity.
        groups/ngroup: explicit aggregation group of each uploaded term (multi-GPU sharding keeps the global
        group numbering on every rank); default = one group per wavelength present."""
        terms = workload.terms if term_ids is None else [workload.terms[i] for i in term_ids]
        keep = []
        omap, copt = {}, []
        for t in terms:
            o = workload.optics[t.optics]
            if id(o) in omap:
                continue
            omap[id(o)] = len(copt)
            arrs = [_f64(o.rmu), _f64(o.ga), _f64(o.alpha), _f64(o.beta), _f64(o.gamma), _f64(o.zeta)]
            keep.extend(arrs)
            co = COptics()
            co.nbmu, co.rmu, co.ga, co.n0, co.tetas, co.os_nb = o.nbmu, _d(arrs[0]), _d(arrs[1]), o.n0, o.tetas, o.os_nb
            co.alpha, co.beta, co.gamma, co.zeta = _d(arrs[2]), _d(arrs[3]), _d(arrs[4]), _d(arrs[5])
            co.a_trunc, co.piz, co.piztr, co.ron, co.rho = o.a_trunc, o.piz, o.piztr, o.ron, o.rho
            co.imat_surf, co.ifresnel, co.ind_surf = o.imat_surf, o.ifresnel, o.ind_surf
            if o.imat_surf == 1:
                s = np.ascontiguousarray(o.surf, dtype=np.float32)
                keep.append(s)
                co.surf, co.n_surf_rec = s.ctypes.data_as(c_fp), s.shape[0]
            else:
                co.surf, co.n_surf_rec = None, 0
            co.igmax, co.ipolar, co.zout = o.igmax, o.ipolar, o.zout
            copt.append(co)
        gmap = {}
        # sosgpu_term descriptors, filled through a numpy view of the same layout (a ctypes field store per term
        # costs more than the H2D copy it describes)
        nterm = len(terms)
        tdesc = np.zeros(nterm, dtype=_TERM_DTYPE)
        plist = []
        h2d = 0
        for t in terms:
            arrs = (_f64(t.zprof), _f64(t.h), _f64(t.pcaer), _f64(t.pcmol))
            keep.append(arrs)
            plist.extend(a.__array_interface__["data"][0] for a in arrs)
            h2d += 4 * arrs[0].nbytes
        ptrs = np.array(plist, dtype=np.uint64).reshape(nterm, 4)
        tdesc["optics"] = [omap[id(workload.optics[t.optics])] for t in terms]
        tdesc["group"] = groups if groups is not None else [gmap.setdefault(t.optics, len(gmap)) for t in terms]
        tdesc["aik"] = [t.aik for t in terms]
        tdesc["nt"] = [t.nt for t in terms]
        for k, name in enumerate(("zprof", "h", "pcaer", "pcmol")):
            tdesc[name] = ptrs[:, k]
        cterms = tdesc.ctypes.data_as(C.POINTER(CTerm))
        h2d += int(sum(a.nbytes for a in keep if isinstance(a, np.ndarray)))
        keep.append(tdesc)
        coptics = (COptics * len(copt))(*copt)
        h = C.c_void_p()
        ng = ngroup if groups is not None else len(gmap)
        if os_level:
            ib = np.ascontiguousarray(iborm, dtype=np.int32)
            rc = self.lib.sosgpu_batch_upload_os(self.ctx, coptics, len(copt), cterms, len(terms), ng,
                                                 ib.ctypes.data_as(c_ip), C.byref(h))
        else:
            rc = self.lib.sosgpu_batch_upload(self.ctx, coptics, len(copt), cterms, len(terms), ng, C.byref(h))
        self._check(rc, "batch_upload")
        rec_stride = max(workload.optics[t.optics].os_nb for t in terms) + 1
        wmax = 2 * max(workload.optics[t.optics].nbmu for t in terms) + 1
        b = Batch(self, h, len(terms), ng, rec_stride, wmax, keep)
        b.h2d_bytes = h2d
        b.nbmu_of_term = [workload.optics[t.optics].nbmu for t in terms]
        return b

    def run(self, batch, want_terms=True, want_groups=True, want_rec=True, part_only=False):
        """Run the resident batch.  With want_* False nothing is copied back (device-resident timing)."""
        nt, ng, rs, w = batch.nterm, batch.ngroup, batch.rec_stride, batch.wmax
        tr = gr = None
        to = go = None
        if want_terms:
            tr = TermResults()
            tr.rec = np.zeros((nt, rs, 3, w)) if want_rec else None
            tr.n_fourier = np.zeros(nt, dtype=np.int32)
            tr.n_scatter = np.zeros((nt, rs), dtype=np.int32)
            tr.stop_reason = np.zeros((nt, rs), dtype=np.int32)
            for n in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
                setattr(tr, n, np.zeros(nt))
            tr.ier = np.zeros(nt, dtype=np.int32)
            to = CTermOut(_d(tr.rec) if want_rec else None, tr.n_fourier.ctypes.data_as(c_ip),
                          tr.n_scatter.ctypes.data_as(c_ip), tr.stop_reason.ctypes.data_as(c_ip),
                          _d(tr.emoins), _d(tr.eplus), _d(tr.ttot_tronc), _d(tr.ttot_vrai), _d(tr.tauout),
                          tr.ier.ctypes.data_as(c_ip))
        if want_groups:
            gr = GroupResults()
            gr.rec = np.zeros((ng, rs, 3, w))
            gr.n_rec = np.zeros(ng, dtype=np.int32)
            for n in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
                setattr(gr, n, np.zeros(ng))
            go = CGroupOut(_d(gr.rec), gr.n_rec.ctypes.data_as(c_ip), _d(gr.emoins), _d(gr.eplus),
                           _d(gr.ttot_tronc), _d(gr.ttot_vrai), _d(gr.tauout))
        rc = self.lib.sosgpu_batch_run(self.ctx, batch.handle, rs, w, int(part_only),
                                       C.byref(to) if to is not None else None,
                                       C.byref(go) if go is not None else None)
        self._check(rc, "batch_run")
        return tr, gr

    def stats(self, batch):
        st = CStats()
        self.lib.sosgpu_batch_stats(batch.handle, C.byref(st))
        return {k: getattr(st, k) for k, _ in CStats._fields_}

    def group_buffer(self, batch):
        p, n = C.c_void_p(), C.c_size_t()
        self.lib.sosgpu_batch_group_buffer(batch.handle, C.byref(p), C.byref(n))
        return p.value, n.value

    def solve(self, workload, term_ids=None, **kw):
        """SOS + SOS_AGGREGATE for every term of the workload (host buffers in, host buffers out)."""
        b = self.upload(workload, term_ids)
        try:
            return self.run(b, **kw)
        finally:
            b.free()

    # ------------------------------------------------------------------ multi-GPU (the library owns the NCCL communicator)
    UNIQUE_ID_BYTES = 128

    def comm_unique_id(self):
        """Rank 0: ncclGetUniqueId; the caller broadcasts the bytes to the other ranks."""
        buf = C.create_string_buffer(self.UNIQUE_ID_BYTES)
        self._check(self.lib.sosgpu_comm_unique_id(buf), "comm_unique_id")
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        self._check(self.lib.sosgpu_comm_init(self.ctx, nranks, rank, C.create_string_buffer(bytes(unique_id), self.UNIQUE_ID_BYTES)),
                    "comm_init")
        self.nranks, self.rank = nranks, rank

    def comm_barrier(self):
        self._check(self.lib.sosgpu_comm_barrier(self.ctx), "comm_barrier")

    def set_group_optics(self, batch, optics_of_group):
        a = np.ascontiguousarray(optics_of_group, dtype=np.int32)
        self._check(self.lib.sosgpu_batch_set_group_optics(batch.handle, a.ctypes.data_as(c_ip)), "set_group_optics")

    def set_group_direct(self, batch, direct):
        """Groups that are one solve the reference does not pass through SOS_AGGREGATE (no gas, -SOS.AbsModeCKD 2): their optical
        thicknesses are the term's own (include/sosgpu.h: sosgpu_batch_set_group_direct)."""
        a = np.ascontiguousarray(direct, dtype=np.int32)
        if a.size != batch.ngroup:
            raise ValueError("set_group_direct: one flag per group")
        self._check(self.lib.sosgpu_batch_set_group_direct(batch.handle, a.ctypes.data_as(c_ip)), "set_group_direct")

    def reduce_groups(self, batch, root=0):
        """Term-sharded layout: ONE in-place ncclReduce of the partial CKD sums + group metadata to `root`."""
        self._check(self.lib.sosgpu_batch_reduce_groups(self.ctx, batch.handle, root), "reduce_groups")

    def groups(self, batch):
        """Band sums on the root after reduce_groups (what the SOS_AGGREGATE chain leaves behind)."""
        ng, rs, w = batch.ngroup, batch.rec_stride, batch.wmax
        gr = GroupResults()
        gr.rec = np.zeros((ng, rs, 3, w))
        gr.n_rec = np.zeros(ng, dtype=np.int32)
        for n in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
            setattr(gr, n, np.zeros(ng))
        go = CGroupOut(_d(gr.rec), gr.n_rec.ctypes.data_as(c_ip), _d(gr.emoins), _d(gr.eplus), _d(gr.ttot_tronc),
                       _d(gr.ttot_vrai), _d(gr.tauout))
        self._check(self.lib.sosgpu_batch_groups(self.ctx, batch.handle, rs, w, C.byref(go)), "batch_groups")
        return gr

    def gather_tables(self, batch, groups_of_rank, nphi, root=0, download=True):
        """Wavelength-sharded layout: tables of the last batch_trphi call of every rank, collected on `root`
        ([sum groups, 7, nphi, Nmax] up and down; None on the other ranks or when download is False)."""
        g = np.ascontiguousarray(groups_of_rank, dtype=np.int32)
        nmax = (batch.wmax - 1) // 2
        is_root = getattr(self, "rank", 0) == root
        up = np.zeros((int(g.sum()), 7, nphi, nmax)) if (download and is_root) else None
        down = np.zeros_like(up) if up is not None else None
        self._check(self.lib.sosgpu_batch_gather_tables(self.ctx, batch.handle, root, g.ctypes.data_as(c_ip), nphi, nmax,
                                                        _d(up) if up is not None else None,
                                                        _d(down) if down is not None else None), "gather_tables")
        return up, down

    # ------------------------------------------------------------------ single routines
    def sos_os(self, o, nt, h, xdel, ydel, zprof, iborm):
        """SOS_OS (SOS_OS.F:303): profile arrays as SOS_OS receives them."""
        from .synth import Term, Workload
        wl = Workload("sos_os", [o], [Term(0, 1.0, _f64(zprof), _f64(h), _f64(xdel), _f64(ydel))])
        b = self.upload(wl, os_level=True, iborm=[iborm])
        try:
            tr, _ = self.run(b, want_groups=False)
        finally:
            b.free()
        return tr

    def noyaux(self, is_, rmu, os_nb, alpha, beta, gamma, zeta):
        """SOS_NOYAUX (SOS_OS.F:1857).  Same return convention as oracle.noyaux."""
        rmu = _f64(rmu)
        W = rmu.size
        N = (W - 1) // 2
        a, b, g, z = (_f64(x) for x in (alpha, beta, gamma, zeta))
        out = {n: np.zeros(W) for n in ("xpl", "xrl", "xtl")}
        ker = {n: np.zeros((W, W)) for n in ("bp", "gr", "gt", "arr", "art", "att")}
        rc = self.lib.sosgpu_noyaux(self.ctx, C.c_int(is_), C.c_int(N), _d(rmu), C.c_int(os_nb), _d(a), _d(b), _d(g),
                                    _d(z), _d(out["xpl"]), _d(out["xrl"]), _d(out["xtl"]), _d(ker["bp"]),
                                    _d(ker["gr"]), _d(ker["gt"]), _d(ker["arr"]), _d(ker["art"]), _d(ker["att"]))
        self._check(rc, "noyaux")
        out.update(ker)
        return out

    def order_step(self, is_, rmu, ga, os_nb, alpha, beta, gamma, zeta, ron, ipolar, nt, h, xdel, ydel, i1, q1, u1):
        """SOS_FSOURCE_ORDREIG + SOS_INTEGR_EPOPT fused (black surface).  Returns (i1n,q1n,u1n,i2,q2,u2)."""
        rmu, ga = _f64(rmu), _f64(ga)
        W = rmu.size
        N = (W - 1) // 2
        arrs = [_f64(x) for x in (alpha, beta, gamma, zeta, h, xdel, ydel, i1, q1, u1)]
        outs = [np.zeros((W, nt + 1)) for _ in range(6)]
        rc = self.lib.sosgpu_order_step(self.ctx, C.c_int(is_), C.c_int(N), _d(rmu), _d(ga), C.c_int(os_nb),
                                        _d(arrs[0]), _d(arrs[1]), _d(arrs[2]), _d(arrs[3]), C.c_double(ron),
                                        C.c_int(ipolar), C.c_int(nt), _d(arrs[4]), _d(arrs[5]), _d(arrs[6]),
                                        _d(arrs[7]), _d(arrs[8]), _d(arrs[9]), None,
                                        _d(outs[0]), _d(outs[1]), _d(outs[2]), _d(outs[3]), _d(outs[4]), _d(outs[5]))
        self._check(rc, "order_step")
        return outs

    def trphi_option(self, rec, nbmu, rmu, tau, tauout, igli, n0, wind, ind_surf, ifresnel, itrphi, phios, pas_phi,
                     ipolar):
        """SOS_TRPHI_OPTION (SOS_TRPHI.F:285).  Same return convention as oracle.trphi_option."""
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        cap = 2 if itrphi == 1 else 360 // max(pas_phi, 1) + 1
        phi_fin, theta = np.zeros(cap), np.zeros(nbmu)
        up, down = np.zeros((7, cap, nbmu)), np.zeros((7, cap, nbmu))
        n = self.lib.sosgpu_trphi_option(self.ctx, _d(rec), C.c_int(rec.shape[0]), C.c_int(nbmu), _d(_f64(rmu)),
                                         C.c_double(tau), C.c_double(tauout), C.c_int(igli), C.c_int(n0),
                                         C.c_double(wind), C.c_double(ind_surf), C.c_int(ifresnel), C.c_int(itrphi),
                                         C.c_double(phios), C.c_int(pas_phi), C.c_int(ipolar), _d(phi_fin), _d(theta),
                                         _d(up), _d(down), C.c_int(cap))
        if n < 0:
            self._check(n, "trphi_option")
        return n, phi_fin[:n], theta, up[:, :n], down[:, :n]

    @property
    def last_kernel_ms(self):
        return float(self.lib.sosgpu_last_kernel_ms(self.ctx))

    def glitter(self, nbmu, rmu, chr_, wind, ind_surf, os_nb, os_ns, os_nm):
        """SOS_GLITTER (SOS_GLITTER.F:229): surface-file records [os_nb+1, 9, N, N] REAL*4 and the G-series lengths."""
        surf = np.zeros((os_nb + 1, 9, nbmu, nbmu), dtype=np.float32)
        il = np.zeros(nbmu * (nbmu + 1) // 2, dtype=np.int32)
        rc = self.lib.sosgpu_glitter(self.ctx, C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_int(os_nb),
                                     C.c_int(os_ns), C.c_int(os_nm), C.c_double(wind), C.c_double(ind_surf),
                                     surf.ctypes.data_as(c_fp), il.ctypes.data_as(c_ip))
        self._check(rc, "glitter")
        return surf, il

    def surface_bpdf(self, isurf, nbmu, rmu, chr_, ind_surf, os_nb, os_ns, os_nm, coef_c=0.0):
        """SOS_SURFACE_BPDF (SOS_SURFACE_BPDF.F:219) for isurf = 4 (Rondeaux), 5 (Breon) or 7 (Maignan, coef_c): records
        [os_nb+1, 9, N, N] REAL*4."""
        surf = np.zeros((os_nb + 1, 9, nbmu, nbmu), dtype=np.float32)
        rc = self.lib.sosgpu_surface_bpdf(self.ctx, C.c_int(isurf), C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_int(os_nb),
                                          C.c_int(os_ns), C.c_int(os_nm), C.c_double(ind_surf), C.c_double(coef_c),
                                          surf.ctypes.data_as(c_fp))
        self._check(rc, "surface_bpdf")
        return surf

    def surface_nadal(self, nbmu, rmu, chr_, ind_surf, alpha, beta, os_nb, os_ns, os_nm, pairing="reference"):
        """SOS_SURFACE_BPDF with ISURF = 6 (SOS_SURFACE_BPDF.F:219, series generator SOS_F21SF_NADAL :686): Nadal's BPDF (alpha,
        beta) -> (records [os_nb+1, 9, N, N] REAL*4, series lengths [N(N+1)/2]).  pairing "reference": the file the reference
        writes (its SOS_MAT_REFLEXION gives pair number p the series of (p / N + 1, p mod N + 1)); "own": each pair its own series."""
        if pairing not in ("reference", "own"):
            raise ValueError("pairing: 'reference' or 'own'")
        surf = np.zeros((os_nb + 1, 9, nbmu, nbmu), dtype=np.float32)
        il = np.zeros(nbmu * (nbmu + 1) // 2, dtype=np.int32)
        rc = self.lib.sosgpu_surface_nadal(self.ctx, C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_int(os_nb), C.c_int(os_ns),
                                           C.c_int(os_nm), C.c_double(ind_surf), C.c_double(alpha), C.c_double(beta),
                                           C.c_int(0 if pairing == "reference" else 1), surf.ctypes.data_as(c_fp), il.ctypes.data_as(c_ip))
        self._check(rc, "surface_nadal")
        return surf, il

    def roujean(self, nbmu, rmu, os_nb, k0, k1, k2):
        """SOS_ROUJEAN (SOS_ROUJEAN.F:212): Fourier series of Roujean's BRDF, records [os_nb+1, 9, N, N] REAL*4."""
        surf = np.zeros((os_nb + 1, 9, nbmu, nbmu), dtype=np.float32)
        rc = self.lib.sosgpu_roujean(self.ctx, C.c_int(nbmu), _d(_f64(rmu)), C.c_int(os_nb), C.c_double(k0), C.c_double(k1),
                                     C.c_double(k2), surf.ctypes.data_as(c_fp))
        self._check(rc, "roujean")
        return surf

    def bpdf_ajout_brdf(self, surf1, surf2):
        """SOS_BPDF_AJOUT_BRDF (SOS_SURFACE.F:2503): sum of two surface files (BPDF + BRDF)."""
        a = np.ascontiguousarray(surf1, dtype=np.float32)
        b = np.ascontiguousarray(surf2, dtype=np.float32)
        assert a.shape == b.shape
        out = np.zeros_like(a)
        rc = self.lib.sosgpu_bpdf_ajout_brdf(self.ctx, a.ctypes.data_as(c_fp), b.ctypes.data_as(c_fp), C.c_int(a.shape[2]),
                                             C.c_int(a.shape[0] - 1), out.ctypes.data_as(c_fp))
        self._check(rc, "bpdf_ajout_brdf")
        return out

    def mat_fresnel(self, nbmu, rmu, chr_, ind_surf, os_ns):
        """SOS_MAT_FRESNEL (SOS_SURFACE.F:1235): alpha, beta, gamma, zeta [os_ns+1] as the RES_FRESNEL file holds them."""
        out = [np.zeros(os_ns + 1) for _ in range(4)]
        rc = self.lib.sosgpu_mat_fresnel(self.ctx, C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_double(ind_surf),
                                         C.c_int(os_ns), _d(out[0]), _d(out[1]), _d(out[2]), _d(out[3]))
        self._check(rc, "mat_fresnel")
        return out

    # ---- the per-term profile chain ----
    @staticmethod
    def profile_terms(terms):
        """list of dicts (lamb1, ik[8], absprofil, iprofil, tr, hr, ta, ha, zmin, zmax) -> C array; a C array passes through."""
        if isinstance(terms, C.Array):
            return terms
        arr = (CProfileTerm * len(terms))()
        for a, t in zip(arr, terms):
            a.lamb1, a.absprofil, a.iprofil = int(t["lamb1"]), int(t["absprofil"]), int(t["iprofil"])
            for k in range(8):
                a.ik[k] = int(t["ik"][k])
            a.tr, a.hr, a.ta, a.ha, a.zmin, a.zmax = (float(t[k]) for k in ("tr", "hr", "ta", "ha", "zmin", "zmax"))
        return arr

    @staticmethod
    def _ckd(tables):
        """tables: dict of Fortran-ordered arrays as READ_CKD_COEFF fills them (keys as tests/profile_cases.ckd_tables)."""
        keep = [np.asfortranarray(tables[k], dtype=np.float64) for k in ("tab_temp", "tab_pres", "tab_conc", "ki", "kh")]
        nexp = np.asfortranarray(tables["nexp"], dtype=np.int32)
        c = CCkd(int(tables["nb_temp"]), int(tables["nb_pres"]), int(tables["nb_conc"]), _d(keep[0]), _d(keep[1]), _d(keep[2]),
                 nexp.ctypes.data_as(c_ip), _d(keep[3]), _d(keep[4]))
        return c, keep + [nexp]

    def absprofile(self, tables, userprofil, ro, terms):
        """SOS_ABSPROFILE (SOS_ABSPROFILE.F:184) for every term: (tauabs [nterm, 50], ier [nterm])."""
        c, keep = self._ckd(tables)
        user, ro = np.asfortranarray(userprofil, dtype=np.float64), np.asfortranarray(ro, dtype=np.float64)
        g = CGasProfile(_d(user), None, _d(ro))
        arr = self.profile_terms(terms)
        tau, ier = np.zeros((len(terms), 50)), np.zeros(len(terms), dtype=np.int32)
        self.lib.sosgpu_absprofile.argtypes = [C.c_void_p, C.POINTER(CCkd), C.POINTER(CGasProfile), C.POINTER(CProfileTerm), C.c_int,
                                               c_dp, c_ip]
        rc = self.lib.sosgpu_absprofile(self.ctx, C.byref(c), C.byref(g), arr, len(terms), _d(tau), ier.ctypes.data_as(c_ip))
        self._check(rc, "absprofile")
        return tau, ier

    def _profile_out(self, n):
        return (np.zeros(n, dtype=np.int32), *(np.zeros((n, NT_MAX + 1)) for _ in range(4)), np.zeros(n, dtype=np.int32))

    def profile(self, altabs, tauabs, terms, text_hop=True):
        """SOS_PROFILE (SOS_PROFIL.F:224) for every term from given absorption profiles [nterm, 50]:
        (nt, zprof, h, pcaer, pcmol [nterm, 601], ier).  text_hop: the values SOS reads back from PROFIL_TMP."""
        n = len(terms)
        nt, z, h, pa, pm, ier = self._profile_out(n)
        tau = _f64(tauabs).reshape(n, 50)
        self.lib.sosgpu_profile.argtypes = [C.c_void_p, c_dp, c_dp, C.POINTER(CProfileTerm), C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp,
                                            c_dp, c_ip]
        rc = self.lib.sosgpu_profile(self.ctx, _d(_f64(altabs)), _d(tau), self.profile_terms(terms), n, int(bool(text_hop)),
                                     nt.ctypes.data_as(c_ip), _d(z), _d(h), _d(pa), _d(pm), ier.ctypes.data_as(c_ip))
        self._check(rc, "profile")
        return nt, z, h, pa, pm, ier

    def profile_chain(self, tables, userprofil, altabs, ro, terms, text_hop=True, want_tauabs=False):
        """SOS_ABSPROFILE -> SOS_PROFILE -> PROFIL_TMP hop for every term of a band, on the device without a host hop in between
        (SOS_PROC.F:3494-3537).  Returns (nt, zprof, h, pcaer, pcmol, ier[, tauabs])."""
        n = len(terms)
        c, keep = self._ckd(tables)
        user, ro, alt = np.asfortranarray(userprofil, dtype=np.float64), np.asfortranarray(ro, dtype=np.float64), _f64(altabs)
        g = CGasProfile(_d(user), _d(alt), _d(ro))
        nt, z, h, pa, pm, ier = self._profile_out(n)
        tau = np.zeros((n, 50)) if want_tauabs else None
        self.lib.sosgpu_profile_chain.argtypes = [C.c_void_p, C.POINTER(CCkd), C.POINTER(CGasProfile), C.POINTER(CProfileTerm), C.c_int,
                                                  C.c_int, c_dp, c_ip, c_dp, c_dp, c_dp, c_dp, c_ip]
        rc = self.lib.sosgpu_profile_chain(self.ctx, C.byref(c), C.byref(g), self.profile_terms(terms), n, int(bool(text_hop)),
                                           _d(tau) if want_tauabs else None, nt.ctypes.data_as(c_ip), _d(z), _d(h), _d(pa), _d(pm),
                                           ier.ctypes.data_as(c_ip))
        self._check(rc, "profile_chain")
        return (nt, z, h, pa, pm, ier) + ((tau,) if want_tauabs else ())

    # ---- aerosol optics per wavelength ----
    def mie_count(self, alpha0, alphaf):
        self.lib.sosgpu_mie_count.argtypes = [C.c_double, C.c_double]
        return int(self.lib.sosgpu_mie_count(alpha0, alphaf))

    def mie(self, nbmu, rmu, rn, in_, alpha0, alphaf):
        """SOS_MIE (SOS_MIE.F:205) without its file: dict rec [nrec, 3] (ALPHA, QEXT, QSCA, REAL*4), g [nrec], imie / qmie / umie
        [nrec, 2 nbmu + 1] (REAL*4), alphaf."""
        n = self.mie_count(alpha0, alphaf)
        if n < 1:
            raise RuntimeError("mie: size-parameter range outside CTE_MIE_DIM (SOS_MIE error 997)")
        nang = 2 * nbmu + 1
        rec, g = np.zeros((n, 3), dtype=np.float32), np.zeros(n)
        im, qm, um = (np.zeros((n, nang), dtype=np.float32) for _ in range(3))
        nrec = C.c_int(0)
        f = lambda a: a.ctypes.data_as(c_fp)
        self.lib.sosgpu_mie.argtypes = [C.c_void_p, C.c_int, c_dp] + [C.c_double] * 4 + [C.c_int, c_fp, c_dp, c_fp, c_fp, c_fp, c_ip]
        rc = self.lib.sosgpu_mie(self.ctx, nbmu, _d(_f64(rmu)), rn, in_, alpha0, alphaf, n, f(rec), _d(g), f(im), f(qm), f(um), C.byref(nrec))
        self._check(rc, "mie")
        assert nrec.value == n
        return dict(rec=rec, g=g, imie=im, qmie=qm, umie=um, alphaf=float(alphaf))

    def granu(self, nbmu, table, igranu, v1, v2, v3, wa):
        """SOS_GRANU (SOS_AEROSOLS.F:4392) on a Mie table (dict as mie() returns): (ier, kmat1, kmat2, somme_nr, p11, p12, p33)."""
        nang = 2 * nbmu + 1
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        rec, im, qm, um = f(table["rec"]), f(table["imie"]), f(table["qmie"]), f(table["umie"])
        k, ier = np.zeros(3), C.c_int(-1)
        p11, p12, p33 = np.zeros(nang), np.zeros(nang), np.zeros(nang)
        p = lambda a: a.ctypes.data_as(c_fp)
        self.lib.sosgpu_granu.argtypes = [C.c_void_p, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_double, C.c_int] + [C.c_double] * 4 + \
                                         [c_dp, c_dp, c_dp, c_dp, c_ip]
        rc = self.lib.sosgpu_granu(self.ctx, nbmu, rec.shape[0], p(rec), p(im), p(qm), p(um), table["alphaf"], igranu, v1, v2, v3, wa,
                                   _d(k), _d(p11), _d(p12), _d(p33), C.byref(ier))
        self._check(rc, "granu")
        return ier.value, k[0], k[1], k[2], p11, p12, p33

    def decompo_legendre(self, itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33):
        """SOS_DECOMPO_LEGENDRE (SOS_AEROSOLS.F:3924): dict with alp, beta11, beta22, gamma12, delta33, zeta (0:os_nb), the
        (possibly truncated) p11, ttt, coef_tronca, z1, itronc on exit, ier."""
        nang = 2 * nbmu + 1
        q11, ttt = _f64(p11).copy(), np.zeros(nang)
        co = {k: np.zeros(os_nb + 1) for k in ("alp", "beta11", "beta22", "gamma12", "delta33", "zeta")}
        it, ct, z1, ier = C.c_int(itronc), C.c_double(0), C.c_double(0), C.c_int(-1)
        self.lib.sosgpu_decompo_legendre.argtypes = [C.c_void_p, c_ip, C.c_int, c_dp, c_dp, C.c_int] + [c_dp] * 5 + [c_dp, c_dp] + \
                                                    [c_dp] * 6 + [c_ip]
        rc = self.lib.sosgpu_decompo_legendre(self.ctx, C.byref(it), nbmu, _d(_f64(xmu)), _d(_f64(xhr)), os_nb, _d(q11), _d(ttt),
                                              _d(_f64(p12)), _d(_f64(p22)), _d(_f64(p33)), C.byref(ct), C.byref(z1), _d(co["alp"]),
                                              _d(co["beta11"]), _d(co["beta22"]), _d(co["gamma12"]), _d(co["delta33"]), _d(co["zeta"]),
                                              C.byref(ier))
        self._check(rc, "decompo_legendre")
        co.update(p11=q11, ttt=ttt, coef_tronca=ct.value, z1=z1.value, itronc=it.value, ier=ier.value)
        return co

    def aerosols(self, nbmu, xmu, xhr, components, models, os_nb, want_phase=True):
        """The whole chain on the device (SOS_AEROSOLS.F:1158-1304, 1706-2123 without their MIE files): components = list of
        (rn, in, alpha0, alphaf, igranu, v1, v2, v3, wa), models = list of (ncomp, comp indices, weights, itronc).  Returns dict:
        comp_k [nc, 3], comp_phase [nc, 3, nang], comp_ier, scal [nm, 8] (KMAT1, KMAT2, PIZ, PIZTR, COEF_TRONCA, asymmetry
        factor, Z1, ITRONC on exit), coef [nm, 6, os_nb+1] (ALPHA, BETA11, GAMMA12, ZETA, BETA22, DELTA33), phase [nm, 4, nang],
        model_ier."""
        nc, nm, nang = len(components), len(models), 2 * nbmu + 1
        ca = (CAerComponent * nc)()
        for a, c in zip(ca, components):
            a.rn, a.in_, a.alpha0, a.alphaf, a.igranu, a.v1, a.v2, a.v3, a.wa = c
        ma = (CAerModel * max(nm, 1))()
        for a, m in zip(ma, models):
            a.ncomp, a.itronc = int(m[0]), int(m[3])
            for i, ci in enumerate(m[1]):
                a.comp[i] = int(ci)
            for i, w in enumerate(m[2]):
                a.weight[i] = float(w)
        out = dict(comp_k=np.zeros((nc, 3)), comp_phase=np.zeros((nc, 3, nang)) if want_phase else None,
                   comp_ier=np.zeros(nc, dtype=np.int32), scal=np.zeros((nm, 8)), coef=np.zeros((nm, 6, os_nb + 1)),
                   phase=np.zeros((nm, 4, nang)) if want_phase else None, model_ier=np.zeros(nm, dtype=np.int32))
        self.lib.sosgpu_aerosols.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int, C.POINTER(CAerComponent), C.c_int,
                                             C.POINTER(CAerModel), C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp, c_dp, c_ip]
        rc = self.lib.sosgpu_aerosols(self.ctx, nbmu, _d(_f64(xmu)), _d(_f64(xhr)), nc, ca, nm, ma, os_nb, _d(out["comp_k"]),
                                      _d(out["comp_phase"]) if want_phase else None, out["comp_ier"].ctypes.data_as(c_ip),
                                      _d(out["scal"]), _d(out["coef"]), _d(out["phase"]) if want_phase else None,
                                      out["model_ier"].ctypes.data_as(c_ip))
        self._check(rc, "aerosols")
        return out

    def batch_trphi(self, batch, igli, wind, ind_surf, ifresnel, itrphi, phios, pas_phi, ipolar, download=True):
        """SOS_TRPHI_OPTION for every wavelength of a resident batch (after run); tables [ngroup, 7, nphi, Nmax]."""
        cap = 2 if itrphi == 1 else 360 // max(pas_phi, 1) + 1
        nmax = (batch.wmax - 1) // 2
        up = np.zeros((batch.ngroup, 7, cap, nmax)) if download else None
        down = np.zeros((batch.ngroup, 7, cap, nmax)) if download else None
        n = self.lib.sosgpu_batch_trphi(self.ctx, batch.handle, C.c_int(igli), C.c_double(wind), C.c_double(ind_surf),
                                        C.c_int(ifresnel), C.c_int(itrphi), C.c_double(phios), C.c_int(pas_phi),
                                        C.c_int(ipolar), C.c_int(cap), _d(up) if download else None,
                                        _d(down) if download else None)
        if n < 0:
            self._check(n, "batch_trphi")
        return n, up, down

    def transmissions(self, workload, term_ids=None):
        """The -SOS.Trans branch of SOS (SOS.F:605-637): diffuse transmissions of the equivalent atmosphere for the
        solar direction (TDIFMUS) and for every Gauss direction (TDIFMUG(1:N)): 1+N extra SOS_OS solves per term
        with a black surface, IBORM=0, the truncation-adapted profile, and N0 = J.  All of them go to the GPU as one
        SOS_OS-level batch.  Returns (tdifmus[nterm], tdifmug[nterm, N])."""
        import copy
        from .synth import Term, Workload
        terms = workload.terms if term_ids is None else [workload.terms[i] for i in term_ids]
        wl2 = Workload("trans")
        variants = {}
        index = []
        for t in terms:
            o = workload.optics[t.optics]
            if id(o) not in variants:
                vs = []
                for n0 in [o.n0] + list(range(1, o.nbmu + 1)):
                    v = copy.copy(o)
                    v.n0, v.rho, v.imat_surf, v.ifresnel, v.zout, v.surf = n0, 0.0, 0, 0, -1.0, None
                    wl2.optics.append(v)
                    vs.append(len(wl2.optics) - 1)
                variants[id(o)] = vs
            # truncation adaptation of the profile (SOS.F:523-543), same statement order
            h, xd, yd = _f64(t.h).copy(), _f64(t.pcaer).copy(), _f64(t.pcmol).copy()
            if o.a_trunc != 0.0:
                htr = np.zeros_like(h)
                htr[0] = h[0]
                for i in range(1, h.size):
                    dh = h[i] - h[i - 1]
                    va = xd[i] * dh
                    vatr = va * (1 - o.piz * 0.5 * o.a_trunc)
                    vr = yd[i] * dh
                    vg = (1 - xd[i] - yd[i]) * dh
                    htr[i] = (vatr + vr + vg) + htr[i - 1]
                    xd[i] = vatr / (vatr + vr + vg)
                    yd[i] = vr / (vatr + vr + vg)
                h = htr
            xd = xd * o.piztr
            first = len(wl2.terms)
            for vi in variants[id(o)]:
                wl2.terms.append(Term(vi, 1.0, t.zprof, h, xd, yd))
            index.append((first, o.nbmu))
        b = self.upload(wl2, os_level=True, iborm=[0] * len(wl2.terms))
        try:
            tr, _ = self.run(b, want_groups=False, want_rec=False)
        finally:
            b.free()
        nmax = max(n for _, n in index)
        tdifmus = np.zeros(len(terms))
        tdifmug = np.zeros((len(terms), nmax))
        for i, (first, n) in enumerate(index):
            tdifmus[i] = tr.emoins[first]                # EMOINS of the black-surface IS=0 solve (SOS.F:611-616)
            tdifmug[i, :n] = tr.emoins[first + 1:first + 1 + n]
        return tdifmus, tdifmug
