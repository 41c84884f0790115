"""ctypes wrapper around the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` leg.  The product package
``sqmc_b200`` never imports this module.

Determinants cross the boundary as arrays of shape (n, 2) uint64 = little-endian
128-bit integers, the memory layout of the reference's ``integer(ik)``
(src/types.f90:26).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sqmc_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_longlong, C.c_double
        L.orc_chem_new.restype = vp
        L.orc_chem_new.argtypes = [C.c_char_p, i32, i32, i32, vp, i32, i32, i32]
        L.orc_heg_new.restype = vp
        L.orc_heg_new.argtypes = [i32, dbl, i32, i32, dbl]
        L.orc_hubbardk_new.restype = vp
        L.orc_hubbardk_new.argtypes = [i32, i32, dbl, dbl, i32, i32]
        L.orc_free.argtypes = [vp]
        L.orc_set_hf_to_psit.argtypes = [vp, i32]
        L.orc_norb.argtypes = [vp]
        L.orc_nint.restype = i64
        L.orc_nint.argtypes = [vp]
        L.orc_get_chem.argtypes = [vp] * 8
        L.orc_get_heg.argtypes = [vp] * 4
        L.orc_get_hubbardk.argtypes = [vp] * 4
        L.orc_elements.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.orc_row.restype = i64
        L.orc_row.argtypes = [vp, i64, vp, vp, i64, i64, vp, vp]
        L.orc_build_upper.restype = i64
        L.orc_build_upper.argtypes = [vp, i64, vp, vp, i32]
        L.orc_get_upper.argtypes = [vp, vp, vp, vp]
        L.orc_matvec_upper.argtypes = [i64, vp, vp, vp, vp, vp]
        L.orc_matvec_upper_mt.argtypes = [i64, vp, vp, vp, vp, vp, i32]
        L.orc_davidson.argtypes = [i64, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        L.orc_lanczos.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, i32]
        L.orc_lanczos.restype = i32
        L.orc_davidson_single.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, i32]
        L.orc_davidson_single.restype = i32
        L.orc_pt2.argtypes = [vp, i64, vp, vp, vp, C.c_double, C.c_double, vp]
        L.orc_pt2.restype = C.c_double
        L.orc_pt2_sample.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, vp, i32, dbl, dbl, dbl, vp]
        L.orc_pt2_sample.restype = dbl
        L.orc_pt2_sample_terms.argtypes = [vp, i64, vp, vp, vp, vp, i32, dbl, dbl, i64, vp, vp, vp]
        L.orc_pt2_sample_terms.restype = i64
        L.orc_pt2_sample_energy.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, i32, dbl]
        L.orc_pt2_sample_energy.restype = dbl
        L.orc_pt2_alias.argtypes = [vp, i64, vp, vp, vp, dbl, dbl, dbl, i32, dbl, vp, i32, vp, vp, vp, vp]
        L.orc_pt2_alias.restype = i32
        L.orc_projector_step.argtypes = [i64, vp, vp, vp, dbl, dbl, vp, vp]
        L.orc_select.restype = i64
        L.orc_select.argtypes = [vp, i64, vp, vp, vp, vp, dbl, i64, vp, vp]
        L.orc_hci.argtypes = [vp, vp, i32, i32, i32]
        L.orc_hci_ndets.restype = i64
        L.orc_hci_ndets.argtypes = [vp]
        L.orc_hci_niter.argtypes = [vp]
        L.orc_hci_get.argtypes = [vp] * 9
        L.orc_hci_nritz.restype = i64
        L.orc_hci_nritz.argtypes = [vp]
        L.orc_hci_get_ritz.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dets_to_u64(dets):
    """list/array of python ints -> (n,2) uint64 little-endian 128-bit."""
    out = np.zeros((len(dets), 2), dtype=np.uint64)
    for k, d in enumerate(dets):
        d = int(d)
        out[k, 0] = d & 0xFFFFFFFFFFFFFFFF
        out[k, 1] = d >> 64
    return out


def u64_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


class System:
    """One model system (chem / heg / hubbardk) held by the oracle."""

    def __init__(self, handle, model):
        if not handle:
            raise RuntimeError("oracle: system construction failed")
        self.h = handle
        self.model = model
        self.norb = lib().orc_norb(handle)

    def __del__(self):
        try:
            if self.h:
                lib().orc_free(self.h)
                self.h = None
        except Exception:
            pass

    # ---- constructors -------------------------------------------------
    @staticmethod
    def chem(fcidump, norb, nelec, nup, orbsym, time_sym=False, z=1, hf_symmetry=999):
        sym = np.ascontiguousarray(orbsym, dtype=np.int32)
        h = lib().orc_chem_new(str(fcidump).encode(), norb, nelec, nup, _p(sym), int(time_sym), z, hf_symmetry)
        s = System(h, "chem")
        s.nelec, s.nup, s.ndn, s.time_sym, s.z = nelec, nup, nelec - nup, bool(time_sym), z
        return s

    @staticmethod
    def heg(n_dim, r_s, nelec, nup, cutoff_radius):
        s = System(lib().orc_heg_new(n_dim, r_s, nelec, nup, cutoff_radius), "heg")
        s.nelec, s.nup, s.ndn, s.n_dim = nelec, nup, nelec - nup, n_dim
        return s

    @staticmethod
    def hubbardk(l_x, l_y, t, U, nup, ndn):
        s = System(lib().orc_hubbardk_new(l_x, l_y, t, U, nup, ndn), "hubbardk")
        s.nelec, s.nup, s.ndn = nup + ndn, nup, ndn
        return s

    # ---- tables -------------------------------------------------------
    def chem_tables(self):
        n1 = self.norb + 1
        nint = lib().orc_nint(self.h)
        integrals = np.zeros(nint)
        c2 = np.zeros(n1 * n1, dtype=np.int32)
        order = np.zeros(n1, dtype=np.int32)
        sym = np.zeros(self.norb, dtype=np.int32)
        oe = np.zeros(self.norb)
        enuc = C.c_double()
        hf = np.zeros(4, dtype=np.uint64)
        lib().orc_get_chem(self.h, _p(integrals), _p(c2), _p(order), _p(sym), _p(oe), C.addressof(enuc), _p(hf))
        return dict(integrals=integrals, combine_2=c2, orb_order=order, orbital_symmetries=sym,
                    orbital_energies=oe, enuc=enuc.value,
                    hf_up=int(hf[0]) | (int(hf[1]) << 64), hf_dn=int(hf[2]) | (int(hf[3]) << 64))

    def heg_tables(self):
        kv = np.zeros(self.n_dim * self.norb)
        L = C.c_double()
        krel = np.zeros(self.norb * 3, dtype=np.int32)
        lib().orc_get_heg(self.h, _p(kv), C.addressof(L), _p(krel))
        return dict(k_vectors=kv.reshape(self.norb, self.n_dim), length_cell=L.value, k_rel=krel.reshape(self.norb, 3))

    def hubbardk_tables(self):
        kv = np.zeros(2 * self.norb, dtype=np.int32)
        ke = np.zeros(self.norb)
        u = C.c_double()
        lib().orc_get_hubbardk(self.h, _p(kv), _p(ke), C.addressof(u))
        return dict(k_vectors=kv.reshape(self.norb, 2), k_energies=ke, ubyn=u.value)

    # ---- hot path -----------------------------------------------------
    def elements(self, iu, id_, ju, jd):
        iu, id_, ju, jd = (np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2) for a in (iu, id_, ju, jd))
        out = np.zeros(len(iu))
        lib().orc_elements(self.h, len(iu), _p(iu), _p(id_), _p(ju), _p(jd), _p(out))
        return out

    def row(self, up, dn, i, cap=1 << 16):
        """brute-force full row i (0-based) -> (cols 1-based ascending, vals)"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        cols = np.zeros(cap, dtype=np.int64)
        vals = np.zeros(cap)
        k = lib().orc_row(self.h, len(up), _p(up), _p(dn), int(i), cap, _p(cols), _p(vals))
        assert k <= cap
        return cols[:k].copy(), vals[:k].copy()

    def build_upper(self, up, dn, incremental=False, hf_to_psit=False):
        """-> (counts int64[n], indices int64[nnz] 1-based, values f64[nnz]) : reference layout."""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        n = len(up)
        lib().orc_set_hf_to_psit(self.h, int(bool(hf_to_psit)))
        nnz = lib().orc_build_upper(self.h, n, _p(up), _p(dn), int(incremental))
        counts = np.zeros(n, dtype=np.int64)
        idx = np.zeros(nnz, dtype=np.int64)
        val = np.zeros(nnz)
        lib().orc_get_upper(self.h, _p(counts), _p(idx), _p(val))
        return counts, idx, val

    def select(self, up, dn, coeffs, min_H, eps_var, cap=5_000_000):
        """one get_next_det_list step -> (new_up, new_dn sorted by label and not in the old list, updated min_H)"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        mh = np.array(min_H, dtype=np.float64)
        nu = np.zeros((cap, 2), dtype=np.uint64)
        nd = np.zeros((cap, 2), dtype=np.uint64)
        nn = lib().orc_select(self.h, len(up), _p(up), _p(dn), _p(c), _p(mh), float(eps_var), cap, _p(nu), _p(nd))
        assert nn <= cap
        return nu[:nn].copy(), nd[:nn].copy(), mh

    def pt2(self, up, dn, wts, var_energy, eps_pt):
        """deterministic second-order PT (second_order_pt, hci.f90:1100-1182) -> (delta_E, ndets_connected)"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        w = np.ascontiguousarray(wts, dtype=np.float64).reshape(-1)
        nconn = C.c_longlong()
        de = lib().orc_pt2(self.h, len(up), _p(up), _p(dn), _p(w), float(var_energy), float(eps_pt), C.addressof(nconn))
        return de, nconn.value

    def pt2_sample(self, up, dn, s_up, s_dn, s_coeffs, s_w_over_p, n_mc, var_energy, eps_pt, eps_pt_big):
        """one sample of second_order_pt_alias (hci.f90:1563-1654): (up, dn) = label-sorted variational list, s_* = the distinct
        sampled determinants -> (e_2pt_this_sample, ndets_connected)"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        su = np.ascontiguousarray(s_up, dtype=np.uint64).reshape(-1, 2)
        sd = np.ascontiguousarray(s_dn, dtype=np.uint64).reshape(-1, 2)
        sc = np.ascontiguousarray(s_coeffs, dtype=np.float64).reshape(-1)
        sw = np.ascontiguousarray(s_w_over_p, dtype=np.float64).reshape(-1)
        nconn = C.c_longlong()
        e = lib().orc_pt2_sample(self.h, len(up), _p(up), _p(dn), len(su), _p(su), _p(sd), _p(sc), _p(sw), int(n_mc), float(var_energy),
                                 float(eps_pt), float(eps_pt_big), C.addressof(nconn))
        return e, nconn.value

    def pt2_sample_terms(self, s_up, s_dn, s_coeffs, s_w_over_p, n_mc, eps_pt, eps_pt_big):
        """merged output of find_doubly_excited for (a share of) one sample (semistoch.f90:2044-2117): the distinct connected
        determinants in label order with their sums term1, term2, term1_big, term2_big -> (up, dn, terms[nd, 4])"""
        su = np.ascontiguousarray(s_up, dtype=np.uint64).reshape(-1, 2)
        sd = np.ascontiguousarray(s_dn, dtype=np.uint64).reshape(-1, 2)
        sc = np.ascontiguousarray(s_coeffs, dtype=np.float64).reshape(-1)
        sw = np.ascontiguousarray(s_w_over_p, dtype=np.float64).reshape(-1)
        args = (self.h, len(su), _p(su), _p(sd), _p(sc), _p(sw), int(n_mc), float(eps_pt), float(eps_pt_big))
        nd = lib().orc_pt2_sample_terms(*args, 0, None, None, None)
        ou = np.zeros((nd, 2), dtype=np.uint64)
        od = np.zeros((nd, 2), dtype=np.uint64)
        t = np.zeros((nd, 4))
        lib().orc_pt2_sample_terms(*args, nd, _p(ou), _p(od), _p(t))
        return ou, od, t

    def pt2_sample_energy(self, up, dn, c_up, c_dn, terms, n_mc, var_energy):
        """the k loop of second_order_pt_alias (hci.f90:1616-1654) over merged per-determinant sums (label-sorted variational list)"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        cu = np.ascontiguousarray(c_up, dtype=np.uint64).reshape(-1, 2)
        cd = np.ascontiguousarray(c_dn, dtype=np.uint64).reshape(-1, 2)
        t = np.ascontiguousarray(terms, dtype=np.float64).reshape(-1, 4)
        return lib().orc_pt2_sample_energy(self.h, len(up), _p(up), _p(dn), len(cu), _p(cu), _p(cd), _p(t), int(n_mc), float(var_energy))

    def pt2_alias(self, up, dn, wts, var_energy, eps_pt, eps_pt_big, n_mc, target_error, seed4, max_samples=1000):
        """the sampling loop of second_order_pt_alias (one core, n_mc > 0) with the reference's rannyu stream ->
        dict(pt_energy, std_dev, e_now[], n_distinct[])"""
        up = np.ascontiguousarray(up, dtype=np.uint64).reshape(-1, 2)
        dn = np.ascontiguousarray(dn, dtype=np.uint64).reshape(-1, 2)
        w = np.ascontiguousarray(wts, dtype=np.float64).reshape(-1)
        seed = np.ascontiguousarray(seed4, dtype=np.int32)
        e_now = np.zeros(max_samples)
        nd = np.zeros(max_samples, dtype=np.int32)
        out2 = np.zeros(2)
        nc = np.zeros(max_samples, dtype=np.int64)
        ns = lib().orc_pt2_alias(self.h, len(up), _p(up), _p(dn), _p(w), float(var_energy), float(eps_pt), float(eps_pt_big), int(n_mc),
                                 float(target_error), _p(seed), int(max_samples), _p(e_now), _p(nd), _p(out2), _p(nc))
        return {"pt_energy": out2[0], "std_dev": out2[1], "e_now": e_now[:ns].copy(), "n_distinct": nd[:ns].copy(), "n_connected": nc[:ns].copy()}

    def hci(self, eps_var, eps_var_sched=(), n_states=1, max_iters=50, max_dets=0):
        sched = np.zeros(30)
        for k, e in enumerate(eps_var_sched):
            sched[k] = e
        sched = np.maximum(sched, eps_var)  # do_walk.f90:425
        lib().orc_hci(self.h, _p(sched), n_states, max_iters, max_dets)
        n = lib().orc_hci_ndets(self.h)
        nit = lib().orc_hci_niter(self.h)
        up = np.zeros((n, 2), dtype=np.uint64)
        dn = np.zeros((n, 2), dtype=np.uint64)
        wts = np.zeros(n * n_states)
        energy = np.zeros(n_states)
        lnd = np.zeros(nit, dtype=np.int64)
        lnz = np.zeros(nit, dtype=np.int64)
        lnr = np.zeros(nit, dtype=np.int64)
        le = np.zeros(nit * n_states)
        lib().orc_hci_get(self.h, _p(up), _p(dn), _p(wts), _p(energy), _p(lnd), _p(lnz), _p(lnr), _p(le))
        nr = lib().orc_hci_nritz(self.h)
        ritz = np.zeros(nr)
        lib().orc_hci_get_ritz(self.h, _p(ritz))
        ritz_per_iter, o = [], 0
        for k in range(nit):
            ritz_per_iter.append(ritz[o:o + lnr[k]].reshape(-1, n_states))
            o += lnr[k]
        return dict(up=up, dn=dn, wts=wts.reshape(n_states, n).T, energy=energy, ndet=lnd, nnz=lnz,
                    ritz=ritz_per_iter, iter_energy=le.reshape(nit, n_states))


def matvec_upper(counts, idx, val, x):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    lib().orc_matvec_upper(len(counts), _p(idx), _p(counts), _p(val), _p(x), _p(y))
    return y


def matvec_upper_mt(counts, idx, val, x, nthreads):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    lib().orc_matvec_upper_mt(len(counts), _p(idx), _p(counts), _p(val), _p(x), _p(y), int(nthreads))
    return y


def davidson(counts, idx, val, n_states=1, v0=None):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(counts)
    evecs = np.zeros(n * n_states)
    evals = np.zeros(n_states)
    ritz = np.zeros(4096)
    nmv = C.c_int()
    v0p = None
    if v0 is not None:
        v0 = np.ascontiguousarray(np.asarray(v0, dtype=np.float64).T.reshape(-1))  # (n,n_states) -> column-major
        v0p = _p(v0)
    nl = lib().orc_davidson(n, n_states, _p(idx), _p(counts), _p(val), v0p, _p(evecs), _p(evals), _p(ritz), len(ritz), C.addressof(nmv))
    return dict(evals=evals, evecs=evecs.reshape(n_states, n).T, ritz=ritz[:min(nl, len(ritz))].reshape(-1, n_states), n_matvec=nmv.value)


def davidson_single(counts, idx, val, v0=None):
    """davidson_sparse_single (more_tools.f90:3055-3233) -> dict(lowest, highest, evec, ritz (printed values), n_iter)"""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(counts)
    evec = np.zeros(n)
    out3 = np.zeros(3)
    ritz = np.zeros(64)
    v0p = None
    if v0 is not None:
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        v0p = _p(v0)
    nit = lib().orc_davidson_single(n, _p(idx), _p(counts), _p(val), v0p, _p(evec), _p(out3), _p(ritz), len(ritz))
    return dict(lowest=out3[0], highest=out3[1], evec=evec, n_iter=nit, ritz=ritz[:min(int(out3[2]), 64)].copy())


def lanczos(counts, idx, val, v0=None):
    """matrix_lanczos_sparse (more_tools.f90:1742-1883) -> dict(lowest, highest, second_lowest, evec, ritz (per step), n_iter)"""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(counts)
    evec = np.zeros(n)
    out3 = np.zeros(4)
    ritz = np.zeros(64)
    v0p = None
    if v0 is not None:
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        v0p = _p(v0)
    nit = lib().orc_lanczos(n, _p(idx), _p(counts), _p(val), v0p, _p(evec), _p(out3), _p(ritz), len(ritz))
    return dict(lowest=out3[0], highest=out3[1], second_lowest=out3[2], evec=evec, n_iter=nit, ritz=ritz[:min(int(out3[3]), 64)].copy())


def projector_step(counts, idx, minus_tau_H, tau, e_trial, w):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    vals = np.ascontiguousarray(minus_tau_H, dtype=np.float64)
    w = np.array(w, dtype=np.float64)
    dw = np.zeros_like(w)
    lib().orc_projector_step(len(counts), _p(idx), _p(counts), _p(vals), tau, e_trial, _p(w), _p(dw))
    return w, dw


def upper_to_scipy(counts, idx, val):
    """reference upper-tri layout -> full symmetric scipy CSR (test helper)."""
    import scipy.sparse as sp
    n = len(counts)
    rows = np.repeat(np.arange(n), counts)
    cols = np.asarray(idx) - 1
    U = sp.coo_matrix((val, (rows, cols)), shape=(n, n)).tocsr()
    D = sp.diags(U.diagonal())
    return (U + U.T - D).tocsr()


# ----------------------------------------------------------------------------- MPI data distribution of the reference
_M128 = (1 << 128) - 1
_DJB_PRIME = 309485009821345068724781371
_DJB_OFFSET = 144066263297769815596495629667062367629


def _s128(x):
    """two's complement value of a 128-bit pattern"""
    return x - (1 << 128) if x >> 127 else x


def djb_hash(up, dn):
    """djb_hash (mpi_routines.f90:354-379) on integer(16) words: 16 rounds per word of
    hash = (ishft(hash,5) + hash) + ishft(ieor(tmp, Z'5555555555555555') * prime, -j), j = 0, 8, ..., 120, with
    tmp = up for word 1 and dn + offset for word 2; 128-bit wrap-around, ishft = logical shift.  Parity unpinned: the
    reference ships no hash values; restated to produce realistic ownership maps for the distributed-slice tests."""
    h = _DJB_PRIME
    for i, w in enumerate((int(up), int(dn))):
        tmp = (w + _DJB_OFFSET) & _M128 if i == 1 else w & _M128
        t = ((tmp ^ 0x5555555555555555) * _DJB_PRIME) & _M128
        for j in range(0, 128, 8):
            h = (((h << 5) & _M128) + h + (t >> j)) & _M128
    return _s128(h)


def det_owner(up, dn, ncores):
    """get_det_owner (mpi_routines.f90:419-445) + hash (:257-289): coreid = abs(mod(djb_hash(det), ncores)); 0 for one core.
    up, dn: (n,2) uint64 arrays or python ints -> int32[n]"""
    def ints(a):
        a = np.asarray(a)
        return u64_to_ints(a) if (a.dtype == np.uint64 and a.ndim == 2) else [int(v) for v in a]
    ups, dns = ints(up), ints(dn)
    if ncores == 1:
        return np.zeros(len(ups), dtype=np.int32)
    return np.array([abs(djb_hash(a, b)) % ncores for a, b in zip(ups, dns)], dtype=np.int32)


def matvec_local_band_redscatt(counts, idx, val, owner, x_slices):
    """What the ncores>1 branch of walk does (do_walk.f90:2259-2260): every core multiplies its column band of the
    shuffled matrix (from_upper_triangular_to_band_shuffle, more_tools.f90:3371-3471: all rows x the columns it owns)
    with its own slice of the vector -> an n-long partial answer (fast_sparse_matrix_multiply_local_band, :3562-3587);
    MPI_REDUCE_SCATTER sums the partials and hands every core the rows it owns (mpi_routines.f90:1592).
    x_slices[c] = core c's slice (its determinants in ascending list order) -> list of result slices."""
    owner = np.asarray(owner)
    n = len(counts)
    ncores = len(x_slices)
    total = np.zeros(n)
    for c in range(ncores):
        xc = np.zeros(n)
        xc[owner == c] = x_slices[c]
        total = total + matvec_upper(counts, idx, val, xc)       # band of core c times its slice; partials summed in core order
    return [total[owner == c].copy() for c in range(ncores)]
