"""ctypes binding of the CPU oracle (oracle/dang_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs -- never by anything under dang_b200/.  PARITY UNPINNED (see dang_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile oracle/_build/*.so with the committed Makefile (gcc only, no GPU)."""
    args = ["make", "-C", _HERE]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def load(omp: bool = False) -> C.CDLL:
    name = "libdang_oracle_omp.so" if omp else "libdang_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_long
    ip = C.POINTER(C.c_int)
    sigs = {
        "ora_create": (vp, [i, i, i, i, i]),
        "ora_destroy": (None, [vp]),
        "ora_set_band": (i, [vp, i, d, i, c_dp, c_dp]),
        "ora_set_maps": (i, [vp, c_dp, c_dp, c_dp, c_dp, c_dp]),
        "ora_set_pol_type": (None, [vp, i, i]),
        "ora_set_component": (i, [vp, i, i, C.c_char_p, d, i, i, c_dp, c_dp]),
        "ora_set_index": (i, [vp, i, i, i, i, i, i, c_dp, c_dp, d, i, i, ip, i]),
        "ora_amplitude": (c_dp, [vp, i]),
        "ora_indices": (c_dp, [vp, i]),
        "ora_sky_model": (c_dp, [vp]),
        "ora_res_map": (c_dp, [vp]),
        "ora_chi_map": (c_dp, [vp]),
        "ora_step_size": (d, [vp, i, i]),
        "ora_nindices": (i, [vp, i]),
        "ora_set_template": (i, [vp, i, c_dp, c_dp, ip, i]),
        "ora_template_map": (c_dp, [vp, i]),
        "ora_template_amplitudes": (c_dp, [vp, i]),
        "ora_set_gain": (None, [vp, i, d]),
        "ora_eval_sed": (d, [vp, i, i, i, i, c_dp]),
        "ora_eval_signal": (d, [vp, i, i, i, i, c_dp]),
        "ora_cg_create": (vp, [vp, i, i, d, ip, i]),
        "ora_cg_destroy": (None, [vp]),
        "ora_cg_n": (ll, [vp, i]),
        "ora_cg_m": (ll, [vp, i]),
        "ora_cg_x": (c_dp, [vp, i]),
        "ora_compute_rhs": (None, [vp, i, c_dp]),
        "ora_compute_Ax": (None, [vp, c_dp, i, c_dp]),
        "ora_compute_sample_vector": (None, [vp, c_dp, i, c_dp, i]),
        "ora_cg_search": (i, [vp, i, c_dp, i, c_dp, i, c_dp, c_dp, i]),
        "ora_unpack_amplitudes": (None, [vp, i]),
        "ora_sample_cg_group": (i, [vp, i, c_dp, i, ip, c_dp]),
        "ora_update_sky_model": (None, [vp]),
        "ora_compute_chisq": (d, [vp, c_dp]),
        "ora_mask_avg": (d, [vp, i, i, i]),
        "ora_eval_normal_prior": (d, [d, d, d]),
        "ora_rand_normal_from_uniform": (d, [d, d, d, d]),
        "ora_sample_index_mh": (i, [vp, i, i, i, i, i, c_dp, c_dp, c_dp, C.POINTER(C.c_ubyte), c_dp]),
        "ora_sample_spectral_parameters": (i, [vp, i, i, c_dp, c_dp]),
        "ora_tune_step": (i, [vp, i, i, i, i, i, c_dp, c_dp, i]),
        "ora_tune_step_from": (i, [vp, i, i, i, i, i, c_dp, c_dp, i, c_dp]),
        "ora_perpixel_tune_start": (None, [vp, i, i, i, c_dp]),
        "ora_fit_band_gain": (d, [vp, i, i, i, d]),
        "ora_philox_uniform2": (None, [C.c_ulonglong, C.c_uint, C.c_ulonglong, c_dp, c_dp]),
        "ora_philox_normals": (None, [C.c_ulonglong, C.c_uint, C.c_ulonglong, ll, c_dp]),
        "ora_philox_uniforms": (None, [C.c_ulonglong, C.c_uint, C.c_ulonglong, ll, c_dp]),
        "ora_philox_raw": (None, [C.POINTER(C.c_uint), C.c_uint, C.c_uint]),
        "ora_num_threads": (i, []),
        "ora_set_num_threads": (None, [i]),
        "ora_nest2ring": (ll, [ll, ll]),
        "ora_ring2nest": (ll, [ll, ll]),
        "ora_udgrade_ring": (None, [c_dp, ll, c_dp, ll, i]),
        "ora_udgrade_rms": (None, [c_dp, ll, c_dp, ll, i]),
        "ora_udgrade_mask": (None, [c_dp, ll, c_dp, ll, i, d]),
        "ora_get_T_CMB": (d, []),
        "ora_set_T_CMB": (None, [d]),
        "ora_offset": (c_dp, [vp]),
    }
    for name_, (res, args) in sigs.items():
        fn = getattr(lib, name_)
        fn.restype = res
        fn.argtypes = args
    _LIBS[name] = lib
    return lib


class Oracle:
    """The reference's objects (ddata, component_list, cg_groups) built from a RunConfig + Sky."""

    def __init__(self, cfg, sky, omp: bool = False):
        from dang_b200.config import (COMP_TYPES, INDEX_MODES, LNL_TYPES, PRIOR_TYPES,
                                      return_poltype_flag)
        self.lib = lib = load(omp)
        self.cfg = cfg
        self.npix, self.nmaps, self.nbands = cfg.npix, cfg.nmaps, cfg.nbands
        self.st = lib.ora_create(cfg.nside, cfg.npix, cfg.nmaps, cfg.nbands, len(cfg.comps))
        for j, b in enumerate(cfg.bands):
            if b.is_delta:
                lib.ora_set_band(self.st, j, b.nu_ghz, 0, None, None)
            else:
                nu = np.ascontiguousarray(b.bp_nu_ghz, dtype=np.float64)
                tau = np.ascontiguousarray(b.bp_tau, dtype=np.float64)
                lib.ora_set_band(self.st, j, b.nu_ghz, len(nu), _dp(nu), _dp(tau))
        lib.ora_set_maps(self.st, _dp(sky.sig), _dp(sky.rms), _dp(sky.mask), _dp(sky.gain),
                         _dp(sky.offset))
        lib.ora_set_pol_type(self.st, *cfg.pol_type)
        for ic, c in enumerate(cfg.comps):
            if c.type in ("template", "monopole", "hi_fit"):
                idx = _dp(sky.indices[c.label]) if c.type == "hi_fit" else None
                rc = lib.ora_set_component(self.st, ic, COMP_TYPES[c.type], c.label.encode(),
                                           c.nu_ref_ghz, c.cg_group, int(c.amp_sample), None, idx)
                assert rc == 0
                corr = (C.c_int * cfg.nbands)(*[int(bool(v)) for v in c.corr])
                if c.type == "monopole":
                    rc = lib.ora_set_template(self.st, ic, None,
                                              _dp(np.ascontiguousarray(sky.template_amplitudes[c.label])), corr,
                                              int(sum(map(bool, c.corr))))
                    assert rc == 0, rc
                    continue
                # the oracle divides by the per-plane maximum itself (the constructor's :574-577); the
                # map handed over here is already normalised, so that division is by 1
                rc = lib.ora_set_template(self.st, ic, _dp(np.ascontiguousarray(sky.template[c.label])),
                                          _dp(np.ascontiguousarray(sky.template_amplitudes[c.label])), corr,
                                          int(sum(map(bool, c.corr))))
                assert rc == 0, rc
                if c.type != "hi_fit":
                    continue
            else:
                rc = lib.ora_set_component(self.st, ic, COMP_TYPES[c.type], c.label.encode(),
                                       c.nu_ref_ghz, c.cg_group, int(c.amp_sample),
                                       _dp(sky.amplitude[c.label]), _dp(sky.indices[c.label]))
            assert rc == 0
            for k, s in enumerate(c.indices):
                flags = return_poltype_flag(s.poltype)
                fl = (C.c_int * len(flags))(*flags)
                g = np.asarray(s.gauss, dtype=np.float64)
                u = np.asarray(s.uni, dtype=np.float64)
                rc = lib.ora_set_index(self.st, ic, k, int(s.sample), INDEX_MODES[s.region],
                                       LNL_TYPES[s.lnl_type], PRIOR_TYPES[s.prior], _dp(g), _dp(u),
                                       s.step, int(not s.tune), s.samp_nside or cfg.nside, fl,
                                       len(flags))
                assert rc == 0
        self.cg = []
        for ig, g in enumerate(cfg.cg_groups):
            flags = return_poltype_flag(g.poltype)
            fl = (C.c_int * len(flags))(*flags)
            self.cg.append(lib.ora_cg_create(self.st, ig + 1, g.max_iter, g.converge, fl, len(flags)))

    def close(self):
        if self.st:
            for g in self.cg:
                self.lib.ora_cg_destroy(g)
            self.lib.ora_destroy(self.st)
            self.st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- views into the oracle's state (no copies) ---
    def _view(self, ptr, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)

    def amplitude(self, ic):
        return self._view(self.lib.ora_amplitude(self.st, ic), (self.nmaps, self.npix))

    def indices(self, ic):
        nind = self.lib.ora_nindices(self.st, ic)
        return self._view(self.lib.ora_indices(self.st, ic), (nind, self.nmaps, self.npix))

    def template_amplitudes(self, ic):
        return self._view(self.lib.ora_template_amplitudes(self.st, ic), (self.nmaps, self.nbands))

    def sky_model(self):
        return self._view(self.lib.ora_sky_model(self.st), (self.nbands, self.nmaps, self.npix))

    def res_map(self):
        return self._view(self.lib.ora_res_map(self.st), (self.nbands, self.nmaps, self.npix))

    def chi_map(self):
        return self._view(self.lib.ora_chi_map(self.st), (self.nmaps, self.npix))

    def cg_x(self, ig=0, flag_n=0):
        n = self.lib.ora_cg_n(self.cg[ig], flag_n)
        return self._view(self.lib.ora_cg_x(self.cg[ig], flag_n), (n,))

    # --- the reference's operations ---
    def eval_sed(self, ic, band, pix, map_n, theta=None):
        th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        return self.lib.ora_eval_sed(self.st, ic, band, pix, map_n, _dp(th))

    def compute_rhs(self, ig=0, flag_n=0):
        b = np.zeros(self.lib.ora_cg_n(self.cg[ig], flag_n))
        self.lib.ora_compute_rhs(self.cg[ig], flag_n, _dp(b))
        return b

    def compute_Ax(self, x, ig=0, flag_n=0):
        res = np.zeros_like(x)
        self.lib.ora_compute_Ax(self.cg[ig], _dp(np.ascontiguousarray(x)), flag_n, _dp(res))
        return res

    def compute_sample_vector(self, eta, ig=0, flag_n=0, fix_q1=False):
        res = np.zeros(self.lib.ora_cg_n(self.cg[ig], flag_n))
        self.lib.ora_compute_sample_vector(self.cg[ig], _dp(np.ascontiguousarray(eta)), flag_n,
                                           _dp(res), int(fix_q1))
        return res

    def sample_cg_group(self, ig=0, ml_mode=1, eta=None, fix_q1=False):
        """rhs -> cg_search -> unpack per flag, then update_sky_model (dang_cg_mod.f90:166-172)."""
        nflag = len(self.cfg.cg_groups[ig].poltype.split(","))
        niter = (C.c_int * 3)()
        delta = np.zeros(3)
        e = None if eta is None else np.ascontiguousarray(eta, dtype=np.float64)
        self.lib.ora_sample_cg_group(self.cg[ig], ml_mode, _dp(e), int(fix_q1), niter, _dp(delta))
        return list(niter)[:nflag], delta[:nflag].copy()

    def cg_search_trace(self, ig=0, flag_n=0, ml_mode=1, eta=None, fix_q1=False, trace_len=256):
        """compute_rhs -> cg_search -> unpack_amplitudes for one (group, flag), with the delta trace."""
        b = self.compute_rhs(ig, flag_n)
        trace = np.full(trace_len, np.nan)
        delta = C.c_double()
        e = None if eta is None else np.ascontiguousarray(eta, dtype=np.float64)
        it = self.lib.ora_cg_search(self.cg[ig], flag_n, _dp(b), ml_mode, _dp(e), int(fix_q1),
                                    C.byref(delta), _dp(trace), trace_len)
        self.lib.ora_unpack_amplitudes(self.cg[ig], flag_n)
        return it, delta.value, trace[:it]

    def update_sky_model(self):
        self.lib.ora_update_sky_model(self.st)

    def compute_chisq(self):
        planes = np.zeros(self.nmaps)
        chisq = self.lib.ora_compute_chisq(self.st, _dp(planes))
        return chisq, planes

    def sample_index_mh(self, ic, nind, map_n, nsample, ml_mode, z, u, want_trace=False):
        z = np.ascontiguousarray(z, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        acc = C.c_double()
        dec = np.zeros(z.size, dtype=np.uint8)
        tr = np.zeros(z.size) if want_trace else None
        rc = self.lib.ora_sample_index_mh(self.st, ic, nind, map_n, nsample, ml_mode, _dp(z),
                                          _dp(u), C.byref(acc),
                                          dec.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(tr))
        if rc != 0:
            raise RuntimeError(f"ora_sample_index_mh rc={rc}")
        return acc.value, dec, tr

    def sample_spectral_parameters(self, nsample, ml_mode, z, u):
        z = np.ascontiguousarray(z, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        return self.lib.ora_sample_spectral_parameters(self.st, nsample, ml_mode, _dp(z), _dp(u))

    def fit_band_gain(self, map_n, band, ml_mode, z):
        return self.lib.ora_fit_band_gain(self.st, map_n, band, ml_mode, z)

    def tune_step(self, ic, nind, map_n, nsample, ml_mode, z, u, max_blocks, perpixel_start=False):
        """tune_spectral_parameter_length; perpixel_start: start at the map's mean as the per-pixel call site does."""
        z = np.ascontiguousarray(z, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        th = None
        if perpixel_start:
            th = np.zeros(2)
            self.lib.ora_perpixel_tune_start(self.st, ic, nind, map_n, _dp(th))
        nb = self.lib.ora_tune_step_from(self.st, ic, nind, map_n, nsample, ml_mode, _dp(z), _dp(u),
                                         max_blocks, _dp(th))
        return nb, self.lib.ora_step_size(self.st, ic, nind)


def philox_normals(seed: int, stream: int, slot0: int, n: int) -> np.ndarray:
    z = np.zeros(n)
    load().ora_philox_normals(seed, stream, slot0, n, _dp(z))
    return z


def philox_uniforms(seed: int, stream: int, slot0: int, n: int) -> np.ndarray:
    u = np.zeros(n)
    load().ora_philox_uniforms(seed, stream, slot0, n, _dp(u))
    return u


def planck_rj(nu_hz, T):
    """B_nu(nu, T) / compute_bnu_prime_RJ(nu) * 1e6 (evaluate_hi_fit / evaluate_T_cmb on a delta band), numpy."""
    h, k_B, c = 1.0545726691251021e-34 * 2.0 * np.pi, 1.3806503e-23, 2.99792458e8
    B = ((2.0 * h * nu_hz ** 3.0) / c ** 2.0) * (1.0 / (np.exp((h * nu_hz) / (k_B * np.asarray(T))) - 1))
    return B / (2.0 * k_B * nu_hz ** 2.0 / c ** 2.0) * 1e6


def udgrade(kind: str, data: np.ndarray, nside_in: int, nside_out: int, threshold: float = 0.5) -> np.ndarray:
    """udgrade_ring / udgrade_rms / udgrade_mask on [nmaps][npix] RING maps (HEALPix udgrade_nr as published +
    src/dang_util_mod.f90:341-376)."""
    lib = load()
    data = np.ascontiguousarray(data, dtype=np.float64)
    nmaps = data.shape[0]
    out = np.zeros((nmaps, 12 * nside_out * nside_out))
    if kind == "ring":
        lib.ora_udgrade_ring(_dp(data), nside_in, _dp(out), nside_out, nmaps)
    elif kind == "rms":
        lib.ora_udgrade_rms(_dp(data), nside_in, _dp(out), nside_out, nmaps)
    else:
        lib.ora_udgrade_mask(_dp(data), nside_in, _dp(out), nside_out, nmaps, threshold)
    return out
