"""Host-side description of a dang run: the subset of `dang_params` / `dang_comps` /
`dang_cg_group` fields the Gibbs hot path reads.

Field names follow the reference's parameter-file keys (src/dang_param_mod.f90:370-400,
461-513, 532-600, 661-687) so a parameter file maps 1:1 onto these objects; nothing here
touches the device.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

MISSVAL = -1.6375e30  # src/dang_util_mod.f90:19

# enums shared with include/dang_gpu.h
COMP_TYPES = {"power-law": 1, "mbb": 2, "freefree": 3, "lognormal": 4, "cmb": 5, "template": 6,
              "T_cmb": 7, "monopole": 8, "hi_fit": 9}
LNL_TYPES = {"chisq": 0, "marginal": 1, "prior": 2}
PRIOR_TYPES = {"uniform": 0, "gaussian": 1, "jeffreys": 2}
ML_MODES = {"optimize": 0, "sample": 1}
INDEX_MODES = {"fullsky": 1, "per-pixel": 2}
NINDICES = {"power-law": 1, "mbb": 2, "freefree": 1, "lognormal": 2, "cmb": 0, "template": 0,
            "T_cmb": 1, "monopole": 0, "hi_fit": 1}


def return_poltype_flag(string: str) -> List[int]:
    """Bit flags of a comma-separated poltype list (src/dang_util_mod.f90:228-292).

    'T'->1, 'Q'->2, 'U'->4, 'Q+U'->8; several entries give several sequential solves,
    returned in ascending bit order exactly as the reference builds its flag array.
    """
    local_flag = 0
    for tok in string.strip().split(","):
        tok = tok.strip()
        if tok == "T":
            local_flag += 1
        elif tok == "Q":
            local_flag += 2
        elif tok == "U":
            local_flag += 4
        elif tok == "Q+U":
            local_flag += 8
        elif tok == "T+Q+U":
            local_flag = 0  # dead in the reference (iand(flag,0))
    return [1 << j for j in range(4) if local_flag & (1 << j)]


def flag_to_map_n(flag: int) -> int:
    """pol flag -> the `map_n` argument of sample_index_mh (src/dang_sample_mod.f90:53-67)."""
    if flag & 1:
        return 1
    if flag & 2:
        return 2
    if flag & 4:
        return 3
    if flag & 8:
        return -1
    raise ValueError("There is something wrong with the poltype flag")


@dataclass
class Band:
    """bp(j) of src/dang_bp_mod.f90:7-12: delta band, or tabulated (nu [GHz], tau)."""

    nu_ghz: float
    label: str = ""
    bp_nu_ghz: Optional[np.ndarray] = None
    bp_tau: Optional[np.ndarray] = None

    @property
    def is_delta(self) -> bool:
        return self.bp_nu_ghz is None


@dataclass
class IndexSpec:
    """One spectral index of a component (COMP_<BETA|T>... keys)."""

    label: str
    init: float
    sample: bool = False
    region: str = "per-pixel"          # COMP_*_REGION: fullsky | per-pixel
    lnl_type: str = "chisq"            # COMP_*_LNL_TYPE
    prior: str = "uniform"             # COMP_*_PRIOR
    gauss: Sequence[float] = (0.0, 1.0)
    uni: Sequence[float] = (-1e30, 1e30)
    step: float = 0.05                 # COMP_*_STEPSIZE
    poltype: str = "Q+U"               # COMP_*_POLTYPE
    samp_nside: Optional[int] = None   # COMP_*_SAMP_NSIDE (None -> map nside)
    tune: bool = False                 # COMP_*_TUNE_STEPSIZE


@dataclass
class Component:
    label: str
    type: str
    nu_ref_ghz: float
    cg_group: int = 1
    amp_sample: bool = True
    indices: List[IndexSpec] = field(default_factory=list)
    # type 'template' (COMPnn_FITbbb keys, src/dang_param_mod.f90:750-767): bands the template is fitted to
    corr: Optional[Sequence[bool]] = None


@dataclass
class CGGroup:
    sample: bool = True
    max_iter: int = 100
    converge: float = 1e-12
    poltype: str = "Q+U"


@dataclass
class RunConfig:
    name: str
    nside: int
    bands: List[Band]
    comps: List[Component]
    cg_groups: List[CGGroup]
    nsample: int = 20
    ngibbs: int = 10
    ml_mode: str = "sample"
    tqu: str = "QU"                    # ddata%pol_type: planes compute_chisq sums over
    nmaps: int = 3

    @property
    def npix(self) -> int:
        return 12 * self.nside * self.nside

    @property
    def nbands(self) -> int:
        return len(self.bands)

    @property
    def pol_type(self):
        """(lo, hi) 1-based plane range of ddata%pol_type (src/dang_data_mod.f90:507)."""
        return {"T": (1, 1), "QU": (2, 3), "TQU": (1, 3)}[self.tqu]
