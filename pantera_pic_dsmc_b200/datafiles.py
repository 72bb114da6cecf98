"""Readers for PANTERA's (unchanged) species / VSS / reaction / wall-reaction data files and the
derived tables the hot path consumes.

In a PANTERA run these tables are built by the Fortran host (READINPUT and friends stay Fortran,
see BASELINE.json north_star) and reach the device through ``pgpu_set_species / set_vss /
set_reactions / set_wall_reactions``.  The harness needs the same numbers, so the derivations are
restated here:

* species file columns ``name molwt mass rotdof rotrel vibdof vibrel vibtemp spwt charge``
  (src/initialization.f90:577-578);
* VSS file ``name diam omega tref alpha`` and the pair tables VSS_SIGMAS/OMEGAS/ALPHAS/GREFS + SIGMAMAX
  (src/initialization.f90:652-702), including the single-precision literal 1.2354 used when omega = 1/2;
* reactions file: ``A + B --> C + D [+ E [+ F]]`` then ``tce A n Ea[eV]`` or ``lxcat Ea[eV] file``
  (src/initialization.f90:1689-1793) and the TCE coefficients C1..C3 of INITREACTIONS (:2681-2724);
* wall reactions ``R --> P|none`` + probability (:1881-1910).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

PI = 3.1415926535897932
KB = 1.380649e-23
QE = 1.602176634e-19
EPS0 = 8.8541878128e-12

FIXED_RATE, TCE, LXCAT = 0, 1, 2


def _lines(path):
    with open(path) as f:
        for raw in f:
            line = raw.split("!")[0].strip()
            if line:
                yield line


@dataclass
class SpeciesTable:
    names: List[str]
    molwt: np.ndarray
    mass: np.ndarray
    rotdof: np.ndarray
    rotrel: np.ndarray
    vibdof: np.ndarray
    vibrel: np.ndarray
    vibtemp: np.ndarray
    spwt: np.ndarray
    charge: np.ndarray
    diam: Optional[np.ndarray] = None
    omega: Optional[np.ndarray] = None
    tref: Optional[np.ndarray] = None
    alpha: Optional[np.ndarray] = None
    sigmas: Optional[np.ndarray] = None  # (ns, ns)
    omegas: Optional[np.ndarray] = None
    alphas: Optional[np.ndarray] = None
    grefs: Optional[np.ndarray] = None
    sigmamax: float = 0.0

    @property
    def ns(self):
        return len(self.names)

    def id(self, name: str) -> int:
        """1-based species id (SPECIES_NAME_TO_ID)."""
        return self.names.index(name) + 1

    def subset(self, names: List[str]) -> "SpeciesTable":
        idx = [self.names.index(n) for n in names]
        kw = {}
        for k in ("molwt", "mass", "rotdof", "rotrel", "vibdof", "vibrel", "vibtemp", "spwt", "charge",
                  "diam", "omega", "tref", "alpha"):
            v = getattr(self, k)
            kw[k] = None if v is None else v[idx].copy()
        t = SpeciesTable(names=list(names), **kw)
        if t.diam is not None:
            t.build_vss_tables()
        return t

    def build_vss_tables(self):
        """Pair tables as READ_VSS builds them (src/initialization.f90:675-702)."""
        ns = self.ns
        self.sigmas = np.zeros((ns, ns))
        self.omegas = np.zeros((ns, ns))
        self.alphas = np.zeros((ns, ns))
        self.grefs = np.zeros((ns, ns))
        sigmamax = 0.0
        for i in range(ns):
            for j in range(ns):
                m1, m2 = self.mass[i], self.mass[j]
                mred = m1 * m2 / (m1 + m2)
                omega = 0.5 * (self.omega[i] + self.omega[j])
                tref = 0.5 * (self.tref[i] + self.tref[j])
                if abs(omega - 0.5) < 1e-6:
                    gref = (2.0 * KB * tref / mred) ** 0.5 * float(np.float32(1.2354))
                else:
                    gref = (2.0 * KB * tref / mred) ** 0.5 * math.gamma(2.5 - omega) ** (-0.5 / (omega - 0.5))
                self.grefs[i, j] = gref
                self.sigmas[i, j] = PI * (0.5 * (self.diam[i] + self.diam[j])) ** 2
                self.alphas[i, j] = 0.5 * (self.alpha[i] + self.alpha[j])
                self.omegas[i, j] = omega
                sigmamax = max(sigmamax, self.sigmas[i, j])
        self.sigmamax = sigmamax


def read_species(path: str) -> SpeciesTable:
    rows = [ln.split() for ln in _lines(path)]
    f = lambda c: np.array([float(r[c]) for r in rows])
    return SpeciesTable(
        names=[r[0] for r in rows], molwt=f(1), mass=f(2), rotdof=f(3).astype(np.int32), rotrel=f(4),
        vibdof=f(5).astype(np.int32), vibrel=f(6), vibtemp=f(7), spwt=f(8), charge=f(9),
    )


def read_vss(path: str, sp: SpeciesTable) -> SpeciesTable:
    ns = sp.ns
    sp.diam, sp.omega, sp.tref, sp.alpha = (np.zeros(ns) for _ in range(4))
    for ln in _lines(path):
        t = ln.split()
        i = sp.names.index(t[0])
        sp.diam[i], sp.omega[i], sp.tref[i], sp.alpha[i] = (float(x) for x in t[1:5])
    sp.build_vss_tables()
    return sp


def read_lxcat_table(path: str):
    """rows ``energy[eV] sigma[m^2]`` between two ``-----`` lines (src/initialization.f90:1753-1785);
    energies converted to joule."""
    e, cs = [], []
    inside = False
    with open(path) as f:
        for raw in f:
            s = raw.strip()
            if s.startswith("-----"):
                if inside:
                    break
                inside = True
                continue
            if inside and s:
                a, b = s.split()[:2]
                e.append(float(a) * QE)
                cs.append(float(b))
    return np.array(e), np.array(cs)


def tce_coefficients(sp: SpeciesTable, r1: int, r2: int, A: float, n: float) -> tuple:
    """C1, C2, C3 of INITREACTIONS (src/initialization.f90:2696-2716); r1, r2 1-based."""
    i, j = r1 - 1, r2 - 1
    zeta = 0.5 * (sp.rotdof[i] + sp.vibdof[i] + sp.rotdof[j] + sp.vibdof[j])
    omega = 0.5 * (sp.omega[i] + sp.omega[j])
    sigma = 0.25 * PI * (sp.diam[i] + sp.diam[j]) ** 2
    tref = 0.5 * (sp.tref[i] + sp.tref[j])
    m1, m2 = sp.mass[i], sp.mass[j]
    mr = m1 * m2 / (m1 + m2)
    eps = 2.0 if r1 == r2 else 1.0
    c1 = (math.sqrt(PI) * eps * A / (2.0 * sigma) * math.gamma(zeta + 2.5 - omega) / math.gamma(zeta + n + 1.5)
          * math.sqrt(mr / (2.0 * KB * tref)) * tref ** (1.0 - omega) / KB ** (n - 1.0 + omega))
    return c1, n - 1.0 + omega, zeta + 1.5 - omega


def read_reactions(path: str, sp: SpeciesTable) -> List[dict]:
    """Reaction dicts in the layout ``Lib.set_reactions`` takes (species ids 1-based)."""
    out: List[dict] = []
    lines = list(_lines(path))
    base = os.path.dirname(path)
    k = 0
    while k < len(lines):
        # the arrow is '-->' or '-CEX->' (charge exchange), src/initialization.f90:1688,1719-1723
        is_cex = "-CEX->" in lines[k]
        lhs, rhs = lines[k].split("-CEX->" if is_cex else "-->")
        reac = [t.strip() for t in lhs.split(" + ")]
        prod = [t.strip() for t in rhs.split(" + ")]
        spec = lines[k + 1].split()
        k += 2
        r = dict(r1=sp.id(reac[0]), r2=sp.id(reac[1]), nprod=len(prod), is_cex=int(is_cex), p3=-1, p4=-1)
        for q, name in enumerate(prod):
            r[f"p{q + 1}"] = sp.id(name)
        if spec[0].lower() == "tce":
            A, n, ea_ev = float(spec[1]), float(spec[2]), float(spec[3])
            r.update(type=TCE, A=A, n=n, ea=ea_ev * QE)
            r["c1"], r["c2"], r["c3"] = tce_coefficients(sp, r["r1"], r["r2"], A, n)
        elif spec[0].lower() == "lxcat":
            te, tcs = read_lxcat_table(os.path.join(base, spec[2]))
            r.update(type=LXCAT, ea=float(spec[1]) * QE, table_energy=te, table_cs=tcs)
            # LXCat tables raise the global SIGMAMAX (src/initialization.f90:1783)
            if len(tcs) and sp.sigmas is not None:
                sp.sigmamax = max(sp.sigmamax, float(np.max(tcs)))
        else:
            raise ValueError(f"unknown reaction kind {spec[0]}")
        out.append(r)
    return out


def read_wall_reactions(path: str, sp: SpeciesTable) -> Dict[str, list]:
    lines = list(_lines(path))
    r_sp, p1_sp, nprod, prob = [], [], [], []
    for k in range(0, len(lines), 2):
        lhs, rhs = (t.strip() for t in lines[k].split("-->"))
        r_sp.append(sp.id(lhs))
        if rhs.lower() == "none":
            p1_sp.append(-1)
            nprod.append(0)
        else:
            p1_sp.append(sp.id(rhs))
            nprod.append(1)
        prob.append(float(lines[k + 1].split()[0]))
    return dict(r_sp=r_sp, p1_sp=p1_sp, nprod=nprod, prob=prob)


def apply_species(lib, sp: SpeciesTable, with_vss: bool = True):
    """Push a species table (and its VSS pair tables) through the C ABI."""
    lib.set_species(sp.mass, sp.charge, sp.rotdof, sp.vibdof, sp.rotrel, sp.vibrel)
    if with_vss and sp.sigmas is not None:
        lib.set_vss(sp.sigmas, sp.omegas, sp.alphas, sp.grefs, sp.sigmamax)
