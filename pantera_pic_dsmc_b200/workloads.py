"""Synthetic workloads (BASELINE.json configs C1..C5, concretised in SURVEY.md section 8d).

A workload is a plain description (config struct, species table, mesh, boundary conditions,
initial particles) that :func:`apply` pushes through the C ABI of either library (the CUDA
product or the CPU oracle), in the order the Fortran host would: config, species, VSS,
reactions, mesh, BCs, particles.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi, datafiles as df, meshgen as mg

DATA_DIR = os.path.join(capi.REPO_ROOT, "tests", "data")
KB = df.KB


@dataclass
class Workload:
    name: str
    cfg: capi.Config
    species: df.SpeciesTable
    particles: np.ndarray
    mesh: Optional[mg.Mesh] = None
    bcs: Optional[dict] = None
    reactions: List[dict] = field(default_factory=list)
    wall_reactions: Optional[dict] = None
    walls: Optional[dict] = None
    mcc_background: Optional[tuple] = None
    mcc_cell_density: Optional[np.ndarray] = None
    efield: Optional[np.ndarray] = None
    notes: str = ""


def apply(lib: capi.Lib, w: Workload, upload: bool = True):
    lib.set_config(w.cfg)
    df.apply_species(lib, w.species)
    if w.reactions:
        lib.set_reactions(w.reactions)
    if w.wall_reactions:
        lib.set_wall_reactions(**w.wall_reactions)
    if w.mcc_background:
        lib.set_mcc_background(*w.mcc_background)
    lib.set_mcc_cell_density(w.mcc_cell_density)
    if w.mesh is not None:
        lib.set_mesh(w.mesh)
    if w.bcs:
        lib.set_bcs(**w.bcs)
    if w.walls:
        lib.set_walls(**w.walls)
    if w.efield is not None:
        lib.set_efield(w.efield)
    if upload:
        lib.upload_particles(w.particles)
    return lib


def load_species(kind: str, names: Optional[List[str]] = None) -> df.SpeciesTable:
    sp = df.read_species(os.path.join(DATA_DIR, f"{kind}.species"))
    df.read_vss(os.path.join(DATA_DIR, f"{kind}.vss"), sp)
    if names is not None:
        sp = sp.subset(names)
    return sp


def maxwellian(rng: np.random.Generator, n: int, mass, T, u=(0.0, 0.0, 0.0)) -> np.ndarray:
    """n velocities from a (possibly anisotropic) Maxwellian; mass scalar or (n,)."""
    T = np.broadcast_to(np.asarray(T, dtype=np.float64), (3,))
    sig = np.sqrt(KB * T[None, :] / np.asarray(mass, dtype=np.float64).reshape(-1, 1))
    return rng.standard_normal((n, 3)) * sig + np.asarray(u)[None, :]


def new_particles(n: int) -> np.ndarray:
    p = np.zeros(n, dtype=capi.PARTICLE_DTYPE)
    p["ID"] = np.arange(n, dtype=np.int64) << 8  # PROC_ID + ISHFT(counter, 8), src/particle.f90:75
    return p


def sample_in_tets(rng, mesh: mg.Mesh, cells0: np.ndarray) -> np.ndarray:
    """Uniform points inside the given tets/triangles/segments (INITIAL_SEED's simplex sampling idea,
    src/initialization.f90:2112-2148; order of draws is not part of the contract)."""
    d = mesh.dims
    cn = np.asarray(mesh.cell_nodes)[:, cells0] - 1  # (d+1, n)
    P = np.asarray(mesh.node_coords)[:, cn]  # (3, d+1, n)
    e = rng.exponential(size=(d + 1, cells0.size))
    lam = e / e.sum(axis=0, keepdims=True)
    return np.einsum("kvn,vn->nk", P, lam)


# ------------------------------------------------------------------------------------------------
# C1: 0-D/1-D argon heat-bath relaxation, VSS collisions only, structured grid, triply periodic
# ------------------------------------------------------------------------------------------------
def c1_argon_box(n: int = 100_000, nx: int = 1, ny: int = 1, seed: int = 1, T=(600.0, 300.0, 300.0),
                 nrho: float = 1e20, dt_frac: float = 0.2, periodic=True, rng_seed_global: int = 65644) -> Workload:
    sp = load_species("ar", ["Ar"])
    m = sp.mass[0]
    # equilibrium VSS collision frequency per molecule, T_eq = mean(T)
    Teq = float(np.mean(T))
    d, om, tref = sp.diam[0], sp.omega[0], sp.tref[0]
    nu = 4.0 * d * d * nrho * np.sqrt(np.pi * KB * tref / m) * (Teq / tref) ** (1.0 - om)
    dt = dt_frac / nu
    vol = 1.0
    fnum = nrho * vol / n
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=nx, ny=ny, xmin=0.0, xmax=1.0, ymin=0.0,
                           ymax=1.0, zmin=-0.5, zmax=0.5, dt=dt, fnum=fnum, collision_type=capi.DSMC,
                           bool_periodic=[periodic] * 4, bool_specular=[not periodic] * 4, bool_z_periodic=1,
                           rng_seed_global=rng_seed_global)
    rng = np.random.default_rng(seed)
    p = new_particles(n)
    p["X"] = rng.random(n)
    p["Y"] = rng.random(n)
    p["Z"] = rng.random(n) - 0.5
    v = maxwellian(rng, n, m, T)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["DTRIM"] = dt
    p["S_ID"] = 1
    dx, dy = 1.0 / nx, 1.0 / ny
    p["IC"] = (p["X"] / dx).astype(np.int32) + nx * (p["Y"] / dy).astype(np.int32) + 1
    return Workload("C1 argon VSS box", cfg, sp, p, notes=f"nu_eq={nu:.6g} dt={dt:.6g}")


# ------------------------------------------------------------------------------------------------
# C5: 3-D unstructured (structured-to-tet) channel, argon, diffuse walls in y, specular in z and x
# ------------------------------------------------------------------------------------------------
def c5_channel(ncx: int = 8, ncy: int = 4, ncz: int = 4, ppc: float = 40.0, seed: int = 2, T0: float = 300.0,
               nrho: float = 1e20, cell_mfp: float = 0.5, dt_frac: float = 0.2, u_drift: float = 100.0,
               wall_bc=("specular", "specular", "diffuse", "diffuse", "specular", "specular"), wall_temp=400.0,
               jitter: float = 0.0, rng_seed_global: int = 65644, collision_type=capi.DSMC,
               n_override: Optional[int] = None) -> Workload:
    """Box channel of ncx x ncy x ncz cubes (6 tets each).  Cube edge = cell_mfp mean free paths,
    dt = dt_frac / nu_eq.  x end caps specular (constant-N run: unstructured grids have no particle-side
    periodic BC in the reference, timecycle.f90:1611-1614), y walls diffuse at ``wall_temp``, z specular."""
    sp = load_species("ar", ["Ar"])
    m = sp.mass[0]
    d, om, tref = sp.diam[0], sp.omega[0], sp.tref[0]
    nu = 4.0 * d * d * nrho * np.sqrt(np.pi * KB * tref / m) * (T0 / tref) ** (1.0 - om)
    cbar = np.sqrt(8.0 * KB * T0 / (np.pi * m))
    mfp = cbar / nu
    h = cell_mfp * mfp
    hi = (ncx * h, ncy * h, ncz * h)
    mesh = mg.box_tets(ncx, ncy, ncz, lo=(0.0, 0.0, 0.0), hi=hi, jitter=jitter, seed=seed)
    nc = mesh.nc
    n = int(n_override) if n_override is not None else int(round(ppc * nc))
    vol = hi[0] * hi[1] * hi[2]
    fnum = nrho * vol / n
    dt = dt_frac / nu
    cfg = capi.make_config(dims=3, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=hi[0], ymin=0.0, ymax=hi[1],
                           zmin=0.0, zmax=hi[2], dt=dt, fnum=fnum, collision_type=collision_type,
                           rng_seed_global=rng_seed_global)
    kinds = {"vacuum": capi.VACUUM, "specular": capi.SPECULAR, "diffuse": capi.DIFFUSE, "cll": capi.CLL}
    bcs = dict(particle_bc=[kinds[k] for k in wall_bc], wall_temp=[wall_temp] * 6, acc_n=[0.8] * 6, acc_t=[0.8] * 6)
    rng = np.random.default_rng(seed)
    # particles: cells drawn with probability ~ volume, position uniform in the tet
    cum = np.cumsum(mesh.cell_volumes)
    cells0 = np.searchsorted(cum, rng.random(n) * cum[-1]).astype(np.int64)
    cells0 = np.minimum(cells0, nc - 1)
    xyz = sample_in_tets(rng, mesh, cells0)
    p = new_particles(n)
    p["X"], p["Y"], p["Z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    v = maxwellian(rng, n, m, T0, u=(u_drift, 0.0, 0.0))
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["DTRIM"] = dt
    p["S_ID"] = 1
    p["IC"] = cells0 + 1
    return Workload("C5 3-D tet channel", cfg, sp, p, mesh=mesh, bcs=bcs,
                    notes=f"h={h:.4g} mfp={mfp:.4g} nu={nu:.5g} dt={dt:.4g} nc={nc}")


# ------------------------------------------------------------------------------------------------
# C2-like: air (N2/O2) with TCE + Borgnakke-Larsen in a periodic structured box (heat bath)
# ------------------------------------------------------------------------------------------------
def c2_air_box(n: int = 50_000, nx: int = 4, ny: int = 4, seed: int = 3, T: float = 8000.0, nrho: float = 1e21,
               dt: float = 2e-7, species=("O2", "N2", "O", "N", "NO"), frac=(0.21, 0.79, 0.0, 0.0, 0.0),
               with_reactions: bool = True, rng_seed_global: int = 65644) -> Workload:
    sp = load_species("air", list(species))
    reactions = df.read_reactions(os.path.join(DATA_DIR, "air.tce"), sp) if with_reactions else []
    fnum = nrho * 1.0 / n
    cfg = capi.make_config(dims=2, grid_type=capi.RECTILINEAR_UNIFORM, nx=nx, ny=ny, xmin=0.0, xmax=1.0, ymin=0.0,
                           ymax=1.0, zmin=-0.5, zmax=0.5, dt=dt, fnum=fnum, collision_type=capi.DSMC,
                           bool_periodic=[1] * 4, bool_z_periodic=1, rng_seed_global=rng_seed_global)
    rng = np.random.default_rng(seed)
    p = new_particles(n)
    p["X"], p["Y"], p["Z"] = rng.random(n), rng.random(n), rng.random(n) - 0.5
    sid = rng.choice(len(species), size=n, p=np.asarray(frac) / np.sum(frac)) + 1
    mass = sp.mass[sid - 1]
    v = maxwellian(rng, n, mass, T)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["S_ID"] = sid
    # internal energies at T: dof 2 -> exponential with mean kT
    for name, dofs in (("EROT", sp.rotdof), ("EVIB", sp.vibdof)):
        has = dofs[sid - 1] == 2
        p[name][has] = rng.exponential(KB * T, size=int(has.sum()))
    p["DTRIM"] = dt
    p["IC"] = (p["X"] * nx).astype(np.int32) + nx * (p["Y"] * ny).astype(np.int32) + 1
    return Workload("C2 air TCE/BL box", cfg, sp, p, reactions=reactions)


def mcc_box(kind: str = "ar", n: int = 40_000, nx: int = 4, ny: int = 4, seed: int = 11, T: float = 3000.0,
            bg_T: float = 300.0, bg_dens: float = 2e21, dt: float = 2e-7, cell_density: bool = False,
            rng_seed_global: int = 65644) -> Workload:
    """Test particles in a periodic box colliding with a background gas through MCC_COLLISIONS
    (Collision_type: MCC, src/collisions.f90:1261-1595).  kind "ar": argon in argon (elastic VSS only);
    kind "air": O2/N2/O/N/NO test particles in an O2/N2 background with the TCE reactions of data/air.tce and
    Borgnakke-Larsen exchange.  cell_density: per-cell background (BOOL_BG_DENSITY_FILE) with empty cells."""
    if kind == "ar":
        w = c1_argon_box(n, nx=nx, ny=ny, seed=seed, T=(T, T, T))
        sp = w.species
        bg = ([sp.id("Ar")], [1.0])
    else:
        w = c2_air_box(n, nx=nx, ny=ny, seed=seed, T=T, dt=dt, frac=(0.2, 0.5, 0.1, 0.1, 0.1))
        sp = w.species
        bg = ([sp.id("O2"), sp.id("N2")], [0.21, 0.79])
    w.cfg.collision_type = capi.MCC
    w.cfg.mcc_bg_dens = bg_dens
    w.cfg.mcc_bg_ttra = bg_T
    w.cfg.dt = dt
    w.cfg.rng_seed_global = rng_seed_global
    w.particles["DTRIM"] = dt
    w.mcc_background = bg
    if cell_density:
        nc = nx * ny
        dens = np.zeros((len(sp.mass), nc), order="F")
        for k, s in enumerate(bg[0]):
            dens[s - 1, :] = bg[1][k] * bg_dens * (np.arange(nc) % 3)   # every third cell is empty
        w.mcc_cell_density = dens
    w.name = f"MCC {kind} box"
    return w


# ------------------------------------------------------------------------------------------------
# C3-like: 2-D (optionally axisymmetric) triangulated box with wall models on the four sides
# ------------------------------------------------------------------------------------------------
def c3_tri_box(nx: int = 8, ny: int = 6, n: int = 20_000, seed: int = 4, T0: float = 300.0, dt: float = 2e-6,
               lo=(0.0, 0.0), hi=(0.1, 0.05), axi: bool = False, species=("N2", "O2"),
               wall_bc=("vacuum", "specular", "cll", "diffuse"), wall_temp=(300.0, 300.0, 500.0, 700.0),
               acc=(0.8, 0.6), u_drift=(300.0, 0.0, 0.0), wall_reactions: bool = False, react=(0, 0, 0, 0),
               collision_type=capi.NO_COLL, nrho: float = 1e20, rng_seed_global: int = 65644,
               dump_bound_avg_every: int = 1) -> Workload:
    """Triangulated rectangle (2 CCW triangles per quad).  Physical groups 1 xmin, 2 xmax, 3 ymin, 4 ymax.
    Axisymmetric: y is the radius and ymin should be > 0 or carry a specular 'axis' (the reference's AXIS
    BC falls through to removal, timecycle.f90:1611-1614)."""
    sp = load_species("air", list(species))
    zmin, zmax = (0.0, 1.0) if axi else (-0.5, 0.5)
    mesh = mg.rect_tris(nx, ny, lo=lo, hi=hi, zmin=zmin, zmax=zmax, axi=axi)
    vol = float(np.sum(mesh.cell_volumes))
    fnum = nrho * vol / n
    cfg = capi.make_config(dims=2, axi=int(axi), grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=lo[0], xmax=hi[0],
                           ymin=lo[1], ymax=hi[1], zmin=zmin, zmax=zmax, dt=dt, fnum=fnum,
                           collision_type=collision_type, bool_z_periodic=0 if axi else 1,
                           rng_seed_global=rng_seed_global, dump_bound_avg_every=dump_bound_avg_every)
    kinds = {"vacuum": capi.VACUUM, "specular": capi.SPECULAR, "diffuse": capi.DIFFUSE, "cll": capi.CLL}
    bcs = dict(particle_bc=[kinds[k] for k in wall_bc], wall_temp=list(wall_temp), acc_n=[acc[0]] * 4,
               acc_t=[acc[1]] * 4, react=list(react))
    rng = np.random.default_rng(seed)
    cum = np.cumsum(mesh.cell_volumes)
    cells0 = np.minimum(np.searchsorted(cum, rng.random(n) * cum[-1]), mesh.nc - 1).astype(np.int64)
    xyz = sample_in_tets(rng, mesh, cells0)
    p = new_particles(n)
    p["X"], p["Y"] = xyz[:, 0], xyz[:, 1]
    p["Z"] = 0.0 if axi else rng.random(n) - 0.5
    sid = rng.integers(1, len(species) + 1, size=n)
    v = maxwellian(rng, n, sp.mass[sid - 1], T0, u=u_drift)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["S_ID"] = sid
    for name, dofs in (("EROT", sp.rotdof), ("EVIB", sp.vibdof)):
        has = dofs[sid - 1] == 2
        p[name][has] = rng.exponential(KB * T0, size=int(has.sum()))
    p["DTRIM"] = dt
    p["IC"] = cells0 + 1
    wr = None
    if wall_reactions:
        # authored in the reference's wall-reaction format: species 1 -> species 2 with p=0.3, species 2 -> none 0.2
        wr = dict(r_sp=[1, 2], p1_sp=[2, -1], nprod=[1, 0], prob=[0.3, 0.2])
    return Workload("C3 2-D tri box" + (" (axisymmetric)" if axi else ""), cfg, sp, p, mesh=mesh, bcs=bcs,
                    wall_reactions=wr)


# ------------------------------------------------------------------------------------------------
# C4-like: 1-D argon plasma slab, explicit electrostatic PIC + null-collision MCC with LXCat-format tables
# ------------------------------------------------------------------------------------------------
def authored_lxcat_reactions(sp: df.SpeciesTable) -> List[dict]:
    """Cross-section tables authored in the LXCat layout the reference reads (energy[eV] sigma[m^2],
    src/initialization.f90:1753-1785).  Shapes are smooth stand-ins (the reference ships no table):
    e-Ar elastic, excitation (11.5 eV), ionisation (15.8 eV); Ar+-Ar isotropic and charge exchange."""
    QE = df.QE
    e_grid = np.concatenate([[0.0], np.logspace(-2, 3, 60)])
    el = 1e-20 * (1.0 + 15.0 * e_grid / (e_grid + 10.0)) / (1.0 + (e_grid / 60.0) ** 1.2)
    exc = np.where(e_grid > 11.5, 4e-21 * np.log(np.maximum(e_grid, 11.5) / 11.5) / (1 + e_grid / 200.0), 0.0)
    ion = np.where(e_grid > 15.8, 2.5e-20 * np.log(np.maximum(e_grid, 15.8) / 15.8) / (1 + e_grid / 150.0), 0.0)
    iso = 4e-19 / (1.0 + e_grid) ** 0.3
    cex = 6e-19 / (1.0 + e_grid) ** 0.2
    e, ar, arp = sp.id("e"), sp.id("Ar"), sp.id("Ar+")
    mk = lambda r1, r2, prods, ea, cs, cex_=0: dict(
        type=capi.LXCAT, r1=r1, r2=r2, p1=prods[0], p2=prods[1], p3=prods[2] if len(prods) > 2 else -1, p4=-1,
        nprod=len(prods), is_cex=cex_, ea=ea * QE, table_energy=e_grid * QE, table_cs=cs)
    return [
        mk(e, ar, (e, ar), 0.0, el),
        mk(e, ar, (e, ar), 11.5, exc),
        mk(e, ar, (e, arp, e), 15.8, ion),
        mk(arp, ar, (arp, ar), 0.0, iso),
        mk(arp, ar, (ar, arp), 0.0, cex, 1),
    ]


def c4_ccp_1d(ncells: int = 64, n_per_species: int = 20_000, seed: int = 5, length: float = 0.0254, Te: float = 30000.0,
              Ti: float = 300.0, nrho: float = 1e15, dt: float = 1e-11, bg_dens: float = 1e21, bg_T: float = 300.0,
              rng_seed_global: int = 65644, wall_bc=("vacuum", "vacuum"), field_bc=(capi.DIRICHLET_BC, capi.DIRICHLET_BC)
              ) -> Workload:
    sp = load_species("ar", ["Ar", "Ar+", "e"])
    mesh = mg.line_1d(ncells, 0.0, length)
    n = 2 * n_per_species
    fnum = nrho * length * 1.0 * 1.0 / n_per_species
    cfg = capi.make_config(dims=1, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=length, ymin=0.0, ymax=1.0,
                           zmin=0.0, zmax=1.0, dt=dt, fnum=fnum, pic_type=capi.PIC_EXPLICIT,
                           collision_type=capi.MCC_VAHEDI, mcc_bg_dens=bg_dens, mcc_bg_ttra=bg_T,
                           bool_z_periodic=1, bool_y_periodic=1, rng_seed_global=rng_seed_global)
    kinds = {"vacuum": capi.VACUUM, "specular": capi.SPECULAR, "diffuse": capi.DIFFUSE}
    bcs = dict(particle_bc=[kinds[k] for k in wall_bc], field_bc=list(field_bc), wall_temp=[300.0, 300.0])
    rng = np.random.default_rng(seed)
    p = new_particles(n)
    p["X"] = rng.random(n) * length
    p["Y"] = rng.random(n)
    p["Z"] = rng.random(n)
    sid = np.where(np.arange(n) % 2 == 0, sp.id("e"), sp.id("Ar+")).astype(np.int32)
    T = np.where(sid == sp.id("e"), Te, Ti)
    sig = np.sqrt(KB * T / sp.mass[sid - 1])
    v = rng.standard_normal((n, 3)) * sig[:, None]
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["S_ID"] = sid
    p["DTRIM"] = dt
    h = length / ncells
    p["IC"] = np.minimum((p["X"] / h).astype(np.int32), ncells - 1) + 1
    efield = np.zeros((3, ncells), order="F")
    efield[0, :] = 2000.0 * np.sin(2 * np.pi * (np.arange(ncells) + 0.5) / ncells)
    reactions = authored_lxcat_reactions(sp)
    sp.sigmamax = max(sp.sigmamax, max(float(np.max(r["table_cs"])) for r in reactions))
    return Workload("C4 1-D argon PIC-MCC", cfg, sp, p, mesh=mesh, bcs=bcs, reactions=reactions,
                    mcc_background=([sp.id("Ar")], [1.0]), efield=efield)


# ------------------------------------------------------------------------------------------------
# injection tasks (host side of SURVEY 8(f) row 1): what READINPUT + INITINJECTION leave in EMIT_TASKS
# ------------------------------------------------------------------------------------------------
_FACE_IND = ((1, 3, 2), (1, 2, 4), (2, 3, 4), (1, 4, 3))   # src/grid_and_partition.f90:1932-1935


def erf1(x: float) -> float:
    """ERF1, src/tools.f90:1126-1143 (Abramowitz-Stegun 7.1.26 with single-precision literals)."""
    f = lambda v: float(np.float32(v))
    p, a1, a2, a3, a4, a5 = f(0.3275911), f(0.254829592), f(-0.284496736), f(1.421413741), f(-1.453152027), f(1.061405429)
    t = 1.0 / (1.0 + p * abs(x))
    out = 1 - (a1 * t + a2 * t ** 2 + a3 * t ** 3 + a4 * t ** 4 + a5 * t ** 5) * np.exp(-x ** 2)
    return float(np.copysign(out, x))


def emit_tasks_for_group(mesh: mg.Mesh, cfg, species: df.SpeciesTable, mixture, pg: int, nrho: float, u=(0.0, 0.0, 0.0),
                         ttra: float = 300.0, trot: float = 300.0, tvib: float = 300.0, mix_id: int = 0, nranks: int = 1):
    """One EMIT task per boundary face of physical group `pg`, with NFS per mixture component as INITINJECTION
    computes it (src/initialization.f90:2575-2675): Maxwellian flux through the face x area x DT / FNUM / ranks.
    mixture: list of (species id 1-based, molar fraction)."""
    d = mesh.dims
    fpg = np.asarray(mesh.cell_faces_pg)
    nb = np.asarray(mesh.cell_neighbors)
    cn = np.asarray(mesh.cell_nodes)
    xyz = np.asarray(mesh.node_coords)
    fn = np.asarray(mesh.face_normal)
    tasks = []
    faces, cells = np.nonzero((fpg == pg) & (nb == -1))
    for j, ic0 in zip(faces, cells):
        normal = -fn[:, j, ic0]
        iv1 = iv2 = 1
        if d == 1:
            area = (cfg.ymax - cfg.ymin) * (cfg.zmax - cfg.zmin)
            iv1 = iv2 = int(cn[j, ic0])
        elif d == 2:
            iv1, iv2 = int(cn[j, ic0]), int(cn[(j + 1) % 3, ic0])     # edge j joins local nodes j, j+1
            x1, y1, x2, y2 = xyz[0, iv1 - 1], xyz[1, iv1 - 1], xyz[0, iv2 - 1], xyz[1, iv2 - 1]
            length = np.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
            area = (cfg.zmax - cfg.zmin) * 0.5 * (y1 + y2) * length if cfg.axi else length * (cfg.zmax - cfg.zmin)
        else:
            area = float(np.asarray(mesh.face_area)[j, ic0])
        u_norm = u[0] * normal[0] + u[1] * normal[1] + u[2] * normal[2]
        nfs = []
        for s_id, frac in mixture:
            m = species.mass[s_id - 1]
            if ttra == 0:
                flux = nrho * frac * u_norm
            else:
                beta = 1.0 / np.sqrt(2.0 * KB / m * ttra)
                sn = u_norm * beta
                flux = nrho * frac / (beta * 2.0 * np.sqrt(np.pi)) * (np.exp(-sn ** 2) + np.sqrt(np.pi) * sn * (1.0 + erf1(sn)))
            nfs.append(flux * area * cfg.dt / cfg.fnum / float(nranks))
        tasks.append(dict(mix_id=mix_id, ic=int(ic0) + 1, iv1=iv1, iv2=iv2, iface=int(j) + 1, nfs=nfs, ux=u[0], uy=u[1], uz=u[2],
                          u_norm=float(u_norm), ttra=ttra, trot=trot, tvib=tvib))
    return tasks


# ------------------------------------------------------------------------------------------------
# C2: hypersonic air flow around a cylinder: 2-D triangulated half annulus, free stream emitted from the
# outer boundary every step (BOUNDARIES_EMIT on the device), diffuse cylinder wall, TCE chemistry + Borgnakke-Larsen
# ------------------------------------------------------------------------------------------------
def c2_cylinder(n: int = 10_000_000, ni: int = 250, nj: int = 400, r_in: float = 0.1524, r_out: float = 0.4572,
                u_inf: float = 7000.0, T_inf: float = 200.0, nrho: float = 1e20, T_wall: float = 1000.0, dt: float = 1e-7,
                seed: int = 6, with_reactions: bool = True, rng_seed_global: int = 65644) -> Workload:
    """Polar grid ni (radial) x nj (angle 0..pi) cut into 2 ni nj triangles.  Physical groups: 1 cylinder (diffuse,
    T_wall), 2 outer boundary (vacuum + one EMIT task per face carrying the free stream N2 0.79 / O2 0.21 at u_inf along
    +x), 3 and 4 the symmetry line y = 0 behind and ahead of the cylinder (specular).  The domain starts filled with
    the free stream.  `w.mixtures` / `w.emit_tasks` hold what READINPUT + INITINJECTION would leave in MIXTURES and
    EMIT_TASKS (apply_sources forwards them)."""
    sp = load_species("air", ["O2", "N2", "O", "N", "NO"])
    reactions = df.read_reactions(os.path.join(DATA_DIR, "air.tce"), sp) if with_reactions else []
    r = np.linspace(r_in, r_out, ni + 1)
    th = np.linspace(0.0, np.pi, nj + 1)
    R, TH = np.meshgrid(r, th, indexing="ij")
    X, Y = R * np.cos(TH), R * np.sin(TH)
    Y[:, 0] = 0.0
    Y[:, -1] = 0.0
    mesh = mg.mapped_quads_to_tris(X, Y, zmin=-0.5, zmax=0.5)
    mesh.pg_names = {"cylinder": 1, "outer": 2, "wake_axis": 3, "stagnation_axis": 4}
    vol = float(np.sum(mesh.cell_volumes))
    fnum = nrho * vol / n
    cfg = capi.make_config(dims=2, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=-r_out, xmax=r_out, ymin=0.0, ymax=r_out,
                           zmin=-0.5, zmax=0.5, dt=dt, fnum=fnum, collision_type=capi.DSMC, bool_z_periodic=1,
                           rng_seed_global=rng_seed_global, dump_bound_avg_every=0)
    bcs = dict(particle_bc=[capi.DIFFUSE, capi.VACUUM, capi.SPECULAR, capi.SPECULAR], wall_temp=[T_wall] * 4,
               acc_n=[1.0] * 4, acc_t=[1.0] * 4)
    rng = np.random.default_rng(seed)
    cum = np.cumsum(mesh.cell_volumes)
    cells0 = np.minimum(np.searchsorted(cum, rng.random(n) * cum[-1]), mesh.nc - 1).astype(np.int64)
    xyz = sample_in_tets(rng, mesh, cells0)
    p = new_particles(n)
    p["X"], p["Y"] = xyz[:, 0], xyz[:, 1]
    p["Z"] = rng.random(n) - 0.5
    sid = np.where(rng.random(n) < 0.79, sp.id("N2"), sp.id("O2")).astype(np.int32)
    v = maxwellian(rng, n, sp.mass[sid - 1], T_inf, u=(u_inf, 0.0, 0.0))
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["S_ID"] = sid
    for name, dofs in (("EROT", sp.rotdof), ("EVIB", sp.vibdof)):
        has = dofs[sid - 1] == 2
        p[name][has] = rng.exponential(KB * T_inf, size=int(has.sum()))
    p["DTRIM"] = dt
    p["IC"] = cells0 + 1
    w = Workload("C2 air flow around a cylinder (2-D triangles, TCE + Borgnakke-Larsen, free stream emitted at the outer boundary)",
                 cfg, sp, p, mesh=mesh, bcs=bcs, reactions=reactions)
    w.mixtures = [[(sp.id("N2"), 0.79), (sp.id("O2"), 0.21)]]
    w.emit_tasks = emit_tasks_for_group(mesh, cfg, sp, w.mixtures[0], pg=2, nrho=nrho, u=(u_inf, 0.0, 0.0), ttra=T_inf, trot=T_inf, tvib=T_inf)
    return w


def c3_nozzle(n: int = 20_000_000, ni: int = 500, nj: int = 200, length: float = 0.3, r_in: float = 0.05, r_throat: float = 0.02,
              r_out: float = 0.08, u_in: float = 400.0, T_in: float = 1000.0, nrho: float = 2e20, T_wall: float = 300.0,
              acc: float = 0.8, dt: float = 2e-7, seed: int = 8, with_reactions: bool = True, rng_seed_global: int = 65644) -> Workload:
    """C3 of SURVEY 8(d): axisymmetric converging-diverging nozzle, ni x nj quads cut into 2 ni nj triangles (y is the
    radius; the axis y = 0 is a mesh boundary that an axisymmetric trajectory never reaches unless VZ = 0: specular).
    Physical groups: 1 inlet (vacuum + EMIT tasks: N2 0.79 / O2 0.21 at u_in, T_in), 2 outlet (vacuum), 3 axis (specular),
    4 nozzle wall (CLL at T_wall with accommodation acc/acc).  TCE + Borgnakke-Larsen collisions.  Starts filled with the
    inlet gas at rest."""
    sp = load_species("air", ["O2", "N2", "O", "N", "NO"])
    reactions = df.read_reactions(os.path.join(DATA_DIR, "air.tce"), sp) if with_reactions else []
    xs = np.linspace(0.0, length, ni + 1)
    xt = 0.4 * length                                            # throat position
    conv = r_throat + (r_in - r_throat) * np.cos(0.5 * np.pi * np.clip(xs / xt, 0.0, 1.0)) ** 2
    div = r_throat + (r_out - r_throat) * np.sin(0.5 * np.pi * np.clip((xs - xt) / (length - xt), 0.0, 1.0)) ** 2
    radius = np.where(xs <= xt, conv, div)
    X = np.repeat(xs[:, None], nj + 1, axis=1)
    Y = radius[:, None] * np.linspace(0.0, 1.0, nj + 1)[None, :]
    Y[:, 0] = 0.0
    mesh = mg.mapped_quads_to_tris(X, Y, zmin=0.0, zmax=1.0, axi=True)
    mesh.pg_names = {"inlet": 1, "outlet": 2, "axis": 3, "wall": 4}
    vol = float(np.sum(mesh.cell_volumes))
    fnum = nrho * vol / n
    cfg = capi.make_config(dims=2, axi=1, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=length, ymin=0.0,
                           ymax=float(radius.max()), zmin=0.0, zmax=1.0, dt=dt, fnum=fnum, collision_type=capi.DSMC,
                           bool_z_periodic=0, rng_seed_global=rng_seed_global, dump_bound_avg_every=0)
    bcs = dict(particle_bc=[capi.VACUUM, capi.VACUUM, capi.SPECULAR, capi.CLL], wall_temp=[T_wall] * 4, acc_n=[acc] * 4, acc_t=[acc] * 4)
    rng = np.random.default_rng(seed)
    cum = np.cumsum(mesh.cell_volumes)
    cells0 = np.minimum(np.searchsorted(cum, rng.random(n) * cum[-1]), mesh.nc - 1).astype(np.int64)
    # uniform in the VOLUME of an axisymmetric cell ~ uniform in the triangle weighted by the radius: rejection on y / y_max
    xyz = sample_in_tets(rng, mesh, cells0)
    ymax_cell = np.asarray(mesh.node_coords)[1, np.asarray(mesh.cell_nodes) - 1].max(axis=0)[cells0]
    for _ in range(6):
        bad = rng.random(n) * ymax_cell > xyz[:, 1]
        if not bad.any():
            break
        idx = np.nonzero(bad)[0]
        xyz[idx] = sample_in_tets(rng, mesh, cells0[idx])
    p = new_particles(n)
    p["X"], p["Y"] = xyz[:, 0], np.maximum(xyz[:, 1], 1e-9)
    p["Z"] = 0.0
    sid = np.where(rng.random(n) < 0.79, sp.id("N2"), sp.id("O2")).astype(np.int32)
    v = maxwellian(rng, n, sp.mass[sid - 1], T_in)
    p["VX"], p["VY"], p["VZ"] = v[:, 0], v[:, 1], v[:, 2]
    p["S_ID"] = sid
    for name, dofs in (("EROT", sp.rotdof), ("EVIB", sp.vibdof)):
        has = dofs[sid - 1] == 2
        p[name][has] = rng.exponential(KB * T_in, size=int(has.sum()))
    p["DTRIM"] = dt
    p["IC"] = cells0 + 1
    w = Workload("C3 axisymmetric converging-diverging nozzle (triangles, CLL wall, inflow emitted at the inlet, TCE + Borgnakke-Larsen)",
                 cfg, sp, p, mesh=mesh, bcs=bcs, reactions=reactions)
    w.mixtures = [[(sp.id("N2"), 0.79), (sp.id("O2"), 0.21)]]
    w.emit_tasks = emit_tasks_for_group(mesh, cfg, sp, w.mixtures[0], pg=1, nrho=nrho, u=(u_in, 0.0, 0.0), ttra=T_in, trot=T_in, tvib=T_in)
    return w


def open_channel(w: Workload, nrho: float = 1e20, u: float = 300.0, T0: float = 300.0) -> Workload:
    """The open variant of a C5 channel workload: vacuum end caps in x, Boundary_emit at x = 0."""
    w.bcs["particle_bc"][0] = capi.VACUUM
    w.bcs["particle_bc"][1] = capi.VACUUM
    w.mixtures = [[(1, 1.0)]]
    w.emit_tasks = emit_tasks_for_group(w.mesh, w.cfg, w.species, w.mixtures[0], pg=1, nrho=nrho, u=(u, 0.0, 0.0), ttra=T0, trot=T0, tvib=T0)
    w.name += " (open: vacuum end caps, inflow emitted at x = 0)"
    return w


def apply_sources(lib: capi.Lib, w: Workload):
    """MIXTURES and EMIT_TASKS of a workload with boundary injection (after `apply`)."""
    lib.set_mixtures(w.mixtures)
    lib.set_emit_tasks(w.emit_tasks)


# ------------------------------------------------------------------------------------------------
# bench-scale C5: one x-strip per rank, particles generated in chunks straight into (pinned) AoS records
# ------------------------------------------------------------------------------------------------
_PINNED_KEEPALIVE = []


def pinned_particles(n: int) -> np.ndarray:
    """n zeroed particle records in page-locked host memory (falls back to pageable without CUDA)."""
    import torch
    nbytes = max(n, 1) * capi.PARTICLE_DTYPE.itemsize
    try:
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    except RuntimeError:
        t = torch.empty(nbytes, dtype=torch.uint8)
    _PINNED_KEEPALIVE.append(t)
    return t.numpy().view(capi.PARTICLE_DTYPE)[:n]


def pinned_array(n: int, dtype) -> np.ndarray:
    """n zeroed records of `dtype` in page-locked host memory."""
    import torch
    nbytes = max(n, 1) * np.dtype(dtype).itemsize
    try:
        t = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    except RuntimeError:
        t = torch.zeros(nbytes, dtype=torch.uint8)
    _PINNED_KEEPALIVE.append(t)
    return t.numpy().view(dtype)[:n]


def c5_channel_strip(ncx: int, ncy: int, ncz: int, n: int, rank: int = 0, world: int = 1, seed: int = 2,
                     T0: float = 300.0, nrho: float = 1e20, cell_mfp: float = 0.5, dt_frac: float = 0.2,
                     u_drift: float = 100.0, wall_temp: float = 400.0, pinned: bool = False,
                     rng_seed_global: int = 65644, mesh: Optional[mg.Mesh] = None, lazy: bool = False) -> Workload:
    """C5 at bench scale.  The whole channel is world*ncx x ncy x ncz cubes (mesh replicated on every rank,
    as the reference does, grid_and_partition.f90:123-353); rank r owns the x-strip r and seeds n particles
    uniformly in it (all Kuhn tets of the regular box have the same volume)."""
    sp = load_species("ar", ["Ar"])
    m = sp.mass[0]
    d, om, tref = sp.diam[0], sp.omega[0], sp.tref[0]
    nu = 4.0 * d * d * nrho * np.sqrt(np.pi * KB * tref / m) * (T0 / tref) ** (1.0 - om)
    cbar = np.sqrt(8.0 * KB * T0 / (np.pi * m))
    h = cell_mfp * cbar / nu
    nxt = ncx * world
    hi = (nxt * h, ncy * h, ncz * h)
    if mesh is None:
        mesh = mg.box_tets(nxt, ncy, ncz, lo=(0.0, 0.0, 0.0), hi=hi)
        if world > 1:
            mesh.cell_procs = mg.assign_cells_to_procs(mesh, world, 0)
    vol_strip = ncx * ncy * ncz * h ** 3
    fnum = nrho * vol_strip / n
    dt = dt_frac / nu
    cfg = capi.make_config(dims=3, grid_type=capi.UNSTRUCTURED, nx=1, ny=1, xmin=0.0, xmax=hi[0], ymin=0.0, ymax=hi[1],
                           zmin=0.0, zmax=hi[2], dt=dt, fnum=fnum, collision_type=capi.DSMC, rank=rank, nranks=world,
                           rng_seed_global=rng_seed_global, dump_bound_avg_every=0)
    bcs = dict(particle_bc=[capi.SPECULAR, capi.SPECULAR, capi.DIFFUSE, capi.DIFFUSE, capi.SPECULAR, capi.SPECULAR],
               wall_temp=[wall_temp] * 6, acc_n=[1.0] * 6, acc_t=[1.0] * 6)
    cn = np.asarray(mesh.cell_nodes)
    xyz = np.asarray(mesh.node_coords)
    sig = np.sqrt(KB * T0 / m)
    chunk = 2_000_000

    def fill_into(q, a, b):
        """particles [a, b) of the strip (a is a multiple of the chunk size) into the record array q of length b - a"""
        # one generator per chunk (seeded by (seed, chunk index)): chunks are independent, so they are filled by a
        # few host threads -- NumPy releases the GIL in the generators, the gathers and the element-wise kernels
        k = b - a
        rng = np.random.default_rng([seed, a // chunk])
        ci = rng.integers(rank * ncx, (rank + 1) * ncx, k)
        cj = rng.integers(0, ncy, k)
        ck = rng.integers(0, ncz, k)
        cell0 = 6 * (ci + nxt * (cj + ncy * ck)) + rng.integers(0, 6, k)
        lam = rng.exponential(size=(4, k))
        lam /= lam.sum(axis=0, keepdims=True)
        nodes = cn[:, cell0] - 1
        for ax, name in enumerate(("X", "Y", "Z")):
            q[name] = (xyz[ax, nodes] * lam).sum(axis=0)
        q["VX"] = rng.standard_normal(k) * sig + u_drift
        q["VY"] = rng.standard_normal(k) * sig
        q["VZ"] = rng.standard_normal(k) * sig
        q["EROT"] = 0.0
        q["EVIB"] = 0.0
        q["DTRIM"] = dt
        q["IC"] = cell0 + 1
        q["S_ID"] = 1
        q["ID"] = (np.arange(a, b, dtype=np.int64) << 8) + rank
        q["DUMP_TRAJ"] = 0

    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))

    def fill_range(p, a0, b0):
        """fills p[0 : b0 - a0] with the particles [a0, b0) (a0 a multiple of the chunk size), a few host threads"""
        starts = list(range(a0, b0, chunk))
        nthreads = max(1, min(8, cores // local_world, len(starts)))
        job = lambda a: fill_into(p[a - a0:min(b0, a + chunk) - a0], a, min(b0, a + chunk))
        if nthreads > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(nthreads) as ex:
                list(ex.map(job, starts))
        else:
            for a in starts:
                job(a)

    if lazy:
        # 10^9 particles do not fit a host array: the caller streams them, `batch` records at a time
        p = np.zeros(0, dtype=capi.PARTICLE_DTYPE)
        w = Workload("C5 3-D tet channel strip", cfg, sp, p, mesh=mesh, bcs=bcs,
                     notes=f"h={h:.4g} nu={nu:.5g} dt={dt:.4g} nc={mesh.nc} world={world}")

        def batches(batch=24_000_000):
            batch = (batch // chunk) * chunk
            buf = pinned_particles(batch)
            for a0 in range(0, n, batch):
                b0 = min(n, a0 + batch)
                fill_range(buf, a0, b0)
                yield buf[: b0 - a0]
        w.batches = batches
        return w
    p = pinned_particles(n) if pinned else np.zeros(n, dtype=capi.PARTICLE_DTYPE)
    fill_range(p, 0, n)
    return Workload("C5 3-D tet channel strip", cfg, sp, p, mesh=mesh, bcs=bcs,
                    notes=f"h={h:.4g} nu={nu:.5g} dt={dt:.4g} nc={mesh.nc} world={world}")
