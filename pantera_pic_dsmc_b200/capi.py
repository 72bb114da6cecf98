"""ctypes binding of the C ABI declared in ``include/pantera_gpu.h``.

``libpantera_gpu.so`` (prefix ``pgpu_``) is the product: hand-written CUDA for sm_100a.  The class :class:`Lib` is
generic in the symbol prefix so that the test infrastructure can drive the CPU oracle (``oracle/binding.py``, prefix
``po_``, the same entry points) with the same methods; nothing in this package loads, imports or names the oracle.

There is no fallback: :func:`load_gpu` raises if the CUDA library is missing or no GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
# PANTERA_GPU_LIB: developer override to A/B-test differently tuned builds of the SAME library (never a fallback)
GPU_LIB_PATH = os.environ.get("PANTERA_GPU_LIB") or os.path.join(_HERE, "csrc", "libpantera_gpu.so")

# ---- enumerations (values of the reference's ENUM, BIND(C) blocks; see pantera_gpu.h) ----------
RECTILINEAR_UNIFORM, RECTILINEAR_NONUNIFORM, QUADTREE, UNSTRUCTURED = 0, 1, 2, 3
VACUUM, SPECULAR, DIFFUSE, CLL, REACT, AXIS, PERIODIC_MASTER, PERIODIC_SLAVE, EMIT, WB_BC = range(10)
DIRICHLET_BC, NEUMANN_BC, DIELECTRIC_BC, ROBIN_BC = 0, 1, 2, 3
RF_VOLTAGE_BC, DECOUPLED_RF_VOLTAGE_BC = 6, 7
NO_BC = 8
FLUID, SOLID = 0, 1
PIC_NONE, PIC_EXPLICIT = 0, 1
NO_COLL, DSMC, BGK, MCC, MCC_VAHEDI, DSMC_VAHEDI = range(6)
FIXED_RATE, TCE, LXCAT = 0, 1, 2
RESTART_UNASSIGNED = 2  # PGPU_RESTART_UNASSIGNED

# host particle record: TYPE PARTICLE_DATA_STRUCTURE, src/particle.f90:28-37 (96-byte stride)
PARTICLE_DTYPE = np.dtype(
    {
        "names": ["X", "Y", "Z", "VX", "VY", "VZ", "EROT", "EVIB", "DTRIM", "IC", "S_ID", "ID", "DUMP_TRAJ"],
        "formats": ["<f8"] * 9 + ["<i4", "<i4", "<i8", "<i4"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 76, 80, 88],
        "itemsize": 96,
    }
)


POS_DTYPE = np.dtype({"names": ["X", "Y", "Z", "IC", "S_ID"], "formats": ["<f8", "<f8", "<f8", "<i4", "<i4"], "offsets": [0, 8, 16, 24, 28], "itemsize": 32})
VEL_DTYPE = np.dtype({"names": ["VX", "VY", "VZ", "S_ID", "pad_"], "formats": ["<f8", "<f8", "<f8", "<i4", "<i4"], "offsets": [0, 8, 16, 24, 28], "itemsize": 32})
IE_DTYPE = np.dtype({"names": ["EROT", "EVIB"], "formats": ["<f8", "<f8"], "offsets": [0, 8], "itemsize": 16})


def to_planes(parts: np.ndarray, pos=None, vel=None):
    """AoS records -> (pos, vel, ie) record planes (host-side helper of the harness)."""
    n = len(parts)
    pos = np.zeros(n, dtype=POS_DTYPE) if pos is None else pos[:n]
    vel = np.zeros(n, dtype=VEL_DTYPE) if vel is None else vel[:n]
    ie = np.zeros(n, dtype=IE_DTYPE)
    for f in ("X", "Y", "Z", "IC", "S_ID"):
        pos[f] = parts[f]
    for f in ("VX", "VY", "VZ", "S_ID"):
        vel[f] = parts[f]
    ie["EROT"], ie["EVIB"] = parts["EROT"], parts["EVIB"]
    return pos, vel, ie


class Config(C.Structure):
    """``pgpu_config`` -- scalar globals read by the hot path."""

    _fields_ = [
        ("dims", C.c_int32),
        ("axi", C.c_int32),
        ("grid_type", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("xmin", C.c_double),
        ("xmax", C.c_double),
        ("ymin", C.c_double),
        ("ymax", C.c_double),
        ("zmin", C.c_double),
        ("zmax", C.c_double),
        ("dt", C.c_double),
        ("fnum", C.c_double),
        ("pic_type", C.c_int32),
        ("collision_type", C.c_int32),
        ("bool_periodic", C.c_int32 * 4),
        ("bool_specular", C.c_int32 * 4),
        ("bool_diffuse", C.c_int32 * 4),
        ("bool_react", C.c_int32 * 4),
        ("bool_z_periodic", C.c_int32),
        ("bool_y_periodic", C.c_int32),
        ("boundtemp", C.c_double),
        ("external_b", C.c_double * 3),
        ("eps_scaling", C.c_double),
        ("mcc_bg_dens", C.c_double),
        ("mcc_bg_ttra", C.c_double),
        ("rng_seed_global", C.c_int64),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
        ("dump_bound_start", C.c_int32),
        ("dump_bound_avg_every", C.c_int32),
        ("capacity", C.c_int64),
        ("track_id", C.c_int32),
        ("reserved_", C.c_int32),
    ]


def make_config(**kw) -> Config:
    cfg = Config()
    cfg.dims = 2
    cfg.grid_type = RECTILINEAR_UNIFORM
    cfg.nx = cfg.ny = 1
    cfg.xmax = cfg.ymax = 1.0
    cfg.zmin, cfg.zmax = -0.5, 0.5
    cfg.dt = 1e-6
    cfg.fnum = 1.0
    cfg.boundtemp = 300.0
    cfg.eps_scaling = 1.0
    cfg.rng_seed_global = 65644  # the example input's seed (reference input:58)
    cfg.nranks = 1
    for k, v in kw.items():
        if k in ("bool_periodic", "bool_specular", "bool_diffuse", "bool_react"):
            arr = getattr(cfg, k)
            for i, b in enumerate(v):
                arr[i] = int(b)
        elif k == "external_b":
            for i, b in enumerate(v):
                cfg.external_b[i] = float(b)
        else:
            if not hasattr(cfg, k):
                raise KeyError(f"unknown config field {k}")
            setattr(cfg, k, v)
    return cfg


class PanteraError(RuntimeError):
    pass


def _f64(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="K"))


def _i32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel(order="K"))


def _fortran_flat(a, dtype) -> Optional[np.ndarray]:
    """Flatten an array whose *numpy* shape is the Fortran shape, in Fortran (column-major) order."""
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=dtype).ravel(order="F"))


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class Lib:
    """One loaded library + one context; methods mirror the C entry points 1:1."""

    def __init__(self, path: str, prefix: str, device: int = 0):
        if not os.path.exists(path):
            raise PanteraError(f"shared library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "pgpu_" else C.RTLD_LOCAL)
        self._keep = []
        self.ctx = C.c_void_p()
        f = self._fn("create")
        f.restype = C.c_int
        rc = f(C.byref(self.ctx), C.c_int(device))
        if rc != 0:
            raise PanteraError(f"{prefix}create failed with code {rc} (is a CUDA device visible?)")
        self._fn("last_error").restype = C.c_char_p

    # -- plumbing ---------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _call(self, name, *args):
        f = self._fn(name)
        f.restype = C.c_int
        rc = f(self.ctx, *args)
        if rc != 0:
            msg = self._fn("last_error")(self.ctx)
            raise PanteraError(f"{self.prefix}{name} -> {rc}: {msg.decode() if msg else ''}")
        return rc

    def close(self):
        if self.ctx:
            self._fn("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters -------------------------------------------------------------------------
    def set_config(self, cfg: Config):
        self.cfg = cfg
        self._call("set_config", C.byref(cfg))

    def set_species(self, mass, charge, rotdof, vibdof, rotrel, vibrel):
        mass = _f64(mass)
        self.ns = len(mass)
        self._call("set_species", C.c_int(self.ns), _p(mass), _p(_f64(charge)), _p(_i32(rotdof)),
                   _p(_i32(vibdof)), _p(_f64(rotrel)), _p(_f64(vibrel)))

    def set_vss(self, sigmas, omegas, alphas, grefs, sigmamax):
        ns = self.ns
        self._call("set_vss", C.c_int(ns), _p(_fortran_flat(sigmas, np.float64)), _p(_fortran_flat(omegas, np.float64)),
                   _p(_fortran_flat(alphas, np.float64)), _p(_fortran_flat(grefs, np.float64)), C.c_double(sigmamax))

    def set_reactions(self, reactions: Sequence[dict]):
        nr = len(reactions)
        if nr == 0:
            self._call("set_reactions", C.c_int(0), *([None] * 16))
            return
        g = lambda k, d=0: [r.get(k, d) for r in reactions]
        tab_off = [0]
        tab_e, tab_cs = [], []
        for r in reactions:
            te = list(r.get("table_energy", []))
            tab_e += te
            tab_cs += list(r.get("table_cs", []))
            tab_off.append(len(tab_e))
        arrs = [_i32(g("type")), _i32(g("r1")), _i32(g("r2")), _i32(g("p1")), _i32(g("p2")), _i32(g("p3", -1)),
                _i32(g("p4", -1)), _i32(g("nprod")), _i32(g("is_cex", 0)), _f64(g("ea", 0.0)), _f64(g("c1", 0.0)),
                _f64(g("c2", 0.0)), _f64(g("c3", 0.0)), _i32(tab_off), _f64(tab_e if tab_e else [0.0]),
                _f64(tab_cs if tab_cs else [0.0])]
        self._call("set_reactions", C.c_int(nr), *[_p(a) for a in arrs])

    def set_wall_reactions(self, r_sp, p1_sp, nprod, prob):
        n = len(r_sp)
        self._call("set_wall_reactions", C.c_int(n), _p(_i32(r_sp)), _p(_i32(p1_sp)), _p(_i32(nprod)), _p(_f64(prob)))

    def set_mcc_background(self, sp_id, molfrac):
        self._call("set_mcc_background", C.c_int(len(sp_id)), _p(_i32(sp_id)), _p(_f64(molfrac)))

    def set_mcc_cell_density(self, nrho):
        """``nrho``: MCC_BG_CELL_NRHO as an (ns, nc) array (Fortran order), or None to clear it."""
        if nrho is None:
            self._call("set_mcc_cell_density", C.c_int(0), C.c_int64(0), None)
            return
        a = np.asfortranarray(nrho, dtype=np.float64)
        self._call("set_mcc_cell_density", C.c_int(a.shape[0]), C.c_int64(a.shape[1]), _p(a))

    def set_mesh(self, mesh):
        """``mesh``: an object/dict with the Fortran-shaped arrays produced by ``meshgen``."""
        m = mesh if isinstance(mesh, dict) else mesh.__dict__
        d = int(m["dims"])
        nn, nc = int(m["nn"]), int(m["nc"])
        self.mesh = m
        ff = _fortran_flat
        bidx = m.get("face_boundary_index")
        self._call(
            "set_mesh", C.c_int(d), C.c_int(nn), C.c_int(nc),
            _p(ff(m["node_coords"], np.float64)), _p(ff(m["cell_nodes"], np.int32)),
            _p(ff(m["cell_neighbors"], np.int32)), _p(ff(m["face_normal"], np.float64)),
            _p(ff(m.get("face_tang1"), np.float64)), _p(ff(m.get("face_tang2"), np.float64)),
            _p(ff(m.get("cell_edges_len"), np.float64)), _p(ff(m["cell_faces_pg"], np.int32)),
            _p(ff(m.get("cell_pg"), np.int32)), _p(ff(m.get("basis_coeffs"), np.float64)),
            _p(ff(m["cell_volumes"], np.float64)), _p(ff(m.get("cell_procs"), np.int32)),
            _p(ff(bidx, np.int32)), C.c_int(int(m.get("nboundcells", 0)) if bidx is not None else 0),
            _p(ff(m.get("boundary_area"), np.float64)),
        )

    def set_cell_procs(self, cell_procs):
        a = _i32(cell_procs)
        self._call("set_cell_procs", C.c_int(len(a)), _p(a))

    def set_cell_fnum(self, cell_fnum):
        if cell_fnum is None:
            self._call("set_cell_fnum", C.c_int(0), None)
        else:
            a = _f64(cell_fnum)
            self._call("set_cell_fnum", C.c_int(len(a)), _p(a))

    def set_bcs(self, particle_bc, field_bc=None, volume_bc=None, wall_temp=None, acc_n=None, acc_t=None,
                react=None, dump_fluxes=None):
        n = len(particle_bc)
        z = lambda a, fill: a if a is not None else [fill] * n
        self._call("set_bcs", C.c_int(n), _p(_i32(particle_bc)), _p(_i32(z(field_bc, NO_BC))),
                   _p(_i32(z(volume_bc, FLUID))), _p(_f64(z(wall_temp, 300.0))), _p(_f64(z(acc_n, 1.0))),
                   _p(_f64(z(acc_t, 1.0))), _p(_i32(z(react, 0))), _p(_i32(z(dump_fluxes, 0))))

    def set_walls(self, x1, y1, x2, y2, normx, normy, temp, specular, diffuse, porous, transmissivity, react):
        n = len(x1)
        self._call("set_walls", C.c_int(n), _p(_f64(x1)), _p(_f64(y1)), _p(_f64(x2)), _p(_f64(y2)), _p(_f64(normx)),
                   _p(_f64(normy)), _p(_f64(temp)), _p(_i32(specular)), _p(_i32(diffuse)), _p(_i32(porous)),
                   _p(_f64(transmissivity)), _p(_i32(react)))

    def set_grid_coords(self, xcoord, ycoord):
        x, y = _f64(xcoord), _f64(ycoord)
        self._call("set_grid_coords", C.c_int(len(x)), _p(x), C.c_int(len(y)), _p(y))

    # -- particles --------------------------------------------------------------------------
    def upload_particles(self, parts: np.ndarray):
        assert parts.dtype == PARTICLE_DTYPE
        if not parts.flags["C_CONTIGUOUS"]:
            parts = np.ascontiguousarray(parts)
        self._call("upload_particles", _p(parts), C.c_int64(len(parts)), C.c_int32(96))

    def append_particles(self, parts: np.ndarray):
        assert parts.dtype == PARTICLE_DTYPE
        parts = np.ascontiguousarray(parts)
        self._call("append_particles", _p(parts), C.c_int64(len(parts)), C.c_int32(96))

    def num_particles(self) -> int:
        n = C.c_int64()
        self._call("num_particles", C.byref(n))
        return n.value

    def download_particles(self) -> np.ndarray:
        n = self.num_particles()
        out = np.zeros(max(n, 1), dtype=PARTICLE_DTYPE)
        nn = C.c_int64()
        self._call("download_particles", _p(out), C.c_int64(len(out)), C.byref(nn), C.c_int32(96))
        return out[: nn.value]

    def download_into(self, out: np.ndarray) -> int:
        """pgpu_download_particles into a caller-owned (e.g. pinned) record array; returns the count."""
        assert out.dtype == PARTICLE_DTYPE and out.flags["C_CONTIGUOUS"]
        nn = C.c_int64()
        self._call("download_particles", _p(out), C.c_int64(len(out)), C.byref(nn), C.c_int32(96))
        return nn.value

    # -- record planes (device layout on the host) ------------------------------------------
    def upload_planes(self, pos: np.ndarray, vel: np.ndarray, ie: Optional[np.ndarray] = None):
        assert pos.dtype == POS_DTYPE and vel.dtype == VEL_DTYPE and len(pos) == len(vel)
        self._call("upload_planes", _p(pos), _p(vel), _p(ie) if ie is not None else None, C.c_int64(len(pos)))

    def download_planes_into(self, pos: np.ndarray, vel: np.ndarray, ie: Optional[np.ndarray] = None) -> int:
        nn = C.c_int64()
        self._call("download_planes", _p(pos), _p(vel), _p(ie) if ie is not None else None, C.c_int64(len(pos)), C.byref(nn))
        return nn.value

    # -- fields -----------------------------------------------------------------------------
    def set_efield(self, e):
        a = _fortran_flat(e, np.float64)
        self._call("set_efield", _p(a), C.c_int64(len(a)))

    def set_bfield(self, b):
        a = _fortran_flat(b, np.float64)
        self._call("set_bfield", _p(a), C.c_int64(len(a)))

    def get_surface_charge(self, n):
        out = np.zeros(n)
        self._call("get_surface_charge", _p(out), C.c_int64(n))
        return out

    def set_surface_charge(self, sc):
        a = _f64(sc)
        self._call("set_surface_charge", _p(a), C.c_int64(len(a)))

    # -- hot path ---------------------------------------------------------------------------
    def deposit_charge(self, nnodes) -> np.ndarray:
        rhs = np.zeros(nnodes)
        self._call("deposit_charge", _p(rhs), C.c_int64(nnodes))
        return rhs

    def advect(self, tID: int):
        self._call("advect", C.c_int32(tID))

    def exchange(self):
        self._call("exchange")

    def exchange_pack(self, nranks):
        """EXCHANGE, local half: (records grouped by destination rank, count per destination)."""
        buf = np.zeros(max(self.num_particles(), 1), dtype=PARTICLE_DTYPE)
        cnt = np.zeros(nranks, dtype=np.int64)
        self._call("exchange_pack", _p(buf), C.c_int64(len(buf)), _p(cnt))
        return buf[: int(cnt.sum())], cnt

    def set_deterministic(self, on: bool = True):
        self._call("set_deterministic", C.c_int32(int(on)))

    def sort(self):
        self._call("sort")

    def collide(self, tID: int):
        a, b = C.c_int64(), C.c_int64()
        self._call("collide", C.c_int32(tID), C.byref(a), C.byref(b))
        return a.value, b.value

    def mcc_vahedi(self, tID: int):
        a, b = C.c_int64(), C.c_int64()
        self._call("mcc_vahedi", C.c_int32(tID), C.byref(a), C.byref(b))
        return a.value, b.value

    def mcc_collisions(self, tID: int):
        a, b = C.c_int64(), C.c_int64()
        self._call("mcc_collisions", C.c_int32(tID), C.byref(a), C.byref(b))
        return a.value, b.value

    def remove_species(self, sp):
        self._call("remove_species", C.c_int32(sp))

    # -- particle creation -----------------------------------------------------------------
    def set_species_weights(self, spwt):
        a = _f64(spwt)
        self._call("set_species_weights", C.c_int(len(a)), _p(a))

    def set_mixtures(self, mixtures):
        """mixtures: list of lists of (species id 1-based, molar fraction)."""
        off, sp, fr = [0], [], []
        for m in mixtures:
            for s_id, f in m:
                sp.append(s_id)
                fr.append(f)
            off.append(len(sp))
        self._call("set_mixtures", C.c_int(len(mixtures)), _p(_i32(off)), _p(_i32(sp if sp else [0])), _p(_f64(fr if fr else [0.0])))

    def _source_tasks(self, name, tasks):
        g = lambda k, d=0.0: _f64([t.get(k, d) for t in tasks] or [0.0])
        self._call(name, C.c_int(len(tasks)), _p(_i32([t["mix_id"] for t in tasks] or [0])), _p(g("rate")), _p(g("ux")), _p(g("uy")),
                   _p(g("uz")), _p(g("ttrax")), _p(g("ttray")), _p(g("ttraz")), _p(g("trot")), _p(g("tvib")))

    def set_initial_particles_tasks(self, tasks):
        """tasks: dicts mix_id, rate (= NRHO), ux, uy, uz, ttrax, ttray, ttraz, trot, tvib."""
        self._source_tasks("set_initial_particles_tasks", tasks)

    def set_volume_inject_tasks(self, tasks):
        """rate = NRHODOT"""
        self._source_tasks("set_volume_inject_tasks", tasks)

    def set_emit_tasks(self, tasks):
        """tasks: dicts mix_id, ic, iv1, iv2, iface, nfs (list per component), ux, uy, uz, u_norm, ttra, trot, tvib."""
        off, nfs = [0], []
        for t in tasks:
            nfs += list(t["nfs"])
            off.append(len(nfs))
        gi = lambda k, d=1: _i32([t.get(k, d) for t in tasks] or [0])
        gf = lambda k: _f64([t.get(k, 0.0) for t in tasks] or [0.0])
        self._call("set_emit_tasks", C.c_int(len(tasks)), _p(gi("mix_id", 0)), _p(gi("ic")), _p(gi("iv1")), _p(gi("iv2")), _p(gi("iface")),
                   _p(_i32(off)), _p(_f64(nfs or [0.0])), _p(gf("ux")), _p(gf("uy")), _p(gf("uz")), _p(gf("u_norm")), _p(gf("ttra")),
                   _p(gf("trot")), _p(gf("tvib")))

    def initial_seed(self):
        self._call("initial_seed")

    def volume_inject(self, tID: int):
        self._call("volume_inject", C.c_int32(tID))

    def boundaries_emit(self, tID: int):
        self._call("boundaries_emit", C.c_int32(tID))

    # -- field step ------------------------------------------------------------------------
    def set_field_bcs(self, wall_potential=None, wall_rf_potential=None, rf_frequency=None, wall_efield=None, eps_rel=None):
        arrs = [wall_potential, wall_rf_potential, rf_frequency, wall_efield, eps_rel]
        n = max(len(a) for a in arrs if a is not None)
        self._call("set_field_bcs", C.c_int(n), *[_p(_f64(a)) if a is not None else None for a in arrs])

    def compute_e_field(self, phi, ncells):
        phi = _f64(phi)
        e = np.zeros(3 * ncells)
        self._call("compute_e_field", _p(phi), C.c_int64(len(phi)), _p(e))
        return e.reshape(ncells, 3).T

    def field_solve_1d(self, tID, nnodes, ncells, rhs=None):
        phi, e = np.zeros(nnodes), np.zeros(3 * ncells)
        r = _f64(rhs) if rhs is not None else None
        self._call("field_solve_1d", C.c_int32(tID), _p(r), _p(phi), _p(e))
        return phi, e.reshape(ncells, 3).T

    def field_solve(self, tID, nnodes, ncells, rhs=None, rtol=0.0, max_iter=0):
        """Any unstructured mesh; returns (phi, E(3, ncells), iterations, relative residual)."""
        phi, e = np.zeros(nnodes), np.zeros(3 * ncells)
        r = _f64(rhs) if rhs is not None else None
        it, rel = C.c_int32(), C.c_double()
        self._call("field_solve", C.c_int32(tID), _p(r), _p(phi), _p(e), C.c_double(rtol), C.c_int32(max_iter), C.byref(it), C.byref(rel))
        return phi, e.reshape(ncells, 3).T, it.value, rel.value

    def grid_avg(self):
        self._call("grid_avg")

    def grid_reset(self):
        self._call("grid_reset")

    def get_grid_avg(self, ncells: int):
        """dict of (ns, ncells) arrays: np, vx, vy, vz, ttrx, ttry, ttrz, ttr, trot, tvib; and 'cumulated'."""
        n = ncells * self.ns
        names = ["np", "vx", "vy", "vz", "ttrx", "ttry", "ttrz", "ttr", "trot", "tvib"]
        arrs = [np.zeros(n) for _ in names]
        cum = C.c_int32()
        self._call("get_grid_avg", C.c_int64(n), *[_p(a) for a in arrs], C.byref(cum))
        out = {k: a.reshape(self.ns, ncells) for k, a in zip(names, arrs)}
        out["cumulated"] = cum.value
        return out

    def checks(self):
        """CHECKS: (tot_num[ns], tot_momentum[ns, 3], tot_ke_part[ns], tot_ie)."""
        num = np.zeros(self.ns, dtype=np.int64)
        mom = np.zeros(3 * self.ns)
        ke = np.zeros(self.ns)
        ie = C.c_double()
        self._call("checks", _p(num), _p(mom), _p(ke), C.byref(ie))
        return num, mom.reshape(self.ns, 3), ke, ie.value

    def dump_particles_file(self, path_prefix: str, timestep: int, binary: bool = True):
        self._call("dump_particles_file", C.c_char_p(path_prefix.encode()), C.c_int32(timestep), C.c_int32(int(binary)))

    def read_particles_file(self, path_prefix: str, timestep: int, binary: bool = True, unassigned: bool = False):
        """unassigned=True: the reference's restart semantics (IC of the dump not trusted; reassign_particles_to_cells next)."""
        self._call("read_particles_file", C.c_char_p(path_prefix.encode()), C.c_int32(timestep),
                   C.c_int32(int(binary) | (RESTART_UNASSIGNED if unassigned else 0)))

    def reassign_particles_to_cells(self) -> int:
        """REASSIGN_PARTICLES_TO_CELLS_1D/2D/3D: returns the number of particles removed (outside the domain / every cell)."""
        n = C.c_int64()
        self._call("reassign_particles_to_cells", C.byref(n))
        return n.value

    def run_steps(self, tID0: int, nsteps: int):
        a, b = C.c_int64(), C.c_int64()
        self._call("run_steps", C.c_int32(tID0), C.c_int32(nsteps), C.byref(a), C.byref(b))
        return a.value, b.value

    # -- diagnostics ------------------------------------------------------------------------
    def get_boundary_counts(self, ns, nw=0):
        b = np.zeros(4 * ns, dtype=np.int64)
        w = np.zeros(max(nw * ns, 1), dtype=np.int64)
        self._call("get_boundary_counts", _p(b), C.c_int64(len(b)), _p(w), C.c_int64(nw * ns))
        return b, w[: nw * ns]

    def get_boundary_tallies(self, nboundcells, ns, reset=False):
        t = np.zeros(10 * nboundcells * ns)
        self._call("get_boundary_tallies", _p(t), C.c_int64(len(t)), C.c_int32(int(reset)))
        return t.reshape(10, ns, nboundcells)

    def get_cell_counts(self, nc):
        out = np.zeros(nc, dtype=np.int32)
        self._call("get_cell_counts", _p(out), C.c_int64(nc))
        return out


class GpuLib(Lib):
    def __init__(self, device: int = 0, path: str = GPU_LIB_PATH):
        super().__init__(path, "pgpu_", device)

    def get_timings(self):
        ms = np.zeros(8)
        n = C.c_int64()
        self._call("get_timings", _p(ms), C.c_int32(8), C.byref(n))
        return ms, n.value

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        f = self.lib.pgpu_comm_unique_id
        f.restype = C.c_int
        rc = f(buf)
        if rc != 0:
            raise PanteraError(f"pgpu_comm_unique_id -> {rc}")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._call("comm_init", C.c_int(nranks), C.c_int(rank), buf)


def load_gpu(device: int = 0) -> GpuLib:
    """Load the CUDA library. Raises (never falls back) if it is missing or no GPU is usable."""
    return GpuLib(device)
