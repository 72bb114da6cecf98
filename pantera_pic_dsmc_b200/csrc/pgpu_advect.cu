/*
 * pgpu_advect.cu -- ADVECT on the device: Boris kick + free flight + boundary interaction +
 * removal + structured re-index + per-cell histogram for the counting sort.
 *
 * Replaces: ADVECT (src/timecycle.f90:1021-2268), MOVE_PARTICLE (:2297-2332),
 * UPDATE_VELOCITY_BORIS (:2271-2291), APPLY_E_FIELD / APPLY_B_FIELD (src/fields.f90:4194-4249),
 * WALL_REACT (src/fields.f90:4302-4341), TALLY_PARTICLE_TO_BOUNDARY (src/postprocess.f90:1098-1147),
 * CELL_FROM_POSITION (src/grid_and_partition.f90:34-96).
 *
 * Two passes:
 *   LEAN  free flight with neighbour hops (unstructured) or periodic wraps (structured) only.  A particle
 *         that reaches any other boundary event is NOT written: its slot goes to a compact hit list.
 *         This pass is the bulk: one 32-byte pos record and one 32-byte vel record read, one pos record
 *         written per particle, mesh records served from L1/L2.  The number of cells a particle crosses
 *         varies from 0 to a handful, so a warp does not own 32 particles but a CHUNK of them: every
 *         lane runs ONE cell crossing per trip of the loop and a lane whose particle has finished takes
 *         the next particle of the chunk (ballot + popc), which keeps the FP64 face tests -- the actual
 *         bound of this kernel, ncu r01a: 12 of 32 lanes active on average without it -- on full warps.
 *   FULL  one thread per hit-list entry; redoes the particle from its start-of-step state with every
 *         boundary model enabled (specular, Maxwell diffuse, CLL, wall reactions, porous walls,
 *         dielectric charging, tallies).  Deterministic arithmetic makes the redo identical up to the
 *         boundary, so LEAN+FULL == one sequential ADVECT.
 * No FMA contraction (-fmad=false) and the reference's evaluation order are kept so that positions and
 * cell indices match the CPU oracle bit for bit away from walls.
 */
#include "pgpu_internal.cuh"
#include "pgpu_physics.cuh"

namespace {

struct PState {
    double x, y, z, vx, vy, vz, erot, evib, dtrim;
    int ic, sid;
    int foreign; /* the current cell belongs to another rank (carried by the hop codes, PG_CODE_FOREIGN) */
};

enum { DIRTY_V = 1, DIRTY_E = 2, DIRTY_S = 4 };

/* ref: src/timecycle.f90:2297-2332 MOVE_PARTICLE */
template <bool AXI>
__device__ __forceinline__ void move_particle(PState& q, double time) {
    if (AXI) {
        q.x = q.x + q.vx * time;
        q.z = 0.0;
        double yy = q.y + q.vy * time;
        double zz = q.vz * time;
        double r = sqrt(yy * yy + zz * zz);
        double sintheta = q.vz * time / r;
        double costheta = copysign(sqrt(1. - sintheta * sintheta), yy);
        q.y = r;
        double vz = q.vz, vy = q.vy;
        q.vz = costheta * vz - sintheta * vy;
        q.vy = sintheta * vz + costheta * vy;
    } else {
        q.x = q.x + q.vx * time;
        q.y = q.y + q.vy * time;
        q.z = q.z + q.vz * time;
    }
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* NORM2 as gfortran expands it inline (scaled accumulation) -- used by the CLL tangent */
__device__ __forceinline__ double norm2_3(const double* x) {
    double result = 0.0, scale = 1.0;
    for (int i = 0; i < 3; i++) {
        if (x[i] != 0.0) {
            double ax = fabs(x[i]);
            if (scale < ax) {
                double val = scale / ax;
                result = 1.0 + result * val * val;
                scale = ax;
            } else {
                double val = ax / scale;
                result += val * val;
            }
        }
    }
    return scale * sqrt(result);
}

/* ref: src/fields.f90:5341-5431 COMPUTE_WEIGHTS (structured grids) */
__device__ __forceinline__ void compute_weights(const PgDev& d, const PState& q, double* wt, int* indi, int* indj) {
    int nx = d.nx;
    int i = (q.ic - 1) % nx;
    int j = (q.ic - 1) / nx;
    double hx, hy, cxmin, cymin;
    if (d.grid_type == PGPU_RECTILINEAR_UNIFORM) {
        hx = (d.xmax - d.xmin) / (double)nx;
        hy = (d.ymax - d.ymin) / (double)d.ny;
        cxmin = i * hx + d.xmin;
        cymin = j * hy + d.ymin;
    } else {
        hx = d.xcoord[i + 1] - d.xcoord[i];
        hy = d.ycoord[j + 1] - d.ycoord[j];
        cxmin = d.xcoord[i];
        cymin = d.ycoord[j];
    }
    double fx = (q.x - cxmin) / hx;
    double fy = (q.y - cymin) / hy;
    if (d.dims == 2) {
        indi[0] = i; indi[1] = i + 1; indi[2] = i; indi[3] = i + 1;
        indj[0] = j; indj[1] = j; indj[2] = j + 1; indj[3] = j + 1;
        wt[0] = (1. - fx) * (1. - fy);
        wt[1] = fx * (1. - fy);
        wt[2] = (1. - fx) * fy;
        wt[3] = fx * fy;
    } else {
        indi[0] = i; indi[1] = i + 1; indi[2] = 0; indi[3] = 0;
        indj[0] = 0; indj[1] = 0; indj[2] = 0; indj[3] = 0;
        wt[0] = (1. - fx); wt[1] = fx; wt[2] = 0; wt[3] = 0;
    }
}

/* ref: src/fields.f90:4194-4223 APPLY_E_FIELD */
__device__ __forceinline__ void apply_e_field(const PgDev& d, const PState& q, double* e) {
    if (!d.e_field) { e[0] = e[1] = e[2] = 0; return; }
    if (d.grid_type == PGPU_UNSTRUCTURED) {
        for (int k = 0; k < 3; k++) e[k] = __ldg(&d.e_field[3 * (size_t)(q.ic - 1) + k]);
    } else {
        double wt[4];
        int indi[4], indj[4];
        compute_weights(d, q, wt, indi, indj);
        int npy = (d.dims == 2) ? d.ny + 1 : 1;
#define EF(k, jj, ii) __ldg(&d.e_field[(k) + 3 * ((jj) + (size_t)npy * (ii))])
        if (d.dims == 2) {
            for (int k = 0; k < 3; k++)
                e[k] = wt[0] * EF(k, indj[0], indi[0]) + wt[1] * EF(k, indj[1], indi[1]) +
                       wt[2] * EF(k, indj[2], indi[2]) + wt[3] * EF(k, indj[3], indi[3]);
        } else {
            for (int k = 0; k < 3; k++) e[k] = wt[0] * EF(k, indj[0], indi[0]) + wt[1] * EF(k, indj[1], indi[1]);
        }
#undef EF
    }
}

/* ref: src/fields.f90:4227-4249 APPLY_B_FIELD (2-D unstructured nodal interpolation) */
__device__ __forceinline__ void apply_b_field(const PgDev& d, const PState& q, double* b) {
    b[0] = b[1] = b[2] = 0.0;
    if (d.grid_type == PGPU_UNSTRUCTURED && d.b_field && d.dims == 2) {
        size_t c = (size_t)(q.ic - 1);
        for (int j = 0; j < 3; j++) {
            int vj = d.cell_nodes[3 * c + j];
            const double* bs = d.basis + 9 * c + 3 * j;
            double psij = q.x * bs[0] + q.y * bs[1] + bs[2];
            for (int k = 0; k < 3; k++) b[k] = b[k] + d.b_field[3 * (size_t)(vj - 1) + k] * psij;
        }
    }
}

/* ref: src/timecycle.f90:2271-2291 UPDATE_VELOCITY_BORIS
 * (Tried r02: answering the 0 / m divisions of an unmagnetised or 1-D push without the ~100-instruction FP64 division
 *  subroutine removes 60 % of the C4 ADVECT instructions and not a microsecond of its time -- the kernel waits on memory
 *  -- while the extra branches cost the tet walk registers it does not have: +0.25 ms at full size.  Not kept.) */
__device__ __forceinline__ void boris(double dtime, PState& q, double charge, double mass, const double* e,
                                      const double* b) {
    double qc = charge * PG_QE;
    double vo[3] = {q.vx, q.vy, q.vz};
    double vm[3], t[3], vp[3], s[3], vpl[3], c[3];
    for (int k = 0; k < 3; k++) vm[k] = vo[k] + 0.5 * qc * e[k] / mass * dtime;
    for (int k = 0; k < 3; k++) t[k] = 0.5 * qc * b[k] / mass * dtime;
    cross3(vm, t, c);
    for (int k = 0; k < 3; k++) vp[k] = vm[k] + c[k];
    double t2 = 1. + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int k = 0; k < 3; k++) s[k] = 2. * t[k] / t2;
    cross3(vp, s, c);
    for (int k = 0; k < 3; k++) vpl[k] = vm[k] + c[k];
    q.vx = vpl[0] + 0.5 * qc * e[0] / mass * dtime;
    q.vy = vpl[1] + 0.5 * qc * e[1] / mass * dtime;
    q.vz = vpl[2] + 0.5 * qc * e[2] / mass * dtime;
}

/* ref: src/fields.f90:4302-4341 WALL_REACT (no EXIT after a successful reaction) */
__device__ void wall_react(const PgDev& d, PgRng& g, PState& q, int& remove, int& dirty) {
    int js = q.sid;
    double prob_scale = 1.;
    remove = 0;
    for (int jr = 0; jr < d.nwr; jr++) {
        if (d.wr_r[jr] == js) {
            if (g.rf() <= d.wr_prob[jr] / prob_scale) {
                if (d.wr_nprod[jr] == 0) {
                    remove = 1;
                    q.dtrim = 0.;
                } else if (d.wr_nprod[jr] == 1) {
                    int jp = d.wr_p1[jr];
                    q.sid = jp;
                    double vel_scale = d.mass[js - 1] / d.mass[jp - 1];
                    q.vx = q.vx * vel_scale;
                    q.vy = q.vy * vel_scale;
                    q.vz = q.vz * vel_scale;
                    dirty |= DIRTY_V | DIRTY_S;
                }
            } else {
                prob_scale = prob_scale - d.wr_prob[jr];
            }
        }
    }
}

/* ref: src/postprocess.f90:1098-1147; layout [NIN,NOUT,PXIN,PYIN,PZIN,PXOUT,PYOUT,PZOUT,EIN,EOUT][s][b] */
__device__ void tally(const PgDev& d, int reflected, const PState& q, int ic, int iface) {
    if (!d.face_bidx || !d.tallies) return;
    int b = d.face_bidx[(size_t)d.nf * (ic - 1) + iface];
    if (b < 1) return;
    size_t nbs = (size_t)d.nboundcells * d.ns;
    size_t index = (size_t)(b - 1) + (size_t)(q.sid - 1) * d.nboundcells;
    double area = d.bound_area[b - 1];
    double molmass = d.mass[q.sid - 1];
    double K = d.fnum * molmass / area / d.dt;
    double* T = d.tallies;
    double v2 = q.vx * q.vx + q.vy * q.vy + q.vz * q.vz;
    if (reflected) {
        atomicAdd(&T[1 * nbs + index], -(d.fnum / area / d.dt));
        atomicAdd(&T[5 * nbs + index], -(K * q.vx));
        atomicAdd(&T[6 * nbs + index], -(K * q.vy));
        atomicAdd(&T[7 * nbs + index], -(K * q.vz));
        atomicAdd(&T[9 * nbs + index], -(K * v2));
    } else {
        atomicAdd(&T[0 * nbs + index], d.fnum / area / d.dt);
        atomicAdd(&T[2 * nbs + index], K * q.vx);
        atomicAdd(&T[3 * nbs + index], K * q.vy);
        atomicAdd(&T[4 * nbs + index], K * q.vz);
        atomicAdd(&T[8 * nbs + index], K * v2);
    }
}

/* ref: src/grid_and_partition.f90:34-96 CELL_FROM_POSITION; BINARY_SEARCH src/tools.f90:1499-1535 */
__device__ __forceinline__ int binary_search(double value, const double* a, int n) {
    int L = 1, R = n, idx;
    if (value < a[L - 1] || value > a[R - 1]) return -1;
    if (R == L + 1) return L;
    /* DEVIATION: value == ARRAY(R) makes the reference loop forever (src/tools.f90:1518-1531) */
    if (value == a[R - 1]) return R - 1;
    for (;;) {
        idx = (L + R) / 2;
        if (a[idx - 1] <= value) {
            if (a[idx] > value) return idx;
            L = idx;
        } else {
            if (a[idx - 2] <= value) return idx - 1;
            R = idx;
        }
    }
}
__device__ __forceinline__ int cell_from_position(const PgDev& d, double xp, double yp) {
    if (d.grid_type == PGPU_RECTILINEAR_UNIFORM) {
        double dx = (d.xmax - d.xmin) / d.nx;
        double dy = (d.ymax - d.ymin) / d.ny;
        int xcell = (int)((xp - d.xmin) / dx);
        int ycell = (int)((yp - d.ymin) / dy);
        return xcell + d.nx * ycell + 1;
    } else {
        int xcell = binary_search(xp, d.xcoord, d.nxp);
        int ycell = binary_search(yp, d.ycoord, d.nyp);
        if (xcell < 0 || ycell < 0) return -1;
        return xcell + d.nx * (ycell - 1);
    }
}

/* diffuse re-emission, structured flavour. ref: src/timecycle.f90:1915-1931, 1983-1999 */
__device__ void diffuse_structured(const PgDev& d, PgRng& g, PState& q, double temp, double nxw, double nyw) {
    int s = q.sid;
    double v3[3];
    pg_maxwell0(g, temp, d.mass[s - 1], v3);
    double erot = pg_internal_energy(g, d.rotdof[s - 1], temp);
    double evib = pg_internal_energy(g, d.vibdof[s - 1], temp);
    double v_norm = pg_flx(g, 0., temp, d.mass[s - 1]);
    q.vx = v_norm * nxw - v3[1] * nyw;
    q.vy = v3[1] * nxw + v_norm * nyw;
    q.vz = v3[2];
    q.erot = erot;
    q.evib = evib;
}

/* quadratic candidate solver of the axisymmetric variants (ref: :1190-1226, :1690-1712, :1793-1815) */
__device__ int axi_roots(double coefa, double coefb, double coefc, double dtcoll, double tol, int use_radial_check,
                         double beta, double alpha, double vx, double y1, double* test) {
    double delta = coefb * coefb - coefa * coefc;
    if (!(delta >= 0)) return 0;
    double sol1 = (-coefb - sqrt(delta)) / coefa;
    double sol2 = (-coefb + sqrt(delta)) / coefa;
    int ok1 = (sol1 >= -tol && sol1 < dtcoll && (!use_radial_check || beta + alpha * vx * sol1 + y1 >= 0));
    int ok2 = (sol2 >= -tol && sol2 < dtcoll && (!use_radial_check || beta + alpha * vx * sol2 + y1 >= 0));
    if (ok1) {
        if (ok2) {
            if (sol1 <= sol2) { test[0] = sol1; test[1] = sol2; }
            else { test[0] = sol2; test[1] = sol1; }
            return 2;
        }
        test[0] = sol1;
        return 1;
    }
    if (ok2) { test[0] = sol2; return 1; }
    return 0;
}

/* boundary interaction on an unstructured grid. ref: src/timecycle.f90:1363-1647 */
template <int DIMS>
__device__ void boundary_unstructured(const PgDev& d, PgRng& g, PState& q, int ic, int iface, int code,
                                      const double* nrm, int& remove, int& dirty) {
    if (code == 0) { /* ref: :1644-1647 boundary without physical group */
        remove = 1;
        q.dtrim = 0.;
        return;
    }
    int gi = -code - 1;
    double fn[3] = {-nrm[0], -nrm[1], -nrm[2]};
    if (d.tally_on) tally(d, 0, q, ic, iface);

    /* dielectric surface charge, ref: :1399-1430 */
    double charge = d.charge[q.sid - 1];
    if (d.bc_field[gi] == PGPU_DIELECTRIC_BC && fabs(charge) >= 1.e-6 && d.surface_charge) {
        double K = PG_QE / (PG_EPS0 * (d.eps_scaling * d.eps_scaling));
        size_t c = (size_t)(ic - 1);
        if (DIMS == 1) {
            double rho_q = K * charge * d.fnum / (d.ymax - d.ymin) / (d.zmax - d.zmin);
            for (int i = 0; i < 2; i++) {
                int vp = d.cell_nodes[2 * c + i];
                const double* bs = d.basis + 4 * c + 2 * i;
                double psip = bs[0] * q.x + bs[1];
                atomicAdd(&d.surface_charge[vp - 1], rho_q * psip);
            }
        } else if (DIMS == 2) {
            double rho_q = K * charge * d.fnum / (d.zmax - d.zmin);
            for (int i = 0; i < 3; i++) {
                int vp = d.cell_nodes[3 * c + i];
                const double* bs = d.basis + 9 * c + 3 * i;
                double psip = bs[0] * q.x + bs[1] * q.y + bs[2];
                atomicAdd(&d.surface_charge[vp - 1], rho_q * psip);
            }
        } else {
            double rho_q = K * charge * d.fnum;
            for (int i = 0; i < 4; i++) {
                int vp = d.cell_nodes[4 * c + i];
                const double* bs = d.basis + 16 * c + 4 * i;
                double psip = bs[0] * q.x + bs[1] * q.y + bs[2] * q.z + bs[3];
                atomicAdd(&d.surface_charge[vp - 1], rho_q * psip);
            }
        }
    }

    int pbc = d.bc_particle[gi];
    if (pbc == PGPU_SPECULAR) {
        /* ref: :1433-1443 */
        if (d.bc_react[gi]) wall_react(d, g, q, remove, dirty);
        double vdotn = q.vx * fn[0] + q.vy * fn[1] + q.vz * fn[2];
        q.vx = q.vx - 2. * vdotn * fn[0];
        q.vy = q.vy - 2. * vdotn * fn[1];
        q.vz = q.vz - 2. * vdotn * fn[2];
        dirty |= DIRTY_V;
    } else if (pbc == PGPU_DIFFUSE) {
        /* ref: :1445-1477 */
        if (d.bc_react[gi]) wall_react(d, g, q, remove, dirty);
        double ft1[3], ft2[3];
        if (DIMS == 1) {
            ft1[0] = 0.; ft1[1] = fn[0]; ft1[2] = 0.;
            ft2[0] = 0.; ft2[1] = 0.; ft2[2] = 1.;
        } else if (DIMS == 2) {
            ft1[0] = -fn[1]; ft1[1] = fn[0]; ft1[2] = 0.;
            ft2[0] = 0.; ft2[1] = 0.; ft2[2] = 1.;
        } else {
            size_t o = 3 * (4 * (size_t)(ic - 1) + iface);
            for (int k = 0; k < 3; k++) {
                ft1[k] = d.face_tang1[o + k];
                ft2[k] = d.face_tang2[o + k];
            }
        }
        int s = q.sid;
        double wt = d.bc_temp[gi];
        double v3[3];
        pg_maxwell0(g, wt, d.mass[s - 1], v3);
        double erot = pg_internal_energy(g, d.rotdof[s - 1], wt);
        double evib = pg_internal_energy(g, d.vibdof[s - 1], wt);
        double v_perp = pg_flx(g, 0., wt, d.mass[s - 1]);
        q.vx = v_perp * fn[0] + v3[1] * ft1[0] + v3[2] * ft2[0];
        q.vy = v_perp * fn[1] + v3[1] * ft1[1] + v3[2] * ft2[1];
        q.vz = v_perp * fn[2] + v3[1] * ft1[2] + v3[2] * ft2[2];
        q.erot = erot;
        q.evib = evib;
        dirty |= DIRTY_V | DIRTY_E;
    } else if (pbc == PGPU_CLL) {
        /* ref: :1485-1550 */
        if (d.bc_react[gi]) wall_react(d, g, q, remove, dirty);
        double vdotn = q.vx * fn[0] + q.vy * fn[1] + q.vz * fn[2];
        double tg1[3], tg2[3];
        tg1[0] = q.vx - vdotn * fn[0];
        tg1[1] = q.vy - vdotn * fn[1];
        tg1[2] = q.vz - vdotn * fn[2];
        double nrm2 = norm2_3(tg1);
        for (int k = 0; k < 3; k++) tg1[k] = tg1[k] / nrm2;
        cross3(fn, tg1, tg2);
        double vdott1 = q.vx * tg1[0] + q.vy * tg1[1] + q.vz * tg1[2];
        int s = q.sid;
        double wt = d.bc_temp[gi];
        double accn = d.bc_accn[gi], acct = d.bc_acct[gi];
        double vrm = sqrt(2. * PG_KB * wt / d.mass[s - 1]);
        double rn = g.rf();
        while (rn < 1.0e-13) rn = g.rf();
        double r1 = sqrt(-accn * log(rn));
        double theta1 = 2. * PG_PI * g.rf();
        double dot_norm = vdotn / vrm * sqrt(1 - accn);
        double v_perp = vrm * sqrt(r1 * r1 + dot_norm * dot_norm + 2. * r1 * dot_norm * cos(theta1));
        rn = g.rf();
        while (rn < 1.0e-13) rn = g.rf();
        double r2 = sqrt(-acct * log(rn));
        double theta2 = 2. * PG_PI * g.rf();
        double vtangent = vdott1 / vrm * sqrt(1 - acct);
        double v_tang1 = vrm * (vtangent + r2 * cos(theta2));
        double v_tang2 = vrm * r2 * sin(theta2);
        double erot = pg_internal_energy(g, d.rotdof[s - 1], wt);
        double evib = pg_internal_energy(g, d.vibdof[s - 1], wt);
        q.vx = v_perp * fn[0] + v_tang1 * tg1[0] + v_tang2 * tg2[0];
        q.vy = v_perp * fn[1] + v_tang1 * tg1[1] + v_tang2 * tg2[1];
        q.vz = v_perp * fn[2] + v_tang1 * tg1[2] + v_tang2 * tg2[2];
        q.erot = erot;
        q.evib = evib;
        dirty |= DIRTY_V | DIRTY_E;
    } else {
        /* VACUUM, REACT, AXIS, PERIODIC_*, EMIT fall through to removal (ref: :1611-1614);
         * WB_BC (washboard) is out of scope: its tables are not shipped with the reference. */
        remove = 1;
        q.dtrim = 0.;
    }
    if (!remove && d.tally_on) tally(d, 1, q, ic, iface);
}

enum { IT_NONE = 0, IT_CONTINUE = 1, IT_DONE = 2, IT_DEFER = 3 };

/*
 * Face selection of one trip on a non-axisymmetric unstructured grid, ref: src/timecycle.f90:1257-1305:
 *     DO I = 1, NFACES:  VN = V.n_I;  DX = (P_I - X).n_I;  IF (VN >= 0 .AND. VN*DTCOLL >= DX) DTCOLL = DX/VN; BOUNDCOLL = I
 * (sequential tightening, '>=' tie rule, last one wins).  The kernel is bound by the L1 data pipe -- the bytes each
 * lane pulls per trip (ncu r01o: l1tex__data_pipe_lsu_wavefronts 76 %) -- so the 48 bytes per face of FP64 operands
 * are not read for every face.  A 16-byte single-precision FILTER record per face {float(n), float(P.n)} gives VN and
 * DX with rigorous error bounds, which classify every face as
 *     fail   it cannot pass the test for any DTCOLL <= DTRIM (VN < 0 for certain, or VN*DTRIM < DX for certain);
 *     pass   it passes the test against DTRIM for certain; its crossing time is then at most t_hi;
 *     open   anything else.
 * With T = min t_hi over the `pass` faces, a face whose crossing time is certainly > T (t_lo > T, relative margin
 * 1e-6) cannot change the outcome of the sequential loop: tested after the face of T it fails against the tightened
 * DTCOLL, tested before it whatever it does leaves DTCOLL above T, where the face of T still passes and overwrites
 * it (the rounded product VN*DTCOLL is monotone in DTCOLL, and the margins are nine orders of magnitude above the
 * FP64 rounding of the test).  So the reference loop restricted to the remaining RELEVANT faces, run on the exact
 * FP64 operands in the original face order, returns the same (DTCOLL, BOUNDCOLL) bit for bit; that is what is
 * executed.  On average one face per crossing is relevant and none on the final flight (5 instead of 13 sixteen-byte
 * loads, or 8 with one exact face); a particle near an edge or a vertex, NaNs and degenerate inputs simply make more
 * faces relevant.  Bounds (u = 2^-24): |VN_f - VN| <= 5u.sum|v_k||n_k| <= 1e-6.(|vx|+|vy|+|vz|) =: EV and
 * |DX_f - DX| <= 6u.sum|x_k| + 2u.|P.n| <= 1e-6.(|x|+|y|+|z| + |c_f|) =: ED, each with a 2.8x margin over the worst case
 * of rounding the inputs and evaluating in single precision in any order.
 */
template <int DIMS>
__device__ __forceinline__ void select_face(const PgDev& d, const PState& q, int ic, double& dtcoll, int& boundcoll,
                                            int& codesel, double* nsel) {
    constexpr int NF = DIMS + 1;
    const char* fb = pg_filt_base(d.face_filt, ic - 1, NF + 1);
    float4 F[NF];
#pragma unroll
    for (int i = 0; i < NF; i++) F[i] = __ldg(reinterpret_cast<const float4*>(fb + i * PG_GEO_LINE));
    const int4 codes = __ldg(reinterpret_cast<const int4*>(fb + NF * PG_GEO_LINE));

    const float INF = __int_as_float(0x7f800000);
    const float vxf = (float)q.vx, vyf = (float)q.vy, vzf = (float)q.vz;
    const float xf = (float)q.x, yf = (float)q.y, zf = (float)q.z;
    float V1 = fabsf(vxf), X1 = fabsf(xf);
    if (DIMS >= 2) { V1 += fabsf(vyf); X1 += fabsf(yf); }
    if (DIMS >= 3) { V1 += fabsf(vzf); X1 += fabsf(zf); }
    const float EV = 1.e-6f * V1;
    const float dtf = (float)dtcoll;
    const float dt_hi = dtf * (1.f + 2.e-6f), dt_lo = dtf * (1.f - 2.e-6f);
    float t_lo[NF];
    float T = INF;
    unsigned int open = 0u;
    if constexpr (DIMS == 3) {
        /* Tetrahedra: the kernel issues on 65 % of its cycles (ncu r02t) and the brackets are a third of its
         * instructions, so the four faces are evaluated two at a time in packed single precision (FFMA2 / FADD2 /
         * FMUL2 of sm_100: the same roundings as the scalar code below, half the instructions), the accurate-range
         * tests of the reciprocals are one integer compare each, and the products VN*DT of the scalar form are replaced
         * by comparisons of the crossing-time brackets that are needed anyway:
         *     t_lo = dx_lo / vn_hi (a lower bound of the crossing time, if dx_lo > 0 and vn_hi is in range, else -inf)
         *     t_hi = max(dx_hi, 0) / vn_lo * (1 + 4e-6)   (an upper bound, if vn_lo is in range)
         *     fail  <=  vn_hi < 0  or  t_lo > dt_hi ;   pass  <=  vn_lo in range  and  t_hi < dt_lo .
         * The approximate reciprocal is good to 1 ulp and the product to half an ulp; t_lo carries no margin of its own,
         * t_hi carries both (4e-6 >> 2 x 2^-22), so that t_lo > T still means "certainly later than the face of T". */
        const float2 vx2 = make_float2(vxf, vxf), vy2 = make_float2(vyf, vyf), vz2 = make_float2(vzf, vzf);
        const float2 x2 = make_float2(xf, xf), y2 = make_float2(yf, yf), z2 = make_float2(zf, zf);
        const float2 pEV = make_float2(EV, EV), nEV = make_float2(-EV, -EV), X12 = make_float2(X1, X1);
        const float2 cp = make_float2(1.e-6f, 1.e-6f), cn = make_float2(-1.e-6f, -1.e-6f), m1 = make_float2(-1.f, -1.f);
        const float2 c4 = make_float2(1.f + 4.e-6f, 1.f + 4.e-6f);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float4 A = F[2 * h], B = F[2 * h + 1];
            const float2 nx = make_float2(A.x, B.x), ny = make_float2(A.y, B.y), nz = make_float2(A.z, B.z), w = make_float2(A.w, B.w);
            float2 vn = __fmul2_rn(vx2, nx), xn = __fmul2_rn(x2, nx);
            vn = __ffma2_rn(vy2, ny, vn); xn = __ffma2_rn(y2, ny, xn);
            vn = __ffma2_rn(vz2, nz, vn); xn = __ffma2_rn(z2, nz, xn);
            const float2 dx = __ffma2_rn(xn, m1, w); /* w - xn, one rounding */
            const float2 Ea = __fadd2_rn(X12, make_float2(fabsf(A.w), fabsf(B.w)));
            const float2 vhi = __fadd2_rn(vn, pEV), vlo = __fadd2_rn(vn, nEV);
            const float2 dhi = __ffma2_rn(Ea, cp, dx), dlo = __ffma2_rn(Ea, cn, dx);
            float2 rhi, rlo;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rhi.x) : "f"(vhi.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rhi.y) : "f"(vhi.y));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rlo.x) : "f"(vlo.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rlo.y) : "f"(vlo.y));
            const float2 tl = __fmul2_rn(dlo, rhi);
            const float2 th = __fmul2_rn(__fmul2_rn(make_float2(fmaxf(dhi.x, 0.f), fmaxf(dhi.y, 0.f)), rlo), c4);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int i = 2 * h + e;
                const float vhi_e = e ? vhi.y : vhi.x, vlo_e = e ? vlo.y : vlo.x, dlo_e = e ? dlo.y : dlo.x;
                const float tl_e = e ? tl.y : tl.x, th_e = e ? th.y : th.x;
                /* x in (1e-30, 1e30) as one unsigned compare of the bit patterns (negative, zero, NaN: false) */
                const bool hi_ok = (__float_as_uint(vhi_e) - 0x0DA24261u) < (0x7149F2CAu - 0x0DA24261u);
                const bool lo_ok = (__float_as_uint(vlo_e) - 0x0DA24261u) < (0x7149F2CAu - 0x0DA24261u);
                t_lo[i] = (dlo_e > 0.f && hi_ok) ? tl_e : -INF;
                const bool fail = (vhi_e < 0.f) || (t_lo[i] > dt_hi);
                const bool pass = lo_ok && (th_e < dt_lo);
                T = fminf(T, pass ? th_e : INF);
                if (!fail) open |= 1u << i; /* (a NaN anywhere compares false: the face stays open) */
            }
        }
    } else {
#pragma unroll
    for (int i = 0; i < NF; i++) {
        /* (explicit fused multiply-adds: the bounds hold for any single-precision evaluation order) */
        float vnf = vxf * F[i].x, xn = xf * F[i].x;
        if (DIMS >= 2) { vnf = __fmaf_rn(vyf, F[i].y, vnf); xn = __fmaf_rn(yf, F[i].y, xn); }
        if (DIMS >= 3) { vnf = __fmaf_rn(vzf, F[i].z, vnf); xn = __fmaf_rn(zf, F[i].z, xn); }
        const float dxf = F[i].w - xn;
        const float ED = 1.e-6f * (X1 + fabsf(F[i].w));
        const float vn_hi = vnf + EV, vn_lo = vnf - EV, dx_lo = dxf - ED, dx_hi = dxf + ED;
        const bool fail = (vn_hi < 0.f) || (dx_lo > 0.f && vn_hi * dt_hi < dx_lo);
        const bool pass = (vn_lo > 1.e-30f) && (vn_lo < 1.e30f) && (vn_lo * dt_lo > dx_hi);
        /* crossing-time bracket with the approximate reciprocal (__fdividef: 2 ulp, no slow path; covered by the
         * 2e-6 margins).  A denominator outside its accurate range leaves the face relevant / unbounded. */
        const float tl = __fdividef(dx_lo, vn_hi) * (1.f - 2.e-6f);
        t_lo[i] = (dx_lo > 0.f && vn_hi > 1.e-30f && vn_hi < 1.e30f) ? tl : -INF;
        const float th = __fdividef(fmaxf(dx_hi, 0.f), vn_lo) * (1.f + 2.e-6f);
        T = pass ? fminf(T, th) : T;
        if (!fail) open |= 1u << i; /* (a NaN anywhere compares false: the face stays open) */
    }
    }
    unsigned int rel = 0u;
#pragma unroll
    for (int i = 0; i < NF; i++)
        if (((open >> i) & 1u) && !(t_lo[i] > T)) rel |= 1u << i;

    /* A single relevant face that passes for certain WILL be taken: its neighbour is the next cell.  Ask for that
     * cell's filter lines now, so that their latency overlaps the exact-operand load below instead of following it. */
    if (PG_ADVECT_PREFETCH && rel && !(rel & (rel - 1u))) {
        const int f = __ffs((int)rel) - 1;
        const int cn = (f == 0) ? codes.x : (f == 1) ? codes.y : (f == 2) ? codes.z : codes.w;
        if (cn > 0) {
            const char* nb = pg_filt_base(d.face_filt, (cn & PG_CODE_CELL_MASK) - 1, NF + 1);
#pragma unroll
            for (int i = 0; i <= NF; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(nb + i * PG_GEO_LINE));
        }
    }

    /* the reference loop on the exact operands, relevant faces only, original order */
    const double* ex = d.face_exact + (size_t)(ic - 1) * (size_t)(NF * 8);
    while (rel) {
        const int i = __ffs((int)rel) - 1;
        rel &= rel - 1u;
        const PgF64x4 a = pg_ldg256(ex + 8 * i);                                           /* nx, ny, nz, Px */
        const double2 b = __ldg(reinterpret_cast<const double2*>(ex + 8 * i + 4));        /* Py, Pz */
        double vn, dx;
        if (DIMS == 1) {
            vn = q.vx * a.a;
            dx = (a.d - q.x) * a.a;
        } else if (DIMS == 2) {
            vn = q.vx * a.a + q.vy * a.b;
            dx = (a.d - q.x) * a.a + (b.x - q.y) * a.b;
        } else {
            vn = q.vx * a.a + q.vy * a.b + q.vz * a.c;
            dx = (a.d - q.x) * a.a + (b.x - q.y) * a.b + (b.y - q.z) * a.c;
        }
        if (vn >= 0. && vn * dtcoll >= dx) {
            dtcoll = dx / vn;
            boundcoll = i;
            nsel[0] = a.a;
            nsel[1] = (DIMS >= 2) ? a.b : 0.;
            nsel[2] = (DIMS >= 3) ? a.c : 0.;
        }
    }
    codesel = (boundcoll == 0) ? codes.x : (boundcoll == 1) ? codes.y : (boundcoll == 2) ? codes.z : codes.w;
}

/* One trip of the ADVECT while-loop (one cell crossing or the final flight) of a particle on an
 * unstructured grid. ref: src/timecycle.f90:1166-1657.  IT_DEFER: LEAN and the particle reached a
 * boundary event; it is left untouched for the FULL pass. */
template <int DIMS, bool AXI, bool LEAN>
__device__ __forceinline__ int unstructured_iter(const PgDev& d, PgRng& g, PState& q, int& remove, int& dirty, int& hops) {
    constexpr int NF = DIMS + 1;
    const double TOL = 1.0e-15;
    int ic = q.ic;
    {
        double dtcoll = q.dtrim;
        int boundcoll = -1;
        int codesel = 0; /* hop code of the face selected so far */
        double nsel[3] = {0., 0., 0.}; /* its outward normal */
        if (AXI) {
            /* ref: :1173-1253.  The axisymmetric walk reads the double-precision chunked record (PG_GEO_*): blocks of
             * 8 cells, chunk-major, so that lanes in nearby cells share their L1 wavefronts. */
            constexpr int NCH = PG_GEO_CHUNKS(NF);
            const char* gb = pg_geo_base(d.face_geo, ic - 1, NCH);
            const int4 codes = __ldg(reinterpret_cast<const int4*>(gb + (NF * 3) * PG_GEO_LINE));
#pragma unroll 1
            for (int i = 0; i < 3; i++) {
                double test[2];
                int goodsol;
                double n1 = pg_geo_at(gb, 6 * i + 0), n2 = pg_geo_at(gb, 6 * i + 1);
                double ex1 = pg_geo_at(gb, 6 * i + 3), ey1 = pg_geo_at(gb, 6 * i + 4);
                if (fabs(n2) < 1.e-10) {
                    double sol1 = (ex1 - q.x) / q.vx;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double alpha = -n1 / n2;
                    double beta = (q.x - ex1) * alpha;
                    double coefa = q.vy * q.vy + q.vz * q.vz - alpha * alpha * (q.vx * q.vx);
                    double coefb = q.y * q.vy - (beta + ey1) * alpha * q.vx;
                    double coefc = q.y * q.y - (beta + ey1) * (beta + ey1);
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, TOL, 1, beta, alpha, q.vx, ey1, test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle<true>(q, test[s]);
                    if ((q.vx * n1 + q.vy * n2) > 0) {
                        double len = d.edge_len[3 * (size_t)(ic - 1) + i];
                        double colldist = -n2 * (q.x - ex1) / len + n1 * (q.y - ey1) / len;
                        if (colldist >= 0 && colldist <= 1) {
                            dtcoll = test[s];
                            boundcoll = i;
                            codesel = (i == 0) ? codes.x : (i == 1) ? codes.y : codes.z;
                            nsel[0] = n1; nsel[1] = n2; nsel[2] = 0.;
                        }
                    }
                    move_particle<true>(q, -test[s]);
                }
            }
        } else {
            select_face<DIMS>(d, q, ic, dtcoll, boundcoll, codesel, nsel);
        }

        if (boundcoll != -1) {
            const int code = codesel;
            if (code > 0) {
                /* ref: :1311-1312, :1648-1652 cell hop */
                move_particle<AXI>(q, dtcoll);
                q.dtrim = q.dtrim - dtcoll;
                q.ic = code & PG_CODE_CELL_MASK;
                q.foreign = (code & PG_CODE_FOREIGN) != 0;
                ic = q.ic;
                if (++hops > PG_MAX_HOPS) {
                    pg_set_err(d, PG_DEVERR_HOPS);
                    remove = 1;
                    q.dtrim = 0.;
                }
            } else if (LEAN) {
                /* A plain specular wall (no reactions, no charging, no tally this step) involves no random numbers
                 * and no side effects: reflect here (ref: :1433-1443, same expressions as boundary_unstructured)
                 * and go on.  Everything else is left untouched for the FULL pass. */
                if (AXI || code > -PG_CODE_SPEC || d.tally_on) return IT_DEFER;
                move_particle<AXI>(q, dtcoll);
                q.dtrim = q.dtrim - dtcoll;
                const double f0 = -nsel[0];
                const double f1 = (DIMS >= 2) ? -nsel[1] : -0.;
                const double f2 = (DIMS >= 3) ? -nsel[2] : -0.;
                const double vdotn = q.vx * f0 + q.vy * f1 + q.vz * f2;
                q.vx = q.vx - 2. * vdotn * f0;
                q.vy = q.vy - 2. * vdotn * f1;
                q.vz = q.vz - 2. * vdotn * f2;
                dirty |= DIRTY_V;
                if (++hops > PG_MAX_HOPS) {
                    pg_set_err(d, PG_DEVERR_HOPS);
                    remove = 1;
                    q.dtrim = 0.;
                }
            } else {
                move_particle<AXI>(q, dtcoll);
                q.dtrim = q.dtrim - dtcoll;
                if (AXI) dirty |= DIRTY_V; /* the axisymmetric move rotates the velocity */
                boundary_unstructured<DIMS>(d, g, q, ic, boundcoll, (code <= -PG_CODE_SPEC) ? code + PG_CODE_SPEC : code, nsel,
                                            remove, dirty);
                if (++hops > PG_MAX_HOPS) {
                    pg_set_err(d, PG_DEVERR_HOPS);
                    remove = 1;
                    q.dtrim = 0.;
                }
            }
        } else {
            /* ref: :1653-1657 */
            move_particle<AXI>(q, q.dtrim);
            q.dtrim = 0.;
        }
    }
    if (AXI) dirty |= DIRTY_V;
    return (q.dtrim > 0.) ? IT_CONTINUE : IT_DONE;
}

/* One trip of the ADVECT while-loop of a particle on a rectilinear grid. ref: src/timecycle.f90:1659-2015 */
template <bool AXI, bool LEAN>
__device__ __forceinline__ int structured_iter(const PgDev& d, PgRng& g, PState& q, int& remove, int& dirty, int* bhits,
                                               int& iters) {
    const double NORMX[4] = {1., -1., 0., 0.};
    const double NORMY[4] = {0., 0., 1., -1.};
    const double XW[4] = {d.xmin, d.xmax, d.xmin, d.xmax};
    const double YW[4] = {d.ymax, d.ymin, d.ymin, d.ymax};
    const double BOUNDPOS[4] = {d.xmin, d.xmax, d.ymin, d.ymax};
    const int dims = d.dims;

    {
        double dtcoll = q.dtrim;
        int boundcoll = -1;
        for (int i = 1; i <= 2 * dims; i++) {
            if (AXI && dims == 2) {
                /* ref: :1662-1729 */
                double test[2];
                int goodsol;
                if (i == 1 || i == 2) {
                    double sol1 = (BOUNDPOS[i - 1] - q.x) / q.vx;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double coefa = q.vy * q.vy + q.vz * q.vz;
                    double coefb = q.y * q.vy;
                    double coefc = q.y * q.y - BOUNDPOS[i - 1] * BOUNDPOS[i - 1];
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, 0.0, 0, 0, 0, 0, 0, test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle<true>(q, test[s]);
                    if ((q.vx * NORMX[i - 1] + q.vy * NORMY[i - 1]) < 0) {
                        dtcoll = test[s];
                        boundcoll = i;
                    }
                    move_particle<true>(q, -test[s]);
                }
            } else {
                /* ref: :1733-1752 -- strict '>' here, unlike the unstructured faces */
                double vn = -(q.vx * NORMX[i - 1] + q.vy * NORMY[i - 1]);
                double dx = (q.x - XW[i - 1]) * NORMX[i - 1] + (q.y - YW[i - 1]) * NORMY[i - 1];
                if (vn >= 0. && vn * dtcoll > dx) {
                    dtcoll = dx / vn;
                    boundcoll = i;
                }
            }
        }

        /* interior walls, ref: :1760-1865 */
        int wallcoll = -1;
        for (int i = 1; i <= d.nw; i++) {
            int j = i - 1;
            double wx1 = d.w_x1[j], wy1 = d.w_y1[j], wx2 = d.w_x2[j], wy2 = d.w_y2[j];
            double wnx = d.w_nx[j], wny = d.w_ny[j];
            if (AXI) {
                double test[2];
                int goodsol;
                if (wx1 == wx2) {
                    double sol1 = (wx1 - q.x) / q.vx;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double alpha = (wy2 - wy1) / (wx2 - wx1);
                    double beta = (q.x - wx1) * alpha;
                    double coefa = q.vy * q.vy + q.vz * q.vz - alpha * alpha * (q.vx * q.vx);
                    double coefb = q.y * q.vy - (beta + wy1) * alpha * q.vx;
                    double coefc = q.y * q.y - (beta + wy1) * (beta + wy1);
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, 0.0, 1, beta, alpha, q.vx, wy1, test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle<true>(q, test[s]);
                    if ((q.vx * wnx + q.vy * wny) < 0) {
                        double ddx = wx2 - wx1, ddy = wy2 - wy1;
                        double colldist = ((q.x - wx1) * ddx + (q.y - wy1) * ddy) / (ddx * ddx + ddy * ddy);
                        if (colldist >= 0 && colldist <= 1) {
                            dtcoll = test[s];
                            boundcoll = -1;
                            wallcoll = i;
                        }
                    }
                    move_particle<true>(q, -test[s]);
                }
            } else {
                double vn = -(q.vx * wnx + q.vy * wny);
                double dx = (q.x - wx1) * wnx + (q.y - wy1) * wny;
                if (dx >= 0. && vn * dtcoll >= dx) {
                    double cand = dx / vn;
                    double xcoll = q.x + q.vx * cand;
                    double ycoll = q.y + q.vy * cand;
                    double ddx = wx2 - wx1, ddy = wy2 - wy1;
                    double colldist = ((xcoll - wx1) * ddx + (ycoll - wy1) * ddy) / (ddx * ddx + ddy * ddy);
                    if (colldist >= 0 && colldist <= 1) {
                        boundcoll = -1;
                        wallcoll = i;
                        dtcoll = cand;
                    }
                }
            }
        }

        if (boundcoll != -1) {
            /* ref: :1868-1939 */
            int b = boundcoll - 1;
            if (LEAN && !d.bper[b]) return IT_DEFER;
            /* LOCAL_BOUNDARY_COLL_COUNT (ref: :1871-1872).  The LEAN pass may still defer the particle,
             * so it counts locally and the caller commits once the particle is final. */
            if (LEAN) bhits[b]++;
            else atomicAdd(&d.bcount[b + 4 * (q.sid - 1)], 1ULL);
            if (d.bper[b]) {
                move_particle<AXI>(q, dtcoll);
                switch (boundcoll) {
                    case 1: q.x = q.x + d.xmax - d.xmin; break;
                    case 2: q.x = q.x - d.xmax + d.xmin; break;
                    case 3: q.y = q.y + d.ymax - d.ymin; break;
                    case 4: q.y = q.y - d.ymax + d.ymin; break;
                }
                q.dtrim = q.dtrim - dtcoll;
            } else if (d.bspec[b]) {
                move_particle<AXI>(q, dtcoll);
                if (boundcoll <= 2) q.vx = -q.vx;
                else q.vy = -q.vy;
                q.dtrim = q.dtrim - dtcoll;
                dirty |= DIRTY_V;
                if (d.breact[b]) wall_react(d, g, q, remove, dirty);
            } else if (d.bdiff[b]) {
                move_particle<AXI>(q, dtcoll);
                q.dtrim = q.dtrim - dtcoll;
                if (d.breact[b]) wall_react(d, g, q, remove, dirty);
                if (!remove) {
                    diffuse_structured(d, g, q, d.boundtemp, NORMX[b], NORMY[b]);
                    dirty |= DIRTY_V | DIRTY_E;
                }
            } else {
                remove = 1;
                q.dtrim = 0.;
            }
        } else if (wallcoll != -1) {
            /* ref: :1940-2008 */
            if (LEAN) return IT_DEFER;
            int j = wallcoll - 1;
            atomicAdd(&d.wcount[j + (size_t)d.nw * (q.sid - 1)], 1ULL);
            if (d.w_spec[j]) {
                double rfp = g.rf();
                if (d.w_porous[j] && rfp <= d.w_transm[j]) {
                    remove = 1;
                    q.dtrim = 0.;
                } else {
                    move_particle<AXI>(q, dtcoll);
                    double vdotn = q.vx * d.w_nx[j] + q.vy * d.w_ny[j];
                    q.vx = q.vx - 2. * vdotn * d.w_nx[j];
                    q.vy = q.vy - 2. * vdotn * d.w_ny[j];
                    q.dtrim = q.dtrim - dtcoll;
                    dirty |= DIRTY_V;
                    if (d.w_react[j]) wall_react(d, g, q, remove, dirty);
                }
            } else if (d.w_diff[j]) {
                double rfp = g.rf();
                if (d.w_porous[j] && rfp <= d.w_transm[j]) {
                    remove = 1;
                    q.dtrim = 0.;
                } else {
                    move_particle<AXI>(q, dtcoll);
                    q.dtrim = q.dtrim - dtcoll;
                    if (d.w_react[j]) wall_react(d, g, q, remove, dirty);
                    if (!remove) {
                        diffuse_structured(d, g, q, d.w_temp[j], d.w_nx[j], d.w_ny[j]);
                        dirty |= DIRTY_V | DIRTY_E;
                    }
                }
            } else {
                remove = 1;
                q.dtrim = 0.;
            }
        } else {
            /* ref: :2010-2015 */
            move_particle<AXI>(q, q.dtrim);
            q.dtrim = 0.;
        }
        if (++iters > PG_MAX_HOPS) {
            pg_set_err(d, PG_DEVERR_HOPS);
            remove = 1;
            q.dtrim = 0.;
        }
    }
    if (AXI) dirty |= DIRTY_V;
    return (q.dtrim > 0.) ? IT_CONTINUE : IT_DONE;
}

/* ------------------------------------------------------------------------------------------
 * kernels.  MODE: 0 structured, 1/2/3 unstructured of that dimension
 * ---------------------------------------------------------------------------------------- */

/* particle slot -> registers; returns false for a tombstone.  Applies the PIC kick at the start-of-step
 * position/cell with the particle's DTRIM (ref: src/timecycle.f90:1105-1131). */
__device__ __forceinline__ bool load_particle(const PgDev& d, long long slot, PState& q, int& dirty) {
    const PgPos P = d.p.pos[slot];
    if (P.ic < 1) return false;
    const PgVel V = d.p.vel[slot];
    q.x = P.x; q.y = P.y; q.z = P.z; q.ic = P.ic; q.sid = P.sid;
    q.vx = V.vx; q.vy = V.vy; q.vz = V.vz;
    q.dtrim = d.p.dtrim ? d.p.dtrim[slot] : d.dt;
    q.erot = 0.; q.evib = 0.;
    q.foreign = 0;
    dirty = 0;
    if (d.pic != PGPU_PIC_NONE) {
        double e[3], b[3];
        apply_e_field(d, q, e);
        b[0] = b[1] = b[2] = 0.;
        if (d.b_field) apply_b_field(d, q, b);
        for (int k = 0; k < 3; k++) b[k] = b[k] + d.extb[k];
        boris(q.dtrim, q, d.charge[q.sid - 1], d.mass[q.sid - 1], e, b);
        dirty |= DIRTY_V;
    }
    return true;
}

/* end of ADVECT for one particle that was not removed: periodic wraps, structured re-index, radial weighting.
 * Returns false (and flags the error) when the particle ends outside the grid, or when the weighting deletes it. */
template <int MODE>
__device__ __forceinline__ bool finish_particle(const PgDev& d, PState& q, long long slot) {
    /* ref: :2048-2068 */
    if (d.zper) {
        while (q.z > d.zmax) q.z = d.zmin + (q.z - d.zmax);
        while (q.z < d.zmin) q.z = d.zmax + (q.z - d.zmin);
    }
    if (d.dims == 1 && d.yper) {
        while (q.y > d.ymax) q.y = d.ymin + (q.y - d.ymax);
        while (q.y < d.ymin) q.y = d.ymax + (q.y - d.ymin);
    }
    /* ref: :2105-2108 structured grids are re-indexed from the final position */
    const int old_ic = q.ic;
    if (MODE == 0) q.ic = cell_from_position(d, q.x, q.y);
    if (q.ic < 1 || q.ic > d.ncells) {
        pg_set_err(d, PG_DEVERR_CELL_RANGE);
        return false;
    }
    if (MODE == 0 && d.cell_fnum && q.ic != old_ic) {
        /* ref: :2110-2128 BOOL_RADIAL_WEIGHTING: a particle that changed cell is deleted or cloned with the ratio of the
         * cells' weights; one draw (stream kind 8, keyed by the slot) */
        double weight_ratio = d.cell_fnum[old_ic - 1] / d.cell_fnum[q.ic - 1];
        if (weight_ratio < 1.) {
            PgRng g;
            g.init(d.seed_local, 8, d.tID, (unsigned long long)slot, 0ULL);
            if (weight_ratio < g.rf()) return false;
        } else if (weight_ratio > 1.) {
            PgRng g;
            g.init(d.seed_local, 8, d.tID, (unsigned long long)slot, 0ULL);
            const double rfp = g.rf();
            const PgIe E = pg_ie_load(d.p.ie, slot);
            while ((weight_ratio - 1.) > rfp) {
                pg_append_particle(d, q.x, q.y, q.z, q.vx, q.vy, q.vz, E.erot, E.evib, q.sid, q.ic);
                weight_ratio = weight_ratio - 1.;
            }
        }
    }
    return true;
}

/* registers -> slot, and the particle's cell is counted for the counting sort: one fire-and-forget reduction
 * (RED.ADD, nothing waits for it); the arrival rank inside the cell is taken by the scatter pass, where the atomic's
 * latency hides behind HBM traffic (pgpu_api.cu scatter_kernel).  Every lane of `peers` must call this. */
__device__ __forceinline__ void store_particle(const PgDev& d, long long slot, const PState& q, int dirty, unsigned int peers,
                                               unsigned int* __restrict__ sh_count = nullptr) {
    /* one 256-bit store (the compiler splits a struct assignment with int members into three stores) */
    pg_store_sector(d.p.pos + slot, q.x, q.y, q.z, q.ic, q.sid);
    if (dirty & (DIRTY_V | DIRTY_S)) {
        PgVel V;
        V.vx = q.vx; V.vy = q.vy; V.vz = q.vz; V.sid = q.sid; V.pad = 0;
        d.p.vel[slot] = V;
    }
    if (dirty & DIRTY_E) pg_ie_store(d.p.ie, slot, q.erot, q.evib, d.status);
    /* nranks > 1: a particle that ends in a cell of another rank is an emigrant.  It is not counted (the sort of this
     * rank will not see it); its slot and destination are recorded for EXCHANGE instead. */
    const int lane = threadIdx.x & 31;
    if (d.emig_list) {
        /* emig_trust: every particle started the step in a cell of this rank, so the ownership flag carried along the
         * face walk decides, and only the emigrants (about 1 %) read CELL_PROCS; otherwise every particle does */
        int dest = d.rank;
        if (!d.emig_trust || q.foreign) dest = __ldg(&d.cell_procs[q.ic - 1]);
        const bool emig = dest != d.rank;
        const unsigned int em = __ballot_sync(peers, emig);
        if (em) {
            unsigned long long ebase = 0;
            const int el = __ffs(em) - 1;
            if (lane == el) ebase = atomicAdd(&d.xrow[PG_XROW_LEN], (unsigned long long)__popc(em));
            ebase = __shfl_sync(peers, ebase, el);
            if (emig) {
                const unsigned long long k = ebase + __popc(em & ((1u << lane) - 1u));
                if (k < (unsigned long long)d.emig_cap) d.emig_list[k] = (unsigned int)slot;
                else atomicAdd(&d.xrow[PG_XROW_OVF], 1ULL);
                /* one increment per (warp, destination): a rank has a neighbour or two, and a million per-lane
                 * increments of the same address queue up in L2 */
                const unsigned int same = __match_any_sync(em, dest);
                if (lane == __ffs(same) - 1) atomicAdd(&d.xrow[dest], (unsigned long long)__popc(same));
            }
            peers &= ~em;
            if (emig) return;
        }
    }
    /* The store is cell-sorted, so the lanes that finish together often share their cell, and coarse grids (the
     * one-cell heat bath, a 512-segment line) send millions of increments to a handful of counters: when the whole
     * group agrees on the cell, one lane adds the group's size (two warp votes); otherwise every lane adds one. */
    /* Coarse grids again (PG_COUNT_SMEM_CELLS or fewer cells): a 512-segment line takes ten million increments on
     * 512 addresses in sixteen L2 lines, and the L2 atomic unit serialises them (measured: the C4 ADVECT spent 3.2 ms
     * on 1e7 particles, 65 % of it waiting behind the reductions).  The block counts in shared memory and adds its
     * handful of non-zero counters once, at its end. */
    unsigned int* cnt = sh_count ? sh_count : d.cell_count;
    const int leader = __ffs(peers) - 1;
    const int ic0 = __shfl_sync(peers, q.ic, leader);
    if (__all_sync(peers, q.ic == ic0)) {
        if (lane == leader) atomicAdd(&cnt[ic0 - 1], (unsigned int)__popc(peers));
    } else {
        atomicAdd(&cnt[q.ic - 1], 1u);
    }
}

__device__ __forceinline__ void tombstone_particle(const PgDev& d, long long slot) {
    d.p.pos[slot].ic = PG_TOMBSTONE;
    atomicAdd(&d.status->n_removed, 1ULL);
}

/*
 * LEAN pass.  One warp works through `chunk` consecutive slots (64..1024, launch_advect); see the file header.
 */
template <int MODE, bool AXI, bool CSM>
__global__ void __launch_bounds__(128, PG_ADVECT_MINBLOCKS) advect_lean_kernel(const __grid_constant__ PgDev d, long long n,
                                                                               int chunk, int count_smem) {
    /* CSM (coarse grids, launch_advect): the block's cell histogram lives in shared memory, [count_smem] counters.
     * A template parameter, not a run-time branch: the barrier and the shared-memory path cost the fine-grid kernel
     * 0.3 ms at full size when they were merely skipped at run time. */
    extern __shared__ unsigned int sh_count_buf[];
    if (CSM) {
        for (int i = threadIdx.x; i < count_smem; i += blockDim.x) sh_count_buf[i] = 0u;
        __syncthreads();
    }
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long next = warp * chunk;
    if (!CSM && next >= n) return;
    const long long end = (next + chunk < n) ? next + chunk : n;

    PState q;
    long long slot = 0;
    bool have = false;
    int dirty = 0, hops = 0;
    int bhits[4] = {0, 0, 0, 0};
    PgRng g; /* never drawn from in the LEAN pass */
    g.init(0ULL, 0, 0, 0ULL, 0ULL);

    for (bool active = next < n; active;) { /* (a warp past the end of the store only joins the block's final flush) */
        /* refill: idle lanes take the next slots of the chunk, in order */
        unsigned int idle = __ballot_sync(FULL, !have);
        if (idle && next < end) {
            long long my = next + __popc(idle & ((1u << lane) - 1u));
            if (!have && my < end) {
                slot = my;
                if (my + 64 < end) { /* the records this lane is likely to take two refills from now */
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.p.pos + my + 64));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.p.vel + my + 64));
                }
                have = load_particle(d, slot, q, dirty);
                hops = 0;
                if (MODE == 0) { bhits[0] = bhits[1] = bhits[2] = bhits[3] = 0; }
            }
            next += __popc(idle);
        }
        if (!__any_sync(FULL, have)) {
            if (next >= end) break; /* (nothing pending: a trip that stored something refills or ends here first) */
            continue;
        }

        int st = IT_NONE, remove = 0;
        if (have) {
            if constexpr (MODE == 0) st = structured_iter<AXI, true>(d, g, q, remove, dirty, bhits, hops);
            else st = unstructured_iter<MODE, AXI, true>(d, g, q, remove, dirty, hops);
        }

        /* deferred to the FULL pass: untouched in the store, slot appended to the hit list */
        unsigned int dm = __ballot_sync(FULL, st == IT_DEFER);
        if (dm) {
            unsigned long long base = 0;
            int leader = __ffs(dm) - 1;
            if (lane == leader) base = atomicAdd(&d.status->n_hits, (unsigned long long)__popc(dm));
            base = __shfl_sync(FULL, base, leader);
            if (st == IT_DEFER) {
                d.hitlist[base + __popc(dm & ((1u << lane) - 1u))] = (unsigned int)slot;
                have = false;
            }
        }

        bool keep = false;
        if (st == IT_DONE) {
            have = false;
            keep = !remove && finish_particle<MODE>(d, q, slot);
            if (!keep) tombstone_particle(d, slot);
        }
        unsigned int fm = __ballot_sync(FULL, keep);
        if (keep) {
            if (MODE == 0) { /* the species cannot change in the LEAN pass */
                for (int b = 0; b < 4; b++)
                    if (bhits[b]) atomicAdd(&d.bcount[b + 4 * (q.sid - 1)], (unsigned long long)bhits[b]);
            }
            store_particle(d, slot, q, dirty, fm, CSM ? sh_count_buf : nullptr);
        }
    }
    if (CSM) {
        __syncthreads();
        for (int i = threadIdx.x; i < count_smem; i += blockDim.x) {
            const unsigned int v = sh_count_buf[i];
            if (v) atomicAdd(&d.cell_count[i], v);
        }
    }
}

/* FULL pass: one thread per hit-list entry, every boundary model enabled.  The length of the list is read from the
 * status block on the device (no host round trip between the two passes); the grid strides over it. */
template <int MODE, bool AXI>
__global__ void __launch_bounds__(128) advect_full_kernel(const __grid_constant__ PgDev d, long long, int) {
    const long long n = (long long)d.status->n_hits;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long tid = base + threadIdx.x;
        bool have = tid < n;
        long long slot = have ? (long long)d.hitlist[tid] : 0;
        PState q;
        int dirty = 0, remove = 0, hops = 0;
        if (have) have = load_particle(d, slot, q, dirty);
        bool keep = false;
        if (have) {
            PgRng g;
            g.init(d.seed_local, 1, d.tID, (unsigned long long)slot, 0ULL);
            int bhits[4] = {0, 0, 0, 0};
            while (q.dtrim > 0.) {
                if constexpr (MODE == 0) structured_iter<AXI, false>(d, g, q, remove, dirty, bhits, hops);
                else unstructured_iter<MODE, AXI, false>(d, g, q, remove, dirty, hops);
            }
            keep = !remove && finish_particle<MODE>(d, q, slot);
            if (!keep) tombstone_particle(d, slot);
        }
        unsigned int fm = __ballot_sync(0xffffffffu, keep);
        if (keep) store_particle(d, slot, q, dirty, fm);
    }
}

template <bool LEAN>
int launch_advect(pgpu_ctx* c, long long n) {
    if (n <= 0) return PGPU_OK;
    const PgDev& d = c->dev;
    /* slots one warp of the LEAN pass works through: long runs of consecutive (cell-sorted) slots keep the cell
     * records a warp needs in L1 (measured at 3.1e7 particles: 256 -> 1024 slots, move -2.7 %, sort -2.6 %), but the
     * grid must still fill the machine several times over */
    int chunk = PG_ADVECT_CHUNK_MAX;
    while (chunk > PG_ADVECT_CHUNK_MIN && n / (4LL * chunk) < 148LL * PG_ADVECT_MINBLOCKS * 8) chunk >>= 1;
    if (c->advect_chunk_override > 0) chunk = c->advect_chunk_override; /* tests: the bench-size chunk on a small store */
    const long long per_block = LEAN ? 4LL * chunk : 128LL;
    unsigned int blocks = (unsigned int)((n + per_block - 1) / per_block);
    if (!LEAN && blocks > 148u * 16u) blocks = 148u * 16u; /* the FULL pass strides over the device-side list length */
    bool axi = d.axi != 0;
    const int csm = (LEAN && d.ncells <= PG_COUNT_SMEM_CELLS) ? d.ncells : 0; /* block-level histogram: store_particle */
#define PG_LAUNCH(MODE, AXI)                                                                   \
    do {                                                                                       \
        if (LEAN && csm) advect_lean_kernel<MODE, AXI, true><<<blocks, 128, csm * sizeof(unsigned int), c->stream>>>(d, n, chunk, csm); \
        else if (LEAN) advect_lean_kernel<MODE, AXI, false><<<blocks, 128, 0, c->stream>>>(d, n, chunk, 0); \
        else advect_full_kernel<MODE, AXI><<<blocks, 128, 0, c->stream>>>(d, n, chunk);        \
    } while (0)
    if (d.grid_type != PGPU_UNSTRUCTURED) {
        if (axi) PG_LAUNCH(0, true);
        else PG_LAUNCH(0, false);
    } else if (d.dims == 1) {
        PG_LAUNCH(1, false);
    } else if (d.dims == 2) {
        if (axi) PG_LAUNCH(2, true);
        else PG_LAUNCH(2, false);
    } else {
        PG_LAUNCH(3, false);
    }
#undef PG_LAUNCH
    c->launches++;
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}

} /* namespace */

/*
 * ADVECT for all slots [0, np).  Asynchronous.
 */
int pg_advect_impl(pgpu_ctx* c, int tID) {
    int rc;
    c->dev.tID = tID;
    c->dev.tally_on = 0;
    if (c->cfg.dump_bound_avg_every > 0 && c->dev.tallies)
        c->dev.tally_on = ((tID - c->cfg.dump_bound_start) % c->cfg.dump_bound_avg_every) == 0;
    /* (with several ranks an empty rank still goes through the emigrant bookkeeping: every rank must take the same
     *  branch of the EXCHANGE that follows) */
    if (c->np == 0 && !(c->cfg.nranks > 1 && c->dev.cell_procs && c->d_xcount && c->d_emig_list)) return PGPU_OK;
    if ((rc = pg_resolve_vel(c))) return rc;

    /* zero the per-call counters and the histogram */
    {
        /* (several ranks, strays shipped: only the span of this rank's own cells can be counted in -- pg_cell_range) */
        long long lo, hi;
        pg_cell_range(c, &lo, &hi);
        PG_CUDA(c, cudaMemsetAsync(c->d_cell_count + lo, 0, sizeof(unsigned int) * (size_t)(hi - lo), c->stream));
        c->hist_lo = lo;
        c->hist_hi = hi;
    }
    PG_CUDA(c, cudaMemsetAsync(c->d_bcount, 0, sizeof(unsigned long long) * 4 * (size_t)c->ns, c->stream));
    if (c->nw) PG_CUDA(c, cudaMemsetAsync(c->d_wcount, 0, sizeof(unsigned long long) * (size_t)(c->nw * c->ns), c->stream));

    /* nranks > 1: the epilogue records the emigrants (slot list + count per destination) for EXCHANGE */
    c->emig_valid = false;
    const bool track = c->cfg.nranks > 1 && c->dev.cell_procs && c->d_xcount && c->d_emig_list;
    if (track) {
        /* room for the immigrants of the EXCHANGE that follows, made BEFORE the emigrant list is recorded (a
         * re-allocation drops it) */
        if ((rc = pg_ensure_capacity(c, c->np + pg_exchange_recv_slots(c)))) return rc;
        pg_refresh_dev(c);
        PG_CUDA(c, cudaMemsetAsync(c->d_xcount, 0, sizeof(unsigned long long) * PG_XROW, c->stream));
        c->dev.emig_trust = (c->owned_clean && c->cfg.grid_type == PGPU_UNSTRUCTURED && !c->cfg.axi) ? 1 : 0;
        c->dev.emig_list = c->d_emig_list;
        c->dev.emig_cap = c->emig_cap;
        c->dev.xrow = c->d_xcount;
    }

    if ((rc = launch_advect<true>(c, c->np))) return rc;
    /* the FULL pass reads the hit-list length on the device: no host round trip inside ADVECT */
    if ((rc = launch_advect<false>(c, c->np))) return rc;

    c->dev.emig_list = nullptr;
    c->dev.xrow = nullptr;
    c->emig_valid = track;
    c->has_dtrim = false; /* DTRIM = DT for every particle after ADVECT, ref: src/timecycle.f90:2070 */
    c->sorted = false;
    c->hist_valid = true;
    c->maybe_tombstones = true;
    /* several ranks: particles may now sit in cells of other ranks until EXCHANGE ships them (it sets the flag again).
     * Until then neither the next ADVECT may trust the ownership flags of the hop codes nor a sort restrict itself to
     * this rank's own cell range (the histogram above, which leaves the emigrants out, was counted over that range). */
    if (c->cfg.nranks > 1) c->owned_clean = false;
    return PGPU_OK;
}
