/*
 * pgpu_inject.cu -- particle creation on the device: INITIAL_SEED (src/initialization.f90:2017-2233),
 * VOLUME_INJECT (src/timecycle.f90:553-737), BOUNDARIES_EMIT (src/timecycle.f90:368-547).
 *
 * On the host these are the step right before the hot path: with open boundaries every timestep ends in a
 * host-side RNG loop and an H2D append.  Here one thread owns one loop body of the reference -- a (cell, task,
 * mixture component) for the volume sources -- and one warp an (emit task, component) of the boundary source; it draws
 * the number of particles (RANDINT / the fractional-part rule) and creates them straight in the append region of the store.
 * Random numbers: Philox streams of kind 5 (seed), 6 (volume inject), keyed e0 = cell, e1 = task*256 + component, and 7
 * (emit), keyed e0 = emit task, e1 = component for the count and e1 = component + 256 (1 + ip) for particle ip;
 * consumed in the order of the reference's loop body; the oracle's rng_mode 1 follows the
 * same contract, so the created particles can be compared value for value (as a set: creation order is the order of
 * the append cursor).  COLOCATED_ELECTRONS, BOOL_INJECT_FROM_FILE and the kappa distribution are not implemented.
 */
#include <algorithm>
#include <cmath>

#include "pgpu_internal.cuh"

namespace {

struct PgSrcTask { double rate, ux, uy, uz, tx, ty, tz, trot, tvib; int mix, pad; };
struct PgEmitTask { double ux, uy, uz, u_norm, ttra, trot, tvib; int mix, ic, iv1, iv2, iface, nfs_off; };

struct PgInject {
    const int* mix_off; const int* mix_sp; const double* mix_frac;
    const double* spwt;          /* may be null (all 1) */
    const double* node_coords;   /* [nn][3] */
    const unsigned char* cell_solid; /* [nc] or null */
    const double* face_tang1; const double* face_tang2;
};

/* ref: src/tools.f90:130-187 MAXWELL */
__device__ __forceinline__ void pg_maxwell(PgRng& g, double ux, double uy, double uz, double tx, double ty, double tz, double m,
                                           double* v3) {
    const double PI2 = 2. * PG_PI;
    const double tt[3] = {tx, ty, tz}, uu[3] = {ux, uy, uz};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double vel = 0.;
        if (tt[i] != 0.0) {
            double r = g.rf();
            double teta = PI2 * r;
            double beta = 1. / sqrt(2. * PG_KB / m * tt[i]);
            double r1 = g.rf();
            while (r1 < 1.0e-13) r1 = g.rf();
            double ro = sqrt(-log(r1)) / beta;
            vel = ro * sin(teta);
        }
        v3[i] = uu[i] + vel;
    }
}

__device__ __forceinline__ void append(const PgDev& d, double x, double y, double z, const double* v, double erot, double evib, int sid,
                                       int ic, double dtrim) {
    unsigned long long k = atomicAdd(&d.status->n_appended, 1ULL);
    long long o = (long long)d.status->np_cur + (long long)k;
    if (o >= d.capacity) { pg_set_err(d, PG_DEVERR_CAPACITY); return; }
    PgPos P;
    P.x = x; P.y = y; P.z = z; P.ic = ic; P.sid = sid;
    PgVel V;
    V.vx = v[0]; V.vy = v[1]; V.vz = v[2]; V.sid = sid; V.pad = 0;
    d.p.pos[o] = P;
    d.p.vel[o] = V;
    pg_ie_store(d.p.ie, o, erot, evib, d.status);
    if (d.p.dtrim) d.p.dtrim[o] = dtrim;
    else if (dtrim != d.dt) d.status->need_dtrim = 1;
    if (d.p.id) d.p.id[o] = (long long)d.rank + ((d.id_base + (long long)k) << 8);
    if (erot != 0. || evib != 0.) d.status->flag_ie = 1;
}

__device__ __forceinline__ int randint(PgRng& g, double x) { /* ref: src/tools.f90:1304-1311 */
    int r = (int)x;
    if (x - r > g.rf()) r = r + 1;
    return r;
}

/* kind 5: INITIAL_SEED (owned fluid cells, rate = NRHO); kind 6: VOLUME_INJECT (all cells, rate = DT*NRHODOT / ranks).
 * One thread per (cell, slot) where slot enumerates the (task, component) pairs. */
template <int DIMS>
__global__ void __launch_bounds__(128) cell_source_kernel(const __grid_constant__ PgDev d, const __grid_constant__ PgInject q, int kind,
                                                          const PgSrcTask* __restrict__ tasks, int ntask, int nslot,
                                                          const int* __restrict__ slot_task, const int* __restrict__ slot_comp) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)d.nc * nslot) return;
    const int ic0 = (int)(t % d.nc), slot = (int)(t / d.nc);
    const int it = slot_task[slot], icomp = slot_comp[slot];
    const PgSrcTask tk = tasks[it];
    if (kind == 5) {
        if (d.cell_procs && d.nranks > 1 && d.cell_procs[ic0] != d.rank) return;
        if (q.cell_solid && q.cell_solid[ic0]) return;
    }
    const int i = q.mix_off[tk.mix] + icomp;
    const int sid = q.mix_sp[i];
    const double frac = q.mix_frac[i], m = d.mass[sid - 1];
    const double spwt = q.spwt ? q.spwt[sid - 1] : 1.0;
    const double rate = (kind == 5) ? tk.rate : d.dt * tk.rate;
    PgRng g;
    g.init(d.seed_local, kind, d.tID, (unsigned long long)ic0, (unsigned long long)(it * 256 + icomp));
    const double vol = d.cell_vol[ic0];
    int np_init;
    if (kind == 5) np_init = randint(g, rate / (d.fnum * spwt) * vol * frac);
    else np_init = randint(g, rate / (d.fnum * spwt) * vol * frac / (d.nranks > 0 ? d.nranks : 1));
    if (np_init == 0) return;
    double V[DIMS + 1][3];
#pragma unroll
    for (int k = 0; k <= DIMS; k++) {
        const int v = d.cell_nodes[(size_t)(DIMS + 1) * ic0 + k] - 1;
        for (int a = 0; a < 3; a++) V[k][a] = q.node_coords[3 * (size_t)v + a];
    }
    for (int ip = 0; ip < np_init; ip++) {
        double S = g.rf(), T = g.rf(), U = g.rf(), xp, yp, zp;
        if (S + T > 1) { S = 1 - S; T = 1 - T; }
        if (DIMS == 1) {
            xp = V[0][0] + (V[1][0] - V[0][0]) * S;
            yp = d.ymin + (d.ymax - d.ymin) * T;
            zp = d.zmin + (d.zmax - d.zmin) * U;
        } else if (DIMS == 2) {
            xp = V[0][0] + (V[1][0] - V[0][0]) * S + (V[DIMS >= 2 ? 2 : 0][0] - V[0][0]) * T;
            yp = V[0][1] + (V[1][1] - V[0][1]) * S + (V[DIMS >= 2 ? 2 : 0][1] - V[0][1]) * T;
            zp = d.axi ? 0.0 : d.zmin + (d.zmax - d.zmin) * U;
        } else {
            double P, Q, R;
            if (S + T + U <= 1) { P = S; Q = T; R = U; }
            else if (T + U > 1) { P = S; Q = 1 - U; R = 1 - S - T; }
            else { P = 1 - T - U; Q = T; R = S + T + U - 1; }
            constexpr int K2 = DIMS >= 2 ? 2 : 0, K3 = DIMS >= 3 ? 3 : 0;
            xp = V[0][0] + (V[1][0] - V[0][0]) * P + (V[K2][0] - V[0][0]) * Q + (V[K3][0] - V[0][0]) * R;
            yp = V[0][1] + (V[1][1] - V[0][1]) * P + (V[K2][1] - V[0][1]) * Q + (V[K3][1] - V[0][1]) * R;
            zp = V[0][2] + (V[1][2] - V[0][2]) * P + (V[K2][2] - V[0][2]) * Q + (V[K3][2] - V[0][2]) * R;
        }
        double v3[3];
        pg_maxwell(g, tk.ux, tk.uy, tk.uz, tk.tx, tk.ty, tk.tz, m, v3);
        const double erot = pg_internal_energy(g, d.rotdof[sid - 1], tk.trot);
        const double evib = pg_internal_energy(g, d.vibdof[sid - 1], tk.tvib);
        append(d, xp, yp, zp, v3, erot, evib, sid, ic0 + 1, d.dt);
    }
}

/* rectilinear grids: the reference seeds the whole domain in one loop; particle ip has its own stream (e0 = 1 + ip) */
__device__ __forceinline__ int cell_from_position_uniform(const PgDev& d, double xp, double yp) {
    double dx = (d.xmax - d.xmin) / d.nx;
    double dy = (d.ymax - d.ymin) / d.ny;
    int xcell = (int)((xp - d.xmin) / dx);
    int ycell = (int)((yp - d.ymin) / dy);
    return xcell + d.nx * ycell + 1;
}
__global__ void __launch_bounds__(128) domain_source_kernel(const __grid_constant__ PgDev d, const __grid_constant__ PgInject q, int kind,
                                                            PgSrcTask tk, int it, int icomp, int np_init) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= np_init) return;
    const int i = q.mix_off[tk.mix] + icomp;
    const int sid = q.mix_sp[i];
    const double m = d.mass[sid - 1];
    PgRng g;
    g.init(d.seed_local, kind, d.tID, (unsigned long long)(1 + ip), (unsigned long long)(it * 256 + icomp));
    double xp = d.xmin + (d.xmax - d.xmin) * g.rf(), yp, zp;
    if (d.axi && !d.cell_fnum) { /* IF (AXI .AND. (.NOT. BOOL_RADIAL_WEIGHTING)) */
        yp = sqrt(d.ymin * d.ymin + g.rf() * (d.ymax * d.ymax - d.ymin * d.ymin));
        zp = 0;
    } else {
        yp = d.ymin + (d.ymax - d.ymin) * g.rf();
        zp = d.zmin + (d.zmax - d.zmin) * g.rf();
    }
    double v3[3];
    pg_maxwell(g, tk.ux, tk.uy, tk.uz, tk.tx, tk.ty, tk.tz, m, v3);
    const double erot = pg_internal_energy(g, d.rotdof[sid - 1], tk.trot);
    const double evib = pg_internal_energy(g, d.vibdof[sid - 1], tk.tvib);
    int cid;
    if (d.grid_type == PGPU_RECTILINEAR_UNIFORM) cid = cell_from_position_uniform(d, xp, yp);
    else {
        /* BINARY_SEARCH on XCOORD / YCOORD (src/tools.f90:1499-1535), as CELL_FROM_POSITION does */
        int xc = -1, yc = -1;
        for (int k = 0; k + 1 < d.nxp; k++) if (d.xcoord[k] <= xp && xp < d.xcoord[k + 1]) { xc = k + 1; break; }
        for (int k = 0; k + 1 < d.nyp; k++) if (d.ycoord[k] <= yp && yp < d.ycoord[k + 1]) { yc = k + 1; break; }
        cid = (xc < 0 || yc < 0) ? -1 : xc + d.nx * (yc - 1);
    }
    if (cid < 1 || cid > d.ncells) { pg_set_err(d, PG_DEVERR_CELL_RANGE); return; }
    append(d, xp, yp, zp, v3, erot, evib, sid, cid, d.dt);
}

/* BOUNDARIES_EMIT: one WARP per (emit task, component).  The count NFS (with its rounding draw) comes from the stream
 * (task, component); particle ip then has its own stream (task, component + 256 (1 + ip)), so the lanes of the warp
 * create the particles of the task side by side (a face of the C2 free-stream boundary emits ~55 particles per step and
 * component: one thread per task walked them one after the other, 0.93 ms per step for 800 tasks) */
template <int DIMS>
__global__ void __launch_bounds__(128) emit_kernel(const __grid_constant__ PgDev d, const __grid_constant__ PgInject q,
                                                   const PgEmitTask* __restrict__ tasks, const double* __restrict__ nfs_all, int nslot,
                                                   const int* __restrict__ slot_task, const int* __restrict__ slot_comp) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= nslot) return;
    const int it = slot_task[t], icomp = slot_comp[t];
    const PgEmitTask tk = tasks[it];
    const int ic = tk.ic, iface = tk.iface;
    const int IND[4][3] = {{1, 3, 2}, {1, 2, 4}, {2, 3, 4}, {1, 4, 3}}; /* src/grid_and_partition.f90:1932-1935 */
    double fn[3], t1[3], t2[3], x1 = 0, y1 = 0, x2 = 0, y2 = 0, V1[3] = {0, 0, 0}, V2[3] = {0, 0, 0}, V3[3] = {0, 0, 0};
    const size_t f = (size_t)(DIMS + 1) * (ic - 1) + (iface - 1);
    for (int a = 0; a < 3; a++) fn[a] = d.face_exact[8 * f + a]; /* the FACE_NORMAL / EDGE_NORMAL operand of the face walk */
    if (DIMS == 1) {
        x1 = q.node_coords[3 * (size_t)(tk.iv1 - 1)]; x2 = q.node_coords[3 * (size_t)(tk.iv2 - 1)];
        t1[0] = 0; t1[1] = fn[0]; t1[2] = 0; t2[0] = 0; t2[1] = 0; t2[2] = 1;
    } else if (DIMS == 2) {
        x1 = q.node_coords[3 * (size_t)(tk.iv1 - 1)]; y1 = q.node_coords[3 * (size_t)(tk.iv1 - 1) + 1];
        x2 = q.node_coords[3 * (size_t)(tk.iv2 - 1)]; y2 = q.node_coords[3 * (size_t)(tk.iv2 - 1) + 1];
        t1[0] = -fn[1]; t1[1] = fn[0]; t1[2] = 0; t2[0] = 0; t2[1] = 0; t2[2] = 1;
    } else {
        for (int a = 0; a < 3; a++) {
            t1[a] = q.face_tang1[3 * f + a];
            t2[a] = q.face_tang2[3 * f + a];
            V1[a] = q.node_coords[3 * (size_t)(d.cell_nodes[4 * (size_t)(ic - 1) + IND[iface - 1][0] - 1] - 1) + a];
            V2[a] = q.node_coords[3 * (size_t)(d.cell_nodes[4 * (size_t)(ic - 1) + IND[iface - 1][1] - 1] - 1) + a];
            V3[a] = q.node_coords[3 * (size_t)(d.cell_nodes[4 * (size_t)(ic - 1) + IND[iface - 1][2] - 1] - 1) + a];
        }
    }
    const int i = q.mix_off[tk.mix] + icomp;
    const int sid = q.mix_sp[i];
    const double m = d.mass[sid - 1];
    PgRng g;
    g.init(d.seed_local, 7, d.tID, (unsigned long long)it, (unsigned long long)icomp);
    const double nfs_x = nfs_all[tk.nfs_off + icomp];
    int nfs = (int)floor(nfs_x);
    if (nfs_x - (double)nfs >= g.rf()) nfs = nfs + 1;
    double beta = 0;
    if (tk.ttra != 0) beta = 1. / sqrt(2. * PG_KB / m * tk.ttra);
    for (int ip = lane; ip < nfs; ip += 32) {
        g.init(d.seed_local, 7, d.tID, (unsigned long long)it, (unsigned long long)(icomp + ((ip + 1) << 8)));
        double v3[3];
        pg_maxwell(g, 0., 0., 0., tk.ttra, tk.ttra, tk.ttra, m, v3);
        const double erot = pg_internal_energy(g, d.rotdof[sid - 1], tk.trot);
        const double evib = pg_internal_energy(g, d.vibdof[sid - 1], tk.tvib);
        double v_norm = 0;
        if (tk.ttra != 0) v_norm = pg_flx(g, tk.u_norm * beta, tk.ttra, m);
        double v[3];
        v[0] = tk.ux - v_norm * fn[0] - v3[1] * t1[0] - v3[2] * t2[0];
        v[1] = tk.uy - v_norm * fn[1] - v3[1] * t1[1] - v3[2] * t2[1];
        v[2] = tk.uz - v_norm * fn[2] - v3[1] * t1[2] - v3[2] * t2[2];
        const double dtfrac = g.rf() * d.dt;
        double x, y, z;
        if (DIMS == 1) {
            double R = g.rf();
            x = x1 + R * (x2 - x1);
            y = d.ymin + (d.ymax - d.ymin) * g.rf();
            z = d.zmin + (d.zmax - d.zmin) * g.rf();
        } else if (DIMS == 2) {
            double R = g.rf();
            if (d.axi) {
                if (y1 == y2) { y = y1; x = x1 + R * (x2 - x1); z = 0; }
                else if (y1 < y2) { y = sqrt(y1 * y1 + R * (y2 * y2 - y1 * y1)); x = x1 + (y - y1) / (y2 - y1) * (x2 - x1); z = 0; }
                else { y = sqrt(y2 * y2 + R * (y1 * y1 - y2 * y2)); x = x2 + (y - y2) / (y1 - y2) * (x1 - x2); z = 0; }
            } else {
                x = x1 + R * (x2 - x1);
                y = y1 + R * (y2 - y1);
                z = d.zmin + (d.zmax - d.zmin) * g.rf();
            }
        } else {
            double S = g.rf(), T = g.rf(), P, Q;
            if (S > 1 - T) { Q = 1 - T; P = 1 - S; }
            else { Q = T; P = S; }
            x = V1[0] + (V2[0] - V1[0]) * P + (V3[0] - V1[0]) * Q;
            y = V1[1] + (V2[1] - V1[1]) * P + (V3[1] - V1[1]) * Q;
            z = V1[2] + (V2[2] - V1[2]) * P + (V3[2] - V1[2]) * Q;
        }
        append(d, x, y, z, v, erot, evib, sid, ic, dtfrac);
    }
}

/* host copy of the Philox stream, for the one draw the host needs (the domain-wide RANDINT of rectilinear seeding) */
void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
double first_rf_host(unsigned long long seed, int kind, int tID, unsigned long long e0, unsigned int e1) {
    uint32_t o[4];
    philox_host(0u, (uint32_t)e0, (uint32_t)(e0 >> 32), e1, (uint32_t)seed ^ ((uint32_t)kind * 0x632BE5ABu), (uint32_t)(seed >> 32) ^ (uint32_t)tID, o);
    unsigned long long u = ((unsigned long long)o[1] << 32) | o[0];
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

template <class T>
int upload(pgpu_ctx* c, T** slot, const std::vector<T>& h) {
    if (*slot) cudaFree(*slot);
    *slot = nullptr;
    if (h.empty()) return PGPU_OK;
    PG_CUDA(c, cudaMalloc((void**)slot, sizeof(T) * h.size()));
    PG_CUDA(c, cudaMemcpy(*slot, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return PGPU_OK;
}

} /* namespace */

struct PgInjectState {
    std::vector<int> mix_off, mix_sp;
    std::vector<double> mix_frac, spwt;
    std::vector<PgSrcTask> seed, vinj;
    std::vector<PgEmitTask> emit;
    std::vector<double> emit_nfs;
    int *d_mix_off = nullptr, *d_mix_sp = nullptr;
    double *d_mix_frac = nullptr, *d_spwt = nullptr, *d_emit_nfs = nullptr;
    PgSrcTask *d_seed = nullptr, *d_vinj = nullptr;
    PgEmitTask* d_emit = nullptr;
    int *d_slot_task[3] = {nullptr, nullptr, nullptr}, *d_slot_comp[3] = {nullptr, nullptr, nullptr};
    int nslot[3] = {0, 0, 0};
    double expected[3] = {0, 0, 0}; /* expected particles per call (capacity planning) */
};

static PgInjectState* inj(pgpu_ctx* c) {
    if (!c->inject) c->inject = new PgInjectState();
    return (PgInjectState*)c->inject;
}
void pg_inject_destroy(pgpu_ctx* c) {
    if (!c->inject) return;
    PgInjectState* s = (PgInjectState*)c->inject;
    cudaFree(s->d_mix_off); cudaFree(s->d_mix_sp); cudaFree(s->d_mix_frac); cudaFree(s->d_spwt); cudaFree(s->d_emit_nfs);
    cudaFree(s->d_seed); cudaFree(s->d_vinj); cudaFree(s->d_emit);
    for (int k = 0; k < 3; k++) { cudaFree(s->d_slot_task[k]); cudaFree(s->d_slot_comp[k]); }
    delete s;
    c->inject = nullptr;
}

extern "C" int pgpu_set_species_weights(pgpu_ctx* c, int ns, const double* spwt) {
    if (!c || !spwt || ns < 1) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    PgInjectState* s = inj(c);
    s->spwt.assign(spwt, spwt + ns);
    return upload(c, &s->d_spwt, s->spwt);
}

/* MIXTURES(:): components of mixture k are comp_sp / comp_molfrac [comp_off[k], comp_off[k+1]) (0-based mixture ids) */
extern "C" int pgpu_set_mixtures(pgpu_ctx* c, int nmix, const int32_t* comp_off, const int32_t* comp_sp, const double* comp_molfrac) {
    if (!c || nmix < 0 || (nmix > 0 && (!comp_off || !comp_sp || !comp_molfrac))) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    PgInjectState* s = inj(c);
    s->mix_off.assign(comp_off, comp_off + nmix + 1);
    s->mix_sp.assign(comp_sp, comp_sp + comp_off[nmix]);
    s->mix_frac.assign(comp_molfrac, comp_molfrac + comp_off[nmix]);
    for (int v : s->mix_sp)
        if (v < 1 || (c->ns && v > c->ns)) PG_FAIL(c, PGPU_ERR_INVALID, "set_mixtures: species id %d outside 1..%lld", v, c->ns);
    int rc;
    if ((rc = upload(c, &s->d_mix_off, s->mix_off)) || (rc = upload(c, &s->d_mix_sp, s->mix_sp)) || (rc = upload(c, &s->d_mix_frac, s->mix_frac))) return rc;
    return PGPU_OK;
}

static int set_slots(pgpu_ctx* c, PgInjectState* s, int which, const std::vector<int>& task_mix) {
    std::vector<int> st, sc;
    for (size_t it = 0; it < task_mix.size(); it++) {
        int mix = task_mix[it];
        if (mix < 0 || mix + 1 >= (int)s->mix_off.size()) PG_FAIL(c, PGPU_ERR_INVALID, "task %zu: mixture %d not defined (pgpu_set_mixtures first)", it, mix);
        int ncomp = s->mix_off[mix + 1] - s->mix_off[mix];
        if (ncomp > 256) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "mixtures of more than 256 components");
        for (int k = 0; k < ncomp; k++) { st.push_back((int)it); sc.push_back(k); }
    }
    s->nslot[which] = (int)st.size();
    int rc;
    if ((rc = upload(c, &s->d_slot_task[which], st)) || (rc = upload(c, &s->d_slot_comp[which], sc))) return rc;
    return PGPU_OK;
}

static int set_src(pgpu_ctx* c, int which, int n, const int32_t* mix_id, const double* rate, const double* ux, const double* uy, const double* uz,
                   const double* tx, const double* ty, const double* tz, const double* trot, const double* tvib) {
    if (!c || n < 0) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    PgInjectState* s = inj(c);
    std::vector<PgSrcTask>& v = which == 0 ? s->seed : s->vinj;
    v.resize((size_t)n);
    std::vector<int> mix((size_t)n);
    for (int i = 0; i < n; i++) {
        PgSrcTask t;
        t.rate = rate[i]; t.ux = ux[i]; t.uy = uy[i]; t.uz = uz[i]; t.tx = tx[i]; t.ty = ty[i]; t.tz = tz[i]; t.trot = trot[i]; t.tvib = tvib[i];
        t.mix = mix_id[i]; t.pad = 0;
        v[i] = t;
        mix[i] = mix_id[i];
    }
    int rc;
    if ((rc = set_slots(c, s, which, mix))) return rc;
    return upload(c, which == 0 ? &s->d_seed : &s->d_vinj, v);
}
extern "C" int pgpu_set_initial_particles_tasks(pgpu_ctx* c, int n, const int32_t* mix_id, const double* nrho, const double* ux,
                                                const double* uy, const double* uz, const double* ttrax, const double* ttray,
                                                const double* ttraz, const double* trot, const double* tvib) {
    return set_src(c, 0, n, mix_id, nrho, ux, uy, uz, ttrax, ttray, ttraz, trot, tvib);
}
extern "C" int pgpu_set_volume_inject_tasks(pgpu_ctx* c, int n, const int32_t* mix_id, const double* nrhodot, const double* ux,
                                            const double* uy, const double* uz, const double* ttrax, const double* ttray,
                                            const double* ttraz, const double* trot, const double* tvib) {
    return set_src(c, 1, n, mix_id, nrhodot, ux, uy, uz, ttrax, ttray, ttraz, trot, tvib);
}
extern "C" int pgpu_set_emit_tasks(pgpu_ctx* c, int n, const int32_t* mix_id, const int32_t* ic, const int32_t* iv1, const int32_t* iv2,
                                   const int32_t* iface, const int32_t* nfs_off, const double* nfs, const double* ux, const double* uy,
                                   const double* uz, const double* u_norm, const double* ttra, const double* trot, const double* tvib) {
    if (!c || n < 0) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    PgInjectState* s = inj(c);
    s->emit.resize((size_t)n);
    std::vector<int> mix((size_t)n);
    double expect = 0;
    for (int i = 0; i < n; i++) {
        PgEmitTask t;
        t.ux = ux[i]; t.uy = uy[i]; t.uz = uz[i]; t.u_norm = u_norm[i]; t.ttra = ttra[i]; t.trot = trot[i]; t.tvib = tvib[i];
        t.mix = mix_id[i]; t.ic = ic[i]; t.iv1 = iv1 ? iv1[i] : 1; t.iv2 = iv2 ? iv2[i] : 1; t.iface = iface[i]; t.nfs_off = nfs_off[i];
        if (c->mesh_set && (t.ic < 1 || t.ic > c->nc || t.iface < 1 || t.iface > c->mesh_dims + 1))
            PG_FAIL(c, PGPU_ERR_INVALID, "set_emit_tasks: task %d has cell %d face %d", i, t.ic, t.iface);
        s->emit[i] = t;
        mix[i] = mix_id[i];
    }
    const int total = n > 0 ? nfs_off[n] : 0;
    s->emit_nfs.assign(nfs, nfs + total);
    for (double x : s->emit_nfs) expect += x + 1.0;
    s->expected[2] = expect;
    int rc;
    if ((rc = set_slots(c, s, 2, mix))) return rc;
    if ((rc = upload(c, &s->d_emit_nfs, s->emit_nfs))) return rc;
    return upload(c, &s->d_emit, s->emit);
}

static PgInject make_inject(pgpu_ctx* c, PgInjectState* s) {
    PgInject q;
    q.mix_off = s->d_mix_off; q.mix_sp = s->d_mix_sp; q.mix_frac = s->d_mix_frac;
    q.spwt = s->d_spwt;
    q.node_coords = c->d_node_coords;
    q.cell_solid = c->d_cell_solid;
    q.face_tang1 = c->dev.face_tang1; q.face_tang2 = c->dev.face_tang2;
    return q;
}

/* make room for `expect` new particles (plus fluctuations), keep the DTRIM plane meaningful, start the call */
static int begin_create(pgpu_ctx* c, double expect, bool own_dtrim) {
    int rc;
    const long long room = (long long)(expect + 8.0 * std::sqrt(expect + 1.0) + 4096.0);
    if (c->capacity == 0) {
        long long cap = c->cfg.capacity > 0 ? (long long)c->cfg.capacity : (long long)(1.5 * (double)room);
        if ((rc = pg_ensure_capacity(c, std::max<long long>(cap, c->np + room)))) return rc;
    } else if ((rc = pg_ensure_capacity(c, c->np + room))) {
        return rc;
    }
    if ((rc = pg_resolve_vel(c))) return rc;
    if (c->any_internal_dof && (rc = pg_need_ie(c))) return rc;
    if (own_dtrim && !c->has_dtrim) {
        /* emitted particles carry DTRIM = rf()*DT (src/timecycle.f90:463): the plane must hold DT for everybody else */
        if ((rc = pg_fill_dtrim(c))) return rc;
        c->has_dtrim = true;
    }
    if ((rc = pg_reset_status(c))) return rc;
    pg_refresh_dev(c);
    return PGPU_OK;
}
static int end_create(pgpu_ctx* c, const char* what) {
    c->sorted = false;
    c->hist_valid = false;
    c->emig_valid = false;
    int rc = pg_finish(c, what);
    if (rc) return rc;
    if (c->h_status->flag_ie) c->ie_seen = true;
    c->id_counter = std::max(c->id_counter, c->np);
    return PGPU_OK;
}

static int cell_source(pgpu_ctx* c, int which, int kind, int tID, const char* what) {
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    PgInjectState* s = inj(c);
    const std::vector<PgSrcTask>& tasks = which == 0 ? s->seed : s->vinj;
    if (tasks.empty()) return PGPU_OK;
    if (s->mix_off.empty()) PG_FAIL(c, PGPU_ERR_STATE, "%s: pgpu_set_mixtures has not been called", what);
    const bool unstructured = c->cfg.grid_type == PGPU_UNSTRUCTURED;
    const int nranks = c->cfg.nranks > 0 ? c->cfg.nranks : 1;
    /* expected number of particles (capacity planning): rate * volume * molfrac / (Fnum * SPWT) */
    double vol_total = unstructured ? c->volume_total : 0.0;
    if (!unstructured) {
        if (c->cfg.axi) vol_total = 0.5 * (c->cfg.xmax - c->cfg.xmin) * (c->cfg.ymax * c->cfg.ymax - c->cfg.ymin * c->cfg.ymin) * (c->cfg.zmax - c->cfg.zmin);
        else vol_total = (c->cfg.xmax - c->cfg.xmin) * (c->cfg.ymax - c->cfg.ymin) * (c->cfg.zmax - c->cfg.zmin);
    }
    double expect = 0;
    for (size_t it = 0; it < tasks.size(); it++)
        for (int i = s->mix_off[tasks[it].mix]; i < s->mix_off[tasks[it].mix + 1]; i++) {
            const double spwt = s->spwt.empty() ? 1.0 : s->spwt[s->mix_sp[i] - 1];
            const double rate = (kind == 5) ? tasks[it].rate : c->cfg.dt * tasks[it].rate;
            double e = rate / (c->cfg.fnum * spwt) * vol_total * s->mix_frac[i];
            if (kind == 6 || !unstructured) e /= nranks;
            expect += e + (unstructured ? 0.5 * c->nc : 1.0);
        }
    if ((rc = begin_create(c, expect, false))) return rc;
    c->dev.tID = tID;
    PgInject q = make_inject(c, s);
    const PgSrcTask* dt = which == 0 ? s->d_seed : s->d_vinj;
    if (unstructured) {
        const long long nthreads = (long long)c->nc * s->nslot[which];
        const unsigned int blocks = (unsigned int)((nthreads + 127) / 128);
        if (c->mesh_dims == 1) cell_source_kernel<1><<<blocks, 128, 0, c->stream>>>(c->dev, q, kind, dt, (int)tasks.size(), s->nslot[which], s->d_slot_task[which], s->d_slot_comp[which]);
        else if (c->mesh_dims == 2) cell_source_kernel<2><<<blocks, 128, 0, c->stream>>>(c->dev, q, kind, dt, (int)tasks.size(), s->nslot[which], s->d_slot_task[which], s->d_slot_comp[which]);
        else cell_source_kernel<3><<<blocks, 128, 0, c->stream>>>(c->dev, q, kind, dt, (int)tasks.size(), s->nslot[which], s->d_slot_task[which], s->d_slot_comp[which]);
        c->launches++;
    } else {
        const unsigned long long seed = (unsigned long long)c->cfg.rng_seed_global * (unsigned long long)(c->cfg.rank + 1);
        for (size_t it = 0; it < tasks.size(); it++)
            for (int i = s->mix_off[tasks[it].mix]; i < s->mix_off[tasks[it].mix + 1]; i++) {
                const int icomp = i - s->mix_off[tasks[it].mix];
                const double spwt = s->spwt.empty() ? 1.0 : s->spwt[s->mix_sp[i] - 1];
                const double rate = (kind == 5) ? tasks[it].rate : c->cfg.dt * tasks[it].rate;
                const double x = rate / (c->cfg.fnum * spwt) * vol_total * s->mix_frac[i] / nranks;
                int np_init = (int)x; /* RANDINT with the first draw of stream (e0 = 0) */
                if (x - np_init > first_rf_host(seed, kind, tID, 0ULL, (unsigned int)(it * 256 + icomp))) np_init++;
                if (np_init <= 0) continue;
                domain_source_kernel<<<(unsigned int)((np_init + 127) / 128), 128, 0, c->stream>>>(c->dev, q, kind, tasks[it], (int)it, icomp, np_init);
                c->launches++;
            }
    }
    PG_CUDA(c, cudaGetLastError());
    c->owned_clean = c->owned_clean && kind == 5 && unstructured; /* INITIAL_SEED on a mesh creates in owned cells only */
    return end_create(c, what);
}

/* INITIAL_SEED without its trailing EXCHANGE (call pgpu_exchange next, as src/initialization.f90:2231 does) */
extern "C" int pgpu_initial_seed(pgpu_ctx* c) {
    if (!c) return PGPU_ERR_INVALID;
    return cell_source(c, 0, 5, 0, "pgpu_initial_seed");
}
extern "C" int pgpu_volume_inject(pgpu_ctx* c, int32_t tID) {
    if (!c) return PGPU_ERR_INVALID;
    return cell_source(c, 1, 6, tID, "pgpu_volume_inject");
}
extern "C" int pgpu_boundaries_emit(pgpu_ctx* c, int32_t tID) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    PgInjectState* s = inj(c);
    if (s->emit.empty()) return PGPU_OK;
    if (c->cfg.grid_type != PGPU_UNSTRUCTURED) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "boundaries_emit: emit tasks exist on unstructured grids only");
    if (s->mix_off.empty()) PG_FAIL(c, PGPU_ERR_STATE, "boundaries_emit: pgpu_set_mixtures has not been called");
    if ((rc = begin_create(c, s->expected[2], true))) return rc;
    c->dev.tID = tID;
    PgInject q = make_inject(c, s);
    const unsigned int blocks = (unsigned int)(((long long)s->nslot[2] * 32 + 127) / 128);
    if (c->mesh_dims == 1) emit_kernel<1><<<blocks, 128, 0, c->stream>>>(c->dev, q, s->d_emit, s->d_emit_nfs, s->nslot[2], s->d_slot_task[2], s->d_slot_comp[2]);
    else if (c->mesh_dims == 2) emit_kernel<2><<<blocks, 128, 0, c->stream>>>(c->dev, q, s->d_emit, s->d_emit_nfs, s->nslot[2], s->d_slot_task[2], s->d_slot_comp[2]);
    else emit_kernel<3><<<blocks, 128, 0, c->stream>>>(c->dev, q, s->d_emit, s->d_emit_nfs, s->nslot[2], s->d_slot_task[2], s->d_slot_comp[2]);
    c->launches++;
    PG_CUDA(c, cudaGetLastError());
    return end_create(c, "pgpu_boundaries_emit");
}
