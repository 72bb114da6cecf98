/*
 * pgpu_pic_mcc.cu -- PIC charge deposition, null-collision MCC and species removal.
 *
 * Replaces: the particle loop of DEPOSIT_CHARGE + SURFACE_CHARGE + the MPI_REDUCE/MPI_BCAST of RHS
 * (src/fields.f90:5440-5535; COMPUTE_WEIGHTS :5341-5431), MCC_COLLISIONS_VAHEDI
 * (src/collisions.f90:1606-1892), REMOVE_PARTICLES_OF_SPECIES (src/timecycle.f90:2340-2353).
 * The Dirichlet/Neumann row fill and the PETSc vector load (src/fields.f90:5538-5561) stay on the host.
 */
#include <algorithm>

#include "pgpu_internal.cuh"
#include "pgpu_physics.cuh"

int pg_allreduce_sum_f64(pgpu_ctx* c, double* buf, long long n); /* pgpu_exchange.cu */

namespace {

/*
 * One thread per particle.  SMEM: the node array is privatised per block in shared memory (1-D / small
 * 2-D meshes: millions of particles land on a few hundred nodes, so global atomics would serialise), and
 * flushed with one FP64 atomic per touched node per block.  Otherwise FP64 atomics go straight to the
 * L2-resident RHS.
 */
/* The store is cell-sorted (or nearly), so the lanes of a warp usually deposit on the SAME nodes: then the warp sums
 * its contributions with shuffles and issues one atomic per node instead of thirty-two (an FP64 atomicAdd on shared
 * memory is a compare-and-swap loop, and thirty-two lanes on one address retry thirty-two times).  Every lane of the
 * warp calls this; key <= 0: nothing to deposit. */
__device__ __forceinline__ void warp_deposit(double* acc, int key, int nv, const int* node, const double* val) {
    const unsigned int FULL = 0xffffffffu;
    const unsigned int live = __ballot_sync(FULL, key > 0);
    if (!live) return;
    const int lane = threadIdx.x & 31, lead = __ffs(live) - 1;
    const int key0 = __shfl_sync(FULL, key, lead);
    if (__all_sync(FULL, key <= 0 || key == key0)) {
        const int nv0 = __shfl_sync(FULL, nv, lead);
        for (int k = 0; k < 4; k++) {
            if (k >= nv0) break;
            double v = key > 0 ? val[k] : 0.;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            const int nd = __shfl_sync(FULL, node[k], lead);
            if (lane == lead) atomicAdd(&acc[nd], v);
        }
    } else if (key > 0) {
        for (int k = 0; k < nv; k++) atomicAdd(&acc[node[k]], val[k]);
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(256) deposit_kernel(const __grid_constant__ PgDev d, double* __restrict__ rhs, int nn) {
    extern __shared__ double srhs[];
    if (SMEM) {
        for (int i = threadIdx.x; i < nn; i += blockDim.x) srhs[i] = 0.;
        __syncthreads();
    }
    double* acc = SMEM ? srhs : rhs;
    const double K = PG_QE / (PG_EPS0 * (d.eps_scaling * d.eps_scaling));
    const long long n = (long long)d.status->np_cur, stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += stride) {
        const long long jp = base + threadIdx.x;
        int key = 0, nv = 0, node[4] = {0, 0, 0, 0};
        double val[4] = {0., 0., 0., 0.};
        PgPos P;
        P.ic = 0;
        if (jp < n) P = d.p.pos[jp];
        const int ic = P.ic;
        double charge = 0.;
        if (ic >= 1) charge = __ldg(&d.charge[P.sid - 1]);
        if (ic >= 1 && !(fabs(charge) < 1.e-6)) {
            key = ic;
            const double x = P.x;
            if (d.grid_type == PGPU_UNSTRUCTURED) {
                const size_t c = (size_t)(ic - 1);
                if (d.dims == 1) {
                    const double rho_q = K * charge * d.fnum / (d.ymax - d.ymin) / (d.zmax - d.zmin);
                    nv = 2;
                    for (int p = 0; p < 2; p++) {
                        node[p] = __ldg(&d.cell_nodes[2 * c + p]) - 1;
                        const double* bs = d.basis + 4 * c + 2 * p;
                        val[p] = rho_q * (__ldg(&bs[0]) * x + __ldg(&bs[1]));
                    }
                } else if (d.dims == 2) {
                    const double y = P.y;
                    const double rho_q = K * charge * d.fnum / (d.zmax - d.zmin);
                    nv = 3;
                    for (int p = 0; p < 3; p++) {
                        node[p] = __ldg(&d.cell_nodes[3 * c + p]) - 1;
                        const double* bs = d.basis + 9 * c + 3 * p;
                        val[p] = rho_q * (__ldg(&bs[0]) * x + __ldg(&bs[1]) * y + __ldg(&bs[2]));
                    }
                } else {
                    const double y = P.y, z = P.z;
                    const double rho_q = K * charge * d.fnum;
                    nv = 4;
                    for (int p = 0; p < 4; p++) {
                        node[p] = __ldg(&d.cell_nodes[4 * c + p]) - 1;
                        const double* bs = d.basis + 16 * c + 4 * p;
                        val[p] = rho_q * (__ldg(&bs[0]) * x + __ldg(&bs[1]) * y + __ldg(&bs[2]) * z + __ldg(&bs[3]));
                    }
                }
            } else {
                /* ref: src/fields.f90:5341-5431 COMPUTE_WEIGHTS + :5494-5520 */
                const double y = P.y;
                const int nx = d.nx;
                const int i = (ic - 1) % nx, j = (ic - 1) / nx;
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
                const double fx = (x - cxmin) / hx;
                const double fy = (y - cymin) / hy;
                const double vol = (d.xmax - d.xmin) * (d.ymax - d.ymin) * (d.zmax - d.zmin) / (d.nx * d.ny);
                const double cfnum = d.cell_fnum ? __ldg(&d.cell_fnum[ic - 1]) : d.fnum; /* ref: src/fields.f90:5508 */
                const double rho_q = -K * charge * cfnum / vol;
                if (d.dims == 2) {
                    const int i0 = ic + j - 1;
                    nv = 4;
                    node[0] = i0;          val[0] = rho_q * ((1. - fx) * (1. - fy));
                    node[1] = i0 + 1;      val[1] = rho_q * (fx * (1. - fy));
                    node[2] = i0 + nx + 1; val[2] = rho_q * ((1. - fx) * fy);
                    node[3] = i0 + nx + 2; val[3] = rho_q * (fx * fy);
                } else {
                    nv = 2;
                    node[0] = ic - 1; val[0] = rho_q * (1. - fx);
                    node[1] = ic;     val[1] = rho_q * fx;
                }
            }
        }
        warp_deposit(acc, key, nv, node, val);
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            double v = srhs[i];
            if (v != 0.) atomicAdd(&rhs[i], v);
        }
    }
}

__global__ void add_surface_charge_kernel(double* __restrict__ rhs, const double* __restrict__ sc, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rhs[i] = rhs[i] + sc[i];
}

/* ref: src/collisions.f90:1646-1890.  Two kernels: the null-collision test of every particle present at entry (one
 * Philox draw, 32 registers) compacts the ~1 % that collide into a list; the collision proper (background partner,
 * cross-section tables, reaction products: 118 registers) then runs on full warps.  As one kernel a warp with a single
 * candidate ran the whole collision path for it (measured: 8 active lanes on average, 0.53 ms per 1e7 particles). */
__global__ void __launch_bounds__(256) mcc_select_kernel(const __grid_constant__ PgDev d, double p_null) {
    const long long jp1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int ic = 0;
    if (jp1 < (long long)d.status->np_cur) ic = d.p.pos[jp1].ic; /* the particles present at entry */
    PgRng g;
    g.init(d.seed_local, 3, d.tID, (unsigned long long)jp1, 0ULL);
    const bool cand = ic >= 1 && !(g.rf() > p_null);
    const unsigned int cm = __ballot_sync(0xffffffffu, cand);
    if (!cm) return;
    const int lane = threadIdx.x & 31, leader = __ffs(cm) - 1;
    unsigned long long base = 0;
    if (lane == leader) {
        base = atomicAdd(&d.status->n_cand, (unsigned long long)__popc(cm));
        atomicAdd(&d.status->ncoll, (unsigned long long)__popc(cm));
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (cand) d.hitlist[base + __popc(cm & ((1u << lane) - 1u))] = (unsigned int)jp1;
}

__device__ void mcc_vahedi_one(const PgDev& d, const long long jp1, double p_null) {
    const PgPos P1 = d.p.pos[jp1];
    const int ic = P1.ic;
    PgRng g;
    g.init(d.seed_local, 3, d.tID, (unsigned long long)jp1, 0ULL);
    (void)g.rf(); /* the draw of the null-collision test */
    const double dt = d.dt;
    int s1 = P1.sid;
    const PgVel V1 = d.p.vel[jp1];
    double c1[3] = {V1.vx, V1.vy, V1.vz};
    for (int j = 0; j < d.mcc_ncomp; j++) {
        int s2 = d.mcc_sp[j];
        double bg_nrho = d.mcc_frac[j] * d.mcc_dens;
        if (bg_nrho == 0) continue;
        double c2[3];
        pg_maxwell0(g, d.mcc_ttra, d.mass[s2 - 1], c2);
        double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
        double vr = sqrt(vr2);
        double r_select = g.rf();
        double m1 = d.mass[s1 - 1], m2 = d.mass[s2 - 1];
        double mred = m1 * m2 / (m1 + m2);
        double etr = 0.5 * mred * vr2;
        double p_cum = 0;
        for (int jr = 0; jr < d.nr; jr++) {
            const PgReaction r = d.reac[jr];
            if (r.r1 != s1 || r.r2 != s2) continue;
            double ea = r.ea;
            if (etr <= ea) continue;
            if (r.type != PGPU_LXCAT) continue;
            double sigma_r = pg_interp_cs(etr, r.tab_n, d.tab_e + r.tab_off, d.tab_cs + r.tab_off);
            p_cum = p_cum + (1. - exp(-bg_nrho * sigma_r * vr * dt)) / p_null;
            if (r_select < p_cum) {
                /* materialise the background partner at the particle's position (ref: :1726-1741) */
                double erot2 = pg_internal_energy(g, d.rotdof[s2 - 1], d.mcc_ttra);
                double evib2 = pg_internal_energy(g, d.vibdof[s2 - 1], d.mcc_ttra);
                double x = P1.x, y = P1.y, z = P1.z;
                atomicAdd(&d.status->nreac, 1ULL);
                double ecoll = etr - ea;
                const PgIe E1 = pg_ie_load(d.p.ie, jp1);
                double erot1 = E1.erot, evib1 = E1.evib;
                int bad = 0;
                if (!r.is_cex) {
                    /* products past the second are appended by the redistribution itself; the oracle
                     * appends the partner first, so creation order differs but not the set of particles */
                    bad = pg_react_redistribute(d, g, r, ecoll, c1, c2, erot1, evib1, erot2, evib2, x, y, z, ic);
                    pg_ie_store(d.p.ie, jp1, erot1, evib1, d.status);
                }
                (void)bad;
                PgVel W;
                W.vx = c1[0]; W.vy = c1[1]; W.vz = c1[2]; W.sid = r.p1; W.pad = 0;
                d.p.vel[jp1] = W;
                d.p.pos[jp1].sid = r.p1;
                pg_append_particle(d, x, y, z, c2[0], c2[1], c2[2], erot2, evib2, r.p2, ic);
                return;
            }
        }
    }
}

__global__ void __launch_bounds__(128) mcc_vahedi_kernel(const __grid_constant__ PgDev d, double p_null) {
    const long long n = (long long)d.status->n_cand, stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        mcc_vahedi_one(d, (long long)d.hitlist[k], p_null);
}

__global__ void __launch_bounds__(256) remove_species_kernel(const __grid_constant__ PgDev d, int sp) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)d.status->np_cur) return;
    const PgPos P = d.p.pos[i];
    if (P.ic < 1) return;
    if (P.sid == sp) {
        d.p.pos[i].ic = PG_TOMBSTONE;
        atomicAdd(&d.status->n_removed, 1ULL);
    }
}

} /* namespace */

static int nnodes_of(const pgpu_ctx* c) {
    if (c->cfg.grid_type == PGPU_UNSTRUCTURED) return c->nn;
    return (c->cfg.dims == 1) ? c->cfg.nx + 1 : (c->cfg.nx + 1) * (c->cfg.ny + 1);
}

extern "C" int pgpu_deposit_charge(pgpu_ctx* c, double* rhs_host, int64_t nnodes) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    int nn = nnodes_of(c);
    if (!rhs_host || nnodes < nn) PG_FAIL(c, PGPU_ERR_INVALID, "deposit_charge: rhs too small (%lld < %d)", (long long)nnodes, nn);
    if (c->cfg.grid_type == PGPU_UNSTRUCTURED && !c->dev.basis) PG_FAIL(c, PGPU_ERR_STATE, "deposit_charge: mesh without BASIS_COEFFS");
    if (c->rhs_n < nn) {
        cudaFree(c->d_rhs);
        c->d_rhs = nullptr;
        PG_CUDA(c, cudaMalloc((void**)&c->d_rhs, sizeof(double) * (size_t)nn));
        c->rhs_n = nn;
    }
    memset(c->ms, 0, sizeof(c->ms));
    c->launches = 0;
    if ((rc = pg_reset_status(c))) return rc;
    cudaEventRecord(c->ev[6], c->stream);
    PG_CUDA(c, cudaMemsetAsync(c->d_rhs, 0, sizeof(double) * (size_t)nn, c->stream));
    if (c->np > 0) {
        const PgDev& d = c->dev;
        int blocks = (int)std::min<long long>((c->np + 255) / 256, 148LL * 8);
        size_t smem = sizeof(double) * (size_t)nn;
        if (smem <= 48 * 1024) deposit_kernel<true><<<blocks, 256, smem, c->stream>>>(d, c->d_rhs, nn);
        else deposit_kernel<false><<<blocks, 256, 0, c->stream>>>(d, c->d_rhs, nn);
        c->launches++;
    }
    /* RHS = RHS + SURFACE_CHARGE on EVERY rank before the reduction: SURFACE_CHARGE is a per-rank partial in the
     * reference (each rank accumulates the dielectric hits of its own particles in ADVECT, src/timecycle.f90:1399-1430,
     * and adds its array before MPI_REDUCE(SUM), src/fields.f90:5525-5535), and d_surface_charge here is filled per
     * rank by boundary_unstructured in the same way */
    if (c->d_surface_charge && c->cfg.grid_type == PGPU_UNSTRUCTURED) {
        long long m = std::min<long long>(nn, c->nsc);
        add_surface_charge_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, c->stream>>>(c->d_rhs, c->d_surface_charge, m);
        c->launches++;
    }
    if (c->cfg.nranks > 1 && (rc = pg_allreduce_sum_f64(c, c->d_rhs, nn))) return rc;
    PG_CUDA(c, cudaMemcpyAsync(rhs_host, c->d_rhs, sizeof(double) * (size_t)nn, cudaMemcpyDeviceToHost, c->stream));
    cudaEventRecord(c->ev[7], c->stream);
    rc = pg_finish(c, "pgpu_deposit_charge");
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]) == cudaSuccess) c->ms[3] = ms;
    c->ms[5] = c->ms[3];
    return rc;
}

int pg_mcc_impl(pgpu_ctx* c, int tID) {
    if (!c->dev.mcc_ncomp) PG_FAIL(c, PGPU_ERR_STATE, "mcc_vahedi: background mixture not set");
    if (c->np_ub == 0) return PGPU_OK;
    { int rc0; if ((rc0 = pg_resolve_vel(c))) return rc0; }
    c->dev.tID = tID;
    pg_refresh_dev(c);
    /* ref: src/collisions.f90:1637-1644 (single-precision literal 1e-12) */
    double dt = c->cfg.dt;
    double null_coll_freq = c->cfg.mcc_bg_dens * (double)1e-12f;
    double p_null = 1 - exp(-dt * null_coll_freq);
    if (p_null > 1) p_null = 1.;
    PG_CUDA(c, cudaMemsetAsync(&c->d_status->n_cand, 0, sizeof(unsigned long long), c->stream));
    mcc_select_kernel<<<(unsigned int)((c->np_ub + 255) / 256), 256, 0, c->stream>>>(c->dev, p_null);
    {
        /* the list's length stays on the device: a grid that fills the machine strides over it */
        const long long want = (c->np_ub + 127) / 128;
        mcc_vahedi_kernel<<<(unsigned int)std::min<long long>(want, 148LL * 8), 128, 0, c->stream>>>(c->dev, p_null);
    }
    c->launches += 2;
    PG_CUDA(c, cudaGetLastError());
    c->sorted = false;
    c->hist_valid = false;
    return PGPU_OK;
}

extern "C" int pgpu_mcc_vahedi(pgpu_ctx* c, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    memset(c->ms, 0, sizeof(c->ms));
    c->launches = 0;
    if ((rc = pg_reset_status(c))) return rc;
    cudaEventRecord(c->ev[4], c->stream);
    if ((rc = pg_mcc_impl(c, tID))) return rc;
    cudaEventRecord(c->ev[5], c->stream);
    rc = pg_finish(c, "pgpu_mcc_vahedi");
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->ms[2] = ms;
    c->ms[5] = c->ms[2];
    if (timestep_coll) *timestep_coll = (int64_t)c->h_status->ncoll;
    if (timestep_reac) *timestep_reac = (int64_t)c->h_status->nreac;
    return rc;
}

extern "C" int pgpu_remove_species(pgpu_ctx* c, int32_t sp_id) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (sp_id < 1 || sp_id > c->ns) PG_FAIL(c, PGPU_ERR_INVALID, "remove_species: species %d outside 1..%lld", sp_id, c->ns);
    if (c->np == 0) return PGPU_OK;
    if ((rc = pg_reset_status(c))) return rc;
    remove_species_kernel<<<(unsigned int)((c->np + 255) / 256), 256, 0, c->stream>>>(c->dev, sp_id);
    c->launches++;
    c->maybe_tombstones = true;
    c->sorted = false;
    c->hist_valid = false;
    return pg_finish(c, "pgpu_remove_species");
}
