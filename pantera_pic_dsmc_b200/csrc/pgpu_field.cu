/*
 * pgpu_field.cu -- the electrostatic field step of the explicit PIC loop on the device, so that the
 * 1-D capacitively coupled plasma case (config C4) runs DEPOSIT_CHARGE -> Poisson -> COMPUTE_E_FIELD -> ADVECT without a
 * host round trip per step.
 *
 * Replaces, for 1-D unstructured (line) meshes: the matrix of ASSEMBLE_POISSON (src/fields.f90:137-253: per segment
 * K11 = K22 = EPS_REL/LENGTH, K12 = -EPS_REL/LENGTH added to the rows of non-Dirichlet nodes; identity rows for
 * Dirichlet nodes, :758-760), SET_WALL_POTENTIAL (:5566-5607, Dirichlet / RF-voltage electrodes at timestep tID), the
 * right-hand side rows of DEPOSIT_CHARGE (:5538-5561: Dirichlet value, RHS + NEUMANN, or RHS) and the linear solve
 * (PETSc KSP in the reference; here an exact tridiagonal elimination along the chain of nodes).
 * For every unstructured mesh (1-D, 2-D, 3-D): COMPUTE_E_FIELD (:5237-5296), E = -sum_p PHI(v_p) BASIS_COEFFS(1:d, p, cell).
 * 2-D/3-D Poisson solves stay with the host's PETSc (north_star: "the PETSc field solve stays outside the hot path").
 */
#include <algorithm>
#include <vector>

#include "pgpu_internal.cuh"

namespace {

/* one thread per cell: stiffness contributions in chain positions (pos = position of a node along the line) */
__global__ void poisson1d_assemble_kernel(int nc, const int* __restrict__ cell_nodes, const int* __restrict__ node_pos,
                                          const double* __restrict__ node_coords, const double* __restrict__ cell_eps,
                                          double* __restrict__ diag, double* __restrict__ lower, double* __restrict__ upper) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int v1 = cell_nodes[2 * i] - 1, v2 = cell_nodes[2 * i + 1] - 1;
    const double length = fabs(node_coords[3 * (size_t)v2] - node_coords[3 * (size_t)v1]); /* SEGMENT_LENGTHS */
    const double eps = cell_eps ? cell_eps[i] : 1.0;
    const double k11 = 1.0 / length * eps, k12 = -1.0 / length * eps;
    const int p1 = node_pos[v1], p2 = node_pos[v2];
    atomicAdd(&diag[p1], k11);
    atomicAdd(&diag[p2], k11);
    /* neighbours along the chain: the coupling of positions (lo, lo+1) */
    const int lo = p1 < p2 ? p1 : p2;
    atomicAdd(&upper[lo], k12);     /* row lo, column lo+1 */
    atomicAdd(&lower[lo + 1], k12); /* row lo+1, column lo */
}

/* rows of Dirichlet nodes become identity, the right-hand side follows src/fields.f90:5538-5561; then the Thomas
 * elimination along the chain (a few hundred nodes: one thread) and the scatter back to node numbering */
template <bool SMEM>
__global__ void __launch_bounds__(256) poisson1d_solve_kernel(int nn, const int* __restrict__ pos_node, const double* __restrict__ rhs,
                                       const unsigned char* __restrict__ is_dir, const double* __restrict__ dir_val,
                                       const double* __restrict__ neumann, double* __restrict__ g_diag, double* __restrict__ g_lower,
                                       double* __restrict__ g_upper, double* __restrict__ g_work, double* __restrict__ phi) {
    /* The elimination is one dependent chain; run from global memory every link costs an L2 round trip (measured:
     * 0.34 ms for 513 nodes).  The block stages the four arrays in shared memory, one thread walks the chain there. */
    extern __shared__ double sm_rows[];
    double* diag = SMEM ? sm_rows : g_diag;
    double* lower = SMEM ? sm_rows + nn : g_lower;
    double* upper = SMEM ? sm_rows + 2 * (size_t)nn : g_upper;
    double* work = SMEM ? sm_rows + 3 * (size_t)nn : g_work;
    /* b in chain order */
    for (int p = threadIdx.x; p < nn; p += blockDim.x) {
        const int v = pos_node[p];
        double b, dg = g_diag[p], lo = g_lower[p], up = g_upper[p];
        if (is_dir[v]) {
            b = dir_val[v];
            dg = 1.0; lo = 0.0; up = 0.0;
        } else {
            b = rhs[v] + neumann[v];
        }
        diag[p] = dg; lower[p] = lo; upper[p] = up; work[p] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* forward elimination */
        for (int p = 1; p < nn; p++) {
            const double m = lower[p] / diag[p - 1];
            diag[p] = diag[p] - m * upper[p - 1];
            work[p] = work[p] - m * work[p - 1];
        }
        work[nn - 1] = work[nn - 1] / diag[nn - 1];
        for (int p = nn - 2; p >= 0; p--) work[p] = (work[p] - upper[p] * work[p + 1]) / diag[p];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nn; p += blockDim.x) phi[pos_node[p]] = work[p];
}

/* COMPUTE_E_FIELD on unstructured grids, ref: src/fields.f90:5253-5296; E_FIELD(3, 1, NCELLS) */
template <int DIMS>
__global__ void efield_kernel(int nc, const int* __restrict__ cell_nodes, const double* __restrict__ basis, const double* __restrict__ phi,
                              double* __restrict__ e) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    constexpr int NF = DIMS + 1;
    double ev[3] = {0., 0., 0.};
    for (int id = 0; id < DIMS; id++)
        for (int p = 0; p < NF; p++) {
            const int vp = cell_nodes[(size_t)NF * i + p] - 1;
            /* BASIS_COEFFS(ID, P, I): first index fastest */
            ev[id] = ev[id] - phi[vp] * basis[(size_t)NF * NF * i + (size_t)NF * p + id];
        }
    e[3 * (size_t)i] = ev[0]; e[3 * (size_t)i + 1] = ev[1]; e[3 * (size_t)i + 2] = ev[2];
}

} /* namespace */

struct PgFieldState {
    std::vector<double> wall_potential, wall_rf_potential, rf_frequency, wall_efield, eps_rel;
    bool chain_ready = false;
    std::vector<int> node_pos, pos_node;
    int *d_node_pos = nullptr, *d_pos_node = nullptr;
    unsigned char* d_is_dir = nullptr;
    double *d_dir = nullptr, *d_neumann = nullptr, *d_diag = nullptr, *d_lower = nullptr, *d_upper = nullptr, *d_work = nullptr,
           *d_phi = nullptr, *d_cell_eps = nullptr, *d_e = nullptr;
    long long e_n = 0;
};
static PgFieldState* fld(pgpu_ctx* c) {
    if (!c->field) c->field = new PgFieldState();
    return (PgFieldState*)c->field;
}
void pg_field_destroy(pgpu_ctx* c) {
    if (!c->field) return;
    PgFieldState* s = (PgFieldState*)c->field;
    cudaFree(s->d_node_pos); cudaFree(s->d_pos_node); cudaFree(s->d_is_dir); cudaFree(s->d_dir); cudaFree(s->d_neumann);
    cudaFree(s->d_diag); cudaFree(s->d_lower); cudaFree(s->d_upper); cudaFree(s->d_work); cudaFree(s->d_phi); cudaFree(s->d_cell_eps);
    /* (d_e is owned by the context's e_field slot once handed over) */
    delete s;
    c->field = nullptr;
}

/* GRID_BC(:)%WALL_POTENTIAL, %WALL_RF_POTENTIAL, %RF_FREQUENCY, %WALL_EFIELD, %EPS_REL (src/global.f90:225-233), one entry
 * per physical group like pgpu_set_bcs; null arrays mean 0 (EPS_REL: 1) */
extern "C" int pgpu_set_field_bcs(pgpu_ctx* c, int nbc, const double* wall_potential, const double* wall_rf_potential,
                                  const double* rf_frequency, const double* wall_efield, const double* eps_rel) {
    if (!c || nbc < 0) return PGPU_ERR_INVALID;
    PgFieldState* s = fld(c);
    auto fill = [&](std::vector<double>& v, const double* p, double dflt) {
        v.assign((size_t)nbc, dflt);
        if (p) std::copy(p, p + nbc, v.begin());
    };
    fill(s->wall_potential, wall_potential, 0.0);
    fill(s->wall_rf_potential, wall_rf_potential, 0.0);
    fill(s->rf_frequency, rf_frequency, 0.0);
    fill(s->wall_efield, wall_efield, 0.0);
    fill(s->eps_rel, eps_rel, 1.0);
    s->chain_ready = false;
    return PGPU_OK;
}

static int build_chain(pgpu_ctx* c, PgFieldState* s) {
    const int nn = c->nn, nc = c->nc;
    std::vector<std::vector<int>> adj((size_t)nn);
    for (int i = 0; i < nc; i++) {
        int v1 = c->h_cell_nodes[2 * (size_t)i] - 1, v2 = c->h_cell_nodes[2 * (size_t)i + 1] - 1;
        adj[v1].push_back(v2);
        adj[v2].push_back(v1);
    }
    int start = -1;
    for (int v = 0; v < nn; v++) {
        if (adj[v].size() > 2 || adj[v].empty()) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: the 1-D mesh is not a simple chain of segments (node %d)", v + 1);
        if (adj[v].size() == 1 && start < 0) start = v;
    }
    if (start < 0) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: periodic 1-D meshes are not supported");
    s->node_pos.assign((size_t)nn, -1);
    s->pos_node.assign((size_t)nn, -1);
    int prev = -1, cur = start;
    for (int p = 0; p < nn; p++) {
        s->node_pos[cur] = p;
        s->pos_node[p] = cur;
        int nxt = -1;
        for (int w : adj[cur])
            if (w != prev) nxt = w;
        prev = cur;
        cur = nxt;
        if (cur < 0 && p != nn - 1) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: the 1-D mesh is not connected");
    }
    auto dalloc = [&](void** p, size_t bytes) { cudaFree(*p); *p = nullptr; return cudaMalloc(p, bytes); };
    PG_CUDA(c, dalloc((void**)&s->d_node_pos, sizeof(int) * (size_t)nn));
    PG_CUDA(c, dalloc((void**)&s->d_pos_node, sizeof(int) * (size_t)nn));
    PG_CUDA(c, dalloc((void**)&s->d_is_dir, (size_t)nn));
    for (double** p : {&s->d_dir, &s->d_neumann, &s->d_diag, &s->d_lower, &s->d_upper, &s->d_work, &s->d_phi})
        PG_CUDA(c, dalloc((void**)p, sizeof(double) * (size_t)(nn + 1)));
    PG_CUDA(c, dalloc((void**)&s->d_cell_eps, sizeof(double) * (size_t)nc));
    PG_CUDA(c, cudaMemcpy(s->d_node_pos, s->node_pos.data(), sizeof(int) * (size_t)nn, cudaMemcpyHostToDevice));
    PG_CUDA(c, cudaMemcpy(s->d_pos_node, s->pos_node.data(), sizeof(int) * (size_t)nn, cudaMemcpyHostToDevice));
    std::vector<double> eps((size_t)nc, 1.0);
    if (!c->h_cell_pg.empty())
        for (int i = 0; i < nc; i++) {
            int pg = c->h_cell_pg[i];
            if (pg != -1 && pg >= 1 && pg <= (int)s->eps_rel.size()) eps[i] = s->eps_rel[pg - 1];
        }
    PG_CUDA(c, cudaMemcpy(s->d_cell_eps, eps.data(), sizeof(double) * (size_t)nc, cudaMemcpyHostToDevice));
    s->chain_ready = true;
    return PGPU_OK;
}

/* COMPUTE_E_FIELD from a nodal potential held by the host (any unstructured mesh); E becomes the field ADVECT gathers */
static int efield_from_phi(pgpu_ctx* c, PgFieldState* s, const double* d_phi, double* e_host) {
    const int nc = c->nc;
    if (!c->dev.basis) PG_FAIL(c, PGPU_ERR_STATE, "compute_e_field: mesh without BASIS_COEFFS");
    if (s->e_n != 3LL * nc || !s->d_e || c->dev.e_field != s->d_e) { /* (pgpu_set_efield may have replaced the slot) */
        /* the array is handed to the context's E-field slot, which frees the previous one */
        double* p = nullptr;
        PG_CUDA(c, cudaMalloc((void**)&p, sizeof(double) * 3 * (size_t)nc));
        s->d_e = p;
        s->e_n = 3LL * nc;
        int rc = pg_adopt_efield(c, p);
        if (rc) return rc;
    }
    const unsigned int blocks = (unsigned int)((nc + 127) / 128);
    if (c->mesh_dims == 1) efield_kernel<1><<<blocks, 128, 0, c->stream>>>(nc, c->dev.cell_nodes, c->dev.basis, d_phi, s->d_e);
    else if (c->mesh_dims == 2) efield_kernel<2><<<blocks, 128, 0, c->stream>>>(nc, c->dev.cell_nodes, c->dev.basis, d_phi, s->d_e);
    else efield_kernel<3><<<blocks, 128, 0, c->stream>>>(nc, c->dev.cell_nodes, c->dev.basis, d_phi, s->d_e);
    c->launches++;
    if (e_host) PG_CUDA(c, cudaMemcpyAsync(e_host, s->d_e, sizeof(double) * 3 * (size_t)nc, cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}

extern "C" int pgpu_compute_e_field(pgpu_ctx* c, const double* phi_host, int64_t nnodes, double* e_host) {
    if (!c || !phi_host) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (c->cfg.grid_type != PGPU_UNSTRUCTURED) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "compute_e_field: unstructured grids only (the rectilinear stencil stays on the host)");
    if (nnodes < c->nn) PG_FAIL(c, PGPU_ERR_INVALID, "compute_e_field: %lld potentials for %d nodes", (long long)nnodes, c->nn);
    PgFieldState* s = fld(c);
    double* d_phi = nullptr;
    PG_CUDA(c, cudaMalloc((void**)&d_phi, sizeof(double) * (size_t)c->nn));
    PG_CUDA(c, cudaMemcpyAsync(d_phi, phi_host, sizeof(double) * (size_t)c->nn, cudaMemcpyHostToDevice, c->stream));
    rc = efield_from_phi(c, s, d_phi, e_host);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_phi);
    return rc;
}

/* SET_WALL_POTENTIAL + Poisson solve + COMPUTE_E_FIELD for a 1-D line mesh.  rhs_host == NULL: the RHS of the last
 * pgpu_deposit_charge, still on the device.  phi_host / e_host receive PHI_FIELD(NNODES) / E_FIELD(3,1,NCELLS) when not
 * NULL.  E stays on the device as the field of the next ADVECT. */
extern "C" int pgpu_field_solve_1d(pgpu_ctx* c, int32_t tID, const double* rhs_host, double* phi_host, double* e_host) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (c->cfg.grid_type != PGPU_UNSTRUCTURED || c->mesh_dims != 1) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve_1d: 1-D unstructured meshes only");
    PgFieldState* s = fld(c);
    const int nn = c->nn, nc = c->nc, nbc = (int)c->h_bc_field.size();
    if ((int)s->wall_potential.size() < nbc) { /* defaults when pgpu_set_field_bcs was not called */
        s->wall_potential.resize(nbc, 0.0); s->wall_rf_potential.resize(nbc, 0.0); s->rf_frequency.resize(nbc, 0.0);
        s->wall_efield.resize(nbc, 0.0); s->eps_rel.resize(nbc, 1.0);
    }
    if (!s->chain_ready && (rc = build_chain(c, s))) return rc;
    if (!rhs_host && (!c->d_rhs || c->rhs_n < nn)) PG_FAIL(c, PGPU_ERR_STATE, "field_solve_1d: no deposited RHS on the device (call pgpu_deposit_charge or pass rhs)");
    /* boundary rows on the host (a handful of end nodes), ref: src/fields.f90:143-176, 236-251, 5575-5607 */
    std::vector<unsigned char> is_dir((size_t)nn, 0);
    std::vector<double> dir((size_t)nn, 0.0), neumann((size_t)nn, 0.0);
    const double PI = 3.1415926535897932;
    for (int i = 0; i < nc; i++)
        for (int j = 0; j < 2; j++) {
            int pg = c->h_cell_faces_pg[2 * (size_t)i + j];
            if (pg == -1) continue;
            int fbc = c->h_bc_field[pg - 1], v = c->h_cell_nodes[2 * (size_t)i + j] - 1;
            if (fbc == PGPU_DIRICHLET_BC || fbc == PGPU_RF_VOLTAGE_BC || fbc == PGPU_DECOUPLED_RF_VOLTAGE_BC) {
                double pot = s->wall_potential[pg - 1];
                if (fbc != PGPU_DIRICHLET_BC) pot = s->wall_potential[pg - 1] + 0.5 * s->wall_rf_potential[pg - 1] * cos(2 * PI * s->rf_frequency[pg - 1] * tID * c->cfg.dt);
                dir[v] = pot;
                is_dir[v] = 1;
            }
        }
    for (int i = 0; i < nc; i++)
        for (int j = 0; j < 2; j++) {
            int pg = c->h_cell_faces_pg[2 * (size_t)i + j];
            if (pg == -1) continue;
            int v = c->h_cell_nodes[2 * (size_t)i + j] - 1;
            if (c->h_bc_field[pg - 1] == PGPU_NEUMANN_BC && !is_dir[v]) neumann[v] = neumann[v] + s->wall_efield[pg - 1];
        }
    bool any_dir = false;
    for (int v = 0; v < nn; v++) any_dir |= is_dir[v] != 0;
    if (!any_dir) PG_FAIL(c, PGPU_ERR_INVALID, "field_solve_1d: no Dirichlet / RF-voltage electrode: the Poisson problem is singular");
    PG_CUDA(c, cudaMemcpyAsync(s->d_is_dir, is_dir.data(), (size_t)nn, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(s->d_dir, dir.data(), sizeof(double) * (size_t)nn, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(s->d_neumann, neumann.data(), sizeof(double) * (size_t)nn, cudaMemcpyHostToDevice, c->stream));
    const double* d_rhs = c->d_rhs;
    if (rhs_host) {
        if (c->rhs_n < nn) {
            cudaFree(c->d_rhs);
            c->d_rhs = nullptr;
            PG_CUDA(c, cudaMalloc((void**)&c->d_rhs, sizeof(double) * (size_t)nn));
            c->rhs_n = nn;
        }
        PG_CUDA(c, cudaMemcpyAsync(c->d_rhs, rhs_host, sizeof(double) * (size_t)nn, cudaMemcpyHostToDevice, c->stream));
        d_rhs = c->d_rhs;
    }
    PG_CUDA(c, cudaMemsetAsync(s->d_diag, 0, sizeof(double) * (size_t)(nn + 1), c->stream));
    PG_CUDA(c, cudaMemsetAsync(s->d_lower, 0, sizeof(double) * (size_t)(nn + 1), c->stream));
    PG_CUDA(c, cudaMemsetAsync(s->d_upper, 0, sizeof(double) * (size_t)(nn + 1), c->stream));
    poisson1d_assemble_kernel<<<(unsigned int)((nc + 127) / 128), 128, 0, c->stream>>>(nc, c->dev.cell_nodes, s->d_node_pos, c->d_node_coords,
                                                                                    s->d_cell_eps, s->d_diag, s->d_lower, s->d_upper);
    {
        const size_t smem = 4 * sizeof(double) * (size_t)nn;
        if (smem <= 200 * 1024) {
            if (smem > 48 * 1024) /* (per device: set where it is needed rather than remembered per process) */
                PG_CUDA(c, cudaFuncSetAttribute(poisson1d_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            poisson1d_solve_kernel<true><<<1, 256, smem, c->stream>>>(nn, s->d_pos_node, d_rhs, s->d_is_dir, s->d_dir, s->d_neumann,
                                                                     s->d_diag, s->d_lower, s->d_upper, s->d_work, s->d_phi);
        } else {
            poisson1d_solve_kernel<false><<<1, 256, 0, c->stream>>>(nn, s->d_pos_node, d_rhs, s->d_is_dir, s->d_dir, s->d_neumann,
                                                                   s->d_diag, s->d_lower, s->d_upper, s->d_work, s->d_phi);
        }
    }
    c->launches += 2;
    if (phi_host) PG_CUDA(c, cudaMemcpyAsync(phi_host, s->d_phi, sizeof(double) * (size_t)nn, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = efield_from_phi(c, s, s->d_phi, e_host))) return rc;
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}
