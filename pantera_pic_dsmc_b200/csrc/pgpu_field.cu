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
 *
 * 2-D (plane and axisymmetric) and 3-D unstructured meshes: pgpu_field_solve assembles the P1 stiffness matrix of
 * ASSEMBLE_POISSON (:256-424 triangles, :425-637 tetrahedra; identity rows for Dirichlet and unused nodes, :639-646,
 * :758-760) in CSR form -- every entry summed over its cells in ascending cell order, the order in which the reference's
 * MatSetValue(ADD_VALUES) calls accumulate it, so the matrix is the reference's bit for bit -- and solves it with a
 * Jacobi-preconditioned conjugate gradient on the device (the reference hands the system to PETSc's KSP, whose solver is
 * a run-time option: the answer is defined by the linear system and the tolerance, not by the method).  Dirichlet nodes
 * are held at their value and the iteration runs on the free nodes, where the matrix is symmetric positive definite.
 * Two kernels per iteration, every reduction in a fixed order (results do not depend on scheduling), the convergence
 * test on the device; the host reads one status word per batch of iterations.
 */
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

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
    /* --- conjugate-gradient solver for 2-D / 3-D meshes --- */
    bool cg_ready = false, phi_valid = false;
    int cg_nnz = 0, cg_blocks = 0;
    std::vector<int> bnd_cell, bnd_face; /* boundary faces with a physical group, in the reference's loop order */
    std::vector<unsigned char> h_is_dir;
    std::vector<double> h_neumann;
    int *d_row_ptr = nullptr, *d_col = nullptr, *d_diag_pos = nullptr, *d_contrib_off = nullptr, *d_contrib = nullptr;
    double *d_val = nullptr, *d_dinv = nullptr, *d_b = nullptr, *d_r = nullptr, *d_q = nullptr, *d_p[2] = {nullptr, nullptr},
           *d_part = nullptr, *d_extra = nullptr; /* d_extra: s (two buffers) and p of the cooperative solver */
    void* d_cgs = nullptr;
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
    cudaFree(s->d_row_ptr); cudaFree(s->d_col); cudaFree(s->d_diag_pos); cudaFree(s->d_contrib_off); cudaFree(s->d_contrib);
    cudaFree(s->d_val); cudaFree(s->d_dinv); cudaFree(s->d_b); cudaFree(s->d_r); cudaFree(s->d_q); cudaFree(s->d_p[0]); cudaFree(s->d_p[1]);
    cudaFree(s->d_part); cudaFree(s->d_cgs); cudaFree(s->d_extra);
    /* (d_e is owned by the context's e_field slot once handed over) */
    delete s;
    c->field = nullptr;
}

void pg_field_invalidate(pgpu_ctx* c) {
    if (!c->field) return;
    PgFieldState* s = (PgFieldState*)c->field;
    s->chain_ready = false;
    s->cg_ready = false;
    s->phi_valid = false;
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
    s->cg_ready = false; /* (EPS_REL enters the matrix, WALL_EFIELD the Neumann vector) */
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

/* ------------------------------------------------------------------------------------------
 * 2-D / 3-D: P1 stiffness matrix of ASSEMBLE_POISSON in CSR form + preconditioned conjugate gradient
 * ---------------------------------------------------------------------------------------- */
namespace {

#define CG_THREADS 256
#define CG_MAX_BLOCKS 592 /* 4 x 148 SMs; the partial sums of a reduction fit one block's second stage */

/* device-resident solver scalars */
struct CgScalars {
    double thresh; /* the iteration stops when |r|^2 <= thresh = max(rtol^2 |b~|^2, atol^2) */
    double bb;     /* |b~|^2, b~ = right-hand side of the free rows with the Dirichlet columns moved over */
    double rr;     /* |r|^2 when the iteration stopped / at the last look */
    int done, iters;
    int par, pad;
};

/* element (P, Q) of the stiffness matrix of triangle I, ref: src/fields.f90:332-372 (AREA = CELL_AREAS(I),
 * src/grid_and_partition.f90:1315) */
__device__ double stiff_tri(int i, int p, int q, const int* __restrict__ cell_nodes, const double* __restrict__ xyz, double eps_rel, int axi) {
    const int v1 = cell_nodes[3 * (size_t)i] - 1, v2 = cell_nodes[3 * (size_t)i + 1] - 1, v3 = cell_nodes[3 * (size_t)i + 2] - 1;
    const double x1 = xyz[3 * (size_t)v1], x2 = xyz[3 * (size_t)v2], x3 = xyz[3 * (size_t)v3];
    const double y1 = xyz[3 * (size_t)v1 + 1], y2 = xyz[3 * (size_t)v2 + 1], y3 = xyz[3 * (size_t)v3 + 1];
    const double area = 0.5 * fabs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));
    const int lo = p < q ? p : q, hi = p < q ? q : p;
    double k;
    if (lo == 0 && hi == 0) k = 0.25 * ((y2 - y3) * (y2 - y3) + (x2 - x3) * (x2 - x3)) / area * eps_rel;
    else if (lo == 1 && hi == 1) k = 0.25 * ((y1 - y3) * (y1 - y3) + (x1 - x3) * (x1 - x3)) / area * eps_rel;
    else if (lo == 2 && hi == 2) k = 0.25 * ((y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1)) / area * eps_rel;
    else if (lo == 0 && hi == 1) k = -(0.25 * ((y2 - y3) * (y1 - y3) + (x2 - x3) * (x1 - x3)) / area * eps_rel);
    else if (lo == 1 && hi == 2) k = 0.25 * ((y1 - y3) * (y2 - y1) + (x1 - x3) * (x2 - x1)) / area * eps_rel;
    else k = -(0.25 * ((y2 - y3) * (y2 - y1) + (x2 - x3) * (x2 - x1)) / area * eps_rel);
    if (axi) k = k * (y1 + y2 + y3) / 3.;
    return k;
}
/* tetrahedron, ref: src/fields.f90:560-577: KIJ = VOLUME*(grad psi_P . grad psi_Q)*EPS_REL */
__device__ double stiff_tet(int i, int p, int q, const double* __restrict__ basis, const double* __restrict__ cell_vol, double eps_rel) {
    const double* bp = basis + 16 * (size_t)i + 4 * p; /* BASIS_COEFFS(1:4, P, I) */
    const double* bq = basis + 16 * (size_t)i + 4 * q;
    return cell_vol[i] * (bp[0] * bq[0] + bp[1] * bq[1] + bp[2] * bq[2]) * eps_rel;
}

/* one thread per stored entry: the sum of its cell contributions in ascending cell order = the order of the reference's
 * MatSetValue(..., ADD_VALUES) calls */
__global__ void __launch_bounds__(128) cg_assemble_kernel(int nnz, int dims, int axi, const int* __restrict__ contrib_off,
                                                          const int* __restrict__ contrib, const int* __restrict__ cell_nodes,
                                                          const double* __restrict__ xyz, const double* __restrict__ basis,
                                                          const double* __restrict__ cell_vol, const double* __restrict__ cell_eps,
                                                          double* __restrict__ val) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    double a = 0.0;
    for (int t = contrib_off[e]; t < contrib_off[e + 1]; t++) {
        const int code = contrib[t], i = code >> 4, p = (code >> 2) & 3, q = code & 3;
        const double k = dims == 2 ? stiff_tri(i, p, q, cell_nodes, xyz, cell_eps[i], axi) : stiff_tet(i, p, q, basis, cell_vol, cell_eps[i]);
        a = a + k;
    }
    val[e] = a;
}
/* identity rows (src/fields.f90:758-760) and the Jacobi preconditioner */
__global__ void __launch_bounds__(128) cg_diag_kernel(int nn, const unsigned char* __restrict__ is_dir, const int* __restrict__ diag_pos,
                                                      double* __restrict__ val, double* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    if (is_dir[i]) {
        val[diag_pos[i]] = 1.0;
        dinv[i] = 0.0; /* z = p = 0 on Dirichlet rows: the iteration never moves them */
    } else {
        dinv[i] = 1.0 / val[diag_pos[i]];
    }
}

/* sum over the block, the same value in every thread; fixed order */
__device__ double cg_block_sum(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads(); /* (sm may still be read from the previous call) */
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < CG_THREADS / 32; w++) t += sm[w];
    return t;
}
__device__ double cg_total(const double* part, int nb, double* sm) {
    double v = 0.0;
    for (int k = threadIdx.x; k < nb; k += CG_THREADS) v += part[k];
    return cg_block_sum(v, sm);
}

/* partial-sum rows of d_part */
enum { CGP_RZ0 = 0, CGP_RZ1 = 1, CGP_RR0 = 2, CGP_RR1 = 3, CGP_PQ = 4, CGP_BB = 5, CGP_ROWS = 6 };

/* right-hand side rows (src/fields.f90:5538-5547) and the starting vector */
__global__ void __launch_bounds__(CG_THREADS) cg_setup_kernel(int nn, const unsigned char* __restrict__ is_dir, const double* __restrict__ dir,
                                                              const double* __restrict__ neumann, const double* __restrict__ rhs, int warm,
                                                              double* __restrict__ b, double* __restrict__ x, double* __restrict__ p0,
                                                              double* __restrict__ p1) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
        if (is_dir[i]) {
            b[i] = dir[i];
            x[i] = dir[i];
        } else {
            b[i] = rhs[i] + neumann[i];
            if (!warm) x[i] = 0.0;
        }
        p0[i] = 0.0;
        p1[i] = 0.0;
    }
}
/* r = b - A x on the free rows, z = r / diag, and |b~|^2 with b~ = b - A_fd x_d */
__global__ void __launch_bounds__(CG_THREADS) cg_residual_kernel(int nn, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                                 const double* __restrict__ val, const unsigned char* __restrict__ is_dir,
                                                                 const double* __restrict__ dinv, const double* __restrict__ b,
                                                                 const double* __restrict__ x, double* __restrict__ r, double* __restrict__ part) {
    __shared__ double sm[CG_THREADS / 32];
    double rz = 0.0, rr = 0.0, bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
        double ri = 0.0;
        if (!is_dir[i]) {
            double ax = 0.0, ad = 0.0;
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
                const int j = col[e];
                const double t = val[e] * x[j];
                ax += t;
                if (is_dir[j]) ad += t;
            }
            ri = b[i] - ax;
            const double bt = b[i] - ad;
            bb += bt * bt;
            rz += ri * (dinv[i] * ri);
            rr += ri * ri;
        }
        r[i] = ri;
    }
    rz = cg_block_sum(rz, sm);
    rr = cg_block_sum(rr, sm);
    bb = cg_block_sum(bb, sm);
    if (threadIdx.x == 0) {
        part[CGP_RZ0 * CG_MAX_BLOCKS + blockIdx.x] = rz;
        part[CGP_RR0 * CG_MAX_BLOCKS + blockIdx.x] = rr;
        part[CGP_BB * CG_MAX_BLOCKS + blockIdx.x] = bb;
    }
}
__global__ void __launch_bounds__(CG_THREADS) cg_thresh_kernel(int nb, double rtol, double atol, const double* __restrict__ part, CgScalars* sc) {
    __shared__ double sm[CG_THREADS / 32];
    const double bb = cg_total(part + CGP_BB * CG_MAX_BLOCKS, nb, sm);
    if (threadIdx.x == 0) {
        const double t = rtol * rtol * bb, a = atol * atol;
        sc->bb = bb;
        sc->thresh = t > a ? t : a;
        sc->rr = 0.0;
        sc->done = 0;
        sc->iters = 0;
    }
}

/* first half of iteration k: convergence test, beta, p = z + beta p, q = A p, partial p.q */
__global__ void __launch_bounds__(CG_THREADS) cg_direction_kernel(int nn, int k, int nb, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                                  const double* __restrict__ val, const double* __restrict__ dinv,
                                                                  const double* __restrict__ r, const double* __restrict__ p_old,
                                                                  double* __restrict__ p_new, double* __restrict__ q, double* __restrict__ part,
                                                                  CgScalars* sc) {
    __shared__ double sm[CG_THREADS / 32];
    if (sc->done) return;
    const int par = k & 1;
    const double rr = cg_total(part + (CGP_RR0 + par) * CG_MAX_BLOCKS, nb, sm);
    if (rr <= sc->thresh) { /* every block takes the same decision from the same sums */
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sc->done = 1;
            sc->iters = k;
            sc->rr = rr;
        }
        return;
    }
    double beta = 0.0;
    if (k > 0) {
        const double rz_new = cg_total(part + (CGP_RZ0 + par) * CG_MAX_BLOCKS, nb, sm);
        const double rz_old = cg_total(part + (CGP_RZ0 + (par ^ 1)) * CG_MAX_BLOCKS, nb, sm);
        beta = rz_new / rz_old;
    }
    double pq = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
        double pi = 0.0, qi = 0.0;
        if (dinv[i] != 0.0) {
            pi = dinv[i] * r[i] + beta * p_old[i];
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
                const int j = col[e];
                qi += val[e] * (dinv[j] * r[j] + beta * p_old[j]);
            }
            pq += pi * qi;
        }
        p_new[i] = pi;
        q[i] = qi;
    }
    pq = cg_block_sum(pq, sm);
    if (threadIdx.x == 0) part[CGP_PQ * CG_MAX_BLOCKS + blockIdx.x] = pq;
}
/* second half: alpha, x += alpha p, r -= alpha q, partial r.z and r.r for iteration k + 1 */
__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(int nn, int k, int nb, const double* __restrict__ dinv, const double* __restrict__ p,
                                                               const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r,
                                                               double* __restrict__ part, CgScalars* sc) {
    __shared__ double sm[CG_THREADS / 32];
    if (sc->done) return;
    const int par = k & 1;
    const double rz = cg_total(part + (CGP_RZ0 + par) * CG_MAX_BLOCKS, nb, sm);
    const double pq = cg_total(part + CGP_PQ * CG_MAX_BLOCKS, nb, sm);
    const double alpha = rz / pq;
    double rz2 = 0.0, rr2 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
        if (dinv[i] != 0.0) {
            x[i] = x[i] + alpha * p[i];
            const double ri = r[i] - alpha * q[i];
            r[i] = ri;
            rz2 += ri * (dinv[i] * ri);
            rr2 += ri * ri;
        }
    }
    rz2 = cg_block_sum(rz2, sm);
    rr2 = cg_block_sum(rr2, sm);
    if (threadIdx.x == 0) {
        part[(CGP_RZ0 + (par ^ 1)) * CG_MAX_BLOCKS + blockIdx.x] = rz2;
        part[(CGP_RR0 + (par ^ 1)) * CG_MAX_BLOCKS + blockIdx.x] = rr2;
    }
}
/* the look after the last iteration of a batch (no further direction kernel would test the residual) */
__global__ void __launch_bounds__(CG_THREADS) cg_check_kernel(int k, int nb, const double* __restrict__ part, CgScalars* sc) {
    __shared__ double sm[CG_THREADS / 32];
    if (sc->done) return;
    const double rr = cg_total(part + (CGP_RR0 + (k & 1)) * CG_MAX_BLOCKS, nb, sm);
    if (threadIdx.x == 0) {
        sc->rr = rr;
        sc->iters = k;
        if (rr <= sc->thresh) sc->done = 1;
    }
}

/* two totals in one pass (same fixed order as cg_total) */
__device__ void cg_total2(const double* pa, const double* pb, int nb, double* sm, double& ta, double& tb) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nb; k += CG_THREADS) {
        a += pa[k];
        b += pb[k];
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
        sm[warp] = a;
        sm[CG_THREADS / 32 + warp] = b;
    }
    __syncthreads();
    ta = 0.0;
    tb = 0.0;
    for (int w = 0; w < CG_THREADS / 32; w++) {
        ta += sm[w];
        tb += sm[CG_THREADS / 32 + w];
    }
}

/* The whole solve in ONE cooperative launch with ONE grid-wide barrier per iteration: the conjugate gradient in the
 * form of Chronopoulos and Gear (both inner products of an iteration are taken together, alpha follows from a
 * recurrence), with the vector updates folded into the row loop of the next matrix product.  Given alpha, beta:
 *     s' = w + beta s      r' = r - alpha s'      p = u + beta p      x = x + alpha p        (u = r / diag)
 *     w' = A u'  with  u'_j = (r_j - alpha (w_j + beta s_j)) / diag_j  recomputed for the neighbours j of the row
 *     gamma' = r'.u'   delta = w'.u'   |r'|^2  ->  barrier  ->  beta' = gamma'/gamma,  alpha' = gamma'/(delta - beta' gamma'/alpha)
 * r, w, s are double-buffered (a row reads its neighbours' old values while their owners write the new ones), and so
 * are the per-block partial sums (a fast block may start the next iteration while a slow one still adds up this
 * one's).  The first pass runs with alpha = beta = 0: it leaves x and r alone and produces w0 = A u0, gamma0, delta0.
 * Every block takes the stop decision from the same sums, so all blocks leave the loop together.
 * (The kernel-per-half-iteration path below keeps the textbook recurrences: 2 launches per iteration.) */
enum { CGQ_G = 0, CGQ_D = 1, CGQ_R = 2 }; /* partial rows: 2 parities x {gamma, delta, |r|^2} */
__global__ void __launch_bounds__(CG_THREADS) cg_coop_kernel(int nn, int max_iter, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                             const double* __restrict__ val, const double* __restrict__ dinv, double* r0,
                                                             double* r1, double* w0, double* w1, double* s0, double* s1, double* p, double* x,
                                                             double* part, CgScalars* sc) {
    /* (r, w, s, p, x and the partial sums are written by other blocks between the barriers: plain pointers, coherent loads) */
    namespace cgr = cooperative_groups;
    cgr::grid_group grid = cgr::this_grid();
    __shared__ double sm[3 * (CG_THREADS / 32)];
    const int nb = gridDim.x, stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double thresh = sc->thresh;
    double alpha = 0.0, beta = 0.0, gamma_old = 0.0;
    for (int pass = 0;; pass++) { /* pass 0 is the set-up pass; pass k >= 1 applies the k-th update of x */
        const int par = pass & 1;
        const double* r_old = par ? r1 : r0;
        const double* w_old = par ? w1 : w0;
        const double* s_old = par ? s1 : s0;
        double* r_new = par ? r0 : r1;
        double* w_new = par ? w0 : w1;
        double* s_new = par ? s0 : s1;
        double g = 0.0, dl = 0.0, rr = 0.0;
        for (int i = first; i < nn; i += stride) {
            double ri = 0.0, wi = 0.0, si = 0.0;
            const double di = dinv[i];
            if (di != 0.0) {
                si = w_old[i] + beta * s_old[i];
                const double pi = di * r_old[i] + beta * p[i];
                p[i] = pi;
                x[i] = x[i] + alpha * pi;
                ri = r_old[i] - alpha * si;
                for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
                    const int j = col[e];
                    wi += val[e] * (dinv[j] * (r_old[j] - alpha * (w_old[j] + beta * s_old[j])));
                }
                const double ui = di * ri;
                g += ri * ui;
                dl += wi * ui;
                rr += ri * ri;
            }
            r_new[i] = ri;
            w_new[i] = wi;
            s_new[i] = si;
        }
        for (int o = 16; o > 0; o >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, o);
            dl += __shfl_xor_sync(0xffffffffu, dl, o);
            rr += __shfl_xor_sync(0xffffffffu, rr, o);
        }
        if (lane == 0) {
            sm[warp] = g;
            sm[CG_THREADS / 32 + warp] = dl;
            sm[2 * (CG_THREADS / 32) + warp] = rr;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0, c3 = 0.0;
            for (int w = 0; w < CG_THREADS / 32; w++) {
                a += sm[w];
                b += sm[CG_THREADS / 32 + w];
                c3 += sm[2 * (CG_THREADS / 32) + w];
            }
            double* row = part + (size_t)(3 * par) * CG_MAX_BLOCKS;
            row[CGQ_G * CG_MAX_BLOCKS + blockIdx.x] = a;
            row[CGQ_D * CG_MAX_BLOCKS + blockIdx.x] = b;
            row[CGQ_R * CG_MAX_BLOCKS + blockIdx.x] = c3;
        }
        grid.sync();
        /* totals: every block adds the same partial sums in the same order */
        {
            const double* row = part + (size_t)(3 * par) * CG_MAX_BLOCKS;
            double a = 0.0, b = 0.0, c3 = 0.0;
            for (int k = threadIdx.x; k < nb; k += CG_THREADS) {
                a += row[CGQ_G * CG_MAX_BLOCKS + k];
                b += row[CGQ_D * CG_MAX_BLOCKS + k];
                c3 += row[CGQ_R * CG_MAX_BLOCKS + k];
            }
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
                c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            __syncthreads(); /* (sm of the block sums above has been read by thread 0) */
            if (lane == 0) {
                sm[warp] = a;
                sm[CG_THREADS / 32 + warp] = b;
                sm[2 * (CG_THREADS / 32) + warp] = c3;
            }
            __syncthreads();
            g = 0.0; dl = 0.0; rr = 0.0;
            for (int w = 0; w < CG_THREADS / 32; w++) {
                g += sm[w];
                dl += sm[CG_THREADS / 32 + w];
                rr += sm[2 * (CG_THREADS / 32) + w];
            }
            __syncthreads(); /* (before the next pass overwrites sm) */
        }
        if (rr <= thresh || pass >= max_iter) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                sc->done = rr <= thresh ? 1 : 0;
                sc->iters = pass;
                sc->rr = rr;
                sc->par = par ^ 1; /* the buffer that holds the final r */
            }
            return;
        }
        if (pass == 0) {
            beta = 0.0;
            alpha = g / dl;
        } else {
            beta = g / gamma_old;
            alpha = g / (dl - beta * g / alpha);
        }
        gamma_old = g;
    }
}

/* nodes of face J (0-based local numbers), ref: src/fields.f90:150-176 (1-D), :268-287 (2-D), :441-477 (3-D) */
const int FACE_NODES_1D[2][1] = {{0}, {1}};
const int FACE_NODES_2D[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int FACE_NODES_3D[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};

double host_norm2_3(const double* x) { /* NORM2 as gfortran expands it inline (scaled accumulation) */
    double result = 0.0, scale = 1.0;
    for (int i = 0; i < 3; i++) {
        if (x[i] != 0.0) {
            const double ax = fabs(x[i]);
            if (scale < ax) {
                const double v = scale / ax;
                result = 1.0 + result * v * v;
                scale = ax;
            } else {
                const double v = ax / scale;
                result += v * v;
            }
        }
    }
    return scale * sqrt(result);
}

} /* namespace */

template <class T>
static cudaError_t cg_upload(T** d, const std::vector<T>& h) {
    cudaFree(*d);
    *d = nullptr;
    cudaError_t e = cudaMalloc((void**)d, sizeof(T) * std::max<size_t>(h.size(), 1));
    if (e != cudaSuccess) return e;
    return h.empty() ? cudaSuccess : cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
}

/* Set-up, once per mesh / set of boundary types: which nodes are Dirichlet (src/fields.f90:262-305, 431-523, 639-646),
 * the Neumann vector (:391-421, :579-633), the CSR pattern with, per entry, the list of its (cell, P, Q) contributions in
 * ascending cell order, the matrix values and the Jacobi preconditioner (kernels). */
static int build_cg(pgpu_ctx* c, PgFieldState* s) {
    const int nn = c->nn, nc = c->nc, d = c->mesh_dims, nf = d + 1;
    if (nc >= (1 << 27)) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: more than 2^27 cells");
    auto face_node = [&](int j, int t) { return d == 1 ? FACE_NODES_1D[j][t] : d == 2 ? FACE_NODES_2D[j][t] : FACE_NODES_3D[j][t]; };
    s->h_is_dir.assign((size_t)nn, 0);
    s->h_neumann.assign((size_t)nn, 0.0);
    s->bnd_cell.clear();
    s->bnd_face.clear();
    std::vector<unsigned char> unused((size_t)nn, 1);
    for (int i = 0; i < nc; i++) {
        for (int p = 0; p < nf; p++) unused[c->h_cell_nodes[(size_t)nf * i + p] - 1] = 0;
        for (int j = 0; j < nf; j++) {
            const int pg = c->h_cell_faces_pg[(size_t)nf * i + j];
            if (pg == -1) continue;
            if (pg < 1 || pg > (int)c->h_bc_field.size()) PG_FAIL(c, PGPU_ERR_INVALID, "field_solve: face group %d without boundary condition", pg);
            s->bnd_cell.push_back(i);
            s->bnd_face.push_back(j);
            const int fbc = c->h_bc_field[pg - 1];
            if (fbc == PGPU_DIRICHLET_BC || fbc == PGPU_RF_VOLTAGE_BC || fbc == PGPU_DECOUPLED_RF_VOLTAGE_BC)
                for (int t = 0; t < d; t++) s->h_is_dir[c->h_cell_nodes[(size_t)nf * i + face_node(j, t)] - 1] = 1;
        }
    }
    for (int v = 0; v < nn; v++)
        if (unused[v]) s->h_is_dir[v] = 1; /* IS_UNUSED -> Dirichlet 0, :639-646 */
    /* NEUMANN(:), in the reference's loop order (cells ascending, faces ascending) */
    for (size_t k = 0; k < s->bnd_cell.size(); k++) {
        const int i = s->bnd_cell[k], j = s->bnd_face[k], pg = c->h_cell_faces_pg[(size_t)nf * i + j];
        if (c->h_bc_field[pg - 1] != PGPU_NEUMANN_BC) continue;
        const double ef = s->wall_efield[pg - 1];
        double add;
        if (d == 2) {
            const int va = c->h_cell_nodes[3 * (size_t)i + FACE_NODES_2D[j][0]] - 1, vb = c->h_cell_nodes[3 * (size_t)i + FACE_NODES_2D[j][1]] - 1;
            const double dx = c->h_node_coords[3 * (size_t)vb] - c->h_node_coords[3 * (size_t)va];
            const double dy = c->h_node_coords[3 * (size_t)vb + 1] - c->h_node_coords[3 * (size_t)va + 1];
            const double edgelength = sqrt(dx * dx + dy * dy); /* J=1: (X2-X1, Y2-Y1), J=2: (X3-X2, Y3-Y2), J=3: (X1-X3, Y1-Y3) */
            add = ef * 0.5 * edgelength;
        } else {
            const double* A = &c->h_node_coords[3 * (size_t)(c->h_cell_nodes[4 * (size_t)i + FACE_NODES_3D[j][0]] - 1)];
            const double* B = &c->h_node_coords[3 * (size_t)(c->h_cell_nodes[4 * (size_t)i + FACE_NODES_3D[j][1]] - 1)];
            const double* C = &c->h_node_coords[3 * (size_t)(c->h_cell_nodes[4 * (size_t)i + FACE_NODES_3D[j][2]] - 1)];
            const double u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, w[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
            const double cr[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
            const double facearea = 0.5 * host_norm2_3(cr); /* FACE_AREA, src/grid_and_partition.f90:1957 */
            add = ef * facearea / 3.;
        }
        for (int t = 0; t < d; t++) {
            const int v = c->h_cell_nodes[(size_t)nf * i + face_node(j, t)] - 1;
            if (!s->h_is_dir[v]) s->h_neumann[v] = s->h_neumann[v] + add;
        }
    }
    /* CSR pattern: one key (row, column) per (cell, P, Q) with a free row, sorted by key then cell */
    struct Item { unsigned long long key; int code; };
    std::vector<Item> items;
    items.reserve((size_t)nc * nf * nf + (size_t)nn);
    for (int i = 0; i < nc; i++)
        for (int p = 0; p < nf; p++) {
            const int vp = c->h_cell_nodes[(size_t)nf * i + p] - 1;
            if (s->h_is_dir[vp]) continue;
            for (int q = 0; q < nf; q++) {
                const int vq = c->h_cell_nodes[(size_t)nf * i + q] - 1;
                items.push_back({((unsigned long long)vp << 32) | (unsigned int)vq, (i << 4) | (p << 2) | q});
            }
        }
    for (int v = 0; v < nn; v++)
        if (s->h_is_dir[v]) items.push_back({((unsigned long long)v << 32) | (unsigned int)v, -1}); /* identity row */
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : a.code < b.code; });
    std::vector<int> row_ptr((size_t)nn + 1, 0), col, contrib_off, contrib, diag_pos((size_t)nn, -1);
    for (size_t k = 0; k < items.size();) {
        size_t e = k;
        const int row = (int)(items[k].key >> 32), cl = (int)(items[k].key & 0xffffffffULL);
        contrib_off.push_back((int)contrib.size());
        while (e < items.size() && items[e].key == items[k].key) {
            if (items[e].code >= 0) contrib.push_back(items[e].code);
            e++;
        }
        if (row == cl) diag_pos[row] = (int)col.size();
        col.push_back(cl);
        row_ptr[(size_t)row + 1]++;
        k = e;
    }
    contrib_off.push_back((int)contrib.size());
    for (int v = 0; v < nn; v++) {
        if (diag_pos[v] < 0) PG_FAIL(c, PGPU_ERR_STATE, "field_solve: node %d has no diagonal entry", v + 1);
        row_ptr[(size_t)v + 1] += row_ptr[v];
    }
    const int nnz = (int)col.size();
    std::vector<double> eps((size_t)nc, 1.0);
    if (!c->h_cell_pg.empty())
        for (int i = 0; i < nc; i++) {
            const int pg = c->h_cell_pg[i];
            if (pg != -1 && pg >= 1 && pg <= (int)s->eps_rel.size()) eps[i] = s->eps_rel[pg - 1];
        }
    PG_CUDA(c, cg_upload(&s->d_row_ptr, row_ptr));
    PG_CUDA(c, cg_upload(&s->d_col, col));
    PG_CUDA(c, cg_upload(&s->d_diag_pos, diag_pos));
    PG_CUDA(c, cg_upload(&s->d_contrib_off, contrib_off));
    PG_CUDA(c, cg_upload(&s->d_contrib, contrib));
    PG_CUDA(c, cg_upload(&s->d_cell_eps, eps));
    PG_CUDA(c, cg_upload(&s->d_is_dir, s->h_is_dir));
    PG_CUDA(c, cg_upload(&s->d_neumann, s->h_neumann));
    auto dalloc = [&](double** p, size_t n) { cudaFree(*p); *p = nullptr; return cudaMalloc((void**)p, sizeof(double) * std::max<size_t>(n, 1)); };
    PG_CUDA(c, dalloc(&s->d_val, (size_t)nnz));
    for (double** p : {&s->d_dinv, &s->d_b, &s->d_r, &s->d_q, &s->d_p[0], &s->d_p[1], &s->d_dir, &s->d_phi}) PG_CUDA(c, dalloc(p, (size_t)nn + 1));
    PG_CUDA(c, dalloc(&s->d_part, (size_t)CGP_ROWS * CG_MAX_BLOCKS));
    PG_CUDA(c, dalloc(&s->d_extra, 3 * ((size_t)nn + 1)));
    cudaFree(s->d_cgs);
    s->d_cgs = nullptr;
    PG_CUDA(c, cudaMalloc(&s->d_cgs, sizeof(CgScalars)));
    PG_CUDA(c, cudaMemsetAsync(s->d_phi, 0, sizeof(double) * ((size_t)nn + 1), c->stream));
    cg_assemble_kernel<<<(unsigned int)((nnz + 127) / 128), 128, 0, c->stream>>>(nnz, d, c->cfg.axi, s->d_contrib_off, s->d_contrib, c->dev.cell_nodes,
                                                                               c->d_node_coords, c->dev.basis, c->dev.cell_vol, s->d_cell_eps, s->d_val);
    cg_diag_kernel<<<(unsigned int)((nn + 127) / 128), 128, 0, c->stream>>>(nn, s->d_is_dir, s->d_diag_pos, s->d_val, s->d_dinv);
    c->launches += 2;
    PG_CUDA(c, cudaGetLastError());
    s->cg_nnz = nnz;
    s->cg_blocks = std::min((nn + CG_THREADS - 1) / CG_THREADS, CG_MAX_BLOCKS);
    s->cg_ready = true;
    s->chain_ready = false; /* (d_is_dir, d_neumann, d_dir, d_phi, d_cell_eps were re-allocated with this mesh's sizes) */
    s->phi_valid = false;
    return PGPU_OK;
}

/* SET_WALL_POTENTIAL + Poisson solve + COMPUTE_E_FIELD on any unstructured mesh.  1-D: the exact tridiagonal path above.
 * 2-D / 3-D: Jacobi-preconditioned conjugate gradient on the free nodes until |r| <= max(rtol |b~|, 1e-50) or max_iter
 * iterations (rtol <= 0: 1e-5, max_iter <= 0: 10000 -- PETSc's KSP defaults, which the reference runs with unless the
 * user passes options); the previous potential is the starting vector.  iters_out / relres_out (optional): iterations
 * used and |r| / |b~| reached; like the reference (which prints KSPConvergedReason and goes on) a solve that ran out of
 * iterations is not an error: the caller compares relres with its tolerance. */
extern "C" int pgpu_field_solve(pgpu_ctx* c, int32_t tID, const double* rhs_host, double* phi_host, double* e_host, double rtol,
                                int32_t max_iter, int32_t* iters_out, double* relres_out) {
    if (!c) return PGPU_ERR_INVALID;
    if (iters_out) *iters_out = 0;
    if (relres_out) *relres_out = 0.0;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (c->cfg.grid_type != PGPU_UNSTRUCTURED) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: unstructured meshes only (the rectilinear stencil stays on the host)");
    if (c->mesh_dims == 1) {
        if ((rc = pgpu_field_solve_1d(c, tID, rhs_host, phi_host, e_host))) return rc;
        if (iters_out) *iters_out = 1;
        return PGPU_OK;
    }
    if (c->cfg.pic_type != PGPU_PIC_NONE && c->cfg.pic_type != PGPU_PIC_EXPLICIT)
        PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "field_solve: PIC_TYPE %d (only the explicit scheme's matrix is assembled)", c->cfg.pic_type);
    if (c->mesh_dims == 3 && (!c->dev.basis || !c->dev.cell_vol)) PG_FAIL(c, PGPU_ERR_STATE, "field_solve: mesh without BASIS_COEFFS / CELL_VOLUMES");
    PgFieldState* s = fld(c);
    const int nn = c->nn, nf = c->mesh_dims + 1, d = c->mesh_dims, nbc = (int)c->h_bc_field.size();
    if ((int)s->wall_potential.size() < nbc) {
        s->wall_potential.resize(nbc, 0.0); s->wall_rf_potential.resize(nbc, 0.0); s->rf_frequency.resize(nbc, 0.0);
        s->wall_efield.resize(nbc, 0.0); s->eps_rel.resize(nbc, 1.0);
        s->cg_ready = false;
    }
    if (!s->cg_ready && (rc = build_cg(c, s))) return rc;
    if (!rhs_host && (!c->d_rhs || c->rhs_n < nn)) PG_FAIL(c, PGPU_ERR_STATE, "field_solve: no deposited RHS on the device (call pgpu_deposit_charge or pass rhs)");
    if (rtol <= 0.) rtol = 1e-5;
    if (max_iter <= 0) max_iter = 10000;
    /* SET_WALL_POTENTIAL at timestep tID (src/fields.f90:5609-5711): the last boundary face in loop order sets a node */
    std::vector<double> dir((size_t)nn, 0.0);
    bool any_dir = false;
    const double PI = 3.1415926535897932;
    for (size_t k = 0; k < s->bnd_cell.size(); k++) {
        const int i = s->bnd_cell[k], j = s->bnd_face[k], pg = c->h_cell_faces_pg[(size_t)nf * i + j], fbc = c->h_bc_field[pg - 1];
        if (!(fbc == PGPU_DIRICHLET_BC || fbc == PGPU_RF_VOLTAGE_BC || fbc == PGPU_DECOUPLED_RF_VOLTAGE_BC)) continue;
        double pot = s->wall_potential[pg - 1];
        if (fbc != PGPU_DIRICHLET_BC) pot = s->wall_potential[pg - 1] + 0.5 * s->wall_rf_potential[pg - 1] * cos(2 * PI * s->rf_frequency[pg - 1] * tID * c->cfg.dt);
        for (int t = 0; t < d; t++) {
            const int lf = d == 2 ? FACE_NODES_2D[j][t] : FACE_NODES_3D[j][t];
            dir[c->h_cell_nodes[(size_t)nf * i + lf] - 1] = pot;
        }
        any_dir = true;
    }
    if (!any_dir) PG_FAIL(c, PGPU_ERR_INVALID, "field_solve: no Dirichlet / RF-voltage boundary: the Poisson problem is singular");
    PG_CUDA(c, cudaMemcpyAsync(s->d_dir, dir.data(), sizeof(double) * (size_t)nn, cudaMemcpyHostToDevice, c->stream));
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
    /* One cooperative launch when the device can hold the whole grid at once (PANTERA_GPU_CG_COOP=0: the
     * kernel-per-half-iteration path below, also the fall-back).  Fewer, fuller blocks make the grid barriers cheaper. */
    int nb = s->cg_blocks, coop = 0;
    {
        const char* e = getenv("PANTERA_GPU_CG_COOP");
        int supported = 0, per_sm = 0, sms = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, c->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        if (supported && !(e && e[0] == '0') &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_coop_kernel, CG_THREADS, 0) == cudaSuccess && per_sm > 0) {
            coop = 1;
            nb = std::min(nb, std::min(2, per_sm) * sms);
        }
    }
    CgScalars* sc = (CgScalars*)s->d_cgs;
    double* x = s->d_phi;
    cg_setup_kernel<<<nb, CG_THREADS, 0, c->stream>>>(nn, s->d_is_dir, s->d_dir, s->d_neumann, d_rhs, s->phi_valid ? 1 : 0, s->d_b, x, s->d_p[0], s->d_p[1]);
    cg_residual_kernel<<<nb, CG_THREADS, 0, c->stream>>>(nn, s->d_row_ptr, s->d_col, s->d_val, s->d_is_dir, s->d_dinv, s->d_b, x, s->d_r, s->d_part);
    cg_thresh_kernel<<<1, CG_THREADS, 0, c->stream>>>(nb, rtol, 1e-50, s->d_part, sc);
    c->launches += 3;
    CgScalars h;
    memset(&h, 0, sizeof(h));
    const int BATCH = 64;
    int k = 0;
    if (coop) {
        /* r: {d_r (filled by cg_residual_kernel), d_q}; w: {d_p[0] (zero), d_p[1]}; s: d_extra[0] (zero), d_extra[1]; p: d_extra[2] (zero) */
        int nn_ = nn, maxit_ = max_iter;
        double *r0 = s->d_r, *r1 = s->d_q, *w0 = s->d_p[0], *w1 = s->d_p[1], *s0 = s->d_extra, *s1 = s->d_extra + ((size_t)nn + 1),
               *pp = s->d_extra + 2 * ((size_t)nn + 1);
        PG_CUDA(c, cudaMemsetAsync(s->d_extra, 0, sizeof(double) * 3 * ((size_t)nn + 1), c->stream));
        void* args[] = {&nn_, &maxit_, &s->d_row_ptr, &s->d_col, &s->d_val, &s->d_dinv, &r0, &r1, &w0, &w1, &s0, &s1, &pp, &x, &s->d_part, &sc};
        PG_CUDA(c, cudaLaunchCooperativeKernel((const void*)cg_coop_kernel, dim3((unsigned int)nb), dim3(CG_THREADS), args, 0, c->stream));
        c->launches++;
        PG_CUDA(c, cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        k = max_iter; /* (skips the loop below) */
    }
    while (k < max_iter) {
        const int kend = std::min(k + BATCH, (int)max_iter);
        c->launches += 2 * (kend - k) + 1;
        for (; k < kend; k++) {
            /* iteration k reads the direction of iteration k-1 from d_p[k & 1] and writes its own to d_p[(k & 1) ^ 1] */
            cg_direction_kernel<<<nb, CG_THREADS, 0, c->stream>>>(nn, k, nb, s->d_row_ptr, s->d_col, s->d_val, s->d_dinv, s->d_r, s->d_p[k & 1],
                                                                 s->d_p[(k & 1) ^ 1], s->d_q, s->d_part, sc);
            cg_update_kernel<<<nb, CG_THREADS, 0, c->stream>>>(nn, k, nb, s->d_dinv, s->d_p[(k & 1) ^ 1], s->d_q, x, s->d_r, s->d_part, sc);
        }
        cg_check_kernel<<<1, CG_THREADS, 0, c->stream>>>(k, nb, s->d_part, sc);
        PG_CUDA(c, cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h.done) break;
    }
    PG_CUDA(c, cudaGetLastError());
    if (!(h.rr == h.rr)) PG_FAIL(c, PGPU_ERR_STATE, "field_solve: the conjugate gradient broke down (NaN residual): is every part of the mesh tied to a Dirichlet node?");
    s->phi_valid = true;
    if (iters_out) *iters_out = h.iters;
    if (relres_out) *relres_out = h.bb > 0. ? sqrt(h.rr / h.bb) : sqrt(h.rr);
    if (phi_host) PG_CUDA(c, cudaMemcpyAsync(phi_host, x, sizeof(double) * (size_t)nn, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = efield_from_phi(c, s, x, e_host))) return rc;
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}
