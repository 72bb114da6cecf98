/*
 * pgpu_restart.cu -- the geometric re-binning of restarted particles on the device.
 *
 * Replaces: REASSIGN_PARTICLES_TO_CELLS_1D / _2D / _3D (src/tools.f90:925-1039, 782-922, 577-778), which PROGRAM
 * PANTERA calls right after READ_PARTICLES_FILE when RESTART_TIMESTEP > 0 (src/pantera.f90:58-70): the IC of the dump
 * is not trusted (the restart may use another mesh); every particle is located again from its position.
 *
 * The reference's search structure: the domain box [XMIN, XMAX] x [YMIN, YMAX] x [ZMIN, ZMAX] (Domain_limits: of the
 * input file) is cut into N_PARTITIONS = 100 slabs per direction; every unstructured cell is entered in all the bins
 * its bounding box overlaps, in ascending cell order; a particle looks up its bin and takes the FIRST cell of the list
 * whose P1 basis functions are all non-negative at the particle position
 *     PSIP = XP*BASIS_COEFFS(1,VP,IC) + YP*BASIS_COEFFS(2,VP,IC) + ZP*BASIS_COEFFS(3,VP,IC) + BASIS_COEFFS(4,VP,IC)
 * (`IF (PSIP < 0) INSIDE = .FALSE.`), i.e. the LOWEST cell id among the cells of the bin that contain it -- a point on
 * a shared face belongs to the lower-numbered neighbour.  A particle outside the box or inside no cell is removed.
 *
 * Here: a kernel counts the bins of every cell, a scan turns the counts into offsets, a second pass fills the lists
 * (arrival order: atomics), and one thread per particle walks its bin keeping the minimum cell id that contains the
 * particle -- the same answer as the ordered list without sorting it.  Same FP64 expression, no FMA (-fmad=false), so
 * the cell indices are bit-exact with the reference's.
 *
 * Deviation, stated: the reference indexes its bin arrays with INT((C - XMIN)/DX) + 1 unclamped, so a cell (or a
 * particle) that touches XMAX exactly addresses bin 101 of 100 -- out of bounds in the Fortran.  Here bin indices are
 * clamped to 1..100 (what the reference computes whenever it stays inside its arrays).
 */
#include <cub/cub.cuh>

#include "pgpu_internal.cuh"

namespace {

#define RB_PARTS 100 /* N_PARTITIONS_X/Y/Z, ref: src/tools.f90:592-594, 798-799, 940 */

struct RbGrid {
    int dims;
    int n[3];
    double lo[3], hi[3], h[3];
};

/* INT((C - XMIN)/DX) + 1 (Fortran INT truncates toward zero), clamped to the array 1..n */
__device__ __forceinline__ int rb_index(double c, double lo, double h, int n) {
    const double q = (c - lo) / h;
    if (!(q < (double)n)) return n;
    if (q < 1.0) return 1;
    return (int)q + 1;
}
__device__ __forceinline__ unsigned int rb_bin(const RbGrid& g, int i, int j, int k) {
    /* the order in which the reference lays out IOFTETRA: I outermost, K innermost (src/tools.f90:655-666) */
    return (unsigned int)(((i - 1) * g.n[1] + (j - 1)) * g.n[2] + (k - 1));
}

/* one thread per cell: the bins its bounding box overlaps.  FILL = false counts, FILL = true enters the cell. */
template <bool FILL>
__global__ void __launch_bounds__(128) rb_cells_kernel(const RbGrid g, int nc, int nf, const int* __restrict__ cell_nodes,
                                                       const double* __restrict__ node_coords, unsigned int* __restrict__ count,
                                                       const unsigned int* __restrict__ off, int* __restrict__ cells) {
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= nc) return;
    int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
    for (int k = 0; k < g.dims; k++) {
        double cmin = 0., cmax = 0.;
        for (int p = 0; p < nf; p++) {
            const double x = node_coords[3 * (size_t)(cell_nodes[(size_t)nf * ic + p] - 1) + k];
            if (p == 0 || x < cmin) cmin = x;
            if (p == 0 || x > cmax) cmax = x;
        }
        lo[k] = rb_index(cmin, g.lo[k], g.h[k], g.n[k]);
        hi[k] = rb_index(cmax, g.lo[k], g.h[k], g.n[k]);
    }
    for (int i = lo[0]; i <= hi[0]; i++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int k = lo[2]; k <= hi[2]; k++) {
                const unsigned int b = rb_bin(g, i, j, k);
                const unsigned int e = atomicAdd(&count[b], 1u);
                if (FILL) cells[off[b] + e] = ic + 1;
            }
}

/* one thread per particle */
__global__ void __launch_bounds__(256) rb_particles_kernel(const __grid_constant__ PgDev d, const RbGrid g,
                                                           const unsigned int* __restrict__ off, const int* __restrict__ cells) {
    const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= (long long)d.status->np_cur) return;
    const PgPos P = d.p.pos[ip];
    if (P.ic == PG_TOMBSTONE) return;
    const double xp[3] = {P.x, P.y, P.z};
    bool outside = false;
    for (int k = 0; k < g.dims; k++)
        if (xp[k] < g.lo[k] || xp[k] > g.hi[k]) outside = true;
    int found = 0x7fffffff;
    if (!outside) {
        int idx[3] = {1, 1, 1};
        for (int k = 0; k < g.dims; k++) idx[k] = rb_index(xp[k], g.lo[k], g.h[k], g.n[k]);
        const unsigned int b = rb_bin(g, idx[0], idx[1], idx[2]);
        const int nf = d.nf;
        for (unsigned int e = off[b]; e < off[b + 1]; e++) {
            const int ic = cells[e];
            if (ic >= found) continue; /* the reference takes the first of the ascending list */
            const double* bc = d.basis + (size_t)nf * nf * (size_t)(ic - 1);
            bool inside = true;
            for (int vp = 0; vp < nf; vp++) {
                const double* q = bc + nf * vp; /* BASIS_COEFFS(1:nf, VP, IC) */
                double psip;
                if (nf == 4) psip = xp[0] * q[0] + xp[1] * q[1] + xp[2] * q[2] + q[3];
                else if (nf == 3) psip = xp[0] * q[0] + xp[1] * q[1] + q[2];
                else psip = xp[0] * q[0] + q[1];
                if (psip < 0) inside = false;
            }
            if (inside) found = ic;
        }
    }
    if (found != 0x7fffffff) {
        d.p.pos[ip].ic = found;
    } else {
        d.p.pos[ip].ic = PG_TOMBSTONE;
        atomicAdd(&d.status->n_removed, 1ULL);
    }
}

} /* namespace */

extern "C" int pgpu_reassign_particles_to_cells(pgpu_ctx* c, int64_t* n_removed) {
    if (!c) return PGPU_ERR_INVALID;
    if (n_removed) *n_removed = 0;
    int rc;
    c->reassigning = true;
    rc = pg_prepare(c);
    c->reassigning = false;
    if (rc) return rc;
    if (c->cfg.grid_type != PGPU_UNSTRUCTURED) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "reassign_particles_to_cells: unstructured grids only ('Not implemented!' in the reference)");
    if (!c->dev.basis) PG_FAIL(c, PGPU_ERR_STATE, "reassign_particles_to_cells: mesh without BASIS_COEFFS");
    RbGrid g;
    memset(&g, 0, sizeof(g));
    g.dims = c->mesh_dims;
    const double lo[3] = {c->cfg.xmin, c->cfg.ymin, c->cfg.zmin}, hi[3] = {c->cfg.xmax, c->cfg.ymax, c->cfg.zmax};
    size_t nbins = 1;
    for (int k = 0; k < 3; k++) {
        g.n[k] = k < g.dims ? RB_PARTS : 1;
        g.lo[k] = lo[k];
        g.hi[k] = hi[k];
        g.h[k] = (hi[k] - lo[k]) / RB_PARTS; /* DX = (XMAX-XMIN)/N_PARTITIONS_X */
        if (k < g.dims && !(g.h[k] > 0.)) PG_FAIL(c, PGPU_ERR_INVALID, "reassign_particles_to_cells: empty Domain_limits in direction %d", k + 1);
        nbins *= (size_t)g.n[k];
    }
    if (c->np == 0) {
        c->cells_unassigned = false;
        return PGPU_OK;
    }
    if ((rc = pg_resolve_vel(c))) return rc;
    if ((rc = pg_reset_status(c))) return rc;
    const int nc = c->nc, nf = c->mesh_dims + 1;
    unsigned int *d_count = nullptr, *d_off = nullptr;
    int* d_cells = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() { cudaFree(d_count); cudaFree(d_off); cudaFree(d_cells); cudaFree(d_tmp); };
#define RB_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            cleanup();                                                                                      \
            PG_FAIL(c, PGPU_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
        }                                                                                                   \
    } while (0)
    RB_CUDA(cudaMalloc((void**)&d_count, sizeof(unsigned int) * (nbins + 1)));
    RB_CUDA(cudaMalloc((void**)&d_off, sizeof(unsigned int) * (nbins + 1)));
    RB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int) * (nbins + 1), c->stream));
    const unsigned int cblocks = (unsigned int)((nc + 127) / 128);
    rb_cells_kernel<false><<<cblocks, 128, 0, c->stream>>>(g, nc, nf, c->dev.cell_nodes, c->d_node_coords, d_count, nullptr, nullptr);
    RB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count, d_off, (int)(nbins + 1), c->stream));
    RB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    RB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_count, d_off, (int)(nbins + 1), c->stream));
    unsigned int total = 0;
    RB_CUDA(cudaMemcpyAsync(&total, d_off + nbins, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(cudaStreamSynchronize(c->stream));
    RB_CUDA(cudaMalloc((void**)&d_cells, sizeof(int) * (size_t)(total ? total : 1)));
    RB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int) * (nbins + 1), c->stream));
    rb_cells_kernel<true><<<cblocks, 128, 0, c->stream>>>(g, nc, nf, c->dev.cell_nodes, c->d_node_coords, d_count, d_off, d_cells);
    rb_particles_kernel<<<(unsigned int)((c->np + 255) / 256), 256, 0, c->stream>>>(c->dev, g, d_off, d_cells);
    c->launches += 5;
    c->maybe_tombstones = true;
    c->sorted = false;
    c->hist_valid = false;
    c->emig_valid = false;
    c->owned_clean = false;
    rc = pg_finish(c, "pgpu_reassign_particles_to_cells");
    cleanup();
#undef RB_CUDA
    if (rc) return rc;
    c->cells_unassigned = false;
    if (n_removed) *n_removed = (int64_t)c->h_status->n_removed;
    return PGPU_OK;
}
