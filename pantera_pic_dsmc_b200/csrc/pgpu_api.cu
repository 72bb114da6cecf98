/*
 * pgpu_api.cu -- C ABI of libpantera_gpu.so (include/pantera_gpu.h): context life cycle, parameter
 * setters (the module globals of src/global.f90 the hot path reads), the device particle store
 * (AoS <-> SoA transposition of TYPE PARTICLE_DATA_STRUCTURE, src/particle.f90:28-37), the counting
 * sort of DSMC_COLLISIONS (src/collisions.f90:61-112) as a physical SoA reorder, and the per-step
 * driver that replays the PIC_TYPE==NONE branch of TIME_LOOP (src/timecycle.f90:271-299).
 *
 * There is no CPU fallback anywhere in this library: every entry point that computes needs a CUDA
 * device and fails with PGPU_ERR_CUDA otherwise.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <new>

#include "pgpu_internal.cuh"

/* named device-table slots */
enum {
    S_MASS, S_CHARGE, S_ROTREL, S_VIBREL, S_ROTDOF, S_VIBDOF,
    S_SIGMAS, S_OMEGAS, S_ALPHAS, S_GREFS,
    S_REAC, S_TAB_E, S_TAB_CS,
    S_WR_R, S_WR_P1, S_WR_NPROD, S_WR_PROB,
    S_MCC_SP, S_MCC_FRAC, S_MCC_CELL_NRHO,
    S_FACE_GEO, S_FACE_FILT, S_FACE_EXACT, S_FACE_CODE, S_TANG1, S_TANG2, S_EDGE_LEN, S_FACE_BIDX, S_BOUND_AREA, S_CELL_NODES, S_BASIS,
    S_CELL_VOL, S_CELL_PROCS, S_CELL_FNUM,
    S_BC_PARTICLE, S_BC_FIELD, S_BC_REACT, S_BC_TEMP, S_BC_ACCN, S_BC_ACCT,
    S_W_X1, S_W_Y1, S_W_X2, S_W_Y2, S_W_NX, S_W_NY, S_W_TEMP, S_W_TRANSM, S_W_SPEC, S_W_DIFF, S_W_POROUS, S_W_REACT,
    S_XCOORD, S_YCOORD, S_EFIELD, S_BFIELD,
    S_COUNT
};
static_assert(S_COUNT <= 64, "slot table too small");

static thread_local char g_create_err[256] = "";

/* replace the device array held in a slot by a copy of host[0..n) */
template <class T>
static int put(pgpu_ctx* c, int slot, const T* host, size_t n, const T** out) {
    if (c->slot[slot]) {
        cudaFree(c->slot[slot]);
        c->slot[slot] = nullptr;
    }
    *out = nullptr;
    if (!host || n == 0) return PGPU_OK;
    void* p = nullptr;
    PG_CUDA(c, cudaMalloc(&p, sizeof(T) * n));
    PG_CUDA(c, cudaMemcpy(p, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    c->slot[slot] = p;
    *out = (const T*)p;
    return PGPU_OK;
}

template <class T>
static int dev_alloc(pgpu_ctx* c, T** p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    PG_CUDA(c, cudaMalloc((void**)p, sizeof(T) * n));
    return PGPU_OK;
}
template <class T>
static void dev_free(T*& p) {
    if (p) cudaFree((void*)p);
    p = nullptr;
}

/* ------------------------------------------------------------------------------------------
 * life cycle
 * ---------------------------------------------------------------------------------------- */
extern "C" int pgpu_version(void) { return 1000; }

extern "C" int pgpu_create(pgpu_ctx** out, int device) {
    if (!out) return PGPU_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_create_err, sizeof(g_create_err), "no usable CUDA device (%s); this library has no CPU fallback",
                 cudaGetErrorString(e));
        return PGPU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return PGPU_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return PGPU_ERR_CUDA;
    pgpu_ctx* c = new (std::nothrow) pgpu_ctx();
    if (!c) return PGPU_ERR_INVALID;
    c->err[0] = 0;
    c->device = device;
    c->cfg_set = false;
    memset(&c->cfg, 0, sizeof(c->cfg));
    memset(&c->dev, 0, sizeof(c->dev));
    memset(c->slot, 0, sizeof(c->slot));
    memset(c->planes, 0, sizeof(c->planes));
    c->mesh_dims = 0; c->nn = 0; c->nc = 0;
    c->mesh_set = false; c->mesh_dirty = false;
    c->cur = 0; c->np = 0; c->capacity = 0;
    c->has_dtrim = false; c->sorted = false; c->hist_valid = false; c->maybe_tombstones = false;
    c->np_ub = 0; c->np_dev_valid = false; c->owned_clean = false;
    c->id_counter = 0;
    c->any_internal_dof = false; c->ie_seen = false;
    c->d_tile_first = nullptr; c->tile_alloc = 0; c->n_big = 0;
    c->d_cell_order = nullptr; c->d_cell_bucket = nullptr;
    c->d_big_list = nullptr; c->d_big_bbox = nullptr; c->big_alloc = 0;
    c->d_cell_count = c->d_cell_off = c->d_hitlist = nullptr;
    c->d_src_of = nullptr; c->vel_deferred = false;
    c->hit_capacity = 0; c->ncells_alloc = 0;
    c->d_status = nullptr; c->h_status = nullptr;
    c->d_scan_tmp = nullptr; c->scan_tmp_bytes = 0;
    c->d_cand = nullptr; c->cand_capacity = 0;
    c->d_owner = nullptr; c->owner_capacity = 0;
    c->d_cell_ncoll = nullptr; c->d_cand_off = nullptr; c->d_cell_vrmax = nullptr; c->d_cell_fcorr = nullptr;
    c->cellscratch_alloc = 0;
    c->d_stage[0] = c->d_stage[1] = nullptr; c->stage_bytes = 0;
    c->h_stage[0] = c->h_stage[1] = nullptr;
    c->d_rhs = nullptr; c->rhs_n = 0;
    c->ns = 0; c->nw = 0;
    c->d_bcount = c->d_wcount = nullptr;
    c->d_tallies = nullptr; c->tallies_n = 0;
    c->d_surface_charge = nullptr; c->nsc = 0;
    memset(c->ms, 0, sizeof(c->ms));
    c->launches = 0;
    c->nccl_lib = nullptr; c->nccl_comm = nullptr; c->comm_nranks = 1; c->comm_rank = 0;
    c->d_sendbuf = c->d_recvbuf = nullptr; c->xbuf_capacity = 0; c->d_xcount = nullptr;
    c->d_emig_list = nullptr; c->emig_cap = 0; c->emig_valid = false;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return PGPU_ERR_CUDA; }
    for (int i = 0; i < 16; i++) cudaEventCreate(&c->ev[i]);
    if (cudaMalloc((void**)&c->d_status, sizeof(PgStatus)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_status, sizeof(PgStatus)) != cudaSuccess) {
        delete c;
        return PGPU_ERR_CUDA;
    }
    cudaMemset(c->d_status, 0, sizeof(PgStatus));
    memset(c->h_status, 0, sizeof(PgStatus));
    c->cfg.nranks = 1;
    c->cfg.eps_scaling = 1.0;
    /* deferred vel reorder (pg_sort_impl defer_vel): off by default -- measured r02c at full size: the sort drops from
     * 2.82 to 1.97 ms but the lane-per-cell gather costs the collision kernel 3.5 ms (29 GB of DRAM reads for 4 GB of
     * records: scattered 32-byte reads do not merge the way the scatter's writes do) */
    if (const char* e = getenv("PANTERA_GPU_DEFER_VEL")) c->defer_vel = atoi(e) != 0;
    c->x_counts_valid = false;
    c->x_headroom = 1.25;
    if (const char* e = getenv("PANTERA_GPU_XCHG_HEADROOM")) { double v = atof(e); if (v >= 1.0 && v <= 64.0) c->x_headroom = v; }
    if (const char* e = getenv("PANTERA_GPU_ADVECT_CHUNK")) {
        int v = atoi(e);
        if (v >= 32 && v <= 65536 && (v & (v - 1)) == 0) c->advect_chunk_override = v;
    }
    *out = c;
    return PGPU_OK;
}

static void free_planes(PgPlanes& p) {
    dev_free(p.pos); dev_free(p.vel); dev_free(p.ie); dev_free(p.dtrim); dev_free(p.id);
}

extern "C" void pg_comm_destroy(pgpu_ctx* c); /* pgpu_exchange.cu */

extern "C" int pgpu_destroy(pgpu_ctx* c) {
    if (!c) return PGPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    pg_comm_destroy(c);
    pg_inject_destroy(c);
    pg_field_destroy(c);
    dev_free(c->d_node_coords); dev_free(c->d_cell_solid);
    for (int i = 0; i < S_COUNT; i++)
        if (c->slot[i]) cudaFree(c->slot[i]);
    free_planes(c->planes[0]);
    free_planes(c->planes[1]);
    dev_free(c->d_cell_count); dev_free(c->d_cell_off); dev_free(c->d_hitlist); dev_free(c->d_src_of);
    dev_free(c->d_tile_first); dev_free(c->d_big_list); dev_free(c->d_big_bbox);
    dev_free(c->d_cell_order); dev_free(c->d_cell_bucket);
    dev_free(c->d_status);
    if (c->h_status) cudaFreeHost(c->h_status);
    dev_free(c->d_scan_tmp); dev_free(c->d_cand); dev_free(c->d_owner);
    dev_free(c->d_cell_ncoll); dev_free(c->d_cand_off); dev_free(c->d_cell_vrmax); dev_free(c->d_cell_fcorr);
    for (int i = 0; i < 2; i++) {
        dev_free(c->d_stage[i]);
        if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    }
    dev_free(c->d_rhs); dev_free(c->d_bcount); dev_free(c->d_wcount); dev_free(c->d_tallies);
    dev_free(c->d_surface_charge);
    for (int k = 0; k < 4; k++) dev_free(c->d_det[k]);
    if (c->d_det_tmp) cudaFree(c->d_det_tmp);
    dev_free(c->d_gavg); dev_free(c->d_gavg_np); dev_free(c->d_gavg_sums); dev_free(c->d_gavg_cum);
    dev_free(c->d_sendbuf); dev_free(c->d_recvbuf); dev_free(c->d_xcount); dev_free(c->d_emig_list);
    for (int i = 0; i < 16; i++) cudaEventDestroy(c->ev[i]);
    if (c->x_trace == 1) {
        if (c->x_traced > 0)
            fprintf(stderr, "[pantera_gpu rank %d] EXCHANGE per step over %lld steps: pack %.3f ms, transfer %.3f ms, unpack %.3f ms\n",
                    c->cfg.rank, c->x_traced, c->x_ms[0] / c->x_traced, c->x_ms[1] / c->x_traced, c->x_ms[2] / c->x_traced);
        for (int i = 0; i < 4; i++) cudaEventDestroy(c->xev[i]);
    }
    cudaStreamDestroy(c->stream);
    delete c;
    return PGPU_OK;
}

extern "C" const char* pgpu_last_error(const pgpu_ctx* c) { return c ? c->err : g_create_err; }

/* ------------------------------------------------------------------------------------------
 * parameters
 * ---------------------------------------------------------------------------------------- */
static int alloc_counters(pgpu_ctx* c) {
    dev_free(c->d_bcount);
    dev_free(c->d_wcount);
    dev_free(c->d_tallies);
    size_t ns = c->ns > 0 ? (size_t)c->ns : 1;
    int rc;
    if ((rc = dev_alloc(c, &c->d_bcount, 4 * ns))) return rc;
    if ((rc = dev_alloc(c, &c->d_wcount, (size_t)(c->nw > 0 ? c->nw : 1) * ns))) return rc;
    PG_CUDA(c, cudaMemset(c->d_bcount, 0, sizeof(unsigned long long) * 4 * ns));
    PG_CUDA(c, cudaMemset(c->d_wcount, 0, sizeof(unsigned long long) * (size_t)(c->nw > 0 ? c->nw : 1) * ns));
    c->tallies_n = 0;
    if (c->dev.nboundcells > 0 && c->ns > 0) {
        c->tallies_n = (long long)10 * c->dev.nboundcells * c->ns;
        if ((rc = dev_alloc(c, &c->d_tallies, (size_t)c->tallies_n))) return rc;
        PG_CUDA(c, cudaMemset(c->d_tallies, 0, sizeof(double) * (size_t)c->tallies_n));
    }
    return PGPU_OK;
}

extern "C" int pgpu_set_config(pgpu_ctx* c, const pgpu_config* cfg) {
    if (!c || !cfg) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (cfg->dims < 1 || cfg->dims > 3) PG_FAIL(c, PGPU_ERR_INVALID, "set_config: dims must be 1..3");
    if (cfg->grid_type == PGPU_QUADTREE) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "set_config: quadtree grids are not supported");
    if (cfg->grid_type != PGPU_UNSTRUCTURED && (cfg->nx < 1 || cfg->ny < 1))
        PG_FAIL(c, PGPU_ERR_INVALID, "set_config: nx, ny must be >= 1 on rectilinear grids");
    if (!(cfg->dt > 0)) PG_FAIL(c, PGPU_ERR_INVALID, "set_config: dt must be > 0");
    if (cfg->reserved_ != 0) PG_FAIL(c, PGPU_ERR_INVALID, "set_config: reserved_ must be 0");
    c->cfg = *cfg;
    if (c->cfg.nranks < 1) c->cfg.nranks = 1;
    if (c->cfg.eps_scaling == 0.0) c->cfg.eps_scaling = 1.0;
    c->cfg_set = true;
    c->mesh_dirty = c->mesh_set;
    return PGPU_OK;
}

extern "C" int pgpu_set_species(pgpu_ctx* c, int ns, const double* mass, const double* charge,
                                const int32_t* rotdof, const int32_t* vibdof, const double* rotrel,
                                const double* vibrel) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (ns < 1 || !mass || !charge || !rotdof || !vibdof || !rotrel || !vibrel)
        PG_FAIL(c, PGPU_ERR_INVALID, "set_species: ns < 1 or null array");
    int rc;
    c->ns = ns;
    c->dev.ns = ns;
    c->any_internal_dof = false;
    for (int i = 0; i < ns; i++)
        if (rotdof[i] != 0 || vibdof[i] != 0) c->any_internal_dof = true;
    if ((rc = put(c, S_MASS, mass, ns, &c->dev.mass))) return rc;
    if ((rc = put(c, S_CHARGE, charge, ns, &c->dev.charge))) return rc;
    if ((rc = put(c, S_ROTDOF, (const int*)rotdof, ns, &c->dev.rotdof))) return rc;
    if ((rc = put(c, S_VIBDOF, (const int*)vibdof, ns, &c->dev.vibdof))) return rc;
    if ((rc = put(c, S_ROTREL, rotrel, ns, &c->dev.rotrel))) return rc;
    if ((rc = put(c, S_VIBREL, vibrel, ns, &c->dev.vibrel))) return rc;
    return alloc_counters(c);
}

extern "C" int pgpu_set_vss(pgpu_ctx* c, int ns, const double* sigmas, const double* omegas, const double* alphas,
                            const double* grefs, double sigmamax) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (ns != c->ns) PG_FAIL(c, PGPU_ERR_INVALID, "set_vss: ns (%d) differs from set_species (%lld)", ns, c->ns);
    int rc;
    size_t n = (size_t)ns * ns;
    if ((rc = put(c, S_SIGMAS, sigmas, n, &c->dev.sigmas))) return rc;
    if ((rc = put(c, S_OMEGAS, omegas, n, &c->dev.omegas))) return rc;
    if ((rc = put(c, S_ALPHAS, alphas, n, &c->dev.alphas))) return rc;
    if ((rc = put(c, S_GREFS, grefs, n, &c->dev.grefs))) return rc;
    c->dev.sigmamax = sigmamax;
    return PGPU_OK;
}

extern "C" int pgpu_set_reactions(pgpu_ctx* c, int nr, const int32_t* type, const int32_t* r1, const int32_t* r2,
                                  const int32_t* p1, const int32_t* p2, const int32_t* p3, const int32_t* p4,
                                  const int32_t* nprod, const int32_t* is_cex, const double* ea, const double* c1,
                                  const double* c2, const double* c3, const int32_t* tab_off, const double* tab_e,
                                  const double* tab_cs) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    c->dev.nr = 0;
    if (nr <= 0) {
        if ((rc = put<PgReaction>(c, S_REAC, nullptr, 0, &c->dev.reac))) return rc;
        return PGPU_OK;
    }
    if (!type || !r1 || !r2 || !p1 || !p2 || !nprod || !ea) PG_FAIL(c, PGPU_ERR_INVALID, "set_reactions: null array");
    std::vector<PgReaction> h((size_t)nr);
    for (int i = 0; i < nr; i++) {
        PgReaction& r = h[i];
        r.type = type[i]; r.r1 = r1[i]; r.r2 = r2[i]; r.p1 = p1[i]; r.p2 = p2[i];
        r.p3 = p3 ? p3[i] : -1; r.p4 = p4 ? p4[i] : -1;
        r.nprod = nprod[i]; r.is_cex = is_cex ? is_cex[i] : 0;
        r.ea = ea[i]; r.c1 = c1 ? c1[i] : 0; r.c2 = c2 ? c2[i] : 0; r.c3 = c3 ? c3[i] : 0;
        r.tab_off = tab_off ? tab_off[i] : 0;
        r.tab_n = tab_off ? tab_off[i + 1] - tab_off[i] : 0;
        r.pad = 0;
        if (r.nprod < 2 || r.nprod > 4) PG_FAIL(c, PGPU_ERR_INVALID, "set_reactions: reaction %d has %d products", i, r.nprod);
        if (r.type == PGPU_LXCAT && r.tab_n < 2) PG_FAIL(c, PGPU_ERR_INVALID, "set_reactions: LXCat reaction %d without table", i);
    }
    int ntab = tab_off ? tab_off[nr] : 0;
    if ((rc = put(c, S_REAC, h.data(), (size_t)nr, &c->dev.reac))) return rc;
    if ((rc = put(c, S_TAB_E, tab_e, (size_t)ntab, &c->dev.tab_e))) return rc;
    if ((rc = put(c, S_TAB_CS, tab_cs, (size_t)ntab, &c->dev.tab_cs))) return rc;
    c->dev.nr = nr;
    return PGPU_OK;
}

extern "C" int pgpu_set_wall_reactions(pgpu_ctx* c, int n, const int32_t* r_sp, const int32_t* p1_sp,
                                       const int32_t* nprod, const double* prob) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    c->dev.nwr = 0;
    if ((rc = put(c, S_WR_R, (const int*)r_sp, (size_t)std::max(n, 0), &c->dev.wr_r))) return rc;
    if ((rc = put(c, S_WR_P1, (const int*)p1_sp, (size_t)std::max(n, 0), &c->dev.wr_p1))) return rc;
    if ((rc = put(c, S_WR_NPROD, (const int*)nprod, (size_t)std::max(n, 0), &c->dev.wr_nprod))) return rc;
    if ((rc = put(c, S_WR_PROB, prob, (size_t)std::max(n, 0), &c->dev.wr_prob))) return rc;
    c->dev.nwr = std::max(n, 0);
    return PGPU_OK;
}

extern "C" int pgpu_set_mcc_background(pgpu_ctx* c, int ncomp, const int32_t* sp_id, const double* molfrac) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    c->dev.mcc_ncomp = 0;
    if ((rc = put(c, S_MCC_SP, (const int*)sp_id, (size_t)std::max(ncomp, 0), &c->dev.mcc_sp))) return rc;
    if ((rc = put(c, S_MCC_FRAC, molfrac, (size_t)std::max(ncomp, 0), &c->dev.mcc_frac))) return rc;
    c->dev.mcc_ncomp = std::max(ncomp, 0);
    return PGPU_OK;
}

/* MCC_BG_CELL_NRHO(SP_ID, IC), ref: src/initialization.f90:2752-2788 (BOOL_BG_DENSITY_FILE); nrho == NULL clears it */
extern "C" int pgpu_set_mcc_cell_density(pgpu_ctx* c, int ns, int64_t nc, const double* nrho) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    if (!nrho || ns <= 0 || nc <= 0) {
        if ((rc = put(c, S_MCC_CELL_NRHO, (const double*)nullptr, 0, &c->dev.mcc_cell_nrho))) return rc;
        c->dev.mcc_cell_nrho = nullptr;
        c->mcc_cell_nrho_cells = 0;
        return PGPU_OK;
    }
    if (c->ns && ns != c->ns) PG_FAIL(c, PGPU_ERR_INVALID, "set_mcc_cell_density: ns %d != %lld species", ns, (long long)c->ns);
    if ((rc = put(c, S_MCC_CELL_NRHO, nrho, (size_t)ns * (size_t)nc, &c->dev.mcc_cell_nrho))) return rc;
    c->mcc_cell_nrho_cells = nc;
    return PGPU_OK;
}

extern "C" int pgpu_set_mesh(pgpu_ctx* c, int dims, int nn, int nc, const double* node_coords,
                             const int32_t* cell_nodes, const int32_t* cell_neighbors, const double* face_normal,
                             const double* face_tang1, const double* face_tang2, const double* cell_edges_len,
                             const int32_t* cell_faces_pg, const int32_t* cell_pg, const double* basis_coeffs,
                             const double* cell_volumes, const int32_t* cell_procs,
                             const int32_t* face_boundary_index, int nboundcells, const double* boundary_area) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (dims < 1 || dims > 3) PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: dims must be 1..3");
    if (nn < 2 || nc < 1) PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: empty mesh");
    if (!node_coords || !cell_nodes || !cell_neighbors || !face_normal || !cell_faces_pg || !cell_volumes)
        PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: null array");
    if (dims == 3 && (!face_tang1 || !face_tang2)) PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: 3-D meshes need FACE_TANG1/2");
    int nf = dims + 1, rc;
    c->mesh_dims = dims; c->nn = nn; c->nc = nc;
    c->h_node_coords.assign(node_coords, node_coords + 3 * (size_t)nn);
    c->h_cell_nodes.assign(cell_nodes, cell_nodes + (size_t)nf * nc);
    c->h_cell_neighbors.assign(cell_neighbors, cell_neighbors + (size_t)nf * nc);
    c->h_face_normal.assign(face_normal, face_normal + 3 * (size_t)nf * nc);
    c->h_cell_faces_pg.assign(cell_faces_pg, cell_faces_pg + (size_t)nf * nc);
    if (cell_pg) c->h_cell_pg.assign(cell_pg, cell_pg + nc);
    else c->h_cell_pg.clear();
    for (size_t i = 0; i < (size_t)nf * nc; i++) {
        int v = cell_nodes[i], nb = cell_neighbors[i];
        if (v < 1 || v > nn) PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: CELL_NODES entry %zu = %d outside 1..%d", i, v, nn);
        if (nb != -1 && (nb < 1 || nb > nc)) PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: CELL_NEIGHBORS entry %zu = %d invalid", i, nb);
    }
    PgDev& d = c->dev;
    d.nn = nn; d.nc = nc; d.nf = nf;
    if ((rc = put(c, S_TANG1, face_tang1, face_tang1 ? 12 * (size_t)nc : 0, &d.face_tang1))) return rc;
    if ((rc = put(c, S_TANG2, face_tang2, face_tang2 ? 12 * (size_t)nc : 0, &d.face_tang2))) return rc;
    if ((rc = put(c, S_EDGE_LEN, cell_edges_len, cell_edges_len ? 3 * (size_t)nc : 0, &d.edge_len))) return rc;
    if ((rc = put(c, S_CELL_NODES, (const int*)cell_nodes, (size_t)nf * nc, &d.cell_nodes))) return rc;
    if ((rc = put(c, S_BASIS, basis_coeffs, basis_coeffs ? (size_t)nf * nf * nc : 0, &d.basis))) return rc;
    if ((rc = put(c, S_CELL_VOL, cell_volumes, (size_t)nc, &d.cell_vol))) return rc;
    if (cell_procs)
        for (int i = 0; i < nc; i++)
            if (cell_procs[i] < 0 || cell_procs[i] >= 64)
                PG_FAIL(c, PGPU_ERR_INVALID, "set_mesh: CELL_PROCS(%d) = %d outside 0..63", i + 1, cell_procs[i]);
    c->cell_procs_max = 0;
    if (cell_procs) {
        c->cell_procs_max = *std::max_element(cell_procs, cell_procs + nc);
        c->h_cell_procs.assign(cell_procs, cell_procs + nc);
    } else {
        c->h_cell_procs.clear();
    }
    c->owned_clean = false;
    if ((rc = put(c, S_CELL_PROCS, (const int*)cell_procs, cell_procs ? (size_t)nc : 0, &d.cell_procs))) return rc;
    d.nboundcells = face_boundary_index ? nboundcells : 0;
    if ((rc = put(c, S_FACE_BIDX, (const int*)face_boundary_index, face_boundary_index ? (size_t)nf * nc : 0, &d.face_bidx))) return rc;
    if ((rc = put(c, S_BOUND_AREA, boundary_area, (size_t)d.nboundcells, &d.bound_area))) return rc;
    dev_free(c->d_node_coords);
    if ((rc = dev_alloc(c, &c->d_node_coords, 3 * (size_t)nn))) return rc;
    PG_CUDA(c, cudaMemcpy(c->d_node_coords, node_coords, sizeof(double) * 3 * (size_t)nn, cudaMemcpyHostToDevice));
    c->volume_total = 0;
    for (int i = 0; i < nc; i++) c->volume_total += cell_volumes[i];
    dev_free(c->d_surface_charge);
    if ((rc = dev_alloc(c, &c->d_surface_charge, (size_t)nn))) return rc;
    PG_CUDA(c, cudaMemset(c->d_surface_charge, 0, sizeof(double) * (size_t)nn));
    c->nsc = nn;
    c->mesh_set = true;
    c->mesh_dirty = true;
    pg_field_invalidate(c);
    return alloc_counters(c);
}

extern "C" int pgpu_set_cell_procs(pgpu_ctx* c, int nc, const int32_t* cell_procs) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->cfg_set) PG_FAIL(c, PGPU_ERR_STATE, "set_cell_procs: set_config first");
    if (nc != pg_ncells(c)) PG_FAIL(c, PGPU_ERR_INVALID, "set_cell_procs: nc (%d) != number of cells (%d)", nc, pg_ncells(c));
    c->emig_valid = false; /* the emigrant list of the last ADVECT was made with the old ownership */
    if (cell_procs) {
        const int nr = c->cfg.nranks > 1 ? c->cfg.nranks : 1;
        for (int i = 0; i < nc; i++)
            if (cell_procs[i] < 0 || cell_procs[i] >= nr)
                PG_FAIL(c, PGPU_ERR_INVALID, "set_cell_procs: CELL_PROCS(%d) = %d outside 0..%d", i + 1, cell_procs[i], nr - 1);
        c->cell_procs_max = *std::max_element(cell_procs, cell_procs + nc);
        c->h_cell_procs.assign(cell_procs, cell_procs + nc);
    } else {
        c->h_cell_procs.clear();
    }
    c->owned_clean = false;
    c->x_counts_valid = false;
    c->mesh_dirty = c->mesh_set; /* the hop codes carry the ownership flags */
    return put(c, S_CELL_PROCS, (const int*)cell_procs, (size_t)nc, &c->dev.cell_procs);
}

/* CELL_FNUM(NX*NY) of BOOL_RADIAL_WEIGHTING (src/initialization.f90:2304-2310); NULL switches the weighting off */
extern "C" int pgpu_set_cell_fnum(pgpu_ctx* c, int nc, const double* cell_fnum) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!cell_fnum || nc <= 0) return put(c, S_CELL_FNUM, (const double*)nullptr, 0, &c->dev.cell_fnum);
    if (!c->cfg_set) PG_FAIL(c, PGPU_ERR_STATE, "set_cell_fnum: set_config first");
    if (c->cfg.grid_type == PGPU_UNSTRUCTURED) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "set_cell_fnum: radial weighting exists on rectilinear grids only (src/timecycle.f90:2105-2128)");
    if (nc != pg_ncells(c)) PG_FAIL(c, PGPU_ERR_INVALID, "set_cell_fnum: nc (%d) != number of cells (%d)", nc, pg_ncells(c));
    return put(c, S_CELL_FNUM, cell_fnum, (size_t)nc, &c->dev.cell_fnum);
}

extern "C" int pgpu_set_bcs(pgpu_ctx* c, int nbc, const int32_t* particle_bc, const int32_t* field_bc,
                            const int32_t* volume_bc, const double* wall_temp, const double* acc_n,
                            const double* acc_t, const int32_t* react, const int32_t* dump_fluxes) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    (void)dump_fluxes;
    if (nbc < 0 || (nbc > 0 && (!particle_bc || !field_bc || !volume_bc || !wall_temp || !acc_n || !acc_t || !react)))
        PG_FAIL(c, PGPU_ERR_INVALID, "set_bcs: null array");
    for (int i = 0; i < nbc; i++)
        if (particle_bc[i] == PGPU_WB_BC)
            PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "set_bcs: the washboard wall model needs tables the reference does not ship");
    int rc;
    PgDev& d = c->dev;
    d.nbc = nbc;
    if ((rc = put(c, S_BC_PARTICLE, (const int*)particle_bc, (size_t)nbc, &d.bc_particle))) return rc;
    if ((rc = put(c, S_BC_FIELD, (const int*)field_bc, (size_t)nbc, &d.bc_field))) return rc;
    if ((rc = put(c, S_BC_REACT, (const int*)react, (size_t)nbc, &d.bc_react))) return rc;
    if ((rc = put(c, S_BC_TEMP, wall_temp, (size_t)nbc, &d.bc_temp))) return rc;
    if ((rc = put(c, S_BC_ACCN, acc_n, (size_t)nbc, &d.bc_accn))) return rc;
    if ((rc = put(c, S_BC_ACCT, acc_t, (size_t)nbc, &d.bc_acct))) return rc;
    c->h_bc_volume.assign(volume_bc, volume_bc + nbc);
    c->h_bc_particle.assign(particle_bc, particle_bc + nbc);
    c->h_bc_field.assign(field_bc, field_bc + nbc);
    c->h_bc_react.assign(react, react + nbc);
    c->mesh_dirty = c->mesh_set;
    pg_field_invalidate(c);
    return PGPU_OK;
}

extern "C" int pgpu_set_walls(pgpu_ctx* c, int nw, const double* x1, const double* y1, const double* x2,
                              const double* y2, const double* normx, const double* normy, const double* temp,
                              const int32_t* specular, const int32_t* diffuse, const int32_t* porous,
                              const double* transmissivity, const int32_t* react) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (nw < 0) PG_FAIL(c, PGPU_ERR_INVALID, "set_walls: nw < 0");
    int rc;
    PgDev& d = c->dev;
    size_t n = (size_t)nw;
    if ((rc = put(c, S_W_X1, x1, n, &d.w_x1))) return rc;
    if ((rc = put(c, S_W_Y1, y1, n, &d.w_y1))) return rc;
    if ((rc = put(c, S_W_X2, x2, n, &d.w_x2))) return rc;
    if ((rc = put(c, S_W_Y2, y2, n, &d.w_y2))) return rc;
    if ((rc = put(c, S_W_NX, normx, n, &d.w_nx))) return rc;
    if ((rc = put(c, S_W_NY, normy, n, &d.w_ny))) return rc;
    if ((rc = put(c, S_W_TEMP, temp, n, &d.w_temp))) return rc;
    if ((rc = put(c, S_W_TRANSM, transmissivity, n, &d.w_transm))) return rc;
    if ((rc = put(c, S_W_SPEC, (const int*)specular, n, &d.w_spec))) return rc;
    if ((rc = put(c, S_W_DIFF, (const int*)diffuse, n, &d.w_diff))) return rc;
    if ((rc = put(c, S_W_POROUS, (const int*)porous, n, &d.w_porous))) return rc;
    if ((rc = put(c, S_W_REACT, (const int*)react, n, &d.w_react))) return rc;
    d.nw = nw;
    c->nw = nw;
    return alloc_counters(c);
}

extern "C" int pgpu_set_grid_coords(pgpu_ctx* c, int nxp, const double* xcoord, int nyp, const double* ycoord) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    if ((rc = put(c, S_XCOORD, xcoord, (size_t)std::max(nxp, 0), &c->dev.xcoord))) return rc;
    if ((rc = put(c, S_YCOORD, ycoord, (size_t)std::max(nyp, 0), &c->dev.ycoord))) return rc;
    c->dev.nxp = nxp;
    c->dev.nyp = nyp;
    return PGPU_OK;
}

extern "C" int pgpu_set_efield(pgpu_ctx* c, const double* e, int64_t n) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    return put(c, S_EFIELD, e, (size_t)std::max<int64_t>(n, 0), &c->dev.e_field);
}
int pg_adopt_efield(pgpu_ctx* c, double* d_e) {
    if (c->slot[S_EFIELD]) cudaFree(c->slot[S_EFIELD]);
    c->slot[S_EFIELD] = d_e;
    c->dev.e_field = d_e;
    return PGPU_OK;
}
extern "C" int pgpu_set_bfield(pgpu_ctx* c, const double* b, int64_t n) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    return put(c, S_BFIELD, b, (size_t)std::max<int64_t>(n, 0), &c->dev.b_field);
}
extern "C" int pgpu_get_surface_charge(pgpu_ctx* c, double* sc, int64_t n) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->d_surface_charge) PG_FAIL(c, PGPU_ERR_STATE, "get_surface_charge: no mesh set");
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    PG_CUDA(c, cudaMemcpy(sc, c->d_surface_charge, sizeof(double) * (size_t)std::min<int64_t>(n, c->nsc), cudaMemcpyDeviceToHost));
    return PGPU_OK;
}
extern "C" int pgpu_set_surface_charge(pgpu_ctx* c, const double* sc, int64_t n) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    dev_free(c->d_surface_charge);
    int rc;
    if ((rc = dev_alloc(c, &c->d_surface_charge, (size_t)n))) return rc;
    PG_CUDA(c, cudaMemcpy(c->d_surface_charge, sc, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    c->nsc = n;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * packed device mesh:  one record per cell = per face { outward normal, a point on the face } + the hop codes.
 * The reference reads FACE_NORMAL/EDGE_NORMAL, CELL_NODES -> NODE_COORDS, CELL_NEIGHBORS,
 * CELL_FACES_PG and the neighbour's CELL_PG/VOLUME_BC separately for every face test
 * (src/timecycle.f90:1257-1360); here they are resolved once so that the face walk is a single
 * 208-byte record per tetrahedron (160 B per triangle) with no dependent index load.  The point of face j is local node j
 * (src/timecycle.f90:1293-1295), the same operand the reference uses, so the arithmetic is unchanged.
 * ---------------------------------------------------------------------------------------- */
static int build_packed_mesh(pgpu_ctx* c) {
    int nf = c->mesh_dims + 1, nc = c->nc, rc;
    /* hop codes: neighbour cell, or boundary physical group, resolved once from CELL_NEIGHBORS, CELL_FACES_PG and the
     * neighbour's CELL_PG / VOLUME_BC (ref: src/timecycle.f90:1307-1360) */
    std::vector<int> code((size_t)nc * nf);
    const int nbc = (int)c->h_bc_volume.size();
    for (int ic = 0; ic < nc; ic++) {
        for (int j = 0; j < nf; j++) {
            size_t f = (size_t)ic * nf + j;
            int nb = c->h_cell_neighbors[f];
            bool fluidboundary = false;
            if (nb == -1) {
                fluidboundary = true;
            } else if (!c->h_cell_pg.empty()) {
                int npg = c->h_cell_pg[nb - 1];
                if (npg != -1) {
                    if (npg < 1 || npg > nbc) PG_FAIL(c, PGPU_ERR_INVALID, "mesh: CELL_PG %d has no boundary condition", npg);
                    if (c->h_bc_volume[npg - 1] == PGPU_SOLID) fluidboundary = true;
                }
            }
            if (!fluidboundary) {
                if (nb > PG_CODE_CELL_MASK) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "mesh: more than %d cells", PG_CODE_CELL_MASK);
                code[f] = nb;
            } else {
                int pg = c->h_cell_faces_pg[f];
                if (pg == -1) code[f] = 0;
                else {
                    if (pg < 1 || pg > nbc) PG_FAIL(c, PGPU_ERR_INVALID, "mesh: face physical group %d has no boundary condition (set_bcs)", pg);
                    code[f] = -pg;
                }
            }
        }
    }
    /* the codes ride in the cell records (no dependent load in the face walk); plain specular walls are marked so
     * that the LEAN pass can reflect in line instead of deferring the particle (PG_CODE_SPEC) */
    const bool owner_flags = c->cfg.nranks > 1 && !c->h_cell_procs.empty();
    /* ranks whose cells share a face with a cell of this rank: the peers of the per-step EXCHANGE (pgpu_exchange.cu).
     * Face adjacency is symmetric, so both ends of a pair come to the same answer. */
    c->x_adjacent.assign((size_t)std::max(c->cfg.nranks, 1), 0);
    c->x_adjacent_known = owner_flags;
    c->own_lo = c->own_hi = 0; /* the span of this rank's own cells (pg_cell_range) */
    if (owner_flags) {
        long long first = -1, last = -1;
        for (long long i = 0; i < (long long)c->h_cell_procs.size(); i++)
            if (c->h_cell_procs[(size_t)i] == c->cfg.rank) {
                if (first < 0) first = i;
                last = i;
            }
        if (first >= 0) { c->own_lo = first; c->own_hi = last + 1; }
    }
    auto packed_codes = [&](int ic, int* rec) {
        rec[0] = rec[1] = rec[2] = rec[3] = 0;
        for (int j = 0; j < nf; j++) {
            int cd = code[(size_t)ic * nf + j];
            if (cd > 0 && owner_flags && c->h_cell_procs[cd - 1] != c->cfg.rank) {
                const int owner = c->h_cell_procs[cd - 1];
                if (c->h_cell_procs[ic] == c->cfg.rank && owner >= 0 && owner < c->cfg.nranks) c->x_adjacent[owner] = 1;
                cd |= PG_CODE_FOREIGN;
            }
            if (cd < 0) {
                int gi = -cd - 1;
                if (c->h_bc_particle[gi] == PGPU_SPECULAR && !c->h_bc_react[gi] && c->h_bc_field[gi] != PGPU_DIELECTRIC_BC)
                    cd -= PG_CODE_SPEC;
            }
            rec[j] = cd;
        }
    };
    /* Filter + exact records of the non-axisymmetric face walk.  The filter holds float(n) and float(P.n) (P.n
     * evaluated in double); pgpu_advect.cu bounds their rounding, so nothing here has to be exact except face_exact,
     * which carries the very operands the reference reads (FACE_NORMAL / EDGE_NORMAL and NODE_COORDS(:,CELL_NODES),
     * the point of face j being local node j, src/timecycle.f90:1293-1295). */
    {
        const int nchf = nf + 1;
        std::vector<float> filt((size_t)((nc + PG_GEO_B - 1) / PG_GEO_B) * (size_t)(nchf * PG_GEO_B * 4), 0.0f);
        std::vector<double> exact((size_t)nc * nf * 8, 0.0);
        for (int ic = 0; ic < nc; ic++) {
            float* blk = &filt[(size_t)(ic / PG_GEO_B) * (size_t)(nchf * PG_GEO_B * 4)];
            for (int j = 0; j < nf; j++) {
                size_t f = (size_t)ic * nf + j;
                int v = c->h_cell_nodes[f];
                const double* n = &c->h_face_normal[3 * f];
                const double* P = &c->h_node_coords[3 * (size_t)(v - 1)];
                double* e = &exact[8 * f]; /* {nx, ny, nz, Px | Py, Pz, 0, 0}: a 256-bit and a 128-bit load */
                for (int k = 0; k < 3; k++) { e[k] = n[k]; e[3 + k] = P[k]; }
                float* q = blk + (size_t)j * (PG_GEO_B * 4) + (size_t)(ic % PG_GEO_B) * 4;
                double pn = 0.0;
                for (int k = 0; k < c->mesh_dims; k++) pn += P[k] * n[k];
                q[0] = (float)n[0]; q[1] = (float)n[1]; q[2] = (float)n[2]; q[3] = (float)pn;
            }
            int rec[4];
            packed_codes(ic, rec);
            memcpy(blk + (size_t)nf * (PG_GEO_B * 4) + (size_t)(ic % PG_GEO_B) * 4, rec, 16);
        }
        if ((rc = put(c, S_FACE_FILT, filt.data(), filt.size(), &c->dev.face_filt))) return rc;
        if ((rc = put(c, S_FACE_EXACT, exact.data(), exact.size(), &c->dev.face_exact))) return rc;
    }
    /* the double-precision chunked record (blocks of PG_GEO_B cells, chunk-major, pgpu_internal.cuh PG_GEO_*) is read
     * by the axisymmetric walk only, which exists for 2-D meshes */
    if (c->mesh_dims == 2) {
        const int bd = PG_GEO_BLOCK_DOUBLES(nf);
        std::vector<double> geo((size_t)((nc + PG_GEO_B - 1) / PG_GEO_B) * bd, 0.0);
        auto slot = [&](int ic, int e) -> double* { /* double e of cell ic's record */
            return &geo[(size_t)(ic / PG_GEO_B) * bd + (size_t)(e >> 1) * (2 * PG_GEO_B) + (size_t)(ic % PG_GEO_B) * 2 + (e & 1)];
        };
        for (int ic = 0; ic < nc; ic++) {
            for (int j = 0; j < nf; j++) {
                size_t f = (size_t)ic * nf + j;
                int v = c->h_cell_nodes[f];
                for (int k = 0; k < 3; k++) {
                    *slot(ic, 6 * j + k) = c->h_face_normal[3 * f + k];
                    *slot(ic, 6 * j + 3 + k) = c->h_node_coords[3 * (size_t)(v - 1) + k];
                }
            }
            int rec[4];
            packed_codes(ic, rec);
            memcpy(slot(ic, 6 * nf), rec, sizeof(int) * (size_t)nf);
        }
        if ((rc = put(c, S_FACE_GEO, geo.data(), geo.size(), &c->dev.face_geo))) return rc;
    } else {
        if ((rc = put(c, S_FACE_GEO, (const double*)nullptr, 0, &c->dev.face_geo))) return rc;
    }
    if ((rc = put(c, S_FACE_CODE, code.data(), code.size(), &c->dev.face_code))) return rc;
    /* cells of a SOLID volume group: INITIAL_SEED skips them (src/initialization.f90:2072-2074) */
    dev_free(c->d_cell_solid);
    if (!c->h_cell_pg.empty()) {
        std::vector<unsigned char> solid((size_t)nc, 0);
        bool any = false;
        for (int ic = 0; ic < nc; ic++) {
            int pg = c->h_cell_pg[ic];
            if (pg != -1 && pg >= 1 && pg <= nbc && c->h_bc_volume[pg - 1] == PGPU_SOLID) { solid[ic] = 1; any = true; }
        }
        if (any) {
            if ((rc = dev_alloc(c, &c->d_cell_solid, (size_t)nc))) return rc;
            PG_CUDA(c, cudaMemcpy(c->d_cell_solid, solid.data(), (size_t)nc, cudaMemcpyHostToDevice));
        }
    }
    c->mesh_dirty = false;
    return PGPU_OK;
}

int pg_ensure_cell_arrays(pgpu_ctx* c) {
    int ncells = pg_ncells(c), rc;
    if (ncells <= c->ncells_alloc) return PGPU_OK;
    dev_free(c->d_cell_count);
    dev_free(c->d_cell_off);
    if ((rc = dev_alloc(c, &c->d_cell_count, (size_t)ncells + 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_cell_off, (size_t)ncells + 1))) return rc;
    dev_free(c->d_cell_order);
    dev_free(c->d_cell_bucket);
    if ((rc = dev_alloc(c, &c->d_cell_order, (size_t)ncells))) return rc;
    if ((rc = dev_alloc(c, &c->d_cell_bucket, (size_t)2 * PG_NBUCKET))) return rc;
    PG_CUDA(c, cudaMemset(c->d_cell_count, 0, sizeof(unsigned int) * ((size_t)ncells + 1)));
    PG_CUDA(c, cudaMemset(c->d_cell_off, 0, sizeof(unsigned int) * ((size_t)ncells + 1)));
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->d_cell_count, c->d_cell_off, ncells + 1, c->stream);
    dev_free(c->d_scan_tmp);
    PG_CUDA(c, cudaMalloc(&c->d_scan_tmp, tmp + 16));
    c->scan_tmp_bytes = tmp + 16;
    c->ncells_alloc = ncells;
    c->hist_valid = false;
    c->sorted = false;
    return PGPU_OK;
}

void pg_refresh_dev(pgpu_ctx* c) {
    PgDev& d = c->dev;
    const pgpu_config& g = c->cfg;
    d.dims = g.dims; d.axi = g.axi; d.grid_type = g.grid_type; d.nx = g.nx; d.ny = g.ny;
    d.xmin = g.xmin; d.xmax = g.xmax; d.ymin = g.ymin; d.ymax = g.ymax; d.zmin = g.zmin; d.zmax = g.zmax;
    d.dt = g.dt; d.fnum = g.fnum; d.pic = g.pic_type; d.coll_type = g.collision_type;
    for (int i = 0; i < 4; i++) {
        d.bper[i] = g.bool_periodic[i]; d.bspec[i] = g.bool_specular[i];
        d.bdiff[i] = g.bool_diffuse[i]; d.breact[i] = g.bool_react[i];
    }
    d.zper = g.bool_z_periodic; d.yper = g.bool_y_periodic;
    d.boundtemp = g.boundtemp;
    for (int i = 0; i < 3; i++) d.extb[i] = g.external_b[i];
    d.eps_scaling = g.eps_scaling; d.mcc_dens = g.mcc_bg_dens; d.mcc_ttra = g.mcc_bg_ttra;
    d.seed_local = (unsigned long long)g.rng_seed_global * (unsigned long long)(g.rank + 1);
    d.rank = g.rank; d.nranks = g.nranks;
    d.surface_charge = c->d_surface_charge;
    d.bcount = c->d_bcount; d.wcount = c->d_wcount; d.tallies = c->d_tallies;
    d.status = c->d_status;
    d.p = c->planes[c->cur];
    if (!c->has_dtrim) d.p.dtrim = nullptr;
    d.np = c->np; d.capacity = c->capacity;
    d.ncells = pg_ncells(c);
    d.cell_count = c->d_cell_count; d.cell_off = c->d_cell_off; d.hitlist = c->d_hitlist;
    d.emig_list = nullptr; d.emig_cap = c->emig_cap; d.xrow = nullptr; /* switched on by ADVECT when it tracks emigrants */
    d.id_base = c->id_counter;
    d.vel_src = c->vel_deferred ? c->planes[c->cur ^ 1].vel : nullptr;
    d.src_of = c->vel_deferred ? c->d_src_of : nullptr;
    d.use_ie = ((c->any_internal_dof || c->dev.nr > 0 || c->ie_seen) && c->planes[c->cur].ie) ? 1 : 0;
    d.tile_first = c->d_tile_first;
    d.cell_order = c->sorted ? c->d_cell_order : nullptr;
    d.cells_lo = (int)c->cells_lo;
    d.order_n = (int)(c->cells_hi - c->cells_lo);
    d.ntiles = c->sorted ? (int)((c->np + PG_TILE_S - 1) / PG_TILE_S) : 0;
}

static bool wants_ie(const pgpu_ctx* c) { return c->any_internal_dof || c->dev.nr > 0 || c->ie_seen; }

int pg_prepare(pgpu_ctx* c) {
    cudaSetDevice(c->device);
    if (!c->cfg_set) PG_FAIL(c, PGPU_ERR_STATE, "pgpu_set_config has not been called");
    if (c->ns < 1) PG_FAIL(c, PGPU_ERR_STATE, "pgpu_set_species has not been called");
    if (c->cells_unassigned && !c->reassigning)
        PG_FAIL(c, PGPU_ERR_STATE, "the restart file was read without trusting its cell indices: call pgpu_reassign_particles_to_cells first");
    int rc;
    if (c->cfg.grid_type == PGPU_UNSTRUCTURED) {
        if (!c->mesh_set) PG_FAIL(c, PGPU_ERR_STATE, "unstructured grid without pgpu_set_mesh");
        if (c->mesh_dims != c->cfg.dims) PG_FAIL(c, PGPU_ERR_INVALID, "mesh dims (%d) != config dims (%d)", c->mesh_dims, c->cfg.dims);
        if (c->cfg.axi && c->cfg.dims != 2) PG_FAIL(c, PGPU_ERR_INVALID, "axisymmetric runs need dims == 2");
        if (c->mesh_dirty && (rc = build_packed_mesh(c))) return rc;
    } else {
        if (c->cfg.dims == 3) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "rectilinear grids are 1-D/2-D in the reference (CELL_FROM_POSITION)");
        if (c->cfg.grid_type == PGPU_RECTILINEAR_NONUNIFORM &&
            (c->dev.nxp != c->cfg.nx + 1 || c->dev.nyp != c->cfg.ny + 1 || !c->dev.xcoord || !c->dev.ycoord))
            PG_FAIL(c, PGPU_ERR_STATE, "non-uniform grid without matching pgpu_set_grid_coords");
    }
    if ((rc = pg_ensure_cell_arrays(c))) return rc;
    if (wants_ie(c) && (rc = pg_need_ie(c))) return rc; /* (species / reactions may have been set after the upload) */
    pg_refresh_dev(c);
    return PGPU_OK;
}

__global__ void status_init_kernel(PgStatus* st, unsigned long long np) {
    if (threadIdx.x == 0) st->np_cur = np;
}

int pg_set_np_cur(pgpu_ctx* c) {
    status_init_kernel<<<1, 32, 0, c->stream>>>(c->d_status, (unsigned long long)c->np);
    return PGPU_OK;
}

int pg_reset_status(pgpu_ctx* c) {
    PG_CUDA(c, cudaMemsetAsync(c->d_status, 0, sizeof(PgStatus), c->stream));
    status_init_kernel<<<1, 32, 0, c->stream>>>(c->d_status, (unsigned long long)c->np);
    c->np_ub = c->np;
    c->np_dev_valid = false;
    return PGPU_OK;
}

static const char* deverr_text(int e) {
    switch (e) {
        case PG_DEVERR_HOPS: return "a particle exceeded the hop limit (malformed mesh or NaN state)";
        case PG_DEVERR_CELL_RANGE: return "a particle ended outside the grid (cell index out of range)";
        case PG_DEVERR_CAPACITY: return "particle capacity exhausted while creating particles (raise pgpu_config.capacity)";
        case PG_DEVERR_XCAP: return "EXCHANGE: more emigrants than the message sized from the previous step holds (PANTERA_GPU_XCHG_HEADROOM), or an emigrant for a rank that shares no face with this one (PANTERA_GPU_XCHG_ALLPAIRS=1)";
        case PG_DEVERR_XWAIT: return "EXCHANGE: a peer's message did not arrive in its window within 30 s (peer stopped?)";
        default: return "unknown device error";
    }
}

int pg_finish(pgpu_ctx* c, const char* what) {
    PG_CUDA(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(PgStatus), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    PG_CUDA(c, cudaGetLastError());
    if (c->np_dev_valid) { /* the count moved on the device (EXCHANGE immigrants, compaction by the sort) */
        c->np = (long long)c->h_status->np_cur;
        c->n_big = c->h_status->n_big;
        c->np_dev_valid = false;
        if (c->h_status->flag_ie) c->ie_seen = true;
    }
    if (c->x_plan_pending) { /* EXCHANGE fast path: this step's counts size the next step's messages */
        const int nr = c->cfg.nranks;
        for (int r = 0; r < nr && r < 64; r++) {
            c->x_prev_send[r] = (long long)c->h_status->x_sent[r];
            c->x_prev_recv[r] = (long long)c->h_status->x_recv[r];
        }
        c->x_plan_pending = false;
        if (c->h_status->err == PG_DEVERR_XCAP) c->x_counts_valid = false; /* (the error is reported below) */
    }
    if (c->h_status->n_appended) {
        c->np += (long long)c->h_status->n_appended;
        c->id_counter += (long long)c->h_status->n_appended;
        if (c->np > c->capacity) c->np = c->capacity;
        c->sorted = false;
        c->hist_valid = false;
    }
    if (c->h_status->err) PG_FAIL(c, PGPU_ERR_STATE, "%s: %s", what, deverr_text(c->h_status->err));
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * particle store
 * ---------------------------------------------------------------------------------------- */
static int alloc_planes(pgpu_ctx* c, PgPlanes& p, long long cap, bool with_ie, bool with_dtrim, bool with_id) {
    int rc;
    memset(&p, 0, sizeof(p));
    size_t n = (size_t)cap;
    if ((rc = dev_alloc(c, &p.pos, n)) || (rc = dev_alloc(c, &p.vel, n))) return rc;
    if (with_ie) {
        if ((rc = dev_alloc(c, &p.ie, n))) return rc;
        PG_CUDA(c, cudaMemsetAsync(p.ie, 0, sizeof(PgIe) * n, c->stream));
    }
    if (with_dtrim && (rc = dev_alloc(c, &p.dtrim, n))) return rc;
    if (with_id && (rc = dev_alloc(c, &p.id, n))) return rc;
    return PGPU_OK;
}

template <class T>
static int copy_plane(pgpu_ctx* c, T* dst, const T* src, long long n) {
    if (dst && src && n > 0) PG_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    return PGPU_OK;
}

int pg_ensure_capacity(pgpu_ctx* c, long long n) {
    if (n <= c->capacity) return PGPU_OK;
    { int rc0; if ((rc0 = pg_resolve_vel(c))) return rc0; }
    long long cap = std::max<long long>(n, (long long)(1.5 * (double)c->capacity));
    cap = std::max<long long>(cap, 1024);
    int rc;
    PgPlanes np0, np1;
    bool with_id = c->cfg.track_id != 0;
    const bool with_ie = wants_ie(c) || c->planes[0].ie != nullptr, with_dtrim = c->has_dtrim || c->planes[0].dtrim != nullptr;
    if ((rc = alloc_planes(c, np0, cap, with_ie, with_dtrim, with_id))) return rc;
    if ((rc = alloc_planes(c, np1, cap, with_ie, with_dtrim, with_id))) return rc;
    PgPlanes& o = c->planes[c->cur];
    long long m = c->np;
    if (m > 0) {
        copy_plane(c, np0.pos, o.pos, m); copy_plane(c, np0.vel, o.vel, m); copy_plane(c, np0.ie, o.ie, m);
        copy_plane(c, np0.dtrim, o.dtrim, m); copy_plane(c, np0.id, o.id, m);
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    free_planes(c->planes[0]);
    free_planes(c->planes[1]);
    c->planes[0] = np0;
    c->planes[1] = np1;
    c->cur = 0;
    dev_free(c->d_hitlist);
    dev_free(c->d_src_of);
    if ((rc = dev_alloc(c, &c->d_hitlist, (size_t)cap))) return rc;
    if (c->defer_vel && (rc = dev_alloc(c, &c->d_src_of, (size_t)cap))) return rc; /* (only the deferred reorder uses it) */
    dev_free(c->d_emig_list);
    c->emig_cap = cap / 8 + 1024;
    if (c->cfg.nranks > 1 && (rc = dev_alloc(c, &c->d_emig_list, (size_t)c->emig_cap))) return rc;
    c->emig_valid = false;
    dev_free(c->d_tile_first);
    dev_free(c->d_big_list);
    dev_free(c->d_big_bbox);
    c->tile_alloc = cap / PG_TILE_S + 4;
    c->big_alloc = cap / PG_TILE_BIG + 4;
    if ((rc = dev_alloc(c, &c->d_tile_first, (size_t)c->tile_alloc))) return rc;
    if ((rc = dev_alloc(c, &c->d_big_list, (size_t)c->big_alloc))) return rc;
    if ((rc = dev_alloc(c, &c->d_big_bbox, (size_t)c->big_alloc * 6))) return rc;
    c->capacity = cap;
    c->hist_valid = false;
    return PGPU_OK;
}

/* host record offsets, ref: src/particle.f90:28-37 as laid out by gfortran (see pantera_gpu.h) */
#define OFF_X 0
#define OFF_V 24
#define OFF_EROT 48
#define OFF_EVIB 56
#define OFF_DTRIM 64
#define OFF_IC 72
#define OFF_SID 76
#define OFF_ID 80
#define OFF_TRAJ 88

__global__ void __launch_bounds__(256) aos_to_soa_kernel(const unsigned char* __restrict__ aos, int stride, long long n,
                                                         PgPlanes p, long long base, double dt, int ncells,
                                                         PgStatus* st) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* r = aos + (size_t)i * stride;
    const double* f = reinterpret_cast<const double*>(r);
    long long o = base + i;
    int ic = *reinterpret_cast<const int*>(r + OFF_IC);
    int sid = *reinterpret_cast<const int*>(r + OFF_SID);
    PgPos P;
    P.x = f[0]; P.y = f[1]; P.z = f[2]; P.ic = ic; P.sid = sid;
    PgVel V;
    V.vx = f[3]; V.vy = f[4]; V.vz = f[5]; V.sid = sid; V.pad = 0;
    PgIe E;
    E.erot = f[6]; E.evib = f[7];
    p.pos[o] = P;
    p.vel[o] = V;
    pg_ie_store(p.ie, o, E.erot, E.evib, st);
    double dtrim = f[8];
    if (p.dtrim) p.dtrim[o] = dtrim;
    else if (dtrim != dt) st->need_dtrim = 1;
    if (p.id) p.id[o] = *reinterpret_cast<const long long*>(r + OFF_ID);
    if (dtrim != dt) st->flag_dtrim = 1;
    if (E.erot != 0. || E.evib != 0.) st->flag_ie = 1;
    if (ncells > 0 && (ic < 1 || ic > ncells)) atomicCAS(&st->err, 0, PG_DEVERR_CELL_RANGE);
}

__global__ void __launch_bounds__(256) soa_to_aos_kernel(unsigned char* __restrict__ aos, int stride, long long n,
                                                         PgPlanes p, long long base, double dt, int rank,
                                                         int has_dtrim) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char* r = aos + (size_t)i * stride;
    double* f = reinterpret_cast<double*>(r);
    long long o = base + i;
    const PgPos P = p.pos[o];
    const PgVel V = p.vel[o];
    const PgIe E = pg_ie_load(p.ie, o);
    f[0] = P.x; f[1] = P.y; f[2] = P.z;
    f[3] = V.vx; f[4] = V.vy; f[5] = V.vz;
    f[6] = E.erot; f[7] = E.evib;
    f[8] = (has_dtrim && p.dtrim) ? p.dtrim[o] : dt;
    *reinterpret_cast<int*>(r + OFF_IC) = P.ic;
    *reinterpret_cast<int*>(r + OFF_SID) = P.sid;
    /* ID = PROC_ID + ISHFT(counter, 8), ref: src/particle.f90:75 (regenerated unless cfg.track_id) */
    *reinterpret_cast<long long*>(r + OFF_ID) = p.id ? p.id[o] : ((long long)rank + (o << 8));
    *reinterpret_cast<int*>(r + OFF_TRAJ) = 0;
    if (stride >= 96) *reinterpret_cast<int*>(r + 92) = 0;
}

__global__ void fill_f64_kernel(double* p, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static const long long STAGE_PARTICLES = 1 << 20;

static int ensure_stage(pgpu_ctx* c, int stride) {
    size_t need = (size_t)STAGE_PARTICLES * (size_t)stride;
    if (need <= c->stage_bytes) return PGPU_OK;
    for (int i = 0; i < 2; i++) {
        dev_free(c->d_stage[i]);
        PG_CUDA(c, cudaMalloc(&c->d_stage[i], need));
    }
    c->stage_bytes = need;
    return PGPU_OK;
}

static int append_impl(pgpu_ctx* c, const void* aos, int64_t n, int32_t stride, const char* what) {
    if (n < 0) PG_FAIL(c, PGPU_ERR_INVALID, "%s: n < 0", what);
    if (stride < 92 || (stride % 8) != 0) PG_FAIL(c, PGPU_ERR_INVALID, "%s: stride %d must be a multiple of 8 and >= 92", what, stride);
    if (n > 0 && !aos) PG_FAIL(c, PGPU_ERR_INVALID, "%s: null buffer", what);
    if (!c->cfg_set) PG_FAIL(c, PGPU_ERR_STATE, "%s: pgpu_set_config first (DT is needed)", what);
    int rc;
    cudaSetDevice(c->device);
    if (c->capacity == 0) {
        long long cap = c->cfg.capacity > 0 ? (long long)c->cfg.capacity : (long long)(1.5 * (double)n);
        if ((rc = pg_ensure_capacity(c, std::max<long long>(cap, c->np + n)))) return rc;
    } else if ((rc = pg_ensure_capacity(c, c->np + n))) {
        return rc;
    }
    if (n == 0) return PGPU_OK;
    if ((rc = ensure_stage(c, stride))) return rc;
    if (wants_ie(c) && (rc = pg_need_ie(c))) return rc;
    if (c->has_dtrim && (rc = pg_need_dtrim(c))) return rc;
    int ncells = (c->cfg.grid_type == PGPU_UNSTRUCTURED) ? (c->mesh_set ? c->nc : 0) : c->cfg.nx * c->cfg.ny;
    cudaEvent_t evs[2] = {c->ev[14], c->ev[15]};
  for (int pass = 0; pass < 2; pass++) { /* a second pass only when the records carry internal energy / their own DTRIM
                                            and the plane for it did not exist yet (it does now) */
    if ((rc = pg_reset_status(c))) return rc;
    PgPlanes p = c->planes[c->cur];
    bool used[2] = {false, false};
    int k = 0;
    for (long long done = 0; done < n; done += STAGE_PARTICLES, k ^= 1) {
        long long m = std::min<long long>(STAGE_PARTICLES, n - done);
        if (used[k]) PG_CUDA(c, cudaEventSynchronize(evs[k])); /* staging buffer k is free again */
        PG_CUDA(c, cudaMemcpyAsync(c->d_stage[k], (const unsigned char*)aos + (size_t)done * stride, (size_t)m * stride,
                                   cudaMemcpyHostToDevice, c->stream));
        aos_to_soa_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, c->stream>>>(
            (const unsigned char*)c->d_stage[k], stride, m, p, c->np + done, c->cfg.dt, ncells, c->d_status);
        PG_CUDA(c, cudaEventRecord(evs[k], c->stream));
        used[k] = true;
    }
    PG_CUDA(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(PgStatus), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!c->h_status->need_ie && !c->h_status->need_dtrim) break;
    if (pass == 1) PG_FAIL(c, PGPU_ERR_STATE, "%s: optional planes could not be provided", what);
    if (c->h_status->need_ie && (rc = pg_need_ie(c))) return rc;
    if (c->h_status->need_dtrim) { /* particles already in the store have DTRIM = DT */
        if ((rc = pg_fill_dtrim(c))) return rc;
    }
  }
    long long np_old = c->np;
    c->np += n;
    c->sorted = false;
    c->hist_valid = false;
    c->owned_clean = false; /* host-supplied particles may sit in cells of other ranks until the next EXCHANGE */
    if ((rc = pg_finish(c, what))) {
        c->np = np_old;
        return rc;
    }
    if (c->h_status->flag_dtrim) c->has_dtrim = true;
    if (c->h_status->flag_ie) c->ie_seen = true;
    c->id_counter = std::max(c->id_counter, c->np);
    return PGPU_OK;
}

extern "C" int pgpu_upload_particles(pgpu_ctx* c, const void* aos, int64_t n, int32_t stride) {
    if (!c) return PGPU_ERR_INVALID;
    c->np = 0;
    c->has_dtrim = false;
    c->maybe_tombstones = false;
    c->id_counter = 0;
    c->cells_unassigned = false; /* (the store is replaced) */
    c->sorted = false; /* (also for n == 0: the offsets of the previous population must not stay marked valid) */
    c->hist_valid = false;
    c->emig_valid = false;
    return append_impl(c, aos, n, stride, "pgpu_upload_particles");
}

int pg_need_ie(pgpu_ctx* c) {
    cudaSetDevice(c->device);
    if (c->capacity == 0 || c->planes[0].ie) return PGPU_OK;
    for (int k = 0; k < 2; k++) {
        int rc;
        if ((rc = dev_alloc(c, &c->planes[k].ie, (size_t)c->capacity))) return rc;
        PG_CUDA(c, cudaMemsetAsync(c->planes[k].ie, 0, sizeof(PgIe) * (size_t)c->capacity, c->stream));
    }
    pg_refresh_dev(c);
    return PGPU_OK;
}
int pg_need_dtrim(pgpu_ctx* c) {
    cudaSetDevice(c->device);
    if (c->capacity == 0 || c->planes[0].dtrim) return PGPU_OK;
    for (int k = 0; k < 2; k++) {
        int rc;
        if ((rc = dev_alloc(c, &c->planes[k].dtrim, (size_t)c->capacity))) return rc;
    }
    pg_refresh_dev(c);
    return PGPU_OK;
}

int pg_fill_dtrim(pgpu_ctx* c) {
    cudaSetDevice(c->device);
    { int rc0; if ((rc0 = pg_need_dtrim(c))) return rc0; }
    if (c->np > 0) fill_f64_kernel<<<(unsigned int)((c->np + 255) / 256), 256, 0, c->stream>>>(c->planes[c->cur].dtrim, c->np, c->cfg.dt);
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}

extern "C" int pgpu_append_particles(pgpu_ctx* c, const void* aos, int64_t n, int32_t stride) {
    if (!c) return PGPU_ERR_INVALID;
    if (n > 0 && c->np > 0 && !c->has_dtrim && c->planes[c->cur].dtrim) {
        /* the DTRIM plane is not maintained while every particle has DTRIM == DT: make it say so before particles that
         * may differ are appended (DTRIM = rf()*DT after an injection, ref: src/timecycle.f90:463).  When the plane does
         * not exist it is created only if the appended records need it (append_impl, second pass). */
        int rc0;
        if ((rc0 = pg_fill_dtrim(c))) return rc0;
    }
    return append_impl(c, aos, n, stride, "pgpu_append_particles");
}

/* ------------------------------------------------------------------------------------------
 * record-plane transfers: the host keeps the particle state in the device's own layout (two planes of 32-byte records,
 * pgpu_pos_record / pgpu_vel_record in pantera_gpu.h, + an optional 16-byte internal-energy plane), so an upload or a
 * download is two or three plain copies at PCIe speed -- 64 bytes per particle instead of the 96-byte AoS record plus
 * a staging transposition.  ID, DUMP_TRAJ and DTRIM (= DT outside an injection step) stay with the host, as SURVEY
 * 8(d) counts them.
 * ---------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) planes_check_kernel(const PgPos* __restrict__ pos, const PgIe* __restrict__ ie, long long n, int ncells,
                                                           PgStatus* st) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ic = pos[i].ic;
    if (ncells > 0 && (ic < 1 || ic > ncells)) atomicCAS(&st->err, 0, PG_DEVERR_CELL_RANGE);
    if (ie) {
        const PgIe E = ie[i];
        if (E.erot != 0. || E.evib != 0.) st->flag_ie = 1;
    }
}

extern "C" int pgpu_upload_planes(pgpu_ctx* c, const void* pos, const void* vel, const void* ie, int64_t n) {
    if (!c) return PGPU_ERR_INVALID;
    if (n < 0 || (n > 0 && (!pos || !vel))) PG_FAIL(c, PGPU_ERR_INVALID, "upload_planes: null plane");
    if (!c->cfg_set) PG_FAIL(c, PGPU_ERR_STATE, "upload_planes: pgpu_set_config first");
    cudaSetDevice(c->device);
    int rc;
    c->np = 0;
    c->has_dtrim = false;
    c->maybe_tombstones = false;
    c->id_counter = 0;
    c->cells_unassigned = false; /* (the store is replaced) */
    c->sorted = false; c->hist_valid = false; c->emig_valid = false; c->vel_deferred = false; c->owned_clean = false;
    if (c->capacity == 0) {
        long long cap = c->cfg.capacity > 0 ? (long long)c->cfg.capacity : (long long)(1.5 * (double)n);
        if ((rc = pg_ensure_capacity(c, std::max<long long>(cap, n)))) return rc;
    } else if ((rc = pg_ensure_capacity(c, n))) {
        return rc;
    }
    if (n == 0) return PGPU_OK;
    if ((rc = pg_reset_status(c))) return rc;
    PgPlanes p = c->planes[c->cur];
    PG_CUDA(c, cudaMemcpyAsync(p.pos, pos, sizeof(PgPos) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(p.vel, vel, sizeof(PgVel) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    if (ie || wants_ie(c)) {
        if ((rc = pg_need_ie(c))) return rc;
        p = c->planes[c->cur];
    }
    if (ie) PG_CUDA(c, cudaMemcpyAsync(p.ie, ie, sizeof(PgIe) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    else if (p.ie) PG_CUDA(c, cudaMemsetAsync(p.ie, 0, sizeof(PgIe) * (size_t)n, c->stream));
    const int ncells = (c->cfg.grid_type == PGPU_UNSTRUCTURED) ? (c->mesh_set ? c->nc : 0) : c->cfg.nx * c->cfg.ny;
    planes_check_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, c->stream>>>(p.pos, ie ? p.ie : nullptr, n, ncells, c->d_status);
    c->np = n;
    if ((rc = pg_finish(c, "pgpu_upload_planes"))) {
        c->np = 0;
        return rc;
    }
    if (c->h_status->flag_ie) c->ie_seen = true;
    c->id_counter = c->np;
    return PGPU_OK;
}

extern "C" int pgpu_download_planes(pgpu_ctx* c, void* pos, void* vel, void* ie, int64_t capacity, int64_t* n) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    if (c->maybe_tombstones && c->np > 0) {
        if ((rc = pg_prepare(c))) return rc;
        if ((rc = pg_reset_status(c))) return rc;
        if ((rc = pg_sort_impl(c))) return rc;
        if ((rc = pg_finish(c, "pgpu_download_planes"))) return rc;
    }
    if ((rc = pg_resolve_vel(c))) return rc;
    if (capacity < c->np) PG_FAIL(c, PGPU_ERR_INVALID, "download_planes: capacity %lld < %lld particles", (long long)capacity, c->np);
    if (n) *n = c->np;
    if (c->np == 0) return PGPU_OK;
    if (!pos || !vel) PG_FAIL(c, PGPU_ERR_INVALID, "download_planes: null plane");
    PgPlanes p = c->planes[c->cur];
    PG_CUDA(c, cudaMemcpyAsync(pos, p.pos, sizeof(PgPos) * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(vel, p.vel, sizeof(PgVel) * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
    if (ie) {
        if (p.ie) PG_CUDA(c, cudaMemcpyAsync(ie, p.ie, sizeof(PgIe) * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
        else memset(ie, 0, sizeof(PgIe) * (size_t)c->np);
    }
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    return PGPU_OK;
}

extern "C" int pgpu_num_particles(pgpu_ctx* c, int64_t* n) {
    if (!c || !n) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->maybe_tombstones && c->np > 0) {
        int rc;
        if ((rc = pg_prepare(c))) return rc;
        if ((rc = pg_reset_status(c))) return rc;
        if ((rc = pg_sort_impl(c))) return rc;
        if ((rc = pg_finish(c, "pgpu_num_particles"))) return rc;
    }
    *n = c->np;
    return PGPU_OK;
}

extern "C" int pgpu_download_particles(pgpu_ctx* c, void* aos, int64_t capacity, int64_t* n, int32_t stride) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (stride < 92 || (stride % 8) != 0) PG_FAIL(c, PGPU_ERR_INVALID, "download: stride %d must be a multiple of 8 and >= 92", stride);
    int rc;
    if (c->maybe_tombstones && c->np > 0) {
        if ((rc = pg_prepare(c))) return rc;
        if ((rc = pg_reset_status(c))) return rc;
        if ((rc = pg_sort_impl(c))) return rc;
        if ((rc = pg_finish(c, "pgpu_download_particles"))) return rc;
    }
    if ((rc = pg_resolve_vel(c))) return rc;
    if (capacity < c->np) PG_FAIL(c, PGPU_ERR_INVALID, "download: capacity %lld < %lld particles", (long long)capacity, c->np);
    if (n) *n = c->np;
    if (c->np == 0) return PGPU_OK;
    if (!aos) PG_FAIL(c, PGPU_ERR_INVALID, "download: null buffer");
    if ((rc = ensure_stage(c, stride))) return rc;
    PgPlanes p = c->planes[c->cur];
    cudaEvent_t evs[2] = {c->ev[14], c->ev[15]};
    bool used[2] = {false, false};
    int k = 0;
    for (long long done = 0; done < c->np; done += STAGE_PARTICLES, k ^= 1) {
        long long m = std::min<long long>(STAGE_PARTICLES, c->np - done);
        if (used[k]) PG_CUDA(c, cudaEventSynchronize(evs[k]));
        soa_to_aos_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, c->stream>>>(
            (unsigned char*)c->d_stage[k], stride, m, p, done, c->cfg.dt, c->cfg.rank, c->has_dtrim ? 1 : 0);
        PG_CUDA(c, cudaMemcpyAsync((unsigned char*)aos + (size_t)done * stride, c->d_stage[k], (size_t)m * stride,
                                   cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaEventRecord(evs[k], c->stream));
        used[k] = true;
    }
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * counting sort by cell (physical reorder of the record planes).
 * The reference builds index permutations NPCALL/IOFALL/INDALL (src/collisions.f90:61-112) and then
 * gathers through them; here the histogram and the arrival rank are produced by the ADVECT epilogue,
 * an exclusive scan gives the cell offsets, and ONE scatter pass moves whole 32-byte records with
 * 256-bit loads and stores (every random write is a full HBM sector).  Tombstones drop out here.
 * The pass that follows the scan also cuts the sorted store into collision tiles (pgpu_collide.cu).
 * ---------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) hist_kernel(const PgPos* __restrict__ pos, const PgStatus* __restrict__ st,
                                                   unsigned int* __restrict__ cell_count) {
    const long long n = (long long)st->np_cur;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c = (i < n) ? pos[i].ic : 0;
    unsigned int live = __ballot_sync(0xffffffffu, c >= 1);
    if (c < 1) return;
    unsigned int mask = __match_any_sync(live, c);
    if ((threadIdx.x & 31) == __ffs(mask) - 1) atomicAdd(&cell_count[c - 1], (unsigned int)__popc(mask));
}

/* the same histogram on a coarse grid (PG_COUNT_SMEM_CELLS or fewer cells): a few hundred counters take the whole
 * store's increments and L2 serialises them, so each block counts in shared memory over a grid-stride sweep and adds
 * its non-zero counters once */
__global__ void __launch_bounds__(256) hist_small_kernel(const PgPos* __restrict__ pos, const PgStatus* __restrict__ st,
                                                         unsigned int* __restrict__ cell_count, int ncells) {
    extern __shared__ unsigned int sh_hist[];
    for (int i = threadIdx.x; i < ncells; i += blockDim.x) sh_hist[i] = 0u;
    __syncthreads();
    const long long n = (long long)st->np_cur;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const int c = (i < n) ? pos[i].ic : 0;
        const unsigned int live = __ballot_sync(0xffffffffu, c >= 1);
        if (c >= 1) {
            const unsigned int mask = __match_any_sync(live, c);
            if ((threadIdx.x & 31) == __ffs(mask) - 1) atomicAdd(&sh_hist[c - 1], (unsigned int)__popc(mask));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncells; i += blockDim.x) {
        const unsigned int v = sh_hist[i];
        if (v) atomicAdd(&cell_count[i], v);
    }
}

/* dst[cell_off[IC-1] + rank] = src: the rank of a particle inside its cell is its arrival order at the cell's cursor
 * (one warp-aggregated atomic per (warp, cell); the store is cell-sorted from the previous step, so the lanes of a warp
 * go to a handful of cells).  The kernel is HBM-bound with most of the issue slots idle, so the atomic's round trip
 * costs nothing here -- it used to sit in the ADVECT epilogue, together with a 4-byte rank per particle written there
 * and read back here. */
__global__ void __launch_bounds__(256) scatter_kernel(PgPlanes src, PgPlanes dst, const PgStatus* __restrict__ st,
                                                      const unsigned int* __restrict__ cell_off,
                                                      unsigned int* __restrict__ cursor, int use_ie,
                                                      unsigned int* __restrict__ src_of) {
    const long long n = (long long)st->np_cur; /* (the grid covers the host's upper bound of it) */
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    /* The source records are read exactly once: streaming (evict-first) loads keep them from displacing the
     * destination lines that L2 is still assembling from the scattered 32-byte stores (measured: -19 % at
     * 3.1e7 particles; streaming the stores as well, or two records per thread, gains nothing more). */
    PgRec32 P;
    int ic = 0;
    if (i < n) {
        P = pg_ld_stream(src.pos + i);
        ic = (int)(P.w & 0xffffffffULL);
    }
    const unsigned int live = __ballot_sync(0xffffffffu, ic >= 1);
    if (ic < 1) return;
    const unsigned int mask = __match_any_sync(live, ic);
    const int leader = __ffs(mask) - 1, lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[ic - 1], (unsigned int)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    const size_t o = (size_t)__ldg(&cell_off[ic - 1]) + base + __popc(mask & ((1u << lane) - 1u));
    *reinterpret_cast<PgRec32*>(dst.pos + o) = P;
    if (src_of) {
        /* deferred: the velocity record stays where it is; the collision kernel fetches it through this index when
         * it sweeps the cell anyway, and writes the sorted plane once (saves a 32-byte read and write per particle) */
        src_of[o] = (unsigned int)i;
    } else {
        const PgRec32 V = pg_ld_stream(src.vel + i);
        *reinterpret_cast<PgRec32*>(dst.vel + o) = V;
    }
    if (use_ie && src.ie) dst.ie[o] = src.ie[i];
    if (src.dtrim) dst.dtrim[o] = src.dtrim[i];
    if (src.id) dst.id[o] = src.id[i];
}

/* The scatter on a coarse grid (PG_COUNT_SMEM_CELLS or fewer cells).  There the per-(warp, cell) cursor atomics above
 * all land on a few hundred addresses and wait for their return values in turn (measured on the 512-segment line:
 * 0.37 ms on a freshly sorted store, 1.4 ms four steps later).  A block takes a tile of 1024 particles, ranks them
 * per cell in shared memory, reserves each touched cell's range with ONE global atomic, then writes. */
#define PG_SCATTER_TILE 4
__global__ void __launch_bounds__(256) scatter_small_kernel(PgPlanes src, PgPlanes dst, const PgStatus* __restrict__ st,
                                                            const unsigned int* __restrict__ cell_off,
                                                            unsigned int* __restrict__ cursor, int use_ie, int ncells) {
    extern __shared__ unsigned int sh_sc[]; /* [ncells] counts, then [ncells] reserved bases */
    unsigned int* const sh_cnt = sh_sc;
    unsigned int* const sh_base = sh_sc + ncells;
    for (int i = threadIdx.x; i < ncells; i += blockDim.x) sh_cnt[i] = 0u;
    __syncthreads();
    const long long n = (long long)st->np_cur;
    const long long tile0 = (long long)blockIdx.x * (256 * PG_SCATTER_TILE);
    const int lane = threadIdx.x & 31;
    PgRec32 P[PG_SCATTER_TILE];
    int ic[PG_SCATTER_TILE];
    unsigned int rk[PG_SCATTER_TILE];
#pragma unroll
    for (int k = 0; k < PG_SCATTER_TILE; k++) {
        const long long i = tile0 + k * 256 + threadIdx.x;
        ic[k] = 0;
        if (i < n) {
            P[k] = pg_ld_stream(src.pos + i);
            ic[k] = (int)(P[k].w & 0xffffffffULL);
        }
    }
#pragma unroll
    for (int k = 0; k < PG_SCATTER_TILE; k++) {
        const unsigned int live = __ballot_sync(0xffffffffu, ic[k] >= 1);
        rk[k] = 0u;
        if (ic[k] >= 1) {
            const unsigned int mask = __match_any_sync(live, ic[k]);
            const int leader = __ffs(mask) - 1;
            unsigned int b = 0;
            if (lane == leader) b = atomicAdd(&sh_cnt[ic[k] - 1], (unsigned int)__popc(mask));
            rk[k] = __shfl_sync(mask, b, leader) + __popc(mask & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncells; i += blockDim.x) {
        const unsigned int v = sh_cnt[i];
        if (v) sh_base[i] = atomicAdd(&cursor[i], v);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PG_SCATTER_TILE; k++) {
        if (ic[k] < 1) continue;
        const long long i = tile0 + k * 256 + threadIdx.x;
        const size_t o = (size_t)__ldg(&cell_off[ic[k] - 1]) + sh_base[ic[k] - 1] + rk[k];
        *reinterpret_cast<PgRec32*>(dst.pos + o) = P[k];
        const PgRec32 V = pg_ld_stream(src.vel + i);
        *reinterpret_cast<PgRec32*>(dst.vel + o) = V;
        if (use_ie && src.ie) dst.ie[o] = src.ie[i];
        if (src.dtrim) dst.dtrim[o] = src.dtrim[i];
        if (src.id) dst.id[o] = src.id[i];
    }
}

/* after the scan: the list of cells too large for one lane of the collision kernel (GLOBAL path), and the new particle
 * count -- the total of the histogram, tombstones left out -- which stays on the device until the call's final read-back */
__global__ void __launch_bounds__(256) cell_tail_kernel(const unsigned int* __restrict__ cell_off, int ncells, unsigned int first_cell,
                                                        unsigned int* __restrict__ big_list, long long big_cap, PgStatus* st) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    unsigned int o = cell_off[c], o1 = cell_off[c + 1];
    if (o1 - o > PG_TILE_BIG) {
        unsigned int k = atomicAdd(&st->n_big, 1u);
        if ((long long)k < big_cap) big_list[k] = first_cell + (unsigned int)c;
    }
    if (c == ncells - 1) st->np_cur = (unsigned long long)o1;
}

/* The collision kernel gives one lane to a cell, and a warp idles behind the lane with the most candidates.  The
 * number of candidates goes with the square of the particle count, so the cells are handed out grouped by particle
 * count (fullest first): three small kernels over the cell offsets -- sizes of the groups, their offsets, and a
 * block-level counting-sort scatter of the cell ids.  Results do not depend on which lane takes which cell. */
__global__ void __launch_bounds__(256) cell_bucket_count_kernel(const unsigned int* __restrict__ cell_off, int ncells,
                                                                unsigned int* __restrict__ bucket) {
    __shared__ unsigned int h[PG_NBUCKET];
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x) h[i] = 0;
    __syncthreads();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) {
        unsigned int n = cell_off[c + 1] - cell_off[c];
        atomicAdd(&h[(n < 2 || n > PG_TILE_BIG) ? 0 : n], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x)
        if (h[i]) atomicAdd(&bucket[i], h[i]);
}
/* bucket[0..NB) sizes -> bucket[NB..2NB) start of each group, groups laid out from the fullest cells down */
__global__ void __launch_bounds__(288) cell_bucket_scan_kernel(unsigned int* __restrict__ bucket) {
    __shared__ unsigned int h[PG_NBUCKET], o[PG_NBUCKET];
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x) h[i] = bucket[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int b = PG_NBUCKET - 1; b >= 0; b--) {
            o[b] = run;
            run += h[b];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x) bucket[PG_NBUCKET + i] = o[i];
}
__global__ void __launch_bounds__(256) cell_bucket_fill_kernel(const unsigned int* __restrict__ cell_off, int ncells, unsigned int first_cell,
                                                               unsigned int* __restrict__ bucket,
                                                               unsigned int* __restrict__ order) {
    __shared__ unsigned int h[PG_NBUCKET], base[PG_NBUCKET];
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x) h[i] = 0;
    __syncthreads();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int b = 0, r = 0;
    if (c < ncells) {
        unsigned int n = cell_off[c + 1] - cell_off[c];
        b = (n < 2 || n > PG_TILE_BIG) ? 0 : n;
        r = atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PG_NBUCKET; i += blockDim.x)
        if (h[i]) base[i] = atomicAdd(&bucket[PG_NBUCKET + i], h[i]);
    __syncthreads();
    if (c < ncells) order[base[b] + r] = first_cell + (unsigned int)c;
}

__global__ void __launch_bounds__(256) resolve_vel_kernel(const PgVel* __restrict__ vel_src, const unsigned int* __restrict__ src_of,
                                                          PgVel* __restrict__ vel, long long n) {
    long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const PgRec32 V = pg_ld_stream(vel_src + __ldcs(src_of + o));
    *reinterpret_cast<PgRec32*>(vel + o) = V;
}

/* a deferred vel reorder that no collision kernel consumed (e.g. a download right after pgpu_sort) */
int pg_resolve_vel(pgpu_ctx* c) {
    if (!c->vel_deferred) return PGPU_OK;
    if (c->np > 0) {
        resolve_vel_kernel<<<(unsigned int)((c->np + 255) / 256), 256, 0, c->stream>>>(c->planes[c->cur ^ 1].vel, c->d_src_of,
                                                                                       c->planes[c->cur].vel, c->np);
        c->launches++;
        PG_CUDA(c, cudaGetLastError());
    }
    c->vel_deferred = false;
    pg_refresh_dev(c);
    return PGPU_OK;
}

/* ---- deterministic mode (pgpu_set_deterministic): a STABLE counting sort -- radix sort of (cell, slot) pairs -- instead
 * of the atomic per-cell cursor, so that the in-cell order, and with it the pair selection and the slot-keyed wall
 * streams, do not depend on the order in which warps happen to reach the atomics: the same input and seed give the
 * same particles bit for bit, run after run (the reference's counting sort is stable too, src/collisions.f90:102-112).
 * Cost at the bench size: see DESIGN.md section 3. */
__global__ void __launch_bounds__(256) det_keys_kernel(const PgPos* __restrict__ pos, long long n, unsigned int ncells,
                                                       unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ic = pos[i].ic;
    keys[i] = (ic >= 1) ? (unsigned int)(ic - 1) : ncells; /* tombstones sort behind the last cell */
    vals[i] = (unsigned int)i;
}
__global__ void __launch_bounds__(256) det_gather_kernel(PgPlanes src, PgPlanes dst, const unsigned int* __restrict__ perm,
                                                         const unsigned int* __restrict__ cell_off, int ncells, int use_ie) {
    const long long n = (long long)cell_off[ncells];
    long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const size_t i = perm[o];
    *reinterpret_cast<PgRec32*>(dst.pos + o) = pg_ld_stream(src.pos + i);
    *reinterpret_cast<PgRec32*>(dst.vel + o) = pg_ld_stream(src.vel + i);
    if (use_ie && src.ie) dst.ie[o] = src.ie[i];
    if (src.dtrim) dst.dtrim[o] = src.dtrim[i];
    if (src.id) dst.id[o] = src.id[i];
}

static int sort_deterministic(pgpu_ctx* c, PgPlanes src, PgPlanes dst, int ncells) {
    /* the item count must be on the host for the radix sort: this mode pays one more synchronisation */
    unsigned long long* hn = reinterpret_cast<unsigned long long*>(&c->h_status->scratch[2]);
    PG_CUDA(c, cudaMemcpyAsync(hn, &c->d_status->np_cur, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long n = (long long)*hn;
    if (c->det_cap < c->capacity) {
        for (int k = 0; k < 4; k++) { cudaFree(c->d_det[k]); c->d_det[k] = nullptr; }
        for (int k = 0; k < 4; k++) PG_CUDA(c, cudaMalloc((void**)&c->d_det[k], sizeof(unsigned int) * (size_t)c->capacity));
        c->det_cap = c->capacity;
    }
    int bits = 1;
    while ((1ULL << bits) <= (unsigned long long)ncells) bits++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->d_det[0], c->d_det[1], c->d_det[2], c->d_det[3], (int)n, 0, bits, c->stream);
    if (tmp > c->det_tmp_bytes) {
        cudaFree(c->d_det_tmp);
        c->d_det_tmp = nullptr;
        PG_CUDA(c, cudaMalloc(&c->d_det_tmp, tmp + 256));
        c->det_tmp_bytes = tmp + 256;
    }
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    if (n > 0) {
        det_keys_kernel<<<blocks, 256, 0, c->stream>>>(src.pos, n, (unsigned int)ncells, c->d_det[0], c->d_det[2]);
        tmp = c->det_tmp_bytes;
        PG_CUDA(c, cub::DeviceRadixSort::SortPairs(c->d_det_tmp, tmp, c->d_det[0], c->d_det[1], c->d_det[2], c->d_det[3], (int)n, 0, bits, c->stream));
        det_gather_kernel<<<blocks, 256, 0, c->stream>>>(src, dst, c->d_det[3], c->d_cell_off, ncells, c->dev.use_ie);
        c->launches += 3;
    }
    return PGPU_OK;
}

extern "C" int pgpu_set_deterministic(pgpu_ctx* c, int32_t on) {
    if (!c) return PGPU_ERR_INVALID;
    c->deterministic = on != 0;
    return PGPU_OK;
}

void pg_cell_range(pgpu_ctx* c, long long* lo, long long* hi) {
    *lo = 0;
    *hi = pg_ncells(c);
    if (c->cfg.nranks > 1 && c->owned_clean && !c->deterministic && !c->h_cell_procs.empty() && c->own_hi > c->own_lo &&
        (long long)c->h_cell_procs.size() == *hi && *hi > PG_COUNT_SMEM_CELLS) {
        *lo = c->own_lo;
        *hi = c->own_hi;
    }
}

int pg_sort_impl(pgpu_ctx* c, bool defer_vel, bool sync_np) {
    int ncells = pg_ncells(c);
    long long lo, hi;
    pg_cell_range(c, &lo, &hi);
    const int nrc = (int)(hi - lo); /* cells in the range */
    { int rc0; if ((rc0 = pg_resolve_vel(c))) return rc0; }
    if (c->deterministic) defer_vel = false;
    c->n_big = 0;
    c->dev.ntiles = 0;
    if (c->np_ub == 0) {
        PG_CUDA(c, cudaMemsetAsync(c->d_cell_off, 0, sizeof(unsigned int) * ((size_t)ncells + 1), c->stream));
        c->sorted = true;
        c->maybe_tombstones = false;
        c->cells_lo = 0;
        c->cells_hi = ncells;
        return PGPU_OK;
    }
    if (c->hist_valid && (c->hist_lo != lo || c->hist_hi != hi)) c->hist_valid = false; /* counted over another range */
    PgPlanes src = c->planes[c->cur], dst = c->planes[c->cur ^ 1];
    if (!c->has_dtrim) src.dtrim = nullptr;
    /* the particle count may have moved on the device since the call began (immigrants): the grids cover the host's
     * upper bound and the kernels read the count itself from the status block */
    unsigned int blocks = (unsigned int)((c->np_ub + 255) / 256);
    if (c->emig_valid) { /* ADVECT left the emigrants out of its histogram and no EXCHANGE has shipped them: recount */
        c->hist_valid = false;
        c->emig_valid = false;
    }
    if (!c->hist_valid) {
        PG_CUDA(c, cudaMemsetAsync(c->d_cell_count + lo, 0, sizeof(unsigned int) * ((size_t)nrc + 1), c->stream));
        if (ncells <= PG_COUNT_SMEM_CELLS) {
            const unsigned int hb = blocks < 148u * 8u ? blocks : 148u * 8u;
            hist_small_kernel<<<hb, 256, sizeof(unsigned int) * (size_t)ncells, c->stream>>>(src.pos, c->d_status, c->d_cell_count, ncells);
        } else {
            hist_kernel<<<blocks, 256, 0, c->stream>>>(src.pos, c->d_status, c->d_cell_count);
        }
        c->launches++;
    }
    size_t tmp = c->scan_tmp_bytes;
    PG_CUDA(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, tmp, c->d_cell_count + lo, c->d_cell_off + lo, nrc + 1, c->stream));
    c->launches++;
    if (c->deterministic) {
        int rcd;
        if ((rcd = sort_deterministic(c, src, dst, ncells))) return rcd;
    } else {
        /* the histogram has been scanned into cell_off: its array now serves as the per-cell cursor of the scatter */
        PG_CUDA(c, cudaMemsetAsync(c->d_cell_count + lo, 0, sizeof(unsigned int) * ((size_t)nrc + 1), c->stream));
        if (ncells <= PG_COUNT_SMEM_CELLS && !defer_vel) {
            const unsigned int sb = (unsigned int)((c->np_ub + 256 * PG_SCATTER_TILE - 1) / (256 * PG_SCATTER_TILE));
            scatter_small_kernel<<<sb, 256, 2 * sizeof(unsigned int) * (size_t)ncells, c->stream>>>(
                src, dst, c->d_status, c->d_cell_off, c->d_cell_count, c->dev.use_ie, ncells);
        } else {
            scatter_kernel<<<blocks, 256, 0, c->stream>>>(src, dst, c->d_status, c->d_cell_off, c->d_cell_count, c->dev.use_ie,
                                                          defer_vel ? c->d_src_of : nullptr);
        }
        c->launches++;
    }
    {
        /* (the kernels see the range as cells 0..nrc-1 of a shifted cell_off and add `lo` to the ids they store) */
        unsigned int cb = (unsigned int)((nrc + 255) / 256);
        cell_tail_kernel<<<cb, 256, 0, c->stream>>>(c->d_cell_off + lo, nrc, (unsigned int)lo, c->d_big_list, c->big_alloc, c->d_status);
        PG_CUDA(c, cudaMemsetAsync(c->d_cell_bucket, 0, sizeof(unsigned int) * 2 * PG_NBUCKET, c->stream));
        cell_bucket_count_kernel<<<cb, 256, 0, c->stream>>>(c->d_cell_off + lo, nrc, c->d_cell_bucket);
        cell_bucket_scan_kernel<<<1, 288, 0, c->stream>>>(c->d_cell_bucket);
        cell_bucket_fill_kernel<<<cb, 256, 0, c->stream>>>(c->d_cell_off + lo, nrc, (unsigned int)lo, c->d_cell_bucket, c->d_cell_order);
        c->launches += 4;
        c->cells_lo = lo;
        c->cells_hi = hi;
    }
    PG_CUDA(c, cudaGetLastError());
    c->cur ^= 1;
    c->sorted = true;
    c->hist_valid = false;
    c->maybe_tombstones = false;
    c->vel_deferred = defer_vel;
    c->np_dev_valid = true; /* the new count (total of the histogram) is in status->np_cur; np_ub still bounds it */
    if (sync_np) {
        /* callers that need the count on the host right away (everything but the per-step loop) */
        unsigned int* h = c->h_status->scratch;
        unsigned long long* hn = reinterpret_cast<unsigned long long*>(&c->h_status->scratch[2]);
        PG_CUDA(c, cudaMemcpyAsync(hn, &c->d_status->np_cur, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaMemcpyAsync(&h[1], &c->d_status->n_big, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        c->np = (long long)*hn;
        c->np_ub = c->np;
        c->n_big = h[1];
    }
    pg_refresh_dev(c);
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * timings
 * ---------------------------------------------------------------------------------------- */
static void tic(pgpu_ctx* c, int g) { cudaEventRecord(c->ev[2 * g], c->stream); }
static void toc(pgpu_ctx* c, int g) { cudaEventRecord(c->ev[2 * g + 1], c->stream); }
static void collect(pgpu_ctx* c, int g) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[2 * g], c->ev[2 * g + 1]) == cudaSuccess) c->ms[g] += ms;
}
static void reset_timings(pgpu_ctx* c) {
    memset(c->ms, 0, sizeof(c->ms));
    c->launches = 0;
}

extern "C" int pgpu_get_timings(pgpu_ctx* c, double* ms, int32_t n, int64_t* launches) {
    if (!c) return PGPU_ERR_INVALID;
    for (int i = 0; i < n && i < 8; i++) ms[i] = c->ms[i];
    if (launches) *launches = c->launches;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * hot-path entry points
 * ---------------------------------------------------------------------------------------- */
extern "C" int pgpu_advect(pgpu_ctx* c, int32_t tID) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    reset_timings(c);
    if ((rc = pg_reset_status(c))) return rc;
    tic(c, 0);
    if ((rc = pg_advect_impl(c, tID))) return rc;
    toc(c, 0);
    rc = pg_finish(c, "pgpu_advect");
    collect(c, 0);
    c->ms[5] = c->ms[0];
    return rc;
}

extern "C" int pgpu_sort(pgpu_ctx* c) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    reset_timings(c);
    if ((rc = pg_reset_status(c))) return rc;
    tic(c, 1);
    if ((rc = pg_sort_impl(c))) return rc;
    toc(c, 1);
    rc = pg_finish(c, "pgpu_sort");
    collect(c, 1);
    c->ms[5] = c->ms[1];
    return rc;
}

extern "C" int pgpu_collide(pgpu_ctx* c, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (!c->dev.sigmas) PG_FAIL(c, PGPU_ERR_STATE, "pgpu_collide: VSS parameters not set");
    reset_timings(c);
    if ((rc = pg_reset_status(c))) return rc;
    bool did_sort = !c->sorted;
    if (did_sort) {
        tic(c, 1);
        if ((rc = pg_sort_impl(c, c->defer_vel))) return rc; /* deferred: the collision kernel finishes the vel reorder */
        toc(c, 1);
    }
    tic(c, 2);
    if ((rc = pg_collide_impl(c, tID))) return rc;
    toc(c, 2);
    rc = pg_finish(c, "pgpu_collide");
    if (did_sort) collect(c, 1);
    collect(c, 2);
    c->ms[5] = c->ms[1] + c->ms[2];
    if (timestep_coll) *timestep_coll = (int64_t)c->h_status->ncoll;
    if (timestep_reac) *timestep_reac = (int64_t)c->h_status->nreac;
    return rc;
}

extern "C" int pgpu_exchange(pgpu_ctx* c) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (c->cfg.nranks <= 1) return PGPU_OK;
    reset_timings(c);
    if ((rc = pg_reset_status(c))) return rc;
    tic(c, 4);
    if ((rc = pg_exchange_impl(c))) return rc;
    toc(c, 4);
    rc = pg_finish(c, "pgpu_exchange");
    collect(c, 4);
    c->ms[5] = c->ms[4];
    return rc;
}

int pg_mcc_impl(pgpu_ctx* c, int tID);     /* pgpu_pic_mcc.cu */
int pg_mcc_vss_impl(pgpu_ctx* c, int tID); /* pgpu_collide.cu */

/* nsteps x [ADVECT -> EXCHANGE -> collisions], ref: src/timecycle.f90:271-299 */
extern "C" int pgpu_run_steps(pgpu_ctx* c, int32_t tID0, int32_t nsteps, int64_t* total_coll, int64_t* total_reac) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if (c->cfg.collision_type == PGPU_DSMC && !c->dev.sigmas) PG_FAIL(c, PGPU_ERR_STATE, "pgpu_run_steps: VSS parameters not set");
    if (c->cfg.collision_type == PGPU_BGK || c->cfg.collision_type == PGPU_DSMC_VAHEDI)
        PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "pgpu_run_steps: collision type %d is outside the hot path (SURVEY 2.1 rows 22, 25)", c->cfg.collision_type);
    reset_timings(c);
    int64_t tc = 0, tr = 0;
    cudaEvent_t e0 = c->ev[12], e1 = c->ev[13];
    cudaEventRecord(e0, c->stream);
    for (int s = 0; s < nsteps; s++) {
        int tID = tID0 + s;
        if ((rc = pg_reset_status(c))) return rc;
        tic(c, 0);
        if ((rc = pg_advect_impl(c, tID))) return rc;
        toc(c, 0);
        if (c->cfg.nranks > 1) {
            tic(c, 4);
            if ((rc = pg_exchange_impl(c))) return rc;
            toc(c, 4);
        }
        bool dsmc = c->cfg.collision_type == PGPU_DSMC;
        bool mcc = c->cfg.collision_type == PGPU_MCC_VAHEDI;
        bool mccv = c->cfg.collision_type == PGPU_MCC;
        /* The whole step is queued without a host round trip: the particle count lives in the status block while the
         * step runs (immigrants, compaction), and ONE read-back at the end brings it, the counters and any device
         * error to the host.  Oversized cells (GLOBAL collision path, host-driven dependency rounds) are only known
         * then; they are rare (the one-cell heat bath), and run after the read-back. */
        tic(c, 1);
        if ((rc = pg_sort_impl(c, dsmc && c->defer_vel, false))) return rc;
        toc(c, 1);
        tic(c, 2);
        if (dsmc && (rc = pg_collide_impl(c, tID, false))) return rc;
        if (mcc && (rc = pg_mcc_impl(c, tID))) return rc;
        if (mccv && (rc = pg_mcc_vss_impl(c, tID))) return rc;
        toc(c, 2);
        rc = pg_finish(c, "pgpu_run_steps");
        collect(c, 0); collect(c, 1); collect(c, 2);
        if (c->cfg.nranks > 1) collect(c, 4);
        if (rc) return rc;
        if (dsmc && (c->n_big > 0 || c->vel_deferred)) {
            tic(c, 2);
            if ((rc = pg_collide_big(c))) return rc;
            toc(c, 2);
            rc = pg_finish(c, "pgpu_run_steps");
            collect(c, 2);
            if (rc) return rc;
        }
        tc += (int64_t)c->h_status->ncoll;
        tr += (int64_t)c->h_status->nreac;
        pg_refresh_dev(c);
    }
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->ms[5] = ms;
    if (total_coll) *total_coll = tc;
    if (total_reac) *total_reac = tr;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * counters / diagnostics
 * ---------------------------------------------------------------------------------------- */
extern "C" int pgpu_get_boundary_counts(pgpu_ctx* c, int64_t* b, int64_t nb, int64_t* w, int64_t nw) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (b && nb > 0 && c->d_bcount)
        PG_CUDA(c, cudaMemcpy(b, c->d_bcount, sizeof(int64_t) * (size_t)std::min<int64_t>(nb, 4 * c->ns), cudaMemcpyDeviceToHost));
    if (w && nw > 0 && c->d_wcount)
        PG_CUDA(c, cudaMemcpy(w, c->d_wcount, sizeof(int64_t) * (size_t)std::min<int64_t>(nw, c->nw * c->ns), cudaMemcpyDeviceToHost));
    return PGPU_OK;
}

extern "C" int pgpu_get_boundary_tallies(pgpu_ctx* c, double* t, int64_t n, int32_t reset) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->d_tallies) return PGPU_OK;
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    PG_CUDA(c, cudaMemcpy(t, c->d_tallies, sizeof(double) * (size_t)std::min<int64_t>(n, c->tallies_n), cudaMemcpyDeviceToHost));
    if (reset) PG_CUDA(c, cudaMemset(c->d_tallies, 0, sizeof(double) * (size_t)c->tallies_n));
    return PGPU_OK;
}

extern "C" int pgpu_get_cell_counts(pgpu_ctx* c, int32_t* npc, int64_t nc) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->sorted) PG_FAIL(c, PGPU_ERR_STATE, "get_cell_counts: the store is not sorted (call pgpu_sort or pgpu_collide first)");
    int64_t m = std::min<int64_t>(nc, pg_ncells(c));
    /* cell_off describes the cells [cells_lo, cells_hi) (the rank's own range once an EXCHANGE has run); the rest are empty */
    const int64_t lo = std::min<int64_t>(c->cells_lo, m), hi = std::min<int64_t>(c->cells_hi, m);
    std::vector<unsigned int> off((size_t)(hi - lo) + 1);
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hi > lo) PG_CUDA(c, cudaMemcpy(off.data(), c->d_cell_off + lo, sizeof(unsigned int) * ((size_t)(hi - lo) + 1), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < m; i++) npc[i] = (i >= lo && i < hi) ? (int32_t)(off[i - lo + 1] - off[i - lo]) : 0;
    return PGPU_OK;
}
