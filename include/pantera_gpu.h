/*
 * pantera_gpu.h -- C ABI of the B200-native particle hot path for PANTERA PIC-DSMC.
 *
 * Boundary.  The reference has NO foreign-function interface: its
 * hot path is a set of argument-less Fortran module procedures called from TIME_LOOP
 * (src/timecycle.f90:106-362) that act on module globals (src/global.f90).  Each entry
 * point below replaces one of those procedures (or carries the globals it reads) and is
 * what an ISO_C_BINDING interface block in the Fortran host binds to (see INTEGRATION.md
 * for the gpu_bridge module and the #ifdef PANTERA_GPU call-site patch).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 (PGPU_OK) on success, non-zero on error; the message is
 *     available from pgpu_last_error().  The Fortran shim forwards it to ERROR_ABORT
 *     (src/screen.f90:113-124), which is the reference's only error convention.
 *   - arrays are passed exactly as Fortran stores them: column-major (first index fastest),
 *     1-based cell/node/species/physical-group ids, -1 = "none".
 *   - host arrays are copied during the call; the library never keeps host pointers.
 *   - one context <-> one CUDA device <-> one host thread; calls are synchronous on return.
 *   - there is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef PANTERA_GPU_H
#define PANTERA_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGPU_OK 0
#define PGPU_ERR_INVALID 1
#define PGPU_ERR_CUDA 2
#define PGPU_ERR_STATE 3
#define PGPU_ERR_UNSUPPORTED 4
#define PGPU_ERR_NCCL 5

/* ---- enumerations: numeric values are those of the reference's ENUM, BIND(C) blocks ---- */
/* GRID_TYPE, src/global.f90:102-105 */
enum { PGPU_RECTILINEAR_UNIFORM = 0, PGPU_RECTILINEAR_NONUNIFORM = 1, PGPU_QUADTREE = 2, PGPU_UNSTRUCTURED = 3 };
/* particle boundary conditions, src/global.f90:205-207 */
enum { PGPU_VACUUM = 0, PGPU_SPECULAR = 1, PGPU_DIFFUSE = 2, PGPU_CLL = 3, PGPU_REACT = 4, PGPU_AXIS = 5,
       PGPU_PERIODIC_MASTER = 6, PGPU_PERIODIC_SLAVE = 7, PGPU_EMIT = 8, PGPU_WB_BC = 9 };
/* field boundary conditions, src/global.f90:209-212 */
enum { PGPU_DIRICHLET_BC = 0, PGPU_NEUMANN_BC = 1, PGPU_DIELECTRIC_BC = 2, PGPU_ROBIN_BC = 3,
       PGPU_PERIODIC_MASTER_BC = 4, PGPU_PERIODIC_SLAVE_BC = 5, PGPU_RF_VOLTAGE_BC = 6,
       PGPU_DECOUPLED_RF_VOLTAGE_BC = 7, PGPU_NO_BC = 8 };
/* volume type of a physical group, src/global.f90:214-216 */
enum { PGPU_FLUID = 0, PGPU_SOLID = 1 };
/* PIC_TYPE, src/global.f90:302-305 (only NONE and EXPLICIT are on the hot path) */
enum { PGPU_PIC_NONE = 0, PGPU_PIC_EXPLICIT = 1 };
/* COLLISION_TYPE, src/global.f90:385-388 */
enum { PGPU_NO_COLL = 0, PGPU_DSMC = 1, PGPU_BGK = 2, PGPU_MCC = 3, PGPU_MCC_VAHEDI = 4, PGPU_DSMC_VAHEDI = 5 };
/* reaction TYPE, src/global.f90:538-540 */
enum { PGPU_FIXED_RATE = 0, PGPU_TCE = 1, PGPU_LXCAT = 2 };
/* PARTITION_STYLE, src/global.f90:108-111 */
enum { PGPU_STRIPSX = 0, PGPU_STRIPSY = 1, PGPU_STRIPSZ = 2, PGPU_SLICESX = 3, PGPU_SLICESZ = 4 };

/*
 * Host particle record = TYPE PARTICLE_DATA_STRUCTURE (src/particle.f90:28-37) as gfortran
 * lays it out (natural alignment, 96-byte stride).  Fields at byte offsets
 * X 0, Y 8, Z 16, VX 24, VY 32, VZ 40, EROT 48, EVIB 56, DTRIM 64, IC 72, S_ID 76,
 * ID 80, DUMP_TRAJ 88.  The upload/download calls take the stride explicitly so that a
 * host compiled with a different padding can pass STORAGE_SIZE(particles(1))/8.
 */
typedef struct pgpu_particle {
    double X, Y, Z;
    double VX, VY, VZ;
    double EROT, EVIB;
    double DTRIM;
    int32_t IC;
    int32_t S_ID;
    int64_t ID;
    int32_t DUMP_TRAJ;
    int32_t pad_;
} pgpu_particle;

/*
 * Scalar globals read by the hot path (src/global.f90; input keywords in SURVEY.md App. D).
 */
typedef struct pgpu_config {
    int32_t dims;             /* DIMS         global.f90, Dimensions:            */
    int32_t axi;              /* AXI          Axisymmetric:                      */
    int32_t grid_type;        /* GRID_TYPE                                        */
    int32_t nx, ny;           /* NX, NY       Number_of_cells: (structured)      */
    double xmin, xmax, ymin, ymax, zmin, zmax; /* Domain_limits:                  */
    double dt;                /* DT           Timestep:                          */
    double fnum;              /* FNUM         Fnum:                              */
    int32_t pic_type;         /* PIC_TYPE                                         */
    int32_t collision_type;   /* COLLISION_TYPE                                   */
    int32_t bool_periodic[4]; /* BOOL_PERIODIC(1:4)  xmin,xmax,ymin,ymax sides    */
    int32_t bool_specular[4]; /* BOOL_SPECULAR(1:4)                               */
    int32_t bool_diffuse[4];  /* BOOL_DIFFUSE(1:4)                                */
    int32_t bool_react[4];    /* BOOL_REACT(1:4)                                  */
    int32_t bool_z_periodic;  /* BOOL_Z_PERIODIC                                  */
    int32_t bool_y_periodic;  /* BOOL_Y_PERIODIC (DIMS==1 only)                   */
    double boundtemp;         /* BOUNDTEMP    Boundary_temperature:              */
    double external_b[3];     /* EXTERNAL_B_FIELD                                 */
    double eps_scaling;       /* EPS_SCALING                                      */
    double mcc_bg_dens;       /* MCC_BG_DENS                                      */
    double mcc_bg_ttra;       /* MCC_BG_TTRA                                      */
    int64_t rng_seed_global;  /* RNG_SEED_GLOBAL; local seed = seed*(rank+1)      */
    int32_t rank, nranks;     /* PROC_ID, N_MPI_THREADS                           */
    int32_t dump_bound_start; /* DUMP_BOUND_START     (tally gating)              */
    int32_t dump_bound_avg_every; /* DUMP_BOUND_AVG_EVERY; <=0 disables tallies   */
    int64_t capacity;         /* device particle capacity (0 = 1.5 x first upload)*/
    int32_t track_id;         /* 1: keep particles(:)%ID on the device (8 B/particle more, permuted
                                 by the sort); 0: IDs are regenerated on download, as the reference
                                 itself does on restart (src/tools.f90:465-517)                  */
    int32_t reserved_;        /* must be 0                                        */
} pgpu_config;

typedef struct pgpu_ctx pgpu_ctx;

/* ---- life cycle ------------------------------------------------------------------------
 * created between INITFIELDS and TIME_LOOP in PROGRAM PANTERA (src/pantera.f90:51-87). */
int pgpu_create(pgpu_ctx** ctx, int device);
/* at the end of PROGRAM PANTERA (src/pantera.f90:91-98) */
int pgpu_destroy(pgpu_ctx* ctx);
/* text of the last error of this context; the shim passes it to ERROR_ABORT (src/screen.f90:113-124) */
const char* pgpu_last_error(const pgpu_ctx* ctx);
/* library/ABI version: major*1000+minor */
int pgpu_version(void);

/* ---- parameters (module globals the path reads) ---------------------------------------- */
/* the scalars of src/global.f90 listed at struct pgpu_config above (set by READINPUT, src/initialization.f90:38-560) */
int pgpu_set_config(pgpu_ctx* ctx, const pgpu_config* cfg);

/* SPECIES(:) -- src/global.f90:488-508: MOLECULAR_MASS, CHARGE, ROTDOF, VIBDOF, ROTREL, VIBREL */
int pgpu_set_species(pgpu_ctx* ctx, int ns, const double* mass, const double* charge,
                     const int32_t* rotdof, const int32_t* vibdof,
                     const double* rotrel, const double* vibrel);

/* VSS_SIGMAS/OMEGAS/ALPHAS/GREFS(NS,NS) and SIGMAMAX -- src/initialization.f90:675-702 */
int pgpu_set_vss(pgpu_ctx* ctx, int ns, const double* sigmas, const double* omegas,
                 const double* alphas, const double* grefs, double sigmamax);

/* REACTIONS(:) -- src/global.f90:542-558; C1..C3 from INITREACTIONS (initialization.f90:2681-2724);
 * LXCat tables concatenated, table of reaction r is tab_e/tab_cs[tab_off[r] .. tab_off[r+1]-1]. */
int pgpu_set_reactions(pgpu_ctx* ctx, int nr, const int32_t* type,
                       const int32_t* r1, const int32_t* r2,
                       const int32_t* p1, const int32_t* p2, const int32_t* p3, const int32_t* p4,
                       const int32_t* nprod, const int32_t* is_cex,
                       const double* ea, const double* c1, const double* c2, const double* c3,
                       const int32_t* tab_off, const double* tab_e, const double* tab_cs);

/* WALL_REACTIONS(:) -- src/fields.f90:4315-4338 reads R_SP_ID, P1_SP_ID, N_PROD, PROB */
int pgpu_set_wall_reactions(pgpu_ctx* ctx, int n, const int32_t* r_sp, const int32_t* p1_sp,
                            const int32_t* nprod, const double* prob);

/* MIXTURES(MCC_BG_MIX) components -- src/collisions.f90:1653-1657 */
int pgpu_set_mcc_background(pgpu_ctx* ctx, int ncomp, const int32_t* sp_id, const double* molfrac);

/* MCC_BG_CELL_NRHO(N_SPECIES, NCELLS) -- src/global.f90:392, filled by src/initialization.f90:2752-2788
 * (BOOL_BG_DENSITY_FILE); column-major as Fortran stores it.  nrho == NULL clears it (uniform background). */
int pgpu_set_mcc_cell_density(pgpu_ctx* ctx, int ns, int64_t nc, const double* nrho);

/* U1D/U2D/U3D_GRID -- src/global.f90:128-193, as computed by READ_*_SU2
 * (src/grid_and_partition.f90:627-2147).  d = dims.  Shapes (Fortran order):
 *   node_coords(3,nn) cell_nodes(d+1,nc) cell_neighbors(d+1,nc)
 *   face_normal(3,d+1,nc) [EDGE_NORMAL in 1-D/2-D] face_tang1/2(3,4,nc) [3-D only, else NULL]
 *   cell_edges_len(3,nc) [2-D only] cell_faces_pg(d+1,nc) [CELL_EDGES_PG in 1-D/2-D]
 *   cell_pg(nc) basis_coeffs(d+1,d+1,nc) cell_volumes(nc) cell_procs(nc) [NULL = all 0]
 *   face_boundary_index(d+1,nc) + boundary_area(nboundcells): boundary sub-mesh element of a
 *   face and its area for TALLY_PARTICLE_TO_BOUNDARY (postprocess.f90:1098-1147); may be NULL. */
int pgpu_set_mesh(pgpu_ctx* ctx, int dims, int nn, int nc,
                  const double* node_coords, const int32_t* cell_nodes, const int32_t* cell_neighbors,
                  const double* face_normal, const double* face_tang1, const double* face_tang2,
                  const double* cell_edges_len, const int32_t* cell_faces_pg, const int32_t* cell_pg,
                  const double* basis_coeffs, const double* cell_volumes, const int32_t* cell_procs,
                  const int32_t* face_boundary_index, int nboundcells, const double* boundary_area);

/* CELL_PROCS re-cut (load balance, src/timecycle.f90:185-189) */
int pgpu_set_cell_procs(pgpu_ctx* ctx, int nc, const int32_t* cell_procs);
/* BOOL_RADIAL_WEIGHTING: CELL_FNUM(NX*NY) of a rectilinear axisymmetric grid (src/initialization.f90:2304-2310).  With
 * it ADVECT deletes / clones a particle that changed cell with the ratio of the cells' weights
 * (src/timecycle.f90:2110-2128), VSS_COLLIS, DEPOSIT_CHARGE and CHECKS use the cell's weight instead of FNUM
 * (src/collisions.f90:192-196, src/fields.f90:5507-5508, src/postprocess.f90:1764-1768).  NULL switches it off. */
int pgpu_set_cell_fnum(pgpu_ctx* ctx, int nc, const double* cell_fnum);

/* GRID_BC(:) -- src/global.f90:218-253 */
int pgpu_set_bcs(pgpu_ctx* ctx, int nbc, const int32_t* particle_bc, const int32_t* field_bc,
                 const int32_t* volume_bc, const double* wall_temp, const double* acc_n,
                 const double* acc_t, const int32_t* react, const int32_t* dump_fluxes);

/* WALLS(:) interior segments of structured grids -- src/initialization.f90:996-1037 */
int pgpu_set_walls(pgpu_ctx* ctx, int nw, const double* x1, const double* y1, const double* x2,
                   const double* y2, const double* normx, const double* normy, const double* temp,
                   const int32_t* specular, const int32_t* diffuse, const int32_t* porous,
                   const double* transmissivity, const int32_t* react);

/* XCOORD(NX+1), YCOORD(NY+1) of RECTILINEAR_NONUNIFORM grids -- initialization.f90:1939-2009 */
int pgpu_set_grid_coords(pgpu_ctx* ctx, int nxp, const double* xcoord, int nyp, const double* ycoord);

/* ---- particle store (particles(1:NP_PROC), src/global.f90:45-47) ----------------------- */
int pgpu_upload_particles(pgpu_ctx* ctx, const void* aos, int64_t n, int32_t stride_bytes);
int pgpu_download_particles(pgpu_ctx* ctx, void* aos, int64_t capacity, int64_t* n, int32_t stride_bytes);
/* ADD_PARTICLE_ARRAY for host-side injection (src/particle.f90:86-136) */
int pgpu_append_particles(pgpu_ctx* ctx, const void* aos, int64_t n, int32_t stride_bytes);
/* NP_PROC (src/global.f90:45) */
int pgpu_num_particles(pgpu_ctx* ctx, int64_t* n);

/* ---- fields ---------------------------------------------------------------------------- */
/* E_FIELD as APPLY_E_FIELD reads it (src/fields.f90:4194-4223): unstructured E_FIELD(3,1,NCELLS); structured
 * E_FIELD(3,0:NPY-1,0:NPX-1)  (n doubles) */
int pgpu_set_efield(pgpu_ctx* ctx, const double* e_field, int64_t n);
/* B_FIELD(3,1,NNODES) (2-D unstructured nodal, src/fields.f90:4227-4249); NULL/0 clears */
int pgpu_set_bfield(pgpu_ctx* ctx, const double* b_field, int64_t n);
/* SURFACE_CHARGE(NNODES) read back / reset (dielectric walls, timecycle.f90:1399-1430) */
int pgpu_get_surface_charge(pgpu_ctx* ctx, double* sc, int64_t n);
int pgpu_set_surface_charge(pgpu_ctx* ctx, const double* sc, int64_t n);

/* Record-plane transfers: the particle state in the device's own layout -- a plane of 32-byte position records
 * and a plane of 32-byte velocity records (+ optionally a plane of {EROT, EVIB}) -- so that an upload / download is
 * plain copies at PCIe speed: 64 (80) bytes per particle instead of the 96-byte AoS record.  ID, DUMP_TRAJ and DTRIM
 * (= DT) stay on the host side.  Semantics of upload / download as pgpu_upload_particles / pgpu_download_particles. */
typedef struct pgpu_pos_record { double X, Y, Z; int32_t IC, S_ID; } pgpu_pos_record;      /* 32 bytes */
typedef struct pgpu_vel_record { double VX, VY, VZ; int32_t S_ID, pad_; } pgpu_vel_record; /* 32 bytes */
typedef struct pgpu_ie_record { double EROT, EVIB; } pgpu_ie_record;                       /* 16 bytes */
int pgpu_upload_planes(pgpu_ctx* ctx, const void* pos, const void* vel, const void* ie_or_null, int64_t n);
int pgpu_download_planes(pgpu_ctx* ctx, void* pos, void* vel, void* ie_or_null, int64_t capacity, int64_t* n);

/* ---- the hot path ---------------------------------------------------------------------- */
/* DEPOSIT_CHARGE particle loop + SURFACE_CHARGE + reduction over ranks (src/fields.f90:5440-5535).
 * rhs_host(0:NNODES-1) receives the all-reduced RHS; the Dirichlet/Neumann row fill and the
 * PETSc vector load (fields.f90:5538-5561) stay on the host. */
int pgpu_deposit_charge(pgpu_ctx* ctx, double* rhs_host, int64_t nnodes);

/* ADVECT (src/timecycle.f90:1021-2268): kick (Boris) + move + boundary conditions + removal +
 * structured re-index.  tID = current timestep (tally gating and RNG stream). */
int pgpu_advect(pgpu_ctx* ctx, int32_t tID);

/* EXCHANGE (src/grid_and_partition.f90:525-622) over NCCL; no-op for nranks == 1 */
int pgpu_exchange(pgpu_ctx* ctx);

/* EXCHANGE, local half only, for a host that carries the all-to-all itself (the reference's MPI_ALLTOALL +
 * MPI_ALLTOALLV, grid_and_partition.f90:590-605): removes every particle whose cell belongs to another rank and
 * returns them as 96-byte host records grouped by destination rank, sendcount(0:NRANKS-1) = records per
 * destination (the pack + COUNTING_SORT of :562-588).  The receiving side is pgpu_append_particles
 * (the ADD_PARTICLE_ARRAY loop of :610-612).  Needs no communicator. */
int pgpu_exchange_pack(pgpu_ctx* ctx, void* sendbuf_aos, int64_t capacity, int64_t* sendcount);

/* Deterministic mode: the counting sort becomes STABLE (radix sort of (cell, slot) pairs instead of the atomic per-cell
 * cursor), so that the in-cell order -- which the pair selection indexes and the slot-keyed wall streams depend on --
 * is a function of the input alone: the same particles and seed reproduce a run bit for bit, as the reference's stable
 * counting sort does (src/collisions.f90:102-112).  Off by default (cost: DESIGN.md).  Particles created on the device
 * within one call (reaction products, injection, immigrants) are appended in cursor order, which this mode does not fix. */
int pgpu_set_deterministic(pgpu_ctx* ctx, int32_t on);

/* counting sort by cell of DSMC_COLLISIONS (src/collisions.f90:61-112), done as a physical
 * SoA reorder.  Called implicitly by pgpu_collide when the store is not sorted. */
int pgpu_sort(pgpu_ctx* ctx);

/* DSMC_COLLISIONS -> VSS_COLLIS (src/collisions.f90:37-558); returns TIMESTEP_COLL/REAC */
int pgpu_collide(pgpu_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);

/* MCC_COLLISIONS_VAHEDI (src/collisions.f90:1606-1892) */
int pgpu_mcc_vahedi(pgpu_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);

/* MCC_COLLISIONS (src/collisions.f90:1261-1595): VSS cross-section MCC against the background mixture,
 * uniform (MCC_BG_DENS x MOLFRAC) or per cell (pgpu_set_mcc_cell_density) */
int pgpu_mcc_collisions(pgpu_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);

/* REMOVE_PARTICLES_OF_SPECIES (src/timecycle.f90:2340-2353) */
int pgpu_remove_species(pgpu_ctx* ctx, int32_t sp_id);

/* nsteps x [ADVECT -> EXCHANGE -> DSMC_COLLISIONS] without returning to the host in between
 * (the PIC_TYPE==NONE branch of TIME_LOOP, src/timecycle.f90:271-299).  First step uses tID0. */
int pgpu_run_steps(pgpu_ctx* ctx, int32_t tID0, int32_t nsteps,
                   int64_t* total_coll, int64_t* total_reac);

/* ---- particle creation on the device -------------------------------- */
/* SPECIES(:)%SPWT (default 1) and MIXTURES(:): components of mixture k (0-based) are comp_sp (1-based species ids) /
 * comp_molfrac [comp_off[k], comp_off[k+1]) */
int pgpu_set_species_weights(pgpu_ctx* ctx, int ns, const double* spwt);
int pgpu_set_mixtures(pgpu_ctx* ctx, int nmix, const int32_t* comp_off, const int32_t* comp_sp, const double* comp_molfrac);
/* INITIAL_PARTICLES_TASKS / VOLUME_INJECT_TASKS (src/global.f90:275-294), mix_id 0-based */
int pgpu_set_initial_particles_tasks(pgpu_ctx* ctx, int n, const int32_t* mix_id, const double* nrho,
                                     const double* ux, const double* uy, const double* uz,
                                     const double* ttrax, const double* ttray, const double* ttraz,
                                     const double* trot, const double* tvib);
int pgpu_set_volume_inject_tasks(pgpu_ctx* ctx, int n, const int32_t* mix_id, const double* nrhodot,
                                 const double* ux, const double* uy, const double* uz,
                                 const double* ttrax, const double* ttray, const double* ttraz,
                                 const double* trot, const double* tvib);
/* EMIT_TASKS (src/global.f90:257-267) as INITINJECTION leaves them (src/initialization.f90:2575-2675): cell, face
 * (1-based), face nodes IV1/IV2 (1-D/2-D), U_NORM, and NFS(component) = particles per step per rank at
 * nfs[nfs_off[task] + component] */
int pgpu_set_emit_tasks(pgpu_ctx* ctx, int n, const int32_t* mix_id, const int32_t* ic, const int32_t* iv1,
                        const int32_t* iv2, const int32_t* iface, const int32_t* nfs_off, const double* nfs,
                        const double* ux, const double* uy, const double* uz, const double* u_norm,
                        const double* ttra, const double* trot, const double* tvib);
/* INITIAL_SEED (src/initialization.f90:2017-2233) without its trailing EXCHANGE; VOLUME_INJECT (src/timecycle.f90:553-737);
 * BOUNDARIES_EMIT (src/timecycle.f90:368-547).  The particles are created in the append region of the device store. */
int pgpu_initial_seed(pgpu_ctx* ctx);
int pgpu_volume_inject(pgpu_ctx* ctx, int32_t tID);
int pgpu_boundaries_emit(pgpu_ctx* ctx, int32_t tID);

/* ---- field step on the device ---------------------------------------- */
/* GRID_BC(:)%WALL_POTENTIAL, %WALL_RF_POTENTIAL, %RF_FREQUENCY, %WALL_EFIELD, %EPS_REL (src/global.f90:225-233), one
 * entry per physical group as in pgpu_set_bcs; NULL arrays mean 0 (EPS_REL: 1) */
int pgpu_set_field_bcs(pgpu_ctx* ctx, int nbc, const double* wall_potential, const double* wall_rf_potential,
                       const double* rf_frequency, const double* wall_efield, const double* eps_rel);
/* COMPUTE_E_FIELD (src/fields.f90:5237-5296) on an unstructured mesh from the host's PHI_FIELD(NNODES): the PETSc solve
 * stays on the host, only phi comes down; E_FIELD(3,1,NCELLS) becomes the field of the next ADVECT (e_host optional) */
int pgpu_compute_e_field(pgpu_ctx* ctx, const double* phi_host, int64_t nnodes, double* e_host);
/* 1-D line meshes: SET_WALL_POTENTIAL (src/fields.f90:5566-5607) + the Poisson system of ASSEMBLE_POISSON
 * (:137-253, 758-760) with the right-hand side rows of DEPOSIT_CHARGE (:5538-5561) + COMPUTE_E_FIELD, all on the
 * device.  rhs_host == NULL uses the RHS of the last pgpu_deposit_charge.  phi_host, e_host optional outputs. */
int pgpu_field_solve_1d(pgpu_ctx* ctx, int32_t tID, const double* rhs_host, double* phi_host, double* e_host);
/* Any unstructured mesh.  1-D: as above.  2-D (plane / axisymmetric triangles) and 3-D (tetrahedra): the P1 stiffness
 * matrix of ASSEMBLE_POISSON (src/fields.f90:256-424, 425-637; Dirichlet, RF-voltage and unused nodes as identity rows,
 * :639-646, :758-760; NEUMANN(:) from the boundary faces, :391-421, :579-633; EPS_REL of the cell's volume group),
 * SET_WALL_POTENTIAL at timestep tID (:5609-5711), the right-hand side rows of DEPOSIT_CHARGE (:5538-5547) and
 * COMPUTE_E_FIELD, with the linear solve (KSPSolve in the reference, src/fields.f90:5205-5234) done on the device by a
 * Jacobi-preconditioned conjugate gradient: it stops when |r| <= max(rtol |b|, 1e-50) or after max_iter iterations
 * (rtol <= 0: 1e-5 and max_iter <= 0: 10000, PETSc's KSP defaults), starting from the previous potential.
 * iters_out / relres_out (optional): iterations used and the relative residual reached; running out of iterations is
 * not an error (the reference prints KSPConvergedReason and goes on) -- compare relres with the tolerance.
 * PIC_TYPE explicit only (EXPLICITLIMITED's DXLDRATIO factors are not assembled). */
int pgpu_field_solve(pgpu_ctx* ctx, int32_t tID, const double* rhs_host, double* phi_host, double* e_host, double rtol,
                     int32_t max_iter, int32_t* iters_out, double* relres_out);

/* ---- dump / restart -------------------------------------------------- */
/* DUMP_PARTICLES_FILE (src/tools.f90:327-365) / READ_PARTICLES_FILE (src/tools.f90:465-517): the file
 * <path_prefix>proc_<rank:05d>_time_<timestep:08d> in the reference's binary layout (Fortran sequential unformatted,
 * big-endian, one 80-byte record per particle: X Y Z VX VY VZ EROT EVIB S_ID IC DTRIM), written from / read into the
 * device store without a host copy of the particle array.  binary must be non-zero (Binary_output: .TRUE., the
 * reference's default); the file must not exist when dumping (STATUS='NEW').  Reading appends and, like the
 * reference, gives every restarted particle DTRIM = DT. */
int pgpu_dump_particles_file(pgpu_ctx* ctx, const char* path_prefix, int32_t timestep, int32_t binary);
int pgpu_read_particles_file(pgpu_ctx* ctx, const char* path_prefix, int32_t timestep, int32_t binary);
/* binary = 1: the IC of every record must be a cell of the current grid (a dump of the same mesh; the run can go on at
 * once).  binary = 3 (1 + PGPU_RESTART_UNASSIGNED): the reference's restart semantics -- IC is not trusted (records whose
 * IC is no cell of this grid are kept with IC = 0) and every hot-path entry point refuses to run until
 * pgpu_reassign_particles_to_cells has located the particles again, as PROGRAM PANTERA does (src/pantera.f90:58-70). */
#define PGPU_RESTART_UNASSIGNED 2
/* REASSIGN_PARTICLES_TO_CELLS_1D / _2D / _3D (src/tools.f90:925-1039, 782-922, 577-778): every particle is located again
 * from its position -- the domain box (pgpu_config xmin..zmax = Domain_limits:) cut into 100 bins per direction, the
 * cells entered in the bins their bounding boxes overlap, the lowest-numbered cell of the particle's bin whose P1
 * basis functions are all >= 0 at the particle (`IF (PSIP < 0) INSIDE = .FALSE.`) becomes its IC; particles outside the
 * box or inside no cell are removed.  n_removed (optional) receives their number.  Unstructured grids only (the
 * reference: 'Not implemented!' otherwise).  Bin indices are clamped to 1..100 where the reference would leave its
 * arrays (a cell or particle exactly at XMAX). */
int pgpu_reassign_particles_to_cells(pgpu_ctx* ctx, int64_t* n_removed);

/* ---- sampling and conservation checks ------------------------------- */
/* GRID_AVG (src/postprocess.f90:34-416, BOOL_DUMP_MOMENTS = .FALSE.): one more sample of the per (cell, species)
 * number, mean velocity, translational temperatures per axis, rotational and vibrational temperature into the running
 * averages AVG_*, which stay on the device.  GRID_RESET (:1070-1094) clears them. */
int pgpu_grid_avg(pgpu_ctx* ctx);
int pgpu_grid_reset(pgpu_ctx* ctx);
/* the running averages, arrays of n >= NCELLS*N_SPECIES doubles indexed JC + NCELLS*(JS-1) (0-based JC, JS here);
 * null pointers are skipped; avg_cumulated = AVG_CUMULATED.  Summed over the ranks (a cell is sampled by its owner). */
int pgpu_get_grid_avg(pgpu_ctx* ctx, int64_t n, double* avg_np, double* avg_vx, double* avg_vy, double* avg_vz,
                      double* avg_ttrx, double* avg_ttry, double* avg_ttrz, double* avg_ttr, double* avg_trot,
                      double* avg_tvib, int32_t* avg_cumulated);
/* CHECKS, particle loop + reduction over ranks (src/postprocess.f90:1752-1796): TOT_NUM(NS), TOT_MOMENTUM(3,NS),
 * TOT_KE_PART(NS), TOT_IE.  The electrostatic terms stay with the field solver on the host. */
int pgpu_checks(pgpu_ctx* ctx, int64_t* tot_num, double* tot_momentum, double* tot_ke_part, double* tot_ie);

/* ---- counters / diagnostics ------------------------------------------------------------ */
/* BOUNDARY_COLL_COUNT(4*NS) and WALL_COLL_COUNT(NW*NS) of the last ADVECT (timecycle.f90:2138-2144) */
int pgpu_get_boundary_counts(pgpu_ctx* ctx, int64_t* boundary_coll_count, int64_t nb,
                             int64_t* wall_coll_count, int64_t nw);
/* TIMESTEP_NIN/NOUT/PXIN.. tallies: 10 arrays x (nboundcells*ns), order
 * NIN,NOUT,PXIN,PYIN,PZIN,PXOUT,PYOUT,PZOUT,EIN,EOUT (postprocess.f90:1098-1147) */
int pgpu_get_boundary_tallies(pgpu_ctx* ctx, double* tallies, int64_t n, int32_t reset);
/* per-cell particle counts NPCALL(1:NCELLS) of the last sort (collisions.f90:63-68) */
int pgpu_get_cell_counts(pgpu_ctx* ctx, int32_t* npc, int64_t nc);
/* device-side time [ms] of the kernels of the last hot-path call, by group (the sections of the reference's TIMER_START /
 * TIMER_STOP, src/tools.f90:1342-1359, src/timecycle.f90:193-300): 0 move, 1 sort, 2 collide, 3 deposit, 4 exchange, 5 total;
 * and launch count */
int pgpu_get_timings(pgpu_ctx* ctx, double* ms, int32_t n, int64_t* launches);

/* ---- multi-GPU plumbing (replaces MPI_COMM_WORLD for the path: MPI_INIT / MPI_COMM_SIZE / MPI_COMM_RANK of
 * src/pantera.f90:35-37 stay with the host, which hands rank and size over in pgpu_config and here) ---- */
/* rank 0 obtains a 128-byte NCCL unique id and the host broadcasts it (MPI_BCAST / torch) */
int pgpu_comm_unique_id(char id128[128]);
int pgpu_comm_init(pgpu_ctx* ctx, int nranks, int rank, const char id128[128]);

#ifdef __cplusplus
}
#endif
#endif /* PANTERA_GPU_H */
