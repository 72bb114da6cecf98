/*
 * pantera_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * A plain-C, sequential, array-of-structures restatement of PANTERA's per-timestep particle
 * hot path, each routine citing the reference file:line it follows (paths relative to the
 * reference tree, e.g. src/timecycle.f90:1021).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and
 * cannot be compiled in this image (needs mpifort + PETSc; no Fortran compiler exists here).
 * The oracle is therefore pinned only by (1) the MT19937-64 / Philox4x32-10 known-answer
 * vectors, (2) closed-form results (specular reflection, periodic wrap, ballistic crossing
 * times, Boris rotation, charge conservation) and (3) analytic kinetic-theory values
 * (equilibrium VSS collision rate, equipartition) -- see tests/test_oracle_*.py.
 *
 * The API mirrors include/pantera_gpu.h one-to-one (prefix po_ instead of pgpu_) so that the
 * same Python driver can run both sides of a parity test.
 *
 * RNG modes (po_set_rng_mode):
 *   0  MT19937-64 genrand64_real1, one sequential stream per rank seeded with
 *      RNG_SEED_GLOBAL*(PROC_ID+1) -- the reference's stream (src/mt19937.f90, tools.f90:33-43,
 *      initialization.f90:2294-2296).
 *   1  counter-based Philox4x32-10 keyed per (kind, timestep, entity): the stream the CUDA
 *      kernels use; lets device results be compared value-by-value with the oracle.
 *   2  scripted: rf() replays a caller-supplied sequence, so that single events (one CLL reflection, one
 *      VSS/Borgnakke-Larsen pair, one TCE dissociation) can be compared with a plain-Python evaluation of the
 *      Fortran expressions on the same draws (tests/golden/fortran_pins.py).
 */
#ifndef PANTERA_ORACLE_H
#define PANTERA_ORACLE_H
#include <stdint.h>
#include "../include/pantera_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct po_ctx po_ctx;

int po_create(po_ctx** ctx, int device_ignored);
int po_destroy(po_ctx* ctx);
const char* po_last_error(const po_ctx* ctx);
int po_set_rng_mode(po_ctx* ctx, int mode);
/* rng mode 2: rf() replays a caller-supplied sequence (independent pins of tests/golden/fortran_pins.py) */
int po_set_rng_script(po_ctx* ctx, const double* draws, int64_t n);
int64_t po_rng_script_used(const po_ctx* ctx);
int po_collision_type(const po_ctx* ctx);
int po_set_deterministic(po_ctx* ctx, int32_t on);
int po_set_cell_fnum(po_ctx* ctx, int nc, const double* cell_fnum);
/* raw generator access for the known-answer tests */
double po_rf(po_ctx* ctx);
int po_pick(po_ctx* ctx, int n); /* pair-selection index INT(N*rf()), one stream word in Philox mode for N <= 65536 */
void po_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t po_mt_next_u64(po_ctx* ctx);
void po_rng_stream(po_ctx* ctx, int kind, int tID, uint64_t e0, uint64_t e1);
/* samplers exposed for distribution tests */
void po_maxwell(po_ctx* ctx, double ux, double uy, double uz, double tx, double ty, double tz,
                double m, double* v3);
double po_internal_energy(po_ctx* ctx, int dof, double temp);
double po_flx(po_ctx* ctx, double sn, double tinf, double m);
double po_interp_cs(double e, int n, const double* te, const double* tcs);

int po_set_config(po_ctx* ctx, const pgpu_config* cfg);
int po_set_species(po_ctx* ctx, int ns, const double* mass, const double* charge,
                   const int32_t* rotdof, const int32_t* vibdof,
                   const double* rotrel, const double* vibrel);
int po_set_vss(po_ctx* ctx, int ns, const double* sigmas, const double* omegas,
               const double* alphas, const double* grefs, double sigmamax);
int po_set_reactions(po_ctx* ctx, int nr, const int32_t* type,
                     const int32_t* r1, const int32_t* r2,
                     const int32_t* p1, const int32_t* p2, const int32_t* p3, const int32_t* p4,
                     const int32_t* nprod, const int32_t* is_cex,
                     const double* ea, const double* c1, const double* c2, const double* c3,
                     const int32_t* tab_off, const double* tab_e, const double* tab_cs);
int po_set_wall_reactions(po_ctx* ctx, int n, const int32_t* r_sp, const int32_t* p1_sp,
                          const int32_t* nprod, const double* prob);
int po_set_mcc_background(po_ctx* ctx, int ncomp, const int32_t* sp_id, const double* molfrac);
int po_set_mcc_cell_density(po_ctx* ctx, int ns, int64_t nc, const double* nrho);
int po_set_mesh(po_ctx* ctx, int dims, int nn, int nc,
                const double* node_coords, const int32_t* cell_nodes, const int32_t* cell_neighbors,
                const double* face_normal, const double* face_tang1, const double* face_tang2,
                const double* cell_edges_len, const int32_t* cell_faces_pg, const int32_t* cell_pg,
                const double* basis_coeffs, const double* cell_volumes, const int32_t* cell_procs,
                const int32_t* face_boundary_index, int nboundcells, const double* boundary_area);
int po_set_cell_procs(po_ctx* ctx, int nc, const int32_t* cell_procs);
int po_set_bcs(po_ctx* ctx, int nbc, const int32_t* particle_bc, const int32_t* field_bc,
               const int32_t* volume_bc, const double* wall_temp, const double* acc_n,
               const double* acc_t, const int32_t* react, const int32_t* dump_fluxes);
int po_set_walls(po_ctx* ctx, int nw, const double* x1, const double* y1, const double* x2,
                 const double* y2, const double* normx, const double* normy, const double* temp,
                 const int32_t* specular, const int32_t* diffuse, const int32_t* porous,
                 const double* transmissivity, const int32_t* react);
int po_set_grid_coords(po_ctx* ctx, int nxp, const double* xcoord, int nyp, const double* ycoord);

int po_upload_particles(po_ctx* ctx, const void* aos, int64_t n, int32_t stride_bytes);
int po_download_particles(po_ctx* ctx, void* aos, int64_t capacity, int64_t* n, int32_t stride_bytes);
int po_append_particles(po_ctx* ctx, const void* aos, int64_t n, int32_t stride_bytes);
int po_num_particles(po_ctx* ctx, int64_t* n);
int po_upload_planes(po_ctx* ctx, const void* pos, const void* vel, const void* ie_or_null, int64_t n);
int po_download_planes(po_ctx* ctx, void* pos, void* vel, void* ie_or_null, int64_t capacity, int64_t* n);

int po_set_efield(po_ctx* ctx, const double* e_field, int64_t n);
int po_set_bfield(po_ctx* ctx, const double* b_field, int64_t n);
int po_get_surface_charge(po_ctx* ctx, double* sc, int64_t n);
int po_set_surface_charge(po_ctx* ctx, const double* sc, int64_t n);

int po_deposit_charge(po_ctx* ctx, double* rhs_host, int64_t nnodes);
/* same, but accumulated in long double with Neumaier compensation */
int po_deposit_charge_compensated(po_ctx* ctx, double* rhs_host, double* abs_sum, int64_t nnodes);
int po_advect(po_ctx* ctx, int32_t tID);
/* EXCHANGE, local part: removes emigrants from the store and returns them packed by
 * destination rank (sendbuf of 96-B records, sendcount[nranks]) -- grid_and_partition.f90:562-588 */
int po_exchange_pack(po_ctx* ctx, void* sendbuf, int64_t capacity, int64_t* sendcount);
int po_sort(po_ctx* ctx); /* no-op: the reference builds index arrays inside DSMC_COLLISIONS */
int po_collide(po_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);
int po_mcc_vahedi(po_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);
int po_mcc_collisions(po_ctx* ctx, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac);
int po_remove_species(po_ctx* ctx, int32_t sp_id);
int po_set_species_weights(po_ctx* ctx, int ns, const double* spwt);
int po_set_mixtures(po_ctx* ctx, int nmix, const int32_t* comp_off, const int32_t* comp_sp, const double* comp_molfrac);
int po_set_initial_particles_tasks(po_ctx* ctx, int n, const int32_t* mix_id, const double* nrho, const double* ux, const double* uy,
                                   const double* uz, const double* ttrax, const double* ttray, const double* ttraz,
                                   const double* trot, const double* tvib);
int po_set_volume_inject_tasks(po_ctx* ctx, int n, const int32_t* mix_id, const double* nrhodot, const double* ux, const double* uy,
                               const double* uz, const double* ttrax, const double* ttray, const double* ttraz,
                               const double* trot, const double* tvib);
int po_set_emit_tasks(po_ctx* ctx, int n, const int32_t* mix_id, const int32_t* ic, const int32_t* iv1, const int32_t* iv2,
                      const int32_t* iface, const int32_t* nfs_off, const double* nfs, const double* ux, const double* uy,
                      const double* uz, const double* u_norm, const double* ttra, const double* trot, const double* tvib);
int po_initial_seed(po_ctx* ctx);
int po_volume_inject(po_ctx* ctx, int32_t tID);
int po_boundaries_emit(po_ctx* ctx, int32_t tID);
int po_set_field_bcs(po_ctx* ctx, int nbc, const double* wall_potential, const double* wall_rf_potential, const double* rf_frequency,
                     const double* wall_efield, const double* eps_rel);
int po_compute_e_field(po_ctx* ctx, const double* phi, int64_t nnodes, double* e_out);
int po_reassign_particles_to_cells(po_ctx* ctx, int64_t* n_removed);
int po_field_solve_1d(po_ctx* ctx, int32_t tID, const double* rhs, double* phi_out, double* e_out);
int po_field_solve(po_ctx* ctx, int32_t tID, const double* rhs, double* phi_out, double* e_out, double rtol, int32_t max_iter,
                   int32_t* iters_out, double* relres_out);
int po_grid_avg(po_ctx* ctx);
int po_grid_reset(po_ctx* ctx);
int po_get_grid_avg(po_ctx* ctx, int64_t n, double* avg_np, double* avg_vx, double* avg_vy, double* avg_vz, double* avg_ttrx,
                    double* avg_ttry, double* avg_ttrz, double* avg_ttr, double* avg_trot, double* avg_tvib, int32_t* avg_cumulated);
int po_checks(po_ctx* ctx, int64_t* tot_num, double* tot_momentum, double* tot_ke_part, double* tot_ie);
int po_dump_particles_file(po_ctx* ctx, const char* path_prefix, int32_t timestep, int32_t binary);
int po_read_particles_file(po_ctx* ctx, const char* path_prefix, int32_t timestep, int32_t binary);
int po_run_steps(po_ctx* ctx, int32_t tID0, int32_t nsteps, int64_t* total_coll, int64_t* total_reac);

int po_get_boundary_counts(po_ctx* ctx, int64_t* boundary_coll_count, int64_t nb,
                           int64_t* wall_coll_count, int64_t nw);
int po_get_boundary_tallies(po_ctx* ctx, double* tallies, int64_t n, int32_t reset);
int po_get_cell_counts(po_ctx* ctx, int32_t* npc, int64_t nc);
/* per-phase CPU seconds accumulated by po_run_steps: 0 move, 1 sort, 2 collide */
int po_get_timings(po_ctx* ctx, double* sec, int32_t n);

/* po_mpi_sim.c: nsteps of [ADVECT -> EXCHANGE -> DSMC_COLLISIONS] on nranks contexts, one host thread per
 * rank (the reference's MPI model); totals = {ncoll, nreac, particles exchanged} */
int po_sim_run(po_ctx** ranks, int nranks, int tID0, int nsteps, double* seconds, int64_t* totals);

#ifdef __cplusplus
}
#endif
#endif
