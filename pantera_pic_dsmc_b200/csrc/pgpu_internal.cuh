/*
 * pgpu_internal.cuh -- shared declarations of libpantera_gpu.so (CUDA, sm_100a).
 *
 * Layout of the device particle store (DESIGN.md "Data layout in HBM"):
 *   structure of arrays of 32-byte VECTOR records, one HBM sector each, moved with single 256-bit
 *   loads/stores (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a):
 *     pos[i] = { X, Y, Z, (IC, S_ID) }      32 B   read+written by ADVECT, read by DEPOSIT
 *     vel[i] = { VX, VY, VZ, (S_ID, -) }    32 B   read by ADVECT, read+written by the collisions
 *     ie[i]  = { EROT, EVIB }               16 B   touched only while some particle can hold internal energy
 *   = 72 B/particle of physics state (the 8 pad bytes of vel replace nothing: S_ID is mirrored there so that
 *   the collision kernels never touch pos).  Why records and not nine scalar planes: the per-step counting
 *   sort moves every particle to a data-dependent slot, and a random 8-byte access costs a whole 32-byte
 *   sector in L2 and HBM; with sector-sized records the reorder moves exactly the bytes it needs.
 *   Two copies (ping-pong) so that the sort is a single pass.
 *   Optional planes: DTRIM (only while some particle has DTRIM != DT, i.e. right after a host-side
 *   injection, src/timecycle.f90:463) and ID (cfg.track_id).
 *   A removed particle is a tombstone IC = -1 until the next sort compacts it away.
 *
 * All "ref:" comments cite files of the reference tree (vonkarmaninstitute/pantera-pic-dsmc).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pantera_gpu.h"

/* physical constants, ref: src/global.f90:31-37 */
#define PG_PI 3.1415926535897932
#define PG_EPS0 8.8541878128e-12
#define PG_KB 1.380649e-23
#define PG_QE 1.602176634e-19

#define PG_TOMBSTONE (-1)
/* collision tiles (pgpu_collide.cu): a tile owns the cells whose first sorted slot lies in a window of
 * PG_TILE_S slots; cells with more than PG_TILE_BIG particles take the global dependency-round path */
#ifndef PG_TILE_S
#define PG_TILE_S 512
#endif
#define PG_TILE_BIG 256
#define PG_NBUCKET (PG_TILE_BIG + 1) /* cells grouped by particle count 0..PG_TILE_BIG (oversized and <2 -> bucket 0) */
#define PG_TILE_CAP (PG_TILE_S + PG_TILE_BIG)  /* >= PG_TILE_S - 1 + PG_TILE_BIG */
#define PG_TILE_MAXC (PG_TILE_CAP / 2)         /* cells of >= 2 particles that fit PG_TILE_CAP slots */
#ifndef PG_TILE_THREADS
#define PG_TILE_THREADS 256
#endif
#ifndef PG_CELL_MINBLOCKS
#define PG_CELL_MINBLOCKS 8
#endif
/* the instantiation with internal energy, reactions and product creation needs more registers: at 8 blocks/SM (64
 * registers) it spills 500 bytes, at 6 (80 registers) almost nothing */
#ifndef PG_CELL_MINBLOCKS_IE
#define PG_CELL_MINBLOCKS_IE 6
#endif
#ifndef PG_CELL_TRIES
#define PG_CELL_TRIES 2 /* acceptance tests a lane runs before the warp re-converges for the deflections */
#endif
#ifndef PG_ADVECT_MINBLOCKS
#define PG_ADVECT_MINBLOCKS 7 /* 72 registers; swept 6/7/8 at full size: r02b 9.00/8.70/9.39 ms, r02v (packed brackets) 7.76/7.28/7.79 ms */
#endif
#define PG_COUNT_SMEM_CELLS 4096 /* grids this coarse count their cells per block in shared memory (pgpu_advect.cu) */
#ifndef PG_ADVECT_PREFETCH
#define PG_ADVECT_PREFETCH 0 /* next-cell filter prefetch: measured r02b, +3 % move time (more L1 traffic than latency saved) */
#endif
/* particles one warp of the LEAN advect pass works through (lanes refill as their particle finishes) */
#ifndef PG_ADVECT_CHUNK_MAX
#define PG_ADVECT_CHUNK_MAX 1024
#endif
#define PG_ADVECT_CHUNK_MIN 64
/* hop codes of the packed cell records: > 0 neighbour cell, 0 boundary without a physical group (remove),
 * -pg boundary with physical group pg; -pg - PG_CODE_SPEC when that group is a plain specular wall (no wall
 * reactions, no dielectric charging), which the LEAN advect pass reflects in line */
#define PG_CODE_SPEC 0x40000000
/* a neighbour code (> 0) carries the cell id in its low 29 bits and this flag when the neighbour belongs to another
 * rank: ownership is then tracked along the face walk for free and only emigrants look up CELL_PROCS */
#define PG_CODE_FOREIGN 0x20000000
#define PG_CODE_CELL_MASK 0x1FFFFFFF
#define PG_MAX_HOPS 1000000 /* guard against a malformed mesh hanging the GPU */

/* status block, device resident, read back at the end of every hot-path call */
struct PgStatus {
    unsigned long long n_removed;   /* particles tombstoned by the current call          */
    unsigned long long n_appended;  /* particles appended past np by the current call    */
    unsigned long long n_hits;      /* entries in the deferred boundary-hit list         */
    unsigned long long ncoll;       /* TIMESTEP_COLL                                     */
    unsigned long long nreac;       /* TIMESTEP_REAC                                     */
    unsigned long long pending;     /* collision candidates still waiting                */
    unsigned long long np_cur;      /* slots in use [0, np_cur): set to the host's count when a call starts, then kept
                                       on the device (immigrants appended by EXCHANGE, the compaction of the sort), so
                                       that a whole timestep needs ONE host synchronisation, at its end            */
    unsigned long long n_immigrants; /* particles received by the last EXCHANGE                                     */
    unsigned long long n_cand;      /* MCC: particles that passed the null-collision test (their slots are in hitlist) */
    unsigned long long x_overflow;  /* emigrants that did not fit the fixed-capacity message of their destination  */
    int err;                        /* first device-side error code (0 = none)           */
    int flag_dtrim;                 /* some uploaded/received particle has DTRIM != DT   */
    int flag_ie;                    /* some uploaded/received particle has EROT/EVIB != 0 */
    unsigned int n_big;             /* cells with more particles than a collision tile holds */
    int need_ie, need_dtrim;        /* a kernel had internal energy / a DTRIM != DT to store but the plane does not exist */
    unsigned int scratch[4];        /* host landing zone of small read-backs (pinned copy only) */
    unsigned long long x_sent[64];  /* EXCHANGE fast path: emigrants per destination ...                            */
    unsigned long long x_recv[64];  /* ... and immigrants per source (they size the next step's messages)          */
};
enum { PG_DEVERR_NONE = 0, PG_DEVERR_HOPS = 1, PG_DEVERR_CELL_RANGE = 2, PG_DEVERR_CAPACITY = 3, PG_DEVERR_XCAP = 4, PG_DEVERR_XWAIT = 5 };

struct __align__(32) PgPos { double x, y, z; int ic, sid; };
struct __align__(32) PgVel { double vx, vy, vz; int sid, pad; };
struct __align__(16) PgIe { double erot, evib; };
/* Packed cell records (pgpu_api.cu build_packed_mesh): per face {outward normal(3), point(3)} = nf*6 doubles, then
 * 16 bytes of hop codes, cut into 16-byte chunks (tetrahedron 13, triangle 10, segment 7).  Cells are stored in
 * blocks of PG_GEO_B, chunk-major inside the block: chunk k of the block's cells is contiguous (PG_GEO_B*16 bytes),
 * so that lanes working in nearby cells share their L1 wavefronts (see pgpu_advect.cu unstructured_iter).  A larger
 * block shares more wavefronts but spreads a record over more 128-byte lines (L1 capacity); PG_GEO_B = 1 is the
 * plain cell-major record. */
#ifndef PG_GEO_B
#define PG_GEO_B 8
#endif
#define PG_GEO_LINE (16 * PG_GEO_B) /* bytes between consecutive chunks of a cell */
#define PG_GEO_CHUNKS(nf) ((nf) * 3 + 1)
#define PG_GEO_BLOCK_DOUBLES(nf) (PG_GEO_CHUNKS(nf) * 2 * PG_GEO_B)
#ifdef __CUDACC__
/* address of chunk 0 of cell cell0 (0-based) */
__device__ __forceinline__ const char* pg_geo_base(const double* geo, int cell0, int nch) {
    return reinterpret_cast<const char*>(geo) + (size_t)(cell0 / PG_GEO_B) * (size_t)(nch * PG_GEO_LINE) +
           (size_t)((cell0 % PG_GEO_B) * 16);
}
/* a 32-byte particle record {3 doubles, 2 ints} as ONE 256-bit store (STG.E.ENL2.256) */
__device__ __forceinline__ void pg_store_sector(void* p, double a, double b, double c, int i0, int i1) {
    const unsigned long long w = ((unsigned long long)(unsigned int)i1 << 32) | (unsigned long long)(unsigned int)i0;
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(__double_as_longlong(a)),
                 "l"(__double_as_longlong(b)), "l"(__double_as_longlong(c)), "l"(w)
                 : "memory");
}
/* 32-byte records with a cache-streaming hint (LDG/STG.E.EF.ENL2.256): evict-first in L2 */
struct __align__(32) PgRec32 { double a, b, c; unsigned long long w; };
__device__ __forceinline__ PgRec32 pg_ld_stream(const void* p) {
    PgRec32 v;
    asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=l"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void pg_st_stream(void* p, const PgRec32& v) {
    asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "l"(v.w) : "memory");
}

/* address of chunk 0 of cell cell0's filter record (nf + 1 chunks) */
__device__ __forceinline__ const char* pg_filt_base(const float* filt, int cell0, int nch) {
    return reinterpret_cast<const char*>(filt) + (size_t)(cell0 / PG_GEO_B) * (size_t)(nch * PG_GEO_LINE) +
           (size_t)((cell0 % PG_GEO_B) * 16);
}
/* 32 read-only bytes as ONE load (LDG.E.256.CONSTANT) */
struct __align__(32) PgF64x4 { double a, b, c, d; };
__device__ __forceinline__ PgF64x4 pg_ldg256(const void* p) {
    PgF64x4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
}
/* double e of the record (e = 6*face + {0,1,2: normal, 3,4,5: point}) */
__device__ __forceinline__ double pg_geo_at(const char* gb, int e) {
    return __ldg(reinterpret_cast<const double*>(gb + (e >> 1) * PG_GEO_LINE + (e & 1) * 8));
}
#endif

struct PgPlanes {
    PgPos* pos;
    PgVel* vel;
    PgIe* ie;
    double* dtrim; /* may be null */
    int64_t* id;   /* may be null */
};

/* one reaction, ref: REACTIONS_DATA_STRUCTURE src/global.f90:542-558 */
struct PgReaction {
    int type, r1, r2, p1, p2, p3, p4, nprod, is_cex;
    int tab_off, tab_n;
    int pad;
    double ea, c1, c2, c3;
};

/* everything a kernel needs, passed by value as a __grid_constant__ argument */
struct PgDev {
    /* --- scalars of pgpu_config --- */
    int dims, axi, grid_type, nx, ny;
    double xmin, xmax, ymin, ymax, zmin, zmax, dt, fnum;
    int pic, coll_type;
    int bper[4], bspec[4], bdiff[4], breact[4];
    int zper, yper;
    double boundtemp, extb[3], eps_scaling, mcc_dens, mcc_ttra;
    unsigned long long seed_local; /* RNG_SEED_GLOBAL*(PROC_ID+1), ref: src/initialization.f90:2294-2296 */
    int rank, nranks;
    int tally_on; /* MOD(tID-DUMP_BOUND_START, DUMP_BOUND_AVG_EVERY)==0, ref: src/timecycle.f90:1375 */
    int tID;
    /* --- species --- */
    int ns;
    const double *mass, *charge, *rotrel, *vibrel;
    const int *rotdof, *vibdof;
    /* --- VSS pair tables (IS,JS) column-major --- */
    const double *sigmas, *omegas, *alphas, *grefs;
    double sigmamax;
    /* --- reactions --- */
    int nr;
    const PgReaction* reac;
    const double *tab_e, *tab_cs;
    /* --- wall reactions --- */
    int nwr;
    const int *wr_r, *wr_p1, *wr_nprod;
    const double* wr_prob;
    /* --- MCC background --- */
    int mcc_ncomp;
    const int* mcc_sp;
    const double* mcc_frac;
    const double* mcc_cell_nrho; /* [nc][ns] MCC_BG_CELL_NRHO(SP_ID, IC) or null */
    /* --- unstructured mesh, device layout --- */
    int nn, nc, nf;
    const double* face_geo;  /* [ceil(nc/8)] blocks of PG_GEO_BLOCK_DOUBLES(nf): per face outward normal (3) + a point on the face (3), then the
                                nf hop codes of face_code as int32 (16 bytes) */
    /* non-axisymmetric face walk (pgpu_advect.cu unstructured_iter): a single-precision FILTER record per cell --
     * nf x {normal (3 floats), plane offset P.n (1 float)} + the 16 bytes of hop codes, stored in blocks of PG_GEO_B
     * cells chunk-major like face_geo -- that brackets every face test rigorously, and the EXACT operands
     * {outward normal (3 doubles), face node (3 doubles)} cell-major, read only for the faces the bracket cannot
     * decide (on average one per crossing, none on the final flight) */
    const float* face_filt;   /* [ceil(nc/8)] blocks of (nf+1)*PG_GEO_B*4 floats */
    const double* face_exact; /* [nc][nf][8]: {nx, ny, nz, Px | Py, Pz, -, -}, one 64-byte slot per face */
    const int* face_code;    /* [nc][nf]: >0 neighbour cell to hop into; 0 boundary without a physical
                                group (remove); <0 boundary with physical group -code        */
    const double *face_tang1, *face_tang2; /* [nc][4][3] (3-D) */
    const double* edge_len;                /* [nc][3] (2-D) */
    const int* face_bidx;                  /* [nc][nf] boundary element (1-based) or -1 */
    const double* bound_area;
    int nboundcells;
    const int* cell_nodes;    /* [nc][nf] 1-based */
    const double* basis;      /* [nc][nf][nf] BASIS_COEFFS(j,i,ic) */
    const double* cell_vol;   /* [nc] */
    const double* cell_fnum;  /* [ncells] CELL_FNUM (BOOL_RADIAL_WEIGHTING, rectilinear axisymmetric grids) or null */
    const int* cell_procs;    /* [ncells] or null */
    /* --- boundary conditions GRID_BC(:) --- */
    int nbc;
    const int *bc_particle, *bc_field, *bc_react;
    const double *bc_temp, *bc_accn, *bc_acct;
    /* --- interior walls of structured grids --- */
    int nw;
    const double *w_x1, *w_y1, *w_x2, *w_y2, *w_nx, *w_ny, *w_temp, *w_transm;
    const int *w_spec, *w_diff, *w_porous, *w_react;
    /* --- non-uniform rectilinear grid --- */
    int nxp, nyp;
    const double *xcoord, *ycoord;
    /* --- fields --- */
    const double* e_field;
    const double* b_field;
    double* surface_charge;
    /* --- counters / tallies --- */
    unsigned long long* bcount; /* 4*ns */
    unsigned long long* wcount; /* nw*ns */
    double* tallies;            /* 10 * ns * nboundcells */
    PgStatus* status;
    /* --- particle store --- */
    PgPlanes p;
    long long np;       /* host's view of the slots in use (an upper bound inside a call: kernels read status->np_cur) */
    long long capacity; /* slots allocated */
    int ncells;         /* nc (unstructured) or nx*ny */
    unsigned int* cell_count; /* [ncells] histogram of IC filled by the ADVECT epilogue */
    unsigned int* cell_off;   /* [ncells+1] exclusive prefix of cell_count (valid while sorted) */
    unsigned int* hitlist;    /* slots deferred to the boundary-interaction pass */
    /* deferred velocity reorder (pg_sort_impl defer_vel): the sort has moved pos (and ie, dtrim, id) but not vel;
     * sorted slot o takes its velocity record from vel_src[src_of[o]].  The collision kernels do that gather while
     * they sweep a cell for its velocity bounding box, and write the sorted vel plane once.  Null when not deferred. */
    const PgVel* vel_src;
    const unsigned int* src_of;
    long long id_base;        /* ID counter for particles created on the device */
    int use_ie;               /* 1: the ie plane is live (some particle can hold internal energy) */
    const unsigned int* tile_first; /* [ntiles+1] first cell of every collision tile (valid while sorted) */
    const unsigned int* cell_order; /* [order_n] the cells grouped by particle count, fullest first (valid while sorted) */
    int order_n, cells_lo;          /* cell_order / cell_off describe the cells [cells_lo, cells_lo + order_n) (pgpu_ctx::own_lo) */
    int ntiles;
    /* emigrant bookkeeping in the ADVECT epilogue (nranks > 1): a particle whose final cell belongs to another
     * rank is not histogrammed; its slot goes to emig_list and its destination is counted in xrow, so that
     * EXCHANGE touches the emigrants only (pgpu_exchange.cu).  Null when not tracking. */
    int emig_trust;                 /* 1: ownership flags of the hop codes are authoritative (owned_clean at step start) */
    unsigned int* emig_list;        /* [emig_cap] */
    long long emig_cap;
    unsigned long long* xrow;       /* [PG_XROW]: [0..63] emigrants per destination, [PG_XROW_LEN] list length,
                                       [PG_XROW_OVF] entries that did not fit the list */
};
#define PG_XROW 72
#define PG_XROW_LEN 64
#define PG_XROW_OVF 65

/* ------------------------------------------------------------------------------------------
 * host context
 * ---------------------------------------------------------------------------------------- */
template <class T>
struct DevBuf {
    T* d = nullptr;
    size_t n = 0;
};

struct pgpu_ctx {
    char err[512];
    int device;
    cudaStream_t stream;
    pgpu_config cfg;
    bool cfg_set;
    PgDev dev; /* host mirror of the kernel argument */

    /* raw host copies needed to (re)build the packed mesh when BCs change */
    int mesh_dims, nn, nc;
    int advect_chunk_override = 0; /* PANTERA_GPU_ADVECT_CHUNK (tuning / test knob; results do not depend on it) */
    int cell_procs_max = 0; /* largest CELL_PROCS entry (validated against nranks before an exchange) */
    int64_t mcc_cell_nrho_cells = 0; /* cells covered by dev.mcc_cell_nrho */
    std::vector<double> h_node_coords, h_face_normal;
    std::vector<int32_t> h_cell_nodes, h_cell_neighbors, h_cell_faces_pg, h_cell_pg, h_cell_procs;
    std::vector<int32_t> h_bc_volume, h_bc_particle, h_bc_field, h_bc_react;
    bool mesh_set, mesh_dirty;

    /* owned device allocations (freed in destroy) */
    std::vector<void*> owned;
    /* named table slots so that a setter can replace its arrays */
    void* slot[64];

    /* particle store */
    PgPlanes planes[2];
    int cur;
    long long np, capacity;
    long long np_ub;      /* upper bound of the device-side count inside a call (sizes the launch grids) */
    bool np_dev_valid;    /* the device-side count differs from np: pg_finish adopts it */
    bool owned_clean;     /* every particle sits in a cell this rank owns (true after an EXCHANGE) */
    bool has_dtrim, sorted, hist_valid, maybe_tombstones;
    bool any_internal_dof, ie_seen; /* species with ROTDOF/VIBDOF != 0; nonzero EROT/EVIB seen in an upload */
    unsigned int* d_tile_first;
    unsigned int* d_cell_order;  /* [ncells] */
    unsigned int* d_cell_bucket; /* [2][PG_NBUCKET] bucket sizes, then bucket cursors */
    long long tile_alloc;
    unsigned int n_big;             /* cells beyond PG_TILE_BIG particles in the current sorted store */
    unsigned int* d_big_list;       /* [n_big] their cell ids */
    unsigned long long* d_big_bbox; /* [n_big][6] order-encoded velocity bounds */
    long long big_alloc;
    long long id_counter;
    unsigned int *d_cell_count, *d_cell_off, *d_hitlist;
    unsigned int* d_src_of; /* [capacity] source slot of every sorted slot while the vel reorder is deferred */
    bool vel_deferred;
    bool defer_vel = false; /* PANTERA_GPU_DEFER_VEL (experiment knob, see pgpu_create) */
    long long hit_capacity;
    int ncells_alloc;
    PgStatus* d_status;
    PgStatus* h_status; /* pinned */
    void* d_scan_tmp;
    size_t scan_tmp_bytes;

    /* collision scratch */
    void* d_cand;      /* uint4 per candidate */
    long long cand_capacity;
    unsigned int* d_owner;
    long long owner_capacity;
    int* d_cell_ncoll;       /* [ncells] */
    unsigned int* d_cand_off; /* [ncells+1] */
    double* d_cell_vrmax;    /* [ncells] */
    double* d_cell_fcorr;    /* [ncells] */
    int cellscratch_alloc;

    /* staging for AoS <-> SoA */
    void* d_stage[2];
    size_t stage_bytes;
    void* h_stage[2]; /* pinned bounce buffers for pageable host memory */

    /* deposit */
    double* d_rhs;
    long long rhs_n;

    /* counters */
    long long ns, nw;
    unsigned long long *d_bcount, *d_wcount;
    double* d_tallies;
    long long tallies_n;
    double* d_surface_charge;
    long long nsc;

    /* deterministic mode: radix-sort scratch (keys in/out, slots in/out) */
    bool deterministic = false;
    unsigned int* d_det[4] = {nullptr, nullptr, nullptr, nullptr};
    long long det_cap = 0;
    void* d_det_tmp = nullptr;
    size_t det_tmp_bytes = 0;

    /* restart (pgpu_io.cu, pgpu_restart.cu): a dump was read with its cell indices unchecked (another mesh); every hot-path
     * entry point refuses to run until pgpu_reassign_particles_to_cells has located the particles again */
    bool cells_unassigned = false, reassigning = false;

    /* injection (pgpu_inject.cu), field step (pgpu_field.cu) */
    void* inject = nullptr;
    void* field = nullptr;
    double* d_node_coords = nullptr;     /* [nn][3] */
    unsigned char* d_cell_solid = nullptr; /* [nc] cells of a SOLID volume group */
    double volume_total = 0;             /* sum of CELL_VOLUMES */

    /* GRID_AVG running averages (pgpu_sample.cu) */
    double* d_gavg = nullptr;
    unsigned int* d_gavg_np = nullptr;
    double* d_gavg_sums = nullptr;
    int* d_gavg_cum = nullptr;
    long long gavg_bins = 0;
    int avg_cumulated = 0;

    /* timings */
    /* Cells this rank can hold particles in.  Every rank carries the whole mesh (hop codes and geometry of all cells:
     * a particle may finish its step anywhere), but once an EXCHANGE has shipped the strays (owned_clean) only the
     * cells [own_lo, own_hi) spanned by this rank's own cells are populated: the histogram, its scan, the cursor, the
     * cell lists and the collision launch then cover that range instead of the global mesh (at N ranks in strips, 1/N
     * of it).  hist_lo/hi: the range a valid histogram was counted over; cells_lo/hi: the range cell_off and
     * cell_order describe. */
    long long own_lo = 0, own_hi = 0, hist_lo = 0, hist_hi = 0, cells_lo = 0, cells_hi = 0;
    std::vector<char> x_adjacent; /* [nranks] shares a face with this rank's cells (build_packed_mesh) */
    bool x_adjacent_known = false;
    int x_allpairs = -1;          /* PANTERA_GPU_XCHG_ALLPAIRS=1: a message for every rank, adjacent or not */
    cudaEvent_t ev[16];
    /* PANTERA_GPU_XCHG_TRACE=1: pack / transfer / unpack split of the per-step EXCHANGE, printed when the context closes */
    cudaEvent_t xev[4] = {nullptr, nullptr, nullptr, nullptr};
    double x_ms[3] = {0., 0., 0.};
    long long x_traced = 0;
    int x_trace = -1;
    double ms[8];
    long long launches;

    /* NCCL (dlopen'ed lazily) */
    void* nccl_lib;
    void* nccl_comm;
    int comm_nranks, comm_rank;
    void* d_sendbuf;
    void* d_recvbuf;
    long long xbuf_capacity; /* particles */
    /* EXCHANGE over peer memory (pgpu_exchange.cu): this rank's window (flags + two parities of one message slot per
     * source rank), the peers' windows mapped into this process (CUDA IPC; null: that pair goes through NCCL) */
    void* x_win = nullptr;
    long long x_win_recs = 0;
    void* x_peer_win[64] = {};
    unsigned long long x_seq = 0;
    unsigned long long* d_xcount; /* 4 rows of PG_XROW (counts, offsets, cursors, spare) + [nranks] gathered rows */
    unsigned int* d_emig_list;    /* [emig_cap] slots of the emigrants found by the last ADVECT */
    long long emig_cap;
    bool emig_valid;              /* d_emig_list / count row describe the current store */
    /* EXCHANGE without host round trips: fixed-capacity messages sized from the previous step's counts */
    bool x_counts_valid, x_plan_pending = false;
    long long x_prev_send[64], x_prev_recv[64];
    double x_headroom; /* message capacity = headroom * previous count + 8192 records (PANTERA_GPU_XCHG_HEADROOM) */
};

#define PG_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

#define PG_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            snprintf((ctx)->err, sizeof((ctx)->err), "CUDA error %s at %s:%d (%s)",                 \
                     cudaGetErrorString(e_), __FILE__, __LINE__, #call);                            \
            return PGPU_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

/* internal entry points shared across translation units */
int pg_prepare(pgpu_ctx* c);                 /* validate state, rebuild packed mesh, refresh c->dev */
int pg_finish(pgpu_ctx* c, const char* what); /* sync stream, read back status, map device errors */
int pg_ensure_capacity(pgpu_ctx* c, long long n);
int pg_ensure_cell_arrays(pgpu_ctx* c);
int pg_reset_status(pgpu_ctx* c);
int pg_set_np_cur(pgpu_ctx* c);              /* status->np_cur = c->np (after the host changed the count inside a call) */
long long pg_exchange_recv_slots(const pgpu_ctx* c);
int pg_sort_impl(pgpu_ctx* c, bool defer_vel = false, bool sync_np = true); /* histogram (if needed) + scan + scatter */
int pg_resolve_vel(pgpu_ctx* c);             /* finishes a deferred vel reorder (no-op otherwise); async */
int pg_advect_impl(pgpu_ctx* c, int tID);    /* async */
int pg_collide_impl(pgpu_ctx* c, int tID, bool big_now = true); /* CELL path (async) [+ GLOBAL path, which syncs] */
int pg_collide_big(pgpu_ctx* c);             /* GLOBAL path for the oversized cells + end of a deferred vel reorder */
int pg_exchange_impl(pgpu_ctx* c);
void pg_refresh_dev(pgpu_ctx* c);
int pg_fill_dtrim(pgpu_ctx* c);               /* DTRIM plane (allocated on first use) = DT for the particles in the store */
void pg_cell_range(pgpu_ctx* c, long long* lo, long long* hi); /* the cells the per-cell passes have to cover now */
int pg_need_ie(pgpu_ctx* c);                  /* allocate the internal-energy planes (zero) if they do not exist yet */
int pg_need_dtrim(pgpu_ctx* c);               /* allocate the DTRIM planes if they do not exist yet */
void pg_inject_destroy(pgpu_ctx* c);
void pg_field_destroy(pgpu_ctx* c);
void pg_field_invalidate(pgpu_ctx* c); /* mesh or boundary types changed: chains, matrices and warm starts are stale */
int pg_adopt_efield(pgpu_ctx* c, double* d_e); /* a device array becomes E_FIELD (owned by the context from then on) */

static inline int pg_ncells(const pgpu_ctx* c) {
    return (c->cfg.grid_type == PGPU_UNSTRUCTURED) ? c->nc : c->cfg.nx * c->cfg.ny;
}

/* ------------------------------------------------------------------------------------------
 * device helpers
 * ---------------------------------------------------------------------------------------- */
#ifdef __CUDACC__

/* Philox4x32-10 (Salmon et al., SC'11) */
__device__ __forceinline__ void pg_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                          uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/*
 * Counter-based stream, one per (kind, timestep, entity): replaces the reference's single
 * sequential MT19937-64 stream rf() (src/tools.f90:33-43).  Contract (also implemented by the
 * oracle's rng_mode 1):  key = { seed_lo ^ kind*0x632BE5AB, seed_hi ^ tID },
 * counter = { word/4, e0_lo, e0_hi, e1 }.  The stream is the blocks' 32-bit words in order:
 * rf() takes two words (lo, hi -> a 53-bit uniform on [0,1)), pick(n) takes one for n <= 65536.
 * kinds: 1 ADVECT (e0 = particle slot), 2 DSMC (e0 = cell 0-based, e1 = candidate), 3 MCC_VAHEDI (e0 = slot),
 * 4 MCC (e0 = slot).
 */
struct PgRng {
    uint32_t k0, k1, e0, e1, e2;
    uint32_t ndraw;          /* words taken from the stream so far */
    uint32_t w0, w1, w2, w3; /* the words of the current block that have not been taken yet, w0 first */
    __device__ __forceinline__ void init(unsigned long long seed, int kind, int tID, unsigned long long ent0,
                                         unsigned long long ent1) {
        k0 = (uint32_t)seed ^ ((uint32_t)kind * 0x632BE5ABu);
        k1 = (uint32_t)(seed >> 32) ^ (uint32_t)tID;
        e0 = (uint32_t)ent0;
        e1 = (uint32_t)(ent0 >> 32);
        e2 = (uint32_t)ent1;
        ndraw = 0;
        w0 = w1 = w2 = w3 = 0;
    }
    __device__ __forceinline__ uint32_t word() {
        if ((ndraw & 3u) == 0) {
            uint32_t o[4];
            pg_philox(ndraw >> 2, e0, e1, e2, k0, k1, o);
            w0 = o[0]; w1 = o[1]; w2 = o[2]; w3 = o[3];
        }
        ndraw++;
        const uint32_t r = w0;
        w0 = w1; w1 = w2; w2 = w3;
        return r;
    }
    __device__ __forceinline__ double rf() {
        const uint32_t lo = word();
        const uint32_t hi = word();
        unsigned long long u = ((unsigned long long)hi << 32) | lo;
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    /* INT(N*rf()) of the pair selection (ref: src/collisions.f90:263-271) clamped to N-1: one word of the stream
     * for N <= 65536 (floor(N*word/2^32), bias <= N/2^32), so that the two indices and the acceptance draw of a
     * candidate are exactly one Philox block; rf() for larger N */
    __device__ __forceinline__ int pick(int n) {
        int k;
        if (n <= 65536) k = (int)__umulhi(word(), (uint32_t)n);
        else k = (int)(n * rf());
        return (k >= n) ? n - 1 : k;
    }
    /* resume a stream after n words (the current block is regenerated when n is not a multiple of 4) */
    __device__ __forceinline__ void seek(uint32_t n) {
        ndraw = n;
        if (n & 3u) {
            uint32_t o[4];
            pg_philox(n >> 2, e0, e1, e2, k0, k1, o);
            const uint32_t s = n & 3u;
            w0 = (s == 1) ? o[1] : (s == 2) ? o[2] : o[3];
            w1 = (s == 1) ? o[2] : o[3];
            w2 = o[3];
        }
    }
};

/* ref: src/tools.f90:130-187 MAXWELL with zero drift and one temperature on all axes */
__device__ __forceinline__ void pg_maxwell0(PgRng& g, double t, double m, double* v3) {
    const double PI2 = 2. * PG_PI;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (t == 0.0) {
            v3[i] = 0.;
        } else {
            double r = g.rf();
            double teta = PI2 * r;
            double beta = 1. / sqrt(2. * PG_KB / m * t);
            double r1 = g.rf();
            while (r1 < 1.0e-13) r1 = g.rf();
            double ro = sqrt(-log(r1)) / beta;
            v3[i] = 0. + ro * sin(teta);
        }
    }
}

/* ref: src/tools.f90:245-285 INTERNAL_ENERGY */
__device__ __forceinline__ double pg_internal_energy(PgRng& g, int dof, double temp) {
    if (dof == 0) return 0.;
    if (dof == 2) {
        double r = g.rf();
        while (r < 1.0e-13) r = g.rf();
        return -log(r) * PG_KB * temp;
    }
    double param = (dof == 1) ? (double)5.0f : (double)0.01f;
    double r = g.rf();
    double x = -log(r);
    r = g.rf();
    while (r > pow(x / param, dof / 2. - 1.)) {
        r = g.rf();
        x = -log(r);
        r = g.rf();
    }
    return x * PG_KB * temp;
}

/* ref: src/tools.f90:1150-1186 FLX */
__device__ __forceinline__ double pg_flx(PgRng& g, double sn, double tinf, double m) {
    double beta = 1. / sqrt(2. * PG_KB / m * tinf);
    double acca = sqrt(sn * sn + 2.);
    double kappa = 2. / (sn + acca) * exp(0.5 + 0.5 * sn * (sn - acca));
    for (;;) {
        double r1 = g.rf();
        double y = -3. + 6. * r1;
        double r2 = g.rf();
        double fm = kappa * (y + sn) * exp(-(y * y));
        if (r2 <= fm) return y / beta;
    }
}

/* ref: src/tools.f90:1544-1593 INTERP_CS (1-based bisection on 0-based arrays) */
__device__ __forceinline__ double pg_interp_cs(double e, int n, const double* __restrict__ te,
                                               const double* __restrict__ tcs) {
    int L = 1, R = n, idx;
    if (e < te[L - 1]) return 0;
    if (e > te[R - 1]) return 0;
    if (R == L + 1) return tcs[L - 1] + (tcs[R - 1] - tcs[L - 1]) * (e - te[L - 1]) / (te[R - 1] - te[L - 1]);
    /* DEVIATION: for e == TABLE_EN(R) exactly the reference's bisection never terminates
     * (src/tools.f90:1574-1586); the table end value is returned instead. */
    if (e == te[R - 1]) return tcs[R - 1];
    for (;;) {
        idx = (L + R) / 2;
        if (te[idx - 1] <= e) {
            if (te[idx] > e) break;
            L = idx;
        } else {
            if (te[idx - 2] <= e) {
                idx = idx - 1;
                break;
            }
            R = idx;
        }
    }
    L = idx;
    R = idx + 1;
    return tcs[L - 1] + (tcs[R - 1] - tcs[L - 1]) * (e - te[L - 1]) / (te[R - 1] - te[L - 1]);
}

__device__ __forceinline__ void pg_set_err(const PgDev& d, int code) { atomicCAS(&d.status->err, 0, code); }

/* The internal-energy and DTRIM planes exist only while something can live in them (pg_need_ie / pg_need_dtrim): a
 * missing plane reads as {0, 0} / DT and swallows stores of exactly that; a store of anything else raises the status
 * flag the host turns into the allocation (10^9 argon particles on one B200: 48 GB that are never touched). */
__device__ __forceinline__ PgIe pg_ie_load(const PgIe* ie, long long i) {
    if (ie) return ie[i];
    PgIe z; z.erot = 0.; z.evib = 0.;
    return z;
}
__device__ __forceinline__ void pg_ie_store(PgIe* ie, long long i, double erot, double evib, PgStatus* st) {
    if (ie) { PgIe E; E.erot = erot; E.evib = evib; ie[i] = E; }
    else if (erot != 0. || evib != 0.) st->need_ie = 1;
}

#endif /* __CUDACC__ */
