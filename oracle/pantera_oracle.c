/*
 * pantera_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).  See pantera_oracle.h.
 *
 * PARITY UNPINNED in the sense of the task brief: the reference ships no tests or golden vectors and cannot be built
 * here (Fortran + MPI + PETSc, no Fortran compiler in the image), so nothing below has been compared with output of the
 * reference itself.  What pins it instead: the RNG known-answer vectors, closed forms and kinetic theory in tests/, and
 * the INDEPENDENT pins -- a second, plain-Python reading of the Fortran, procedure by procedure, on scripted draws
 * (tests/golden/fortran_pins.py, fortran_pins2.py; the table in DESIGN.md section 5 lists the reference lines each covers).
 *
 * Sequential, AoS (96-byte records), double precision, no FMA contraction (build with
 * -ffp-contract=off, mirroring gfortran -O3 on baseline x86-64: SURVEY App. B3).
 * All "ref:" comments cite files of the reference tree.
 */
#include "pantera_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ref: src/global.f90:31-37 */
static const double PI = 3.1415926535897932;
static const double EPS0 = 8.8541878128e-12;
static const double KB = 1.380649e-23;
static const double QE = 1.602176634e-19;

typedef pgpu_particle part_t;

typedef struct {
    int type, r1, r2, p1, p2, p3, p4, nprod, is_cex;
    double ea, c1, c2, c3;
    int tab_n;
    const double *tab_e, *tab_cs;
    int64_t counts;
} reaction_t;

struct po_ctx {
    char err[256];
    pgpu_config c;
    /* species */
    int ns;
    double *mass, *charge, *rotrel, *vibrel;
    int *rotdof, *vibdof;
    /* vss (Fortran (IS,JS), column-major) */
    double *sigmas, *omegas, *alphas, *grefs, sigmamax;
    /* reactions */
    int nr;
    reaction_t* reac;
    int32_t* tab_off;
    double *tab_e, *tab_cs;
    /* wall reactions */
    int nwr;
    int *wr_r, *wr_p1, *wr_nprod;
    double* wr_prob;
    /* mcc background mixture */
    int mcc_ncomp;
    int* mcc_sp;
    double* mcc_frac;
    double* mcc_cell_nrho; /* MCC_BG_CELL_NRHO(N_SPECIES, NCELLS), NULL = BOOL_BG_DENSITY_FILE false */
    int64_t mcc_cell_nrho_n;
    /* mesh */
    int dims_mesh, nn, nc;
    double* node_coords;
    int *cell_nodes, *cell_neighbors;
    double *face_normal, *face_tang1, *face_tang2, *cell_edges_len;
    int *cell_faces_pg, *cell_pg;
    double *basis_coeffs, *cell_volumes;
    int* cell_procs;
    int* face_bidx;
    int nboundcells;
    double* bound_area;
    /* bcs */
    int nbc;
    int *bc_particle, *bc_field, *bc_volume, *bc_react, *bc_dump;
    double *bc_temp, *bc_accn, *bc_acct;
    /* walls */
    int nw;
    double *w_x1, *w_y1, *w_x2, *w_y2, *w_nx, *w_ny, *w_temp, *w_transm;
    int *w_spec, *w_diff, *w_porous, *w_react;
    /* nonuniform grid */
    int nxp, nyp;
    double *xcoord, *ycoord;
    /* fields */
    double* e_field;
    int64_t ne;
    double* b_field;
    int64_t nb;
    double* surface_charge;
    int64_t nsc;
    /* particles */
    part_t* p;
    int64_t np, cap;
    int64_t id_counter;
    /* counters */
    int64_t* bcount; /* 4*ns */
    int64_t* wcount; /* nw*ns */
    double* tallies; /* 10 * nboundcells*ns */
    int32_t* npcall;
    int64_t timestep_coll, timestep_reac;
    double tsec[3];
    /* rng */
    int rng_mode;
    uint64_t mt[312];
    int mti;
    int seeded;
    uint32_t key[2], ent[3], ndraw;
    uint32_t phx_buf[4];
    /* injection / seeding tasks */
    double* spwt;
    int nmix;
    int* mix_off; int* mix_sp; double* mix_frac;
    int n_seed, n_vinj, n_emit;
    double* seed_par; int* seed_mix;   /* [n][9]: rate, UX, UY, UZ, TTRAX, TTRAY, TTRAZ, TROT, TVIB */
    double* vinj_par; int* vinj_mix;
    double* emit_par; int* emit_int;   /* [n][7]: UX, UY, UZ, U_NORM, TTRA, TROT, TVIB; [n][6]: MIX_ID, IC, IV1, IV2, IFACE, nfs offset */
    double* emit_nfs;
    double* cell_fnum; /* CELL_FNUM (BOOL_RADIAL_WEIGHTING) or NULL */
    /* field boundary parameters (GRID_BC%WALL_POTENTIAL, ...) */
    double *fb_pot, *fb_rfpot, *fb_freq, *fb_efield, *fb_eps;
    /* GRID_AVG running averages */
    double* gavg;   /* [9][nbins]: NP, VX, VY, VZ, TTRX, TTRY, TTRZ, TROT, TVIB */
    int *gcum1, *gcum2;
    int64_t gbins;
    int avg_cumulated;
    /* rng_mode 2: rf() returns the next value of a caller-supplied script (tests/golden/fortran_pins.py) */
    double* script;
    int64_t script_n, script_pos;
};

#define FAIL(ctx, code, ...)                                  \
    do {                                                      \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return (code);                                        \
    } while (0)

static void* dupmem(const void* src, size_t bytes) {
    if (!src || bytes == 0) return NULL;
    void* d = malloc(bytes);
    memcpy(d, src, bytes);
    return d;
}
#define REPL(dst, src, count)                                  \
    do {                                                       \
        free(dst);                                             \
        (dst) = dupmem((src), sizeof(*(dst)) * (size_t)(count)); \
    } while (0)

/* ------------------------------------------------------------------------------------------
 * RNG
 * ---------------------------------------------------------------------------------------- */

/* ref: src/mt19937.f90:117-129 init_genrand64 */
static void mt_init(po_ctx* w, uint64_t seed) {
    w->mt[0] = seed;
    for (int i = 1; i < 312; i++)
        w->mt[i] = 6364136223846793005ULL * (w->mt[i - 1] ^ (w->mt[i - 1] >> 62)) + (uint64_t)i;
    w->mti = 312;
}

/* ref: src/mt19937.f90:176-214 genrand64_int64 (standard MT19937-64 of Matsumoto & Nishimura) */
uint64_t po_mt_next_u64(po_ctx* w) {
    static const uint64_t mag01[2] = {0ULL, 0xB5026F5AA96619E9ULL};
    const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    uint64_t x;
    if (!w->seeded) { /* seed = RNG_SEED_GLOBAL*(PROC_ID+1), ref: src/initialization.f90:2294-2296 */
        mt_init(w, (uint64_t)w->c.rng_seed_global * (uint64_t)(w->c.rank + 1));
        w->seeded = 1;
    }
    if (w->mti >= 312) {
        int i;
        for (i = 0; i < 312 - 156; i++) {
            x = (w->mt[i] & UM) | (w->mt[i + 1] & LM);
            w->mt[i] = w->mt[i + 156] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
        }
        for (; i < 311; i++) {
            x = (w->mt[i] & UM) | (w->mt[i + 1] & LM);
            w->mt[i] = w->mt[i + (156 - 312)] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
        }
        x = (w->mt[311] & UM) | (w->mt[0] & LM);
        w->mt[311] = w->mt[155] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
        w->mti = 0;
    }
    x = w->mt[w->mti++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
}

/* Philox4x32-10 (Salmon et al., SC'11); standard constants. */
void po_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint64_t seed_local(const po_ctx* w) {
    /* ref: src/initialization.f90:2294-2296  RNG_SEED_LOCAL = RNG_SEED_GLOBAL*(PROC_ID+1) */
    return (uint64_t)w->c.rng_seed_global * (uint64_t)(w->c.rank + 1);
}

/* Keyed stream selection (device RNG contract, DESIGN.md "RNG"):
 *   key = { seed_lo ^ kind*0x632BE5AB, seed_hi ^ tID },  ctr = { word/4, e0_lo, e0_hi, e1 }; the stream is
 *   the blocks' 32-bit words in order: rf() takes two (lo, hi -> 53 bits), po_pick() one for N <= 65536.
 * kinds: 1 ADVECT (e0 = particle slot), 2 DSMC (e0 = cell (0-based), e1 = candidate),
 *        3 MCC_VAHEDI (e0 = particle slot), 4 MCC (e0 = particle slot).  No-op in MT mode. */
void po_rng_stream(po_ctx* w, int kind, int tID, uint64_t e0, uint64_t e1) {
    if (w->rng_mode != 1) return;
    uint64_t s = seed_local(w);
    w->key[0] = (uint32_t)s ^ ((uint32_t)kind * 0x632BE5ABu);
    w->key[1] = (uint32_t)(s >> 32) ^ (uint32_t)tID;
    w->ent[0] = (uint32_t)e0;
    w->ent[1] = (uint32_t)(e0 >> 32);
    w->ent[2] = (uint32_t)e1;
    w->ndraw = 0;
}

/* next 32-bit word of the keyed Philox stream: word n is word (n mod 4) of block (n / 4) */
static uint32_t philox_word(po_ctx* w) {
    uint32_t n = w->ndraw++;
    if ((n & 3u) == 0) {
        uint32_t ctr[4] = {n >> 2, w->ent[0], w->ent[1], w->ent[2]};
        po_philox4x32_10(ctr, w->key, w->phx_buf);
    }
    return w->phx_buf[n & 3u];
}

/* ref: src/tools.f90:33-43 rf() = genrand64_real1 on [0,1] (src/mt19937.f90:220-225) */
double po_rf(po_ctx* w) {
    if (w->rng_mode == 2) {
        /* injected draws: every rf() of the restated routines consumes the script in call order, exactly as the
         * reference consumes its sequential stream; past the end the script repeats 0.5 (and the overrun is counted) */
        if (w->script_pos < w->script_n) return w->script[w->script_pos++];
        w->script_pos++;
        return 0.5;
    }
    if (w->rng_mode == 0) {
        return (double)(po_mt_next_u64(w) >> 11) * (1.0 / 9007199254740991.0);
    }
    /* Philox: a stream of 32-bit words, four per block; a uniform takes two (53 bits on [0,1)) */
    uint32_t lo = philox_word(w), hi = philox_word(w);
    uint64_t u = ((uint64_t)hi << 32) | lo;
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

/* INT(N*rf()) of the pair selection (ref: src/collisions.f90:263-271), clamped to N-1 (rf() is [0,1], App. B4).
 * Philox mode: for N <= 65536 the index takes ONE word of the stream, floor(N*word/2^32) (bias <= N/2^32); the
 * two indices and the acceptance draw of a candidate are then exactly one Philox block.  Larger N use rf(). */
int po_pick(po_ctx* w, int n) {
    int k;
    if (w->rng_mode == 1 && n <= 65536) k = (int)(((uint64_t)philox_word(w) * (uint64_t)(uint32_t)n) >> 32);
    else k = (int)(n * po_rf(w));
    if (k >= n) k = n - 1;
    return k;
}

/* ------------------------------------------------------------------------------------------
 * samplers
 * ---------------------------------------------------------------------------------------- */

/* ref: src/tools.f90:130-187 MAXWELL */
void po_maxwell(po_ctx* w, double ux, double uy, double uz, double tx, double ty, double tz,
                double m, double* v3) {
    const double PI2 = 2. * PI;
    double tt[3] = {tx, ty, tz}, vel[3];
    for (int i = 0; i < 3; i++) {
        if (tt[i] == 0.0) {
            vel[i] = 0;
        } else {
            double r = po_rf(w);
            double teta = PI2 * r;
            double beta = 1. / sqrt(2. * KB / m * tt[i]);
            double r1 = po_rf(w);
            while (r1 < 1.0e-13) r1 = po_rf(w);
            double ro = sqrt(-log(r1)) / beta;
            vel[i] = ro * sin(teta);
        }
    }
    v3[0] = ux + vel[0];
    v3[1] = uy + vel[1];
    v3[2] = uz + vel[2];
}

/* ref: src/tools.f90:245-285 INTERNAL_ENERGY */
double po_internal_energy(po_ctx* w, int dof, double temp) {
    if (dof == 0) return 0.;
    if (dof == 2) {
        double r = po_rf(w);
        while (r < 1.0e-13) r = po_rf(w);
        return -log(r) * KB * temp;
    }
    double param = (dof == 1) ? (double)5.0f : (double)0.01f; /* single-precision literals, App. B2 */
    double r = po_rf(w);
    double x = -log(r);
    r = po_rf(w);
    while (r > pow(x / param, dof / 2. - 1.)) {
        r = po_rf(w);
        x = -log(r);
        r = po_rf(w);
    }
    return x * KB * temp;
}

/* ref: src/tools.f90:1150-1186 FLX */
double po_flx(po_ctx* w, double sn, double tinf, double m) {
    double beta = 1. / sqrt(2. * KB / m * tinf);
    double acca = sqrt(sn * sn + 2.);
    double kappa = 2. / (sn + acca) * exp(0.5 + 0.5 * sn * (sn - acca));
    for (;;) {
        double r1 = po_rf(w);
        double y = -3. + 6. * r1;
        double r2 = po_rf(w);
        double fm = kappa * (y + sn) * exp(-(y * y));
        if (r2 <= fm) return y / beta;
    }
}

/* ref: src/tools.f90:1544-1593 INTERP_CS (1-based L,R emulated on 0-based arrays) */
double po_interp_cs(double e, int n, const double* te, const double* tcs) {
    int L = 1, R = n, idx;
#define TE(i) te[(i)-1]
#define TC(i) tcs[(i)-1]
    if (e < TE(L)) return 0;
    if (e > TE(R)) return 0;
    if (R == L + 1) return TC(L) + (TC(R) - TC(L)) * (e - TE(L)) / (TE(R) - TE(L));
    /* DEVIATION: for e == TE(R) exactly the reference's bisection loops forever (tools.f90:1574-1586);
     * the table end value is returned instead (same in the CUDA path). */
    if (e == TE(R)) return TC(R);
    for (;;) {
        idx = (L + R) / 2;
        if (TE(idx) <= e) {
            if (TE(idx + 1) > e) break;
            L = idx;
        } else {
            if (TE(idx - 1) <= e) {
                idx = idx - 1;
                break;
            }
            R = idx;
        }
    }
    L = idx;
    R = idx + 1;
    return TC(L) + (TC(R) - TC(L)) * (e - TE(L)) / (TE(R) - TE(L));
#undef TE
#undef TC
}

/* ref: src/tools.f90:1499-1535 BINARY_SEARCH (returns 1-based index, -1 out of bounds) */
static int binary_search(double value, const double* a, int n) {
    int L = 1, R = n, idx;
#define A(i) a[(i)-1]
    if (value < A(L) || value > A(R)) return -1;
    if (R == L + 1) return L;
    /* DEVIATION: value == ARRAY(R) makes the reference loop forever (tools.f90:1518-1531); last cell returned */
    if (value == A(R)) return R - 1;
    for (;;) {
        idx = (L + R) / 2;
        if (A(idx) <= value) {
            if (A(idx + 1) > value) return idx;
            L = idx;
        } else {
            if (A(idx - 1) <= value) return idx - 1;
            R = idx;
        }
    }
#undef A
}

/* NORM2 as gfortran expands it inline (scaled accumulation; trans-intrinsic.c gfc_conv_intrinsic_arith) */
static double norm2_3(const double* x) {
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

/* ref: src/tools.f90:1284-1292 CROSS */
static void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------------------------------
 * particle container
 * ---------------------------------------------------------------------------------------- */

static void reserve(po_ctx* w, int64_t n) {
    if (n <= w->cap) return;
    int64_t nc = w->cap ? w->cap : 16;
    while (nc < n) nc *= 2;
    w->p = (part_t*)realloc(w->p, sizeof(part_t) * (size_t)nc);
    w->cap = nc;
}

/* ref: src/particle.f90:49-80 INIT_PARTICLE + :86-136 ADD_PARTICLE_ARRAY */
static void add_particle(po_ctx* w, double x, double y, double z, double vx, double vy, double vz,
                         double erot, double evib, int s_id, int ic, double dtrim) {
    reserve(w, w->np + 1);
    part_t* q = &w->p[w->np++];
    q->X = x; q->Y = y; q->Z = z;
    q->VX = vx; q->VY = vy; q->VZ = vz;
    q->EROT = erot; q->EVIB = evib;
    q->S_ID = s_id; q->IC = ic; q->DTRIM = dtrim;
    q->ID = (int64_t)w->c.rank + (w->id_counter << 8);
    w->id_counter++;
    q->DUMP_TRAJ = 0;
    q->pad_ = 0;
}

/* ref: src/particle.f90:142-179 REMOVE_PARTICLE_ARRAY (ip 0-based here) */
static void remove_particle(po_ctx* w, int64_t ip) {
    w->p[ip] = w->p[w->np - 1];
    w->np--;
}

/* ------------------------------------------------------------------------------------------
 * mesh accessors (Fortran shapes, see pantera_gpu.h)
 * ---------------------------------------------------------------------------------------- */
#define NF(w) ((w)->dims_mesh + 1)
#define CELL_NODE(w, i, ic) ((w)->cell_nodes[NF(w) * ((ic)-1) + ((i)-1)])
#define NODE_X(w, k, v) ((w)->node_coords[3 * ((v)-1) + ((k)-1)])
#define FNORMAL(w, k, i, ic) ((w)->face_normal[3 * (NF(w) * ((ic)-1) + ((i)-1)) + ((k)-1)])
#define FTANG1(w, k, i, ic) ((w)->face_tang1[3 * (4 * ((ic)-1) + ((i)-1)) + ((k)-1)])
#define FTANG2(w, k, i, ic) ((w)->face_tang2[3 * (4 * ((ic)-1) + ((i)-1)) + ((k)-1)])
#define NEIGHBOR(w, i, ic) ((w)->cell_neighbors[NF(w) * ((ic)-1) + ((i)-1)])
#define FACE_PG(w, i, ic) ((w)->cell_faces_pg[NF(w) * ((ic)-1) + ((i)-1)])
#define EDGE_LEN(w, i, ic) ((w)->cell_edges_len[3 * ((ic)-1) + ((i)-1)])
#define BASIS(w, j, i, ic) ((w)->basis_coeffs[NF(w) * NF(w) * ((ic)-1) + NF(w) * ((i)-1) + ((j)-1)])

/* ------------------------------------------------------------------------------------------
 * ADVECT and its helpers
 * ---------------------------------------------------------------------------------------- */

/* ref: src/timecycle.f90:2297-2332 MOVE_PARTICLE */
static void move_particle(const po_ctx* w, part_t* q, double time) {
    if (w->c.axi) {
        q->X = q->X + q->VX * time;
        q->Z = 0.0;
        double yy = q->Y + q->VY * time;
        double zz = q->VZ * time;
        double r = sqrt(yy * yy + zz * zz);
        double sintheta = q->VZ * time / r;
        double costheta = copysign(sqrt(1. - sintheta * sintheta), yy);
        q->Y = r;
        double vz = q->VZ, vy = q->VY;
        q->VZ = costheta * vz - sintheta * vy;
        q->VY = sintheta * vz + costheta * vy;
    } else {
        q->X = q->X + q->VX * time;
        q->Y = q->Y + q->VY * time;
        q->Z = q->Z + q->VZ * time;
    }
}

/* ref: src/timecycle.f90:2271-2291 UPDATE_VELOCITY_BORIS */
static void boris(double dtime, const double* v_old, double* v_new, double charge, double mass,
                  const double* e, const double* b) {
    double qc = charge * QE;
    double vm[3], t[3], vp[3], s[3], vpl[3], c[3];
    for (int k = 0; k < 3; k++) vm[k] = v_old[k] + 0.5 * qc * e[k] / mass * dtime;
    for (int k = 0; k < 3; k++) t[k] = 0.5 * qc * b[k] / mass * dtime;
    cross3(vm, t, c);
    for (int k = 0; k < 3; k++) vp[k] = vm[k] + c[k];
    double t2 = 1. + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int k = 0; k < 3; k++) s[k] = 2. * t[k] / t2;
    cross3(vp, s, c);
    for (int k = 0; k < 3; k++) vpl[k] = vm[k] + c[k];
    for (int k = 0; k < 3; k++) v_new[k] = vpl[k] + 0.5 * qc * e[k] / mass * dtime;
}

/* ref: src/fields.f90:5341-5431 COMPUTE_WEIGHTS (structured grids) */
static void compute_weights(const po_ctx* w, const part_t* q, double* wt, int* idx, int* indi, int* indj) {
    int nx = w->c.nx, ny = w->c.ny;
    int ic = q->IC, i, j;
    double hx, hy, cxmin, cymin;
    i = (ic - 1) % nx;
    j = (ic - 1) / nx;
    if (w->c.grid_type == PGPU_RECTILINEAR_UNIFORM) {
        hx = (w->c.xmax - w->c.xmin) / (double)nx;
        hy = (w->c.ymax - w->c.ymin) / (double)ny;
        cxmin = i * hx + w->c.xmin;
        cymin = j * hy + w->c.ymin;
    } else {
        hx = w->xcoord[i + 1] - w->xcoord[i]; /* XSIZE(I+1) */
        hy = w->ycoord[j + 1] - w->ycoord[j]; /* YSIZE(J+1) */
        cxmin = w->xcoord[i];
        cymin = w->ycoord[j];
    }
    double fx = (q->X - cxmin) / hx;
    double fy = (q->Y - cymin) / hy;
    if (w->c.dims == 2) {
        indi[0] = i; indi[1] = i + 1; indi[2] = i; indi[3] = i + 1;
        indj[0] = j; indj[1] = j; indj[2] = j + 1; indj[3] = j + 1;
        idx[0] = ic + j - 1;
        idx[1] = idx[0] + 1;
        idx[2] = idx[0] + nx + 1;
        idx[3] = idx[2] + 1;
        wt[0] = (1. - fx) * (1. - fy);
        wt[1] = fx * (1. - fy);
        wt[2] = (1. - fx) * fy;
        wt[3] = fx * fy;
    } else {
        indi[0] = i; indi[1] = i + 1; indi[2] = -1; indi[3] = -1;
        indj[0] = 0; indj[1] = 0; indj[2] = -1; indj[3] = -1;
        idx[0] = ic - 1; idx[1] = ic; idx[2] = -1; idx[3] = -1;
        wt[0] = (1. - fx); wt[1] = fx; wt[2] = -1; wt[3] = -1;
    }
    (void)ny;
}

/* ref: src/fields.f90:4194-4223 APPLY_E_FIELD */
static void apply_e_field(const po_ctx* w, const part_t* q, double* e) {
    if (!w->e_field) { e[0] = e[1] = e[2] = 0; return; }
    if (w->c.grid_type == PGPU_UNSTRUCTURED) {
        for (int k = 0; k < 3; k++) e[k] = w->e_field[3 * (q->IC - 1) + k];
    } else {
        double wt[4];
        int idx[4], indi[4], indj[4];
        compute_weights(w, q, wt, idx, indi, indj);
        /* E_FIELD(3, 0:NPY-1, 0:NPX-1) indexed (:, INDJ, INDI) */
        int npy = (w->c.dims == 2) ? w->c.ny + 1 : 1;
#define EF(k, jj, ii) w->e_field[(k) + 3 * ((jj) + npy * (ii))]
        if (w->c.dims == 2) {
            for (int k = 0; k < 3; k++)
                e[k] = wt[0] * EF(k, indj[0], indi[0]) + wt[1] * EF(k, indj[1], indi[1]) +
                       wt[2] * EF(k, indj[2], indi[2]) + wt[3] * EF(k, indj[3], indi[3]);
        } else {
            for (int k = 0; k < 3; k++)
                e[k] = wt[0] * EF(k, indj[0], indi[0]) + wt[1] * EF(k, indj[1], indi[1]);
        }
#undef EF
    }
}

/* ref: src/fields.f90:4227-4249 APPLY_B_FIELD (2-D unstructured nodal interpolation) */
static void apply_b_field(const po_ctx* w, const part_t* q, double* b) {
    b[0] = b[1] = b[2] = 0.0;
    if (w->c.grid_type == PGPU_UNSTRUCTURED && w->b_field && w->dims_mesh == 2) {
        int ic = q->IC;
        for (int j = 1; j <= 3; j++) {
            int vj = CELL_NODE(w, j, ic);
            double psij = q->X * BASIS(w, 1, j, ic) + q->Y * BASIS(w, 2, j, ic) + BASIS(w, 3, j, ic);
            for (int k = 0; k < 3; k++) b[k] = b[k] + w->b_field[3 * (vj - 1) + k] * psij;
        }
    }
}

/* ref: src/fields.f90:4302-4341 WALL_REACT (note: no EXIT after a successful reaction, App. A6) */
static void wall_react(po_ctx* w, part_t* q, int* remove) {
    int js = q->S_ID;
    double prob_scale = 1.;
    *remove = 0;
    for (int jr = 0; jr < w->nwr; jr++) {
        if (w->wr_r[jr] == js) {
            if (po_rf(w) <= w->wr_prob[jr] / prob_scale) {
                if (w->wr_nprod[jr] == 0) {
                    *remove = 1;
                    q->DTRIM = 0.;
                } else if (w->wr_nprod[jr] == 1) {
                    int jp = w->wr_p1[jr];
                    q->S_ID = jp;
                    double vel_scale = w->mass[js - 1] / w->mass[jp - 1];
                    q->VX = q->VX * vel_scale;
                    q->VY = q->VY * vel_scale;
                    q->VZ = q->VZ * vel_scale;
                }
            } else {
                prob_scale = prob_scale - w->wr_prob[jr];
            }
        }
    }
}

/* ref: src/postprocess.f90:1098-1147 TALLY_PARTICLE_TO_BOUNDARY
 * tallies layout: [NIN,NOUT,PXIN,PYIN,PZIN,PXOUT,PYOUT,PZOUT,EIN,EOUT][s][b] */
static void tally(po_ctx* w, int reflected, const part_t* q, int ic, int iface) {
    if (!w->face_bidx || !w->tallies) return;
    int b = w->face_bidx[NF(w) * (ic - 1) + (iface - 1)];
    if (b < 1) return;
    size_t nbs = (size_t)w->nboundcells * w->ns;
    size_t index = (size_t)(b - 1) + (size_t)(q->S_ID - 1) * w->nboundcells;
    double area = w->bound_area[b - 1];
    double molmass = w->mass[q->S_ID - 1];
    double K = w->c.fnum * molmass / area / w->c.dt;
    double* T = w->tallies;
    if (reflected) {
        T[1 * nbs + index] -= w->c.fnum / area / w->c.dt;
        T[5 * nbs + index] -= K * q->VX;
        T[6 * nbs + index] -= K * q->VY;
        T[7 * nbs + index] -= K * q->VZ;
        T[9 * nbs + index] -= K * (q->VX * q->VX + q->VY * q->VY + q->VZ * q->VZ);
    } else {
        T[0 * nbs + index] += w->c.fnum / area / w->c.dt;
        T[2 * nbs + index] += K * q->VX;
        T[3 * nbs + index] += K * q->VY;
        T[4 * nbs + index] += K * q->VZ;
        T[8 * nbs + index] += K * (q->VX * q->VX + q->VY * q->VY + q->VZ * q->VZ);
    }
}

static int tally_active(const po_ctx* w, int tID) {
    /* ref: src/timecycle.f90:1375 MOD(tID-DUMP_BOUND_START, DUMP_BOUND_AVG_EVERY) .EQ. 0 */
    if (w->c.dump_bound_avg_every <= 0) return 0;
    return ((tID - w->c.dump_bound_start) % w->c.dump_bound_avg_every) == 0;
}

/* ref: src/grid_and_partition.f90:34-96 CELL_FROM_POSITION */
static int cell_from_position(const po_ctx* w, double xp, double yp) {
    if (w->c.grid_type == PGPU_RECTILINEAR_UNIFORM) {
        double dx = (w->c.xmax - w->c.xmin) / w->c.nx;
        double dy = (w->c.ymax - w->c.ymin) / w->c.ny;
        int xcell = (int)((xp - w->c.xmin) / dx);
        int ycell = (int)((yp - w->c.ymin) / dy);
        return xcell + w->c.nx * ycell + 1;
    } else {
        int xcell = binary_search(xp, w->xcoord, w->nxp);
        int ycell = binary_search(yp, w->ycoord, w->nyp);
        return xcell + w->c.nx * (ycell - 1);
    }
}

/* diffuse re-emission, structured flavour. ref: src/timecycle.f90:1915-1931, 1983-1999 */
static void diffuse_structured(po_ctx* w, part_t* q, double temp, double nxw, double nyw) {
    int s = q->S_ID;
    double v3[3];
    po_maxwell(w, 0., 0., 0., temp, temp, temp, w->mass[s - 1], v3); /* VDUMMY, V_PERP, VZ */
    double erot = po_internal_energy(w, w->rotdof[s - 1], temp);
    double evib = po_internal_energy(w, w->vibdof[s - 1], temp);
    double v_norm = po_flx(w, 0., temp, w->mass[s - 1]);
    q->VX = v_norm * nxw - v3[1] * nyw;
    q->VY = v3[1] * nxw + v_norm * nyw;
    q->VZ = v3[2];
    q->EROT = erot;
    q->EVIB = evib;
}

/* Quadratic candidate solver shared by the three axisymmetric variants.
 * tol = 1e-15 on unstructured edges (timecycle.f90:1068,1203), 0 on structured (1685,1788).
 * Returns number of good solutions, TEST sorted as the reference orders them. */
static int axi_roots(double coefa, double coefb, double coefc, double dtcoll, double tol,
                     int use_radial_check, double beta, double alpha, double vx, double y1,
                     double* test) {
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

/* One particle of ADVECT on an unstructured grid. ref: src/timecycle.f90:1166-1657 */
static void advect_unstructured(po_ctx* w, part_t* q, int tID, int* remove_part) {
    const int dims = w->c.dims;
    const double TOL = 1.0e-15;
    int ic = q->IC;
    while (q->DTRIM > 0.) {
        double dtcoll = q->DTRIM;
        int boundcoll = -1;
        if (w->c.axi) {
            /* ref: :1173-1253 */
            for (int i = 1; i <= 3; i++) {
                double test[2];
                int goodsol;
                double ex1 = NODE_X(w, 1, CELL_NODE(w, i, ic));
                double ey1 = NODE_X(w, 2, CELL_NODE(w, i, ic));
                double n1 = FNORMAL(w, 1, i, ic), n2 = FNORMAL(w, 2, i, ic);
                if (fabs(n2) < 1.e-10) {
                    double sol1 = (ex1 - q->X) / q->VX;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double alpha = -n1 / n2;
                    double beta = (q->X - ex1) * alpha;
                    double coefa = q->VY * q->VY + q->VZ * q->VZ - alpha * alpha * (q->VX * q->VX);
                    double coefb = q->Y * q->VY - (beta + ey1) * alpha * q->VX;
                    double coefc = q->Y * q->Y - (beta + ey1) * (beta + ey1);
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, TOL, 1, beta, alpha, q->VX, ey1, test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle(w, q, test[s]);
                    if ((q->VX * n1 + q->VY * n2) > 0) {
                        double len = EDGE_LEN(w, i, ic);
                        double colldist = -n2 * (q->X - ex1) / len + n1 * (q->Y - ey1) / len;
                        if (colldist >= 0 && colldist <= 1) {
                            dtcoll = test[s];
                            boundcoll = i;
                        }
                    }
                    move_particle(w, q, -test[s]);
                }
            }
        } else {
            /* ref: :1257-1305 (1-D), (2-D), (3-D): sequential tightening, '>=' tie rule */
            for (int i = 1; i <= dims + 1; i++) {
                int v = CELL_NODE(w, i, ic);
                double vn, dx;
                if (dims == 1) {
                    vn = q->VX * FNORMAL(w, 1, i, ic);
                    dx = (NODE_X(w, 1, v) - q->X) * FNORMAL(w, 1, i, ic);
                } else if (dims == 2) {
                    vn = q->VX * FNORMAL(w, 1, i, ic) + q->VY * FNORMAL(w, 2, i, ic);
                    dx = (NODE_X(w, 1, v) - q->X) * FNORMAL(w, 1, i, ic) +
                         (NODE_X(w, 2, v) - q->Y) * FNORMAL(w, 2, i, ic);
                } else {
                    vn = q->VX * FNORMAL(w, 1, i, ic) + q->VY * FNORMAL(w, 2, i, ic) +
                         q->VZ * FNORMAL(w, 3, i, ic);
                    dx = (NODE_X(w, 1, v) - q->X) * FNORMAL(w, 1, i, ic) +
                         (NODE_X(w, 2, v) - q->Y) * FNORMAL(w, 2, i, ic) +
                         (NODE_X(w, 3, v) - q->Z) * FNORMAL(w, 3, i, ic);
                }
                if (vn >= 0. && vn * dtcoll >= dx) {
                    dtcoll = dx / vn;
                    boundcoll = i;
                }
            }
        }

        if (boundcoll != -1) {
            /* ref: :1309-1652 */
            move_particle(w, q, dtcoll);
            q->DTRIM = q->DTRIM - dtcoll;

            int fluidboundary = 0;
            int neighbor = NEIGHBOR(w, boundcoll, ic);
            if (neighbor == -1) {
                fluidboundary = 1;
            } else {
                int npg = w->cell_pg ? w->cell_pg[neighbor - 1] : -1;
                if (npg != -1 && w->bc_volume[npg - 1] == PGPU_SOLID) fluidboundary = 1;
            }
            int face_pg = FACE_PG(w, boundcoll, ic);
            double fn[3], ft1[3], ft2[3];
            for (int k = 0; k < 3; k++) fn[k] = -FNORMAL(w, k + 1, boundcoll, ic);
            if (dims == 1) {
                ft1[0] = 0.; ft1[1] = fn[0]; ft1[2] = 0.;
                ft2[0] = 0.; ft2[1] = 0.; ft2[2] = 1.;
            } else if (dims == 2) {
                ft1[0] = -fn[1]; ft1[1] = fn[0]; ft1[2] = 0.;
                ft2[0] = 0.; ft2[1] = 0.; ft2[2] = 1.;
            } else {
                for (int k = 0; k < 3; k++) {
                    ft1[k] = FTANG1(w, k + 1, boundcoll, ic);
                    ft2[k] = FTANG2(w, k + 1, boundcoll, ic);
                }
            }

            if (fluidboundary) {
                if (face_pg != -1) {
                    int g = face_pg - 1;
                    /* boundary-particle dump (:1368-1371) and drag force (:1381-1397) are
                     * output-side features outside the path. */
                    if (tally_active(w, tID)) tally(w, 0, q, ic, boundcoll);

                    /* dielectric surface charge, ref: :1399-1430 */
                    double charge = w->charge[q->S_ID - 1];
                    if (w->bc_field[g] == PGPU_DIELECTRIC_BC && fabs(charge) >= 1.e-6 && w->surface_charge) {
                        double K = QE / (EPS0 * (w->c.eps_scaling * w->c.eps_scaling));
                        double rho_q;
                        if (dims == 1) {
                            rho_q = K * charge * w->c.fnum / (w->c.ymax - w->c.ymin) / (w->c.zmax - w->c.zmin);
                            for (int i = 1; i <= 2; i++) {
                                int vp = CELL_NODE(w, i, ic);
                                /* the reference indexes BASIS_COEFFS(3,I,IC) here, out of bounds for the
                                 * (2,2,NC) 1-D array (:1406-1407); the intended constant term (2,I,IC) is used. */
                                double psip = BASIS(w, 1, i, ic) * q->X + BASIS(w, 2, i, ic);
                                w->surface_charge[vp - 1] += rho_q * psip; /* SURFACE_CHARGE(VP), 1-based */
                            }
                        } else if (dims == 2) {
                            rho_q = K * charge * w->c.fnum / (w->c.zmax - w->c.zmin);
                            for (int i = 1; i <= 3; i++) {
                                int vp = CELL_NODE(w, i, ic);
                                double psip = BASIS(w, 1, i, ic) * q->X + BASIS(w, 2, i, ic) * q->Y + BASIS(w, 3, i, ic);
                                w->surface_charge[vp - 1] += rho_q * psip;
                            }
                        } else {
                            rho_q = K * charge * w->c.fnum;
                            for (int i = 1; i <= 4; i++) {
                                int vp = CELL_NODE(w, i, ic);
                                double psip = BASIS(w, 1, i, ic) * q->X + BASIS(w, 2, i, ic) * q->Y +
                                              BASIS(w, 3, i, ic) * q->Z + BASIS(w, 4, i, ic);
                                w->surface_charge[vp - 1] += rho_q * psip;
                            }
                        }
                    }

                    int pbc = w->bc_particle[g];
                    if (pbc == PGPU_SPECULAR) {
                        /* ref: :1433-1443 */
                        if (w->bc_react[g]) wall_react(w, q, remove_part);
                        double vdotn = q->VX * fn[0] + q->VY * fn[1] + q->VZ * fn[2];
                        q->VX = q->VX - 2. * vdotn * fn[0];
                        q->VY = q->VY - 2. * vdotn * fn[1];
                        q->VZ = q->VZ - 2. * vdotn * fn[2];
                    } else if (pbc == PGPU_DIFFUSE) {
                        /* ref: :1445-1477 */
                        if (w->bc_react[g]) wall_react(w, q, remove_part);
                        int s = q->S_ID;
                        double wt = w->bc_temp[g];
                        double v3[3];
                        po_maxwell(w, 0., 0., 0., wt, wt, wt, w->mass[s - 1], v3);
                        double erot = po_internal_energy(w, w->rotdof[s - 1], wt);
                        double evib = po_internal_energy(w, w->vibdof[s - 1], wt);
                        double v_perp = po_flx(w, 0., wt, w->mass[s - 1]);
                        q->VX = v_perp * fn[0] + v3[1] * ft1[0] + v3[2] * ft2[0];
                        q->VY = v_perp * fn[1] + v3[1] * ft1[1] + v3[2] * ft2[1];
                        q->VZ = v_perp * fn[2] + v3[1] * ft1[2] + v3[2] * ft2[2];
                        q->EROT = erot;
                        q->EVIB = evib;
                    } else if (pbc == PGPU_CLL) {
                        /* ref: :1485-1550 */
                        if (w->bc_react[g]) wall_react(w, q, remove_part);
                        double vdotn = q->VX * fn[0] + q->VY * fn[1] + q->VZ * fn[2];
                        double tg1[3], tg2[3];
                        tg1[0] = q->VX - vdotn * fn[0];
                        tg1[1] = q->VY - vdotn * fn[1];
                        tg1[2] = q->VZ - vdotn * fn[2];
                        double nrm = norm2_3(tg1);
                        for (int k = 0; k < 3; k++) tg1[k] = tg1[k] / nrm;
                        cross3(fn, tg1, tg2);
                        double vdott1 = q->VX * tg1[0] + q->VY * tg1[1] + q->VZ * tg1[2];
                        int s = q->S_ID;
                        double wt = w->bc_temp[g];
                        double vrm = sqrt(2. * KB * wt / w->mass[s - 1]);
                        double rn = po_rf(w);
                        while (rn < 1.0e-13) rn = po_rf(w);
                        double r1 = sqrt(-w->bc_accn[g] * log(rn));
                        double theta1 = 2. * PI * po_rf(w);
                        double dot_norm = vdotn / vrm * sqrt(1 - w->bc_accn[g]);
                        double v_perp = vrm * sqrt(r1 * r1 + dot_norm * dot_norm + 2. * r1 * dot_norm * cos(theta1));
                        rn = po_rf(w);
                        while (rn < 1.0e-13) rn = po_rf(w);
                        double r2 = sqrt(-w->bc_acct[g] * log(rn));
                        double theta2 = 2. * PI * po_rf(w);
                        double vtangent = vdott1 / vrm * sqrt(1 - w->bc_acct[g]);
                        double v_tang1 = vrm * (vtangent + r2 * cos(theta2));
                        double v_tang2 = vrm * r2 * sin(theta2);
                        double erot = po_internal_energy(w, w->rotdof[s - 1], wt);
                        double evib = po_internal_energy(w, w->vibdof[s - 1], wt);
                        q->VX = v_perp * fn[0] + v_tang1 * tg1[0] + v_tang2 * tg2[0];
                        q->VY = v_perp * fn[1] + v_tang1 * tg1[1] + v_tang2 * tg2[1];
                        q->VZ = v_perp * fn[2] + v_tang1 * tg1[2] + v_tang2 * tg2[2];
                        q->EROT = erot;
                        q->EVIB = evib;
                    } else {
                        /* VACUUM, REACT, AXIS, PERIODIC_*, EMIT fall through to removal (:1611-1614);
                         * WB_BC (washboard) is out of scope: its tables are not shipped. */
                        *remove_part = 1;
                        q->DTRIM = 0.;
                    }
                    if (!*remove_part) {
                        if (tally_active(w, tID)) tally(w, 1, q, ic, boundcoll);
                    }
                } else {
                    /* ref: :1644-1647 */
                    *remove_part = 1;
                    q->DTRIM = 0.;
                }
            } else {
                /* ref: :1648-1652 */
                q->IC = neighbor;
                ic = neighbor;
            }
        } else {
            /* ref: :1653-1657 */
            move_particle(w, q, q->DTRIM);
            q->DTRIM = 0.;
        }
    }
}

/* One particle of ADVECT on a rectilinear grid. ref: src/timecycle.f90:1659-2015 */
static void advect_structured(po_ctx* w, part_t* q, int* remove_part) {
    const double NORMX[4] = {1., -1., 0., 0.};
    const double NORMY[4] = {0., 0., 1., -1.};
    const double XW[4] = {w->c.xmin, w->c.xmax, w->c.xmin, w->c.xmax};
    const double YW[4] = {w->c.ymax, w->c.ymin, w->c.ymin, w->c.ymax};
    const double BOUNDPOS[4] = {w->c.xmin, w->c.xmax, w->c.ymin, w->c.ymax};
    const int dims = w->c.dims, axi = w->c.axi;

    while (q->DTRIM > 0.) {
        double dtcoll = q->DTRIM;
        int boundcoll = -1;
        for (int i = 1; i <= 2 * dims; i++) {
            if (dims == 2 && axi) {
                /* ref: :1662-1729 */
                double test[2];
                int goodsol;
                if (i == 1 || i == 2) {
                    double sol1 = (BOUNDPOS[i - 1] - q->X) / q->VX;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double coefa = q->VY * q->VY + q->VZ * q->VZ;
                    double coefb = q->Y * q->VY;
                    double coefc = q->Y * q->Y - BOUNDPOS[i - 1] * BOUNDPOS[i - 1];
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, 0.0, 0, 0, 0, 0, 0, test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle(w, q, test[s]);
                    if ((q->VX * NORMX[i - 1] + q->VY * NORMY[i - 1]) < 0) {
                        dtcoll = test[s];
                        boundcoll = i;
                    }
                    move_particle(w, q, -test[s]);
                }
            } else {
                /* ref: :1733-1752 -- strict '>' here, unlike the unstructured faces */
                double vn = -(q->VX * NORMX[i - 1] + q->VY * NORMY[i - 1]);
                double dx = (q->X - XW[i - 1]) * NORMX[i - 1] + (q->Y - YW[i - 1]) * NORMY[i - 1];
                if (vn >= 0. && vn * dtcoll > dx) {
                    dtcoll = dx / vn;
                    boundcoll = i;
                }
            }
        }

        /* interior walls, ref: :1760-1865 */
        int wallcoll = -1;
        for (int i = 1; i <= w->nw; i++) {
            int j = i - 1;
            if (axi) {
                double test[2];
                int goodsol;
                if (w->w_x1[j] == w->w_x2[j]) {
                    double sol1 = (w->w_x1[j] - q->X) / q->VX;
                    if (sol1 >= 0 && sol1 < dtcoll) { goodsol = 1; test[0] = sol1; }
                    else goodsol = 0;
                } else {
                    double alpha = (w->w_y2[j] - w->w_y1[j]) / (w->w_x2[j] - w->w_x1[j]);
                    double beta = (q->X - w->w_x1[j]) * alpha;
                    double coefa = q->VY * q->VY + q->VZ * q->VZ - alpha * alpha * (q->VX * q->VX);
                    double coefb = q->Y * q->VY - (beta + w->w_y1[j]) * alpha * q->VX;
                    double coefc = q->Y * q->Y - (beta + w->w_y1[j]) * (beta + w->w_y1[j]);
                    goodsol = axi_roots(coefa, coefb, coefc, dtcoll, 0.0, 1, beta, alpha, q->VX, w->w_y1[j], test);
                }
                for (int s = 0; s < goodsol; s++) {
                    move_particle(w, q, test[s]);
                    if ((q->VX * w->w_nx[j] + q->VY * w->w_ny[j]) < 0) {
                        double ddx = w->w_x2[j] - w->w_x1[j], ddy = w->w_y2[j] - w->w_y1[j];
                        double colldist = ((q->X - w->w_x1[j]) * ddx + (q->Y - w->w_y1[j]) * ddy) /
                                          (ddx * ddx + ddy * ddy);
                        if (colldist >= 0 && colldist <= 1) {
                            dtcoll = test[s];
                            boundcoll = -1;
                            wallcoll = i;
                        }
                    }
                    move_particle(w, q, -test[s]);
                }
            } else {
                double vn = -(q->VX * w->w_nx[j] + q->VY * w->w_ny[j]);
                double dx = (q->X - w->w_x1[j]) * w->w_nx[j] + (q->Y - w->w_y1[j]) * w->w_ny[j];
                if (dx >= 0. && vn * dtcoll >= dx) {
                    double cand = dx / vn;
                    double xcoll = q->X + q->VX * cand;
                    double ycoll = q->Y + q->VY * cand;
                    double ddx = w->w_x2[j] - w->w_x1[j], ddy = w->w_y2[j] - w->w_y1[j];
                    double colldist = ((xcoll - w->w_x1[j]) * ddx + (ycoll - w->w_y1[j]) * ddy) /
                                      (ddx * ddx + ddy * ddy);
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
            w->bcount[b + 4 * (q->S_ID - 1)] += 1;
            if (w->c.bool_periodic[b]) {
                move_particle(w, q, dtcoll);
                switch (boundcoll) {
                    case 1: q->X = q->X + w->c.xmax - w->c.xmin; break;
                    case 2: q->X = q->X - w->c.xmax + w->c.xmin; break;
                    case 3: q->Y = q->Y + w->c.ymax - w->c.ymin; break;
                    case 4: q->Y = q->Y - w->c.ymax + w->c.ymin; break;
                }
                q->DTRIM = q->DTRIM - dtcoll;
            } else if (w->c.bool_specular[b]) {
                move_particle(w, q, dtcoll);
                if (boundcoll <= 2) q->VX = -q->VX;
                else q->VY = -q->VY;
                q->DTRIM = q->DTRIM - dtcoll;
                if (w->c.bool_react[b]) wall_react(w, q, remove_part);
            } else if (w->c.bool_diffuse[b]) {
                move_particle(w, q, dtcoll);
                q->DTRIM = q->DTRIM - dtcoll;
                if (w->c.bool_react[b]) wall_react(w, q, remove_part);
                if (!*remove_part) diffuse_structured(w, q, w->c.boundtemp, NORMX[b], NORMY[b]);
            } else {
                *remove_part = 1;
                q->DTRIM = 0.;
            }
        } else if (wallcoll != -1) {
            /* ref: :1940-2008 */
            int j = wallcoll - 1;
            w->wcount[j + w->nw * (q->S_ID - 1)] += 1;
            if (w->w_spec[j]) {
                double rfp = po_rf(w);
                if (w->w_porous[j] && rfp <= w->w_transm[j]) {
                    *remove_part = 1;
                    q->DTRIM = 0.;
                } else {
                    move_particle(w, q, dtcoll);
                    double vdotn = q->VX * w->w_nx[j] + q->VY * w->w_ny[j];
                    q->VX = q->VX - 2. * vdotn * w->w_nx[j];
                    q->VY = q->VY - 2. * vdotn * w->w_ny[j];
                    q->DTRIM = q->DTRIM - dtcoll;
                    if (w->w_react[j]) wall_react(w, q, remove_part);
                }
            } else if (w->w_diff[j]) {
                double rfp = po_rf(w);
                if (w->w_porous[j] && rfp <= w->w_transm[j]) {
                    *remove_part = 1;
                    q->DTRIM = 0.;
                } else {
                    move_particle(w, q, dtcoll);
                    q->DTRIM = q->DTRIM - dtcoll;
                    if (w->w_react[j]) wall_react(w, q, remove_part);
                    if (!*remove_part) diffuse_structured(w, q, w->w_temp[j], w->w_nx[j], w->w_ny[j]);
                }
            } else {
                *remove_part = 1;
                q->DTRIM = 0.;
            }
        } else {
            /* ref: :2010-2015 */
            move_particle(w, q, q->DTRIM);
            q->DTRIM = 0.;
        }
    }
}

/* ref: src/timecycle.f90:1021-2268 ADVECT */
int po_advect(po_ctx* w, int32_t tID) {
    clock_t t0 = clock();
    const int uns = (w->c.grid_type == PGPU_UNSTRUCTURED);
    if (uns && !w->cell_nodes) FAIL(w, PGPU_ERR_STATE, "po_advect: unstructured grid without mesh");
    if (!w->ns) FAIL(w, PGPU_ERR_STATE, "po_advect: species not set");
    int64_t np = w->np;
    char* remove_part = (char*)calloc((size_t)(np > 0 ? np : 1), 1);
    memset(w->bcount, 0, sizeof(int64_t) * 4 * (size_t)w->ns);
    if (w->nw) memset(w->wcount, 0, sizeof(int64_t) * (size_t)w->nw * w->ns);

    for (int64_t ip = 0; ip < np; ip++) {
        part_t* q = &w->p[ip];
        po_rng_stream(w, 1, tID, (uint64_t)ip, 0);
        /* ref: :1105-1131 kick at the start-of-step position/cell with the particle's DTRIM */
        if (w->c.pic_type != PGPU_PIC_NONE) {
            double e[3], b[3], v_old[3], v_new[3];
            apply_e_field(w, q, e);
            b[0] = b[1] = b[2] = 0;
            if (w->b_field) apply_b_field(w, q, b);
            for (int k = 0; k < 3; k++) b[k] = b[k] + w->c.external_b[k];
            v_old[0] = q->VX; v_old[1] = q->VY; v_old[2] = q->VZ;
            boris(q->DTRIM, v_old, v_new, w->charge[q->S_ID - 1], w->mass[q->S_ID - 1], e, b);
            q->VX = v_new[0]; q->VY = v_new[1]; q->VZ = v_new[2];
        }
        int rem = 0;
        if (uns) advect_unstructured(w, q, tID, &rem);
        else advect_structured(w, q, &rem);
        remove_part[ip] = (char)rem;

        /* ref: :2048-2068 */
        if (w->c.bool_z_periodic) {
            while (q->Z > w->c.zmax) q->Z = w->c.zmin + (q->Z - w->c.zmax);
            while (q->Z < w->c.zmin) q->Z = w->c.zmax + (q->Z - w->c.zmin);
        }
        if (w->c.dims == 1 && w->c.bool_y_periodic) {
            while (q->Y > w->c.ymax) q->Y = w->c.ymin + (q->Y - w->c.ymax);
            while (q->Y < w->c.ymin) q->Y = w->c.ymax + (q->Y - w->c.ymin);
        }
        q->DTRIM = w->c.dt;
    }

    /* ref: :2099-2132 reverse sweep: swap-remove + structured re-index + radial weighting (:2110-2128; Philox mode:
     * stream kind 8 keyed by the particle's index) */
    for (int64_t ip = np - 1; ip >= 0; ip--) {
        if (remove_part[ip]) {
            remove_particle(w, ip);
        } else if (!uns) {
            int ic = cell_from_position(w, w->p[ip].X, w->p[ip].Y);
            int old_ic = w->p[ip].IC;
            w->p[ip].IC = ic;
            if (w->cell_fnum && ic != old_ic) {
                double weight_ratio = w->cell_fnum[old_ic - 1] / w->cell_fnum[ic - 1];
                if (weight_ratio < 1.) {
                    po_rng_stream(w, 8, tID, (uint64_t)ip, 0);
                    double rfp = po_rf(w);
                    if (weight_ratio < rfp) remove_particle(w, ip);
                } else if (weight_ratio > 1.) {
                    po_rng_stream(w, 8, tID, (uint64_t)ip, 0);
                    double rfp = po_rf(w);
                    while ((weight_ratio - 1.) > rfp) {
                        part_t q = w->p[ip];
                        add_particle(w, q.X, q.Y, q.Z, q.VX, q.VY, q.VZ, q.EROT, q.EVIB, q.S_ID, ic, w->c.dt);
                        weight_ratio = weight_ratio - 1.;
                    }
                }
            }
        }
    }
    free(remove_part);
    w->tsec[0] += (double)(clock() - t0) / CLOCKS_PER_SEC;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * collisions
 * ---------------------------------------------------------------------------------------- */

/* ref: src/collisions.f90:892-908 FUNCTION_I */
static double function_i(double ecoll, double ei, double trdof, int idof) {
    if (idof == 2) return pow(1. - ei / ecoll, trdof / 2. - 1.);
    return pow((trdof + idof - 4.) / (idof - 2.), idof / 2. - 1.) *
           pow((trdof + idof - 4.) / (trdof - 2.), trdof / 2. - 1.) *
           pow(ei / ecoll, idof / 2. - 1.) * pow(1. - ei / ecoll, trdof / 2. - 1.);
}

/* ref: src/collisions.f90:915-937 COLL_INTERNAL_ENERGY */
static double coll_internal_energy(po_ctx* w, double ecoll, double trdof, int intdof) {
    if (trdof < 1.e-6) return ecoll;
    if (intdof == 0) return 0.0;
    for (;;) {
        double r1 = po_rf(w);
        double ei = r1 * ecoll;
        double r2 = po_rf(w);
        if (r2 <= function_i(ecoll, ei, trdof, intdof)) return ei;
    }
}

/* ref: src/collisions.f90:956-992 HS_SCATTER */
static void hs_scatter(po_ctx* w, double ecoll, double m1, double m2, double* c1, double* c2) {
    const double PI2 = 2. * PI;
    double mred = m1 * m2 / (m1 + m2);
    double g = sqrt(2. * ecoll / mred);
    double cosphi = 2. * po_rf(w) - 1.;
    double sinphi = sqrt(1 - cosphi * cosphi);
    double theta = PI2 * po_rf(w);
    double grel[3], wc[3];
    grel[0] = g * sinphi * cos(theta);
    grel[1] = g * sinphi * sin(theta);
    grel[2] = g * cosphi;
    for (int i = 0; i < 3; i++) wc[i] = m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i];
    for (int i = 0; i < 3; i++) {
        c1[i] = wc[i] + m2 / (m1 + m2) * grel[i];
        c2[i] = wc[i] - m1 / (m1 + m2) * grel[i];
    }
}

#define VSS2(arr, s1, s2) ((arr)[((s1)-1) + (size_t)w->ns * ((s2)-1)])

/* Borgnakke-Larsen redistribution after a reaction; shared by VSS_COLLIS (:374-444) and
 * MCC_COLLISIONS_VAHEDI (:1749-1879).  jp1/jp2 index the particle array; new products are appended. */
static void react_redistribute(po_ctx* w, const reaction_t* r, int64_t jp1, int64_t jp2,
                               double ecoll, double* c1, double* c2, int new_ic_from_jp2) {
    int p1 = r->p1, p2 = r->p2, p3 = r->p3, p4 = r->p4;
    (void)new_ic_from_jp2;
    double totdof = 3. + w->rotdof[p2 - 1] + w->vibdof[p2 - 1] + w->rotdof[p1 - 1];
    if (r->nprod >= 3) totdof = totdof + 3. + w->rotdof[p3 - 1] + w->vibdof[p3 - 1];
    if (r->nprod >= 4) totdof = totdof + 3. + w->rotdof[p4 - 1] + w->vibdof[p4 - 1];

    double ei = coll_internal_energy(w, ecoll, totdof, w->vibdof[p1 - 1]);
    w->p[jp1].EVIB = ei;
    ecoll = ecoll - ei;
    totdof = totdof - w->rotdof[p1 - 1];
    ei = coll_internal_energy(w, ecoll, totdof, w->rotdof[p1 - 1]);
    w->p[jp1].EROT = ei;
    ecoll = ecoll - ei;
    totdof = totdof - w->vibdof[p2 - 1];
    ei = coll_internal_energy(w, ecoll, totdof, w->vibdof[p2 - 1]);
    w->p[jp2].EVIB = ei;
    ecoll = ecoll - ei;
    totdof = totdof - w->rotdof[p2 - 1];
    ei = coll_internal_energy(w, ecoll, totdof, w->rotdof[p2 - 1]);
    w->p[jp2].EROT = ei;
    ecoll = ecoll - ei;

    double m1 = w->mass[p1 - 1], m2 = 0;
    if (r->nprod == 2) m2 = w->mass[p2 - 1];
    else if (r->nprod == 3) m2 = w->mass[p2 - 1] + w->mass[p3 - 1];
    else if (r->nprod == 4) m2 = w->mass[p2 - 1] + w->mass[p3 - 1] + w->mass[p4 - 1];
    totdof = totdof - 3.;
    ei = coll_internal_energy(w, ecoll, totdof, 3);
    hs_scatter(w, ei, m1, m2, c1, c2);
    ecoll = ecoll - ei;

    w->p[jp1].VX = c1[0]; w->p[jp1].VY = c1[1]; w->p[jp1].VZ = c1[2];

    if (r->nprod == 2) {
        w->p[jp2].VX = c2[0]; w->p[jp2].VY = c2[1]; w->p[jp2].VZ = c2[2];
    } else if (r->nprod == 3) {
        totdof = totdof - w->vibdof[p3 - 1];
        double evib = coll_internal_energy(w, ecoll, totdof, w->vibdof[p3 - 1]);
        ecoll = ecoll - evib;
        totdof = totdof - w->rotdof[p3 - 1];
        double erot = coll_internal_energy(w, ecoll, totdof, w->rotdof[p3 - 1]);
        ecoll = ecoll - erot;
        m1 = w->mass[p2 - 1];
        m2 = w->mass[p3 - 1];
        for (int k = 0; k < 3; k++) c1[k] = c2[k];
        hs_scatter(w, ecoll, m1, m2, c1, c2);
        w->p[jp2].VX = c1[0]; w->p[jp2].VY = c1[1]; w->p[jp2].VZ = c1[2];
        add_particle(w, w->p[jp2].X, w->p[jp2].Y, w->p[jp2].Z, c2[0], c2[1], c2[2], erot, evib, p3,
                     w->p[jp2].IC, w->c.dt);
    } else if (r->nprod == 4) {
        /* ref: src/collisions.f90:1826-1879 (MCC only) */
        totdof = totdof - w->vibdof[p3 - 1];
        double evib = coll_internal_energy(w, ecoll, totdof, w->vibdof[p3 - 1]);
        ecoll = ecoll - evib;
        totdof = totdof - w->rotdof[p3 - 1];
        double erot = coll_internal_energy(w, ecoll, totdof, w->rotdof[p3 - 1]);
        ecoll = ecoll - erot;
        m1 = w->mass[p2 - 1];
        m2 = w->mass[p3 - 1] + w->mass[p4 - 1];
        for (int k = 0; k < 3; k++) c1[k] = c2[k];
        hs_scatter(w, ecoll, m1, m2, c1, c2);
        w->p[jp2].VX = c1[0]; w->p[jp2].VY = c1[1]; w->p[jp2].VZ = c1[2];
        add_particle(w, w->p[jp2].X, w->p[jp2].Y, w->p[jp2].Z, c2[0], c2[1], c2[2], erot, evib, p3,
                     w->p[jp2].IC, w->c.dt);
        int64_t jp3 = w->np - 1;
        totdof = totdof - w->vibdof[p4 - 1];
        evib = coll_internal_energy(w, ecoll, totdof, w->vibdof[p4 - 1]);
        ecoll = ecoll - evib;
        totdof = totdof - w->rotdof[p4 - 1];
        erot = coll_internal_energy(w, ecoll, totdof, w->rotdof[p4 - 1]);
        ecoll = ecoll - erot;
        m1 = w->mass[p3 - 1];
        m2 = w->mass[p4 - 1];
        for (int k = 0; k < 3; k++) c1[k] = c2[k];
        hs_scatter(w, ecoll, m1, m2, c1, c2);
        w->p[jp3].VX = c1[0]; w->p[jp3].VY = c1[1]; w->p[jp3].VZ = c1[2];
        /* the reference passes P3_SP_ID (not P4) for the fourth product (:1875-1876); kept */
        add_particle(w, w->p[jp3].X, w->p[jp3].Y, w->p[jp3].Z, c2[0], c2[1], c2[2], erot, evib, p3,
                     w->p[jp3].IC, w->c.dt);
    }
}

/* An ACCEPTED pair (jp1, jp2): reaction test, Borgnakke-Larsen exchange, VSS deflection.
 * ref: src/collisions.f90:327-546 (VSS_COLLIS) and :1362-1590 (MCC_COLLISIONS); s1/s2 are the species the
 * caller looked up BEFORE the pair was accepted (MCC_COLLISIONS keeps SP_ID1 across background components even
 * after a reaction changed the particle's S_ID).  Returns 1 when the pair reacted. */
static int accepted_pair(po_ctx* w, int64_t jp1, int64_t jp2, int s1, int s2, double vr, double vr2, int gate_all) {
    const double PI2 = 2. * PI;
    part_t *q1 = &w->p[jp1], *q2 = &w->p[jp2];
    double cref = VSS2(w->grefs, s1, s2), sigma = VSS2(w->sigmas, s1, s2);
    double omega = VSS2(w->omegas, s1, s2), alpha = VSS2(w->alphas, s1, s2);
    double m1 = w->mass[s1 - 1], m2 = w->mass[s2 - 1];
    double mred = m1 * m2 / (m1 + m2);
    double etr = 0.5 * mred * vr2;
    double c1[3] = {q1->VX, q1->VY, q1->VZ}, c2[3] = {q2->VX, q2->VY, q2->VZ};
    double rfp;
    int skip = 0;
    /* ref: :331-449 TCE / LXCat reaction test (MCC_COLLISIONS :1362-1495 is the same code plus an EA gate
     * in front of the reaction-type switch, :1383-1384) */
    double ecoll = etr + q1->EROT + q2->EROT + q1->EVIB + q2->EVIB;
    for (int jr = 0; jr < w->nr; jr++) {
        const reaction_t* r = &w->reac[jr];
        if (r->r1 != s1) {
            if (r->r1 != s2) continue;
            else if (r->r2 != s1) continue;
        } else {
            if (r->r2 != s2) continue;
        }
        double ea = r->ea, ptce;
        if (gate_all && ecoll <= ea) continue;
        if (r->type == PGPU_TCE) {
            if (ecoll <= ea) continue;
            ptce = r->c1 * pow(ecoll - ea, r->c2) * pow(1. - ea / ecoll, r->c3);
        } else if (r->type == PGPU_LXCAT) {
            double sigma_r = po_interp_cs(etr, r->tab_n, r->tab_e, r->tab_cs);
            ptce = sigma_r / (sigma * pow(vr / cref, 1. - 2. * omega));
        } else {
            ptce = 0;
        }
        rfp = po_rf(w);
        if (rfp < ptce) {
            skip = 1;
            w->timestep_reac++;
            w->reac[jr].counts++;
            ecoll = ecoll - ea;
            /* roles are NOT swapped when the pair order is opposite to the reaction (App. A9) */
            q1->S_ID = r->p1;
            q2->S_ID = r->p2;
            react_redistribute(w, r, jp1, jp2, ecoll, c1, c2, 1);
            break;
        }
    }
    if (skip) return 1;
    /* pointers may be stale after a realloc in add_particle only in the reacting branch */
    q1 = &w->p[jp1];
    q2 = &w->p[jp2];

    /* ref: :454-490 Borgnakke-Larsen, at most one mode */
    double trdof = 3.;
    double prot1 = (trdof + w->rotdof[s1 - 1]) / trdof * w->rotrel[s1 - 1];
    double prot2 = (trdof + w->rotdof[s2 - 1]) / trdof * w->rotrel[s2 - 1];
    double pvib1 = (trdof + w->vibdof[s1 - 1]) / trdof * w->vibrel[s1 - 1];
    double pvib2 = (trdof + w->vibdof[s2 - 1]) / trdof * w->vibrel[s2 - 1];
    trdof = 5. - 2. * omega;
    /* Philox mode: when neither species can exchange internal energy the selector cannot fall in any branch and
     * its draw is not taken from the stream (an accepted elastic pair then needs exactly one block: deflection
     * angle and azimuth).  The sequential MT19937 mode consumes it, as the reference does (:472). */
    if (w->rng_mode == 1 && w->rotrel[s1 - 1] + w->rotrel[s2 - 1] + w->vibrel[s1 - 1] + w->vibrel[s2 - 1] == 0.) rfp = 1.;
    else rfp = po_rf(w);
    if (rfp < prot1) {
        ecoll = etr + q1->EROT;
        q1->EROT = coll_internal_energy(w, ecoll, trdof, w->rotdof[s1 - 1]);
        etr = ecoll - q1->EROT;
    } else if (rfp < prot1 + prot2) {
        ecoll = etr + q2->EROT;
        q2->EROT = coll_internal_energy(w, ecoll, trdof, w->rotdof[s2 - 1]);
        etr = ecoll - q2->EROT;
    } else if (rfp < prot1 + prot2 + pvib1) {
        ecoll = etr + q1->EVIB;
        q1->EVIB = coll_internal_energy(w, ecoll, trdof, w->vibdof[s1 - 1]);
        etr = ecoll - q1->EVIB;
    } else if (rfp < prot1 + prot2 + pvib1 + pvib2) {
        ecoll = etr + q2->EVIB;
        q2->EVIB = coll_internal_energy(w, ecoll, trdof, w->vibdof[s2 - 1]);
        etr = ecoll - q2->EVIB;
    }

    /* ref: :495-546 VSS deflection */
    double coschi = 2. * pow(po_rf(w), 1. / alpha) - 1.;
    double sinchi = sqrt(1. - coschi * coschi);
    double theta = PI2 * po_rf(w);
    double costheta = cos(theta), sintheta = sin(theta);
    double gx, gy, gz, g;
    if (rfp < prot1 + prot2 + pvib1 + pvib2) {
        g = sqrt(2. * etr / mred);
        double cosa = 2. * po_rf(w) - 1.0;
        double sina = sqrt(1 - cosa * cosa);
        double bb = PI2 * po_rf(w);
        gx = g * sina * cos(bb);
        gy = g * sina * sin(bb);
        gz = g * cosa;
    } else {
        gx = c1[0] - c2[0];
        gy = c1[1] - c2[1];
        gz = c1[2] - c2[2];
        g = sqrt(gx * gx + gy * gy + gz * gz);
    }
    double grel[3], wc[3];
    double gyz = sqrt(gy * gy + gz * gz);
    grel[0] = gx * coschi + gyz * sintheta * sinchi;
    grel[1] = gy * coschi + (g * gz * costheta - gx * gy * sintheta) / gyz * sinchi;
    grel[2] = gz * coschi - (g * gy * costheta + gx * gz * sintheta) / gyz * sinchi;
    for (int i = 0; i < 3; i++) wc[i] = m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i];
    for (int i = 0; i < 3; i++) {
        c1[i] = wc[i] + m2 / (m1 + m2) * grel[i];
        c2[i] = wc[i] - m1 / (m1 + m2) * grel[i];
    }
    q1->VX = c1[0]; q1->VY = c1[1]; q1->VZ = c1[2];
    q2->VX = c2[0]; q2->VY = c2[1]; q2->VZ = c2[2];
    return 0;
}

/* ref: src/collisions.f90:155-558 VSS_COLLIS for cell jc (1-based); ind = INDALL slice (0-based particle ids) */
static int64_t vss_collis(po_ctx* w, int jc, int tID, int npcj, const int32_t* ind) {
    double cfnum = w->cell_fnum ? w->cell_fnum[jc - 1] : w->c.fnum; /* ref: :192-196 CFNUM = CELL_FNUM(JC) */
    double vxmax = -1.e+38, vxmin = 1.e+38, vymax = -1.e+38, vymin = 1.e+38, vzmax = -1.e+38, vzmin = 1.e+38;
    for (int j = 0; j < npcj; j++) {
        const part_t* q = &w->p[ind[j]];
        vxmin = fmin(vxmin, q->VX); vxmax = fmax(vxmax, q->VX);
        vymin = fmin(vymin, q->VY); vymax = fmax(vymax, q->VY);
        vzmin = fmin(vzmin, q->VZ); vzmax = fmax(vzmax, q->VZ);
    }
    double vrmax = sqrt((vxmax - vxmin) * (vxmax - vxmin) + (vymax - vymin) * (vymax - vymin) +
                        (vzmax - vzmin) * (vzmax - vzmin));
    /* ref: :224-230 (VOL is undefined for other grid kinds in the reference; CELL_VOL used as fallback) */
    double vol;
    if (w->c.grid_type == PGPU_UNSTRUCTURED) vol = w->cell_volumes[jc - 1];
    else vol = (w->c.xmax - w->c.xmin) * (w->c.ymax - w->c.ymin) * (w->c.zmax - w->c.zmin) / (w->c.nx * w->c.ny);

    /* ref: :232  0.5*NPC*(NPC-1) is evaluated in single precision (App. B1) */
    float pairs_f = (0.5f * (float)npcj) * (float)(npcj - 1);
    double ncollmax = (double)pairs_f * w->sigmamax * vrmax * cfnum * w->c.dt / vol;
    long ncollmax_int = (long)floor(ncollmax + 0.5);
    long half = (long)floorf(0.5f * (float)npcj);
    long ncoll;
    if (ncollmax_int < 1) ncoll = 1;
    else if (ncollmax_int > half) ncoll = half;
    else ncoll = ncollmax_int;
    double fcorr = ncollmax / ncoll;
    int64_t ncollreal = 0;

    for (long jcol = 0; jcol < ncoll; jcol++) {
        po_rng_stream(w, 2, tID, (uint64_t)(jc - 1), (uint64_t)jcol);
        /* ref: :263-271; rf() is [0,1] so INT(NPCJ*rf()) may equal NPCJ with probability 2^-53 (App. B4):
         * clamped here, cannot happen with the [0,1) Philox stream */
        int k1 = po_pick(w, npcj);
        int64_t jp1 = ind[k1];
        int64_t jp2 = jp1;
        while (jp2 == jp1) {
            int k2 = po_pick(w, npcj);
            jp2 = ind[k2];
        }
        part_t *q1 = &w->p[jp1], *q2 = &w->p[jp2];
        double vr2 = (q2->VX - q1->VX) * (q2->VX - q1->VX) + (q2->VY - q1->VY) * (q2->VY - q1->VY) +
                     (q2->VZ - q1->VZ) * (q2->VZ - q1->VZ);
        double vr = sqrt(vr2);
        int s1 = q1->S_ID, s2 = q2->S_ID;
        double cref = VSS2(w->grefs, s1, s2), sigma = VSS2(w->sigmas, s1, s2);
        double omega = VSS2(w->omegas, s1, s2);

        double rfp = po_rf(w);
        double collsigma;
        if (fabs(omega - 0.5) < 1.e-6) {
            collsigma = sigma;
        } else {
            if (vr < (double)1e-20f) continue; /* ref: :314 (single-precision literal) */
            collsigma = sigma * pow(vr / cref, 1. - 2. * omega);
        }
        double collprob = fcorr / (w->sigmamax * vrmax) * vr * collsigma;
        if (!(rfp < collprob)) continue;
        ncollreal++;

        accepted_pair(w, jp1, jp2, s1, s2, vr, vr2, 0);
    }
    return ncollreal;
}

static int ncells(const po_ctx* w) {
    return (w->c.grid_type == PGPU_UNSTRUCTURED) ? w->nc : w->c.nx * w->c.ny;
}

/* ref: src/collisions.f90:37-149 DSMC_COLLISIONS (index counting sorts, then per-cell VSS_COLLIS).
 * Only the per-cell arrays NPCALL/IOFALL/INDALL are used by VSS_COLLIS; the per-(species,cell) arrays NPC/IOF/IND feed
 * VAHEDI_COLLIS, which is out of scope -- but the reference allocates, counts and fills BOTH sets every step (:55-112,
 * :141-147), so this restatement does too: it is also the CPU arm of the benchmark and must not be flattered. */
int po_collide(po_ctx* w, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!w->sigmas) FAIL(w, PGPU_ERR_STATE, "po_collide: VSS parameters not set");
    clock_t t0 = clock();
    int nc = ncells(w);
    int64_t np = w->np;
    free(w->npcall);
    w->npcall = (int32_t*)calloc((size_t)nc, sizeof(int32_t));
    const int ns = w->ns;
    int64_t* iofall = (int64_t*)malloc(sizeof(int64_t) * (size_t)nc);
    int32_t* indall = (int32_t*)malloc(sizeof(int32_t) * (size_t)(np > 0 ? np : 1));
    int32_t* npc = (int32_t*)calloc((size_t)nc * ns, sizeof(int32_t));      /* NPC(N_SPECIES, NCELLS) */
    int64_t* iof = (int64_t*)malloc(sizeof(int64_t) * (size_t)nc * ns);     /* IOF(N_SPECIES, NCELLS) */
    int32_t* ind = (int32_t*)malloc(sizeof(int32_t) * (size_t)(np > 0 ? np : 1));
    for (int64_t jp = 0; jp < np; jp++) {
        int jc = w->p[jp].IC, js = w->p[jp].S_ID;
        if (jc < 1 || jc > nc || js < 1 || js > ns) {
            free(iofall); free(indall); free(npc); free(iof); free(ind);
            FAIL(w, PGPU_ERR_INVALID, "po_collide: particle %lld has IC=%d outside 1..%d or S_ID=%d outside 1..%d", (long long)jp, jc, nc, js, ns);
        }
        npc[(size_t)(jc - 1) * ns + (js - 1)]++;
        w->npcall[jc - 1]++;
    }
    int64_t idx = 0, idxall = 0;
    for (int jc = 0; jc < nc; jc++) {
        iofall[jc] = idxall;
        idxall += w->npcall[jc];
        for (int js = 0; js < ns; js++) {
            iof[(size_t)jc * ns + js] = idx;
            idx += npc[(size_t)jc * ns + js];
        }
    }
    int32_t* fill = (int32_t*)calloc((size_t)nc, sizeof(int32_t));
    memset(npc, 0, sizeof(int32_t) * (size_t)nc * ns);
    for (int64_t jp = 0; jp < np; jp++) {
        int jc = w->p[jp].IC - 1, js = w->p[jp].S_ID - 1;
        ind[iof[(size_t)jc * ns + js] + npc[(size_t)jc * ns + js]] = (int32_t)jp;
        npc[(size_t)jc * ns + js]++;
        indall[iofall[jc] + fill[jc]] = (int32_t)jp;
        fill[jc]++;
    }
    free(fill);
    w->tsec[1] += (double)(clock() - t0) / CLOCKS_PER_SEC;
    t0 = clock();

    w->timestep_coll = 0;
    w->timestep_reac = 0;
    for (int jc = 1; jc <= nc; jc++) {
        if (w->npcall[jc - 1] > 1)
            w->timestep_coll += vss_collis(w, jc, tID, w->npcall[jc - 1], indall + iofall[jc - 1]);
    }
    free(iofall);
    free(indall);
    free(npc);
    free(iof);
    free(ind);
    if (timestep_coll) *timestep_coll = w->timestep_coll;
    if (timestep_reac) *timestep_reac = w->timestep_reac;
    w->tsec[2] += (double)(clock() - t0) / CLOCKS_PER_SEC;
    return PGPU_OK;
}

int po_sort(po_ctx* w) { (void)w; return PGPU_OK; }

/* ref: src/collisions.f90:1606-1892 MCC_COLLISIONS_VAHEDI */
int po_mcc_vahedi(po_ctx* w, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!w->mcc_ncomp) FAIL(w, PGPU_ERR_STATE, "po_mcc_vahedi: background mixture not set");
    w->timestep_coll = 0;
    w->timestep_reac = 0;
    for (int jr = 0; jr < w->nr; jr++) w->reac[jr].counts = 0;
    double dt = w->c.dt;
    double null_coll_freq = w->c.mcc_bg_dens * (double)1e-12f; /* single-precision literal (:1637, App. B2) */
    double p_null = 1 - exp(-dt * null_coll_freq);
    if (p_null > 1) { p_null = 1.; null_coll_freq = 1. / dt; }

    int64_t np_initial = w->np;
    for (int64_t jp1 = 0; jp1 < np_initial; jp1++) {
        po_rng_stream(w, 3, tID, (uint64_t)jp1, 0);
        int has_reacted = 0;
        if (po_rf(w) > p_null) continue;
        w->timestep_coll++;
        int s1 = w->p[jp1].S_ID;
        for (int j = 0; j < w->mcc_ncomp; j++) {
            int s2 = w->mcc_sp[j];
            double bg_nrho = w->mcc_frac[j] * w->c.mcc_bg_dens;
            if (bg_nrho == 0) continue;
            double c2[3], c1[3];
            po_maxwell(w, 0., 0., 0., w->c.mcc_bg_ttra, w->c.mcc_bg_ttra, w->c.mcc_bg_ttra, w->mass[s2 - 1], c2);
            part_t* q1 = &w->p[jp1];
            double vr2 = (c2[0] - q1->VX) * (c2[0] - q1->VX) + (c2[1] - q1->VY) * (c2[1] - q1->VY) +
                         (c2[2] - q1->VZ) * (c2[2] - q1->VZ);
            double vr = sqrt(vr2);
            double r_select = po_rf(w);
            c1[0] = q1->VX; c1[1] = q1->VY; c1[2] = q1->VZ;
            double m1 = w->mass[s1 - 1], m2 = w->mass[s2 - 1];
            double mred = m1 * m2 / (m1 + m2);
            double etr = 0.5 * mred * vr2;
            double p_cum = 0;
            for (int jr = 0; jr < w->nr; jr++) {
                reaction_t* r = &w->reac[jr];
                if (r->r1 != s1 || r->r2 != s2) continue;
                double ea = r->ea;
                if (etr <= ea) continue;
                if (r->type == PGPU_LXCAT) {
                    double sigma_r = po_interp_cs(etr, r->tab_n, r->tab_e, r->tab_cs);
                    p_cum = p_cum + (1. - exp(-bg_nrho * sigma_r * vr * dt)) / p_null;
                } else {
                    continue;
                }
                if (r_select < p_cum) {
                    r->counts++;
                    double erot = po_internal_energy(w, w->rotdof[s2 - 1], w->c.mcc_bg_ttra);
                    double evib = po_internal_energy(w, w->vibdof[s2 - 1], w->c.mcc_bg_ttra);
                    add_particle(w, q1->X, q1->Y, q1->Z, c2[0], c2[1], c2[2], erot, evib, s2, q1->IC, dt);
                    int64_t jp2 = w->np - 1;
                    w->timestep_reac++;
                    double ecoll = etr - ea;
                    w->p[jp1].S_ID = r->p1;
                    w->p[jp2].S_ID = r->p2;
                    if (!r->is_cex) react_redistribute(w, r, jp1, jp2, ecoll, c1, c2, 1);
                    has_reacted = 1;
                    break;
                }
            }
            if (has_reacted) break;
        }
    }
    if (timestep_coll) *timestep_coll = w->timestep_coll;
    if (timestep_reac) *timestep_reac = w->timestep_reac;
    return PGPU_OK;
}

/* ref: src/collisions.f90:1261-1595 MCC_COLLISIONS: every particle present at entry tests one collision per
 * background component against a partner sampled from the background Maxwellian; an accepted partner becomes
 * a real particle (appended) and the pair goes through the same reaction / Borgnakke-Larsen / VSS code as a
 * DSMC pair.  Philox mode: one stream per particle, kind 4, keyed by the particle's index. */
int po_mcc_collisions(po_ctx* w, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!w->mcc_ncomp) FAIL(w, PGPU_ERR_STATE, "po_mcc_collisions: background mixture not set");
    if (!w->sigmas) FAIL(w, PGPU_ERR_STATE, "po_mcc_collisions: VSS tables not set");
    w->timestep_coll = 0;
    w->timestep_reac = 0;
    for (int jr = 0; jr < w->nr; jr++) w->reac[jr].counts = 0;
    double dt = w->c.dt, ttra = w->c.mcc_bg_ttra;
    int64_t np_initial = w->np; /* ref: :1289-1290 */
    for (int64_t jp1 = 0; jp1 < np_initial; jp1++) {
        po_rng_stream(w, 4, tID, (uint64_t)jp1, 0);
        int s1 = w->p[jp1].S_ID; /* ref: :1291, NOT refreshed after a reaction */
        for (int j = 0; j < w->mcc_ncomp; j++) {
            int s2 = w->mcc_sp[j];
            double bg_nrho;
            if (w->mcc_cell_nrho) bg_nrho = w->mcc_cell_nrho[(s2 - 1) + (size_t)w->ns * (w->p[jp1].IC - 1)]; /* :1296-1297 */
            else bg_nrho = w->mcc_frac[j] * w->c.mcc_bg_dens;                                              /* :1299-1300 */
            if (bg_nrho == 0) continue;
            double cref = VSS2(w->grefs, s1, s2), sigma = VSS2(w->sigmas, s1, s2), omega = VSS2(w->omegas, s1, s2);
            double c2[3];
            po_maxwell(w, 0., 0., 0., ttra, ttra, ttra, w->mass[s2 - 1], c2); /* :1314-1316 */
            part_t* q1 = &w->p[jp1];
            double vr2 = (c2[0] - q1->VX) * (c2[0] - q1->VX) + (c2[1] - q1->VY) * (c2[1] - q1->VY) +
                         (c2[2] - q1->VZ) * (c2[2] - q1->VZ);
            double vr = sqrt(vr2);
            double p_coll = 1 - exp(-dt * bg_nrho * vr * sigma * pow(vr / cref, 1. - 2. * omega)); /* :1327 */
            if (!(po_rf(w) < p_coll)) continue;
            w->timestep_coll++;
            /* the partner becomes a particle at the position of JP1, ref: :1344-1351 */
            double erot = po_internal_energy(w, w->rotdof[s2 - 1], ttra);
            double evib = po_internal_energy(w, w->vibdof[s2 - 1], ttra);
            add_particle(w, q1->X, q1->Y, q1->Z, c2[0], c2[1], c2[2], erot, evib, s2, q1->IC, dt);
            int64_t jp2 = w->np - 1;
            accepted_pair(w, jp1, jp2, s1, s2, vr, vr2, 1);
        }
    }
    if (timestep_coll) *timestep_coll = w->timestep_coll;
    if (timestep_reac) *timestep_reac = w->timestep_reac;
    return PGPU_OK;
}

/* ref: src/timecycle.f90:2340-2353 REMOVE_PARTICLES_OF_SPECIES */
int po_remove_species(po_ctx* w, int32_t sp_id) {
    for (int64_t ip = w->np - 1; ip >= 0; ip--)
        if (w->p[ip].S_ID == sp_id) remove_particle(w, ip);
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * DEPOSIT_CHARGE
 * ---------------------------------------------------------------------------------------- */
static int nnodes_of(const po_ctx* w) {
    if (w->c.grid_type == PGPU_UNSTRUCTURED) return w->nn;
    return (w->c.dims == 1) ? w->c.nx + 1 : (w->c.nx + 1) * (w->c.ny + 1);
}

typedef struct { long double s, c; } kahan_t;
static void kadd(kahan_t* k, long double x) { /* Neumaier */
    long double t = k->s + x;
    if (fabsl(k->s) >= fabsl(x)) k->c += (k->s - t) + x;
    else k->c += (x - t) + k->s;
    k->s = t;
}

/* ref: src/fields.f90:5440-5535 (particle loop + SURFACE_CHARGE; the MPI reduction is the caller's) */
static int deposit_impl(po_ctx* w, double* rhs, double* abs_sum, int64_t nnodes, int compensated) {
    int nn = nnodes_of(w);
    if (nnodes < nn) FAIL(w, PGPU_ERR_INVALID, "po_deposit_charge: rhs too small (%lld < %d)", (long long)nnodes, nn);
    kahan_t* acc = compensated ? (kahan_t*)calloc((size_t)nn, sizeof(kahan_t)) : NULL;
    for (int i = 0; i < nn; i++) rhs[i] = 0.0;
    if (abs_sum) for (int i = 0; i < nn; i++) abs_sum[i] = 0.0;
    double K = QE / (EPS0 * (w->c.eps_scaling * w->c.eps_scaling));
#define ADD(node, val)                               \
    do {                                             \
        if (compensated) kadd(&acc[node], (val));    \
        else rhs[node] = rhs[node] + (val);          \
        if (abs_sum) abs_sum[node] += fabs(val);     \
    } while (0)
    for (int64_t jp = 0; jp < w->np; jp++) {
        const part_t* q = &w->p[jp];
        double charge = w->charge[q->S_ID - 1];
        if (fabs(charge) < 1.e-6) continue;
        if (w->c.grid_type == PGPU_UNSTRUCTURED) {
            int ic = q->IC;
            if (w->c.dims == 1) {
                double rho_q = K * charge * w->c.fnum / (w->c.ymax - w->c.ymin) / (w->c.zmax - w->c.zmin);
                for (int p = 1; p <= 2; p++) {
                    int vp = CELL_NODE(w, p, ic) - 1;
                    double psip = BASIS(w, 1, p, ic) * q->X + BASIS(w, 2, p, ic);
                    ADD(vp, rho_q * psip);
                }
            } else if (w->c.dims == 2) {
                double rho_q = K * charge * w->c.fnum / (w->c.zmax - w->c.zmin);
                for (int p = 1; p <= 3; p++) {
                    int vp = CELL_NODE(w, p, ic) - 1;
                    double psip = BASIS(w, 1, p, ic) * q->X + BASIS(w, 2, p, ic) * q->Y + BASIS(w, 3, p, ic);
                    ADD(vp, rho_q * psip);
                }
            } else {
                double rho_q = K * charge * w->c.fnum;
                for (int p = 1; p <= 4; p++) {
                    int vp = CELL_NODE(w, p, ic) - 1;
                    double psip = BASIS(w, 1, p, ic) * q->X + BASIS(w, 2, p, ic) * q->Y +
                                  BASIS(w, 3, p, ic) * q->Z + BASIS(w, 4, p, ic);
                    ADD(vp, rho_q * psip);
                }
            }
        } else {
            double wt[4];
            int idx[4], indi[4], indj[4];
            compute_weights(w, q, wt, idx, indi, indj);
            double vol = (w->c.xmax - w->c.xmin) * (w->c.ymax - w->c.ymin) * (w->c.zmax - w->c.zmin) /
                         (w->c.nx * w->c.ny); /* CELL_VOL, initialization.f90:2299 */
            double cfnum = w->cell_fnum ? w->cell_fnum[q->IC - 1] : w->c.fnum; /* ref: src/fields.f90:5507-5508 */
            double rho_q = -K * charge * cfnum / vol;
            int nw = (w->c.dims == 2) ? 4 : 2;
            for (int k = 0; k < nw; k++) ADD(idx[k], rho_q * wt[k]);
        }
    }
    if (w->surface_charge)
        for (int i = 0; i < nn && i < w->nsc; i++) ADD(i, w->surface_charge[i]);
#undef ADD
    if (compensated) {
        for (int i = 0; i < nn; i++) rhs[i] = (double)(acc[i].s + acc[i].c);
        free(acc);
    }
    return PGPU_OK;
}
int po_deposit_charge(po_ctx* w, double* rhs, int64_t nnodes) { return deposit_impl(w, rhs, NULL, nnodes, 0); }
int po_deposit_charge_compensated(po_ctx* w, double* rhs, double* abs_sum, int64_t nnodes) {
    return deposit_impl(w, rhs, abs_sum, nnodes, 1);
}

/* ------------------------------------------------------------------------------------------
 * EXCHANGE (local half). ref: src/grid_and_partition.f90:525-588 + COUNTING_SORT :482-518
 * ---------------------------------------------------------------------------------------- */
int po_exchange_pack(po_ctx* w, void* sendbuf, int64_t capacity, int64_t* sendcount) {
    int nr = w->c.nranks > 0 ? w->c.nranks : 1;
    for (int i = 0; i < nr; i++) sendcount[i] = 0;
    if (nr == 1) return PGPU_OK;
    if (!w->cell_procs) FAIL(w, PGPU_ERR_STATE, "po_exchange_pack: CELL_PROCS not set");
    part_t* tmp = (part_t*)malloc(sizeof(part_t) * (size_t)(w->np > 0 ? w->np : 1));
    int64_t jp = 0;
    for (int64_t ip = w->np - 1; ip >= 0; ip--) {
        int jproc = w->cell_procs[w->p[ip].IC - 1];
        if (jproc != w->c.rank) {
            sendcount[jproc]++;
            tmp[jp++] = w->p[ip];
            remove_particle(w, ip);
        }
    }
    if (jp > capacity) { free(tmp); FAIL(w, PGPU_ERR_INVALID, "po_exchange_pack: send buffer too small"); }
    /* COUNTING_SORT: inclusive displacements, filled from the back */
    int64_t* disp = (int64_t*)malloc(sizeof(int64_t) * (size_t)nr);
    disp[0] = sendcount[0];
    for (int i = 1; i < nr; i++) disp[i] = sendcount[i] + disp[i - 1];
    part_t* out = (part_t*)sendbuf;
    for (int64_t i = jp - 1; i >= 0; i--) {
        int iproc = w->cell_procs[tmp[i].IC - 1];
        out[disp[iproc] - 1] = tmp[i];
        disp[iproc]--;
    }
    free(disp);
    free(tmp);
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * DUMP_PARTICLES_FILE / READ_PARTICLES_FILE, binary flavour. ref: src/tools.f90:327-365, 465-517.
 * gfortran SEQUENTIAL UNFORMATTED with CONVERT='BIG_ENDIAN': every WRITE is one record framed by its byte length as
 * a 4-byte integer before and after, markers and payload all byte-swapped.
 * ---------------------------------------------------------------------------------------- */
static void put_be(unsigned char* dst, const void* src, int nbytes) {
    const unsigned char* s = (const unsigned char*)src;
    for (int i = 0; i < nbytes; i++) dst[i] = s[nbytes - 1 - i];
}
static void dump_file_name(const po_ctx* w, char* out, size_t n, const char* prefix, int timestep) {
    snprintf(out, n, "%sproc_%05d_time_%08d", prefix ? prefix : "", w->c.rank, timestep); /* "(A,A,I0.5,A6,I0.8)" */
}
int po_dump_particles_file(po_ctx* w, const char* path_prefix, int32_t timestep, int32_t binary) {
    if (!binary) FAIL(w, PGPU_ERR_UNSUPPORTED, "po_dump_particles_file: binary flavour only");
    char name[1024];
    dump_file_name(w, name, sizeof(name), path_prefix, timestep);
    FILE* probe = fopen(name, "rb");
    if (probe) { fclose(probe); FAIL(w, PGPU_ERR_INVALID, "po_dump_particles_file: %.200s exists (STATUS='NEW')", name); }
    FILE* fp = fopen(name, "wb");
    if (!fp) FAIL(w, PGPU_ERR_INVALID, "po_dump_particles_file: cannot create %.200s", name);
    for (int64_t ip = 0; ip < w->np; ip++) {
        const part_t* q = &w->p[ip];
        unsigned char rec[88];
        const int32_t len = 80;
        const double f[8] = {q->X, q->Y, q->Z, q->VX, q->VY, q->VZ, q->EROT, q->EVIB};
        put_be(rec, &len, 4);
        for (int k = 0; k < 8; k++) put_be(rec + 4 + 8 * k, &f[k], 8);
        put_be(rec + 68, &q->S_ID, 4);
        put_be(rec + 72, &q->IC, 4);
        put_be(rec + 76, &q->DTRIM, 8);
        put_be(rec + 84, &len, 4);
        if (fwrite(rec, 1, 88, fp) != 88) { fclose(fp); FAIL(w, PGPU_ERR_INVALID, "po_dump_particles_file: short write"); }
    }
    fclose(fp);
    return PGPU_OK;
}
int po_read_particles_file(po_ctx* w, const char* path_prefix, int32_t timestep, int32_t binary) {
    if (!binary) FAIL(w, PGPU_ERR_UNSUPPORTED, "po_read_particles_file: binary flavour only");
    char name[1024];
    dump_file_name(w, name, sizeof(name), path_prefix, timestep);
    FILE* fp = fopen(name, "rb");
    if (!fp) FAIL(w, PGPU_ERR_INVALID, "po_read_particles_file: particle restart file %.200s not found", name);
    unsigned char rec[88];
    while (fread(rec, 1, 88, fp) == 88) {
        pgpu_particle q;
        memset(&q, 0, sizeof(q));
        double f[8];
        for (int k = 0; k < 8; k++) put_be((unsigned char*)&f[k], rec + 4 + 8 * k, 8);
        int32_t sid, ic;
        put_be((unsigned char*)&sid, rec + 68, 4);
        put_be((unsigned char*)&ic, rec + 72, 4);
        q.X = f[0]; q.Y = f[1]; q.Z = f[2]; q.VX = f[3]; q.VY = f[4]; q.VZ = f[5]; q.EROT = f[6]; q.EVIB = f[7];
        q.S_ID = sid; q.IC = ic;
        q.DTRIM = w->c.dt; /* INIT_PARTICLE(..., DT, ...), ref: src/tools.f90:497 */
        q.ID = (int64_t)w->c.rank + (w->id_counter << 8);
        int rc = po_append_particles(w, &q, 1, (int32_t)sizeof(q));
        if (rc) { fclose(fp); return rc; }
        w->id_counter++;
    }
    fclose(fp);
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * REASSIGN_PARTICLES_TO_CELLS_1D / _2D / _3D, ref: src/tools.f90:925-1039, 782-922, 577-778 (called after
 * READ_PARTICLES_FILE on a restart, src/pantera.f90:58-70).  One body for the three copies: the domain box cut into
 * N_PARTITIONS = 100 bins per direction, every cell entered (ascending IC) in the bins its bounding box overlaps, the
 * first cell of a particle's bin with all PSIP >= 0 wins; removal by the reverse swap-with-last sweep.
 * Bin indices are clamped to 1..100 where the Fortran would index past its arrays (see pgpu_restart.cu).
 * ---------------------------------------------------------------------------------------- */
static int rb_index(double c, double lo, double h, int n) {
    /* INT((C - XMIN)/DX) + 1 with INT() truncating toward zero, clamped to 1..n */
    const double q = (c - lo) / h;
    if (!(q < (double)n)) return n;
    if (q < 1.0) return 1;
    return (int)q + 1;
}
int po_reassign_particles_to_cells(po_ctx* w, int64_t* n_removed) {
    if (n_removed) *n_removed = 0;
    if (w->c.grid_type != PGPU_UNSTRUCTURED || !w->basis_coeffs) FAIL(w, PGPU_ERR_UNSUPPORTED, "po_reassign_particles_to_cells: Not implemented!");
    const int dims = w->dims_mesh, nf = dims + 1, NPART = 100;
    const double lo[3] = {w->c.xmin, w->c.ymin, w->c.zmin}, hi[3] = {w->c.xmax, w->c.ymax, w->c.zmax};
    int n[3] = {1, 1, 1};
    double h[3] = {1, 1, 1};
    for (int k = 0; k < dims; k++) {
        n[k] = NPART;
        h[k] = (hi[k] - lo[k]) / NPART;
        if (!(h[k] > 0.)) FAIL(w, PGPU_ERR_INVALID, "po_reassign_particles_to_cells: empty Domain_limits");
    }
    const size_t nbins = (size_t)n[0] * n[1] * n[2];
    int* ningrid = (int*)calloc(nbins, sizeof(int));
    int* iof = (int*)calloc(nbins + 1, sizeof(int));
    int* ind = NULL;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int idx = 0; /* I outermost, K innermost, :655-666 */
            for (size_t b = 0; b < nbins; b++) { iof[b] = idx; idx += ningrid[b]; }
            iof[nbins] = idx;
            ind = (int*)calloc((size_t)(idx > 0 ? idx : 1), sizeof(int));
            memset(ningrid, 0, nbins * sizeof(int));
        }
        for (int ic = 1; ic <= w->nc; ic++) {
            int imin[3] = {1, 1, 1}, imax[3] = {1, 1, 1};
            for (int k = 0; k < dims; k++) {
                double cmin = 0, cmax = 0;
                for (int p = 1; p <= nf; p++) {
                    double x = NODE_X(w, k + 1, CELL_NODE(w, p, ic));
                    if (p == 1 || x < cmin) cmin = x;
                    if (p == 1 || x > cmax) cmax = x;
                }
                imin[k] = rb_index(cmin, lo[k], h[k], n[k]);
                imax[k] = rb_index(cmax, lo[k], h[k], n[k]);
            }
            for (int i = imin[0]; i <= imax[0]; i++)
                for (int j = imin[1]; j <= imax[1]; j++)
                    for (int k = imin[2]; k <= imax[2]; k++) {
                        size_t b = ((size_t)(i - 1) * n[1] + (j - 1)) * n[2] + (k - 1);
                        if (pass == 1) ind[iof[b] + ningrid[b]] = ic;
                        ningrid[b]++;
                    }
        }
    }
    char* remove_part = (char*)calloc((size_t)(w->np > 0 ? w->np : 1), 1);
    for (int64_t ip = 0; ip < w->np; ip++) {
        part_t* q = &w->p[ip];
        const double xp[3] = {q->X, q->Y, q->Z};
        int outside = 0;
        for (int k = 0; k < dims; k++)
            if (xp[k] < lo[k] || xp[k] > hi[k]) outside = 1;
        if (outside) { remove_part[ip] = 1; continue; }
        int idx[3] = {1, 1, 1};
        for (int k = 0; k < dims; k++) idx[k] = rb_index(xp[k], lo[k], h[k], n[k]);
        size_t b = ((size_t)(idx[0] - 1) * n[1] + (idx[1] - 1)) * n[2] + (idx[2] - 1);
        int cellfound = 0;
        for (int e = iof[b]; e < iof[b] + ningrid[b]; e++) {
            int ic = ind[e], inside = 1;
            for (int vp = 1; vp <= nf; vp++) {
                double psip;
                if (dims == 3) psip = xp[0] * BASIS(w, 1, vp, ic) + xp[1] * BASIS(w, 2, vp, ic) + xp[2] * BASIS(w, 3, vp, ic) + BASIS(w, 4, vp, ic);
                else if (dims == 2) psip = xp[0] * BASIS(w, 1, vp, ic) + xp[1] * BASIS(w, 2, vp, ic) + BASIS(w, 3, vp, ic);
                else psip = xp[0] * BASIS(w, 1, vp, ic) + BASIS(w, 2, vp, ic);
                if (psip < 0) inside = 0;
            }
            if (inside) { q->IC = ic; cellfound = 1; break; }
        }
        if (!cellfound) remove_part[ip] = 1;
    }
    int64_t removed = 0;
    for (int64_t ip = w->np - 1; ip >= 0; ip--)
        if (remove_part[ip]) { remove_particle(w, ip); removed++; }
    free(remove_part); free(ningrid); free(iof); free(ind);
    if (n_removed) *n_removed = removed;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * INITIAL_SEED (src/initialization.f90:2017-2233), VOLUME_INJECT (src/timecycle.f90:553-737), BOUNDARIES_EMIT
 * (src/timecycle.f90:368-547); RANDINT src/tools.f90:1304-1311.  COLOCATED_ELECTRONS, BOOL_INJECT_FROM_FILE and the
 * kappa distribution are outside the scope.
 * MT mode: the draws come from the rank's sequential stream in the reference's loop order.  Philox mode (device
 * contract): one stream per (cell | emit task, task, component) -- kind 5 INITIAL_SEED, 6 VOLUME_INJECT, 7
 * BOUNDARIES_EMIT; e0 = cell (0-based) or emit task; e1 = task * 256 + component (kinds 5, 6), component for the count of
 * an emit task and component + 256 (1 + ip) for its particle ip (kind 7) -- whose draws are consumed in the
 * reference's order for that loop body: the RANDINT draw, then per particle position, MAXWELL, INTERNAL_ENERGY (x2).
 * On rectilinear grids the whole domain is one loop; there the count comes from stream (e0 = 0) and particle ip draws
 * from stream (e0 = 1 + ip).
 * ---------------------------------------------------------------------------------------- */
static int randint(po_ctx* w, double x) {
    int r = (int)x;
    if (x - r > po_rf(w)) r = r + 1;
    return r;
}
static const int FACE_IND[4][3] = {{1, 3, 2}, {1, 2, 4}, {2, 3, 4}, {1, 4, 3}}; /* src/grid_and_partition.f90:1932-1935 */

/* kind: 5 seed (rate = NRHO, owned fluid cells only, no division by the rank count on unstructured grids),
 *       6 volume inject (rate = DT*NRHODOT, all cells, divided by N_MPI_THREADS) */
static int cell_source(po_ctx* w, int kind, int tID, int ntask, const double* par, const int* mixid) {
    const int nranks = w->c.nranks > 0 ? w->c.nranks : 1;
    for (int it = 0; it < ntask; it++) {
        const double* tp = par + 9 * it;
        const int mix = mixid[it];
        for (int i = w->mix_off[mix]; i < w->mix_off[mix + 1]; i++) {
            const int icomp = i - w->mix_off[mix];
            const int s_id = w->mix_sp[i];
            const double frac = w->mix_frac[i], m = w->mass[s_id - 1];
            const double spwt = w->spwt ? w->spwt[s_id - 1] : 1.0;
            const double rate = (kind == 5) ? tp[0] : w->c.dt * tp[0];
            if (w->c.grid_type == PGPU_UNSTRUCTURED) {
                const int d = w->dims_mesh;
                for (int ic = 1; ic <= w->nc; ic++) {
                    if (kind == 5) {
                        if (w->cell_procs && w->cell_procs[ic - 1] != w->c.rank) continue;
                        if (w->cell_pg && w->cell_pg[ic - 1] != -1 && w->bc_volume && w->bc_volume[w->cell_pg[ic - 1] - 1] == PGPU_SOLID) continue;
                    }
                    po_rng_stream(w, kind, tID, (uint64_t)(ic - 1), (uint64_t)(it * 256 + icomp));
                    const double vol = w->cell_volumes[ic - 1];
                    int np_init;
                    if (kind == 5) np_init = randint(w, rate / (w->c.fnum * spwt) * vol * frac);
                    else np_init = randint(w, rate / (w->c.fnum * spwt) * vol * frac / nranks);
                    if (np_init == 0) continue;
                    double V[4][3] = {{0}};
                    for (int k = 0; k <= d; k++)
                        for (int a = 0; a < 3; a++) V[k][a] = w->node_coords[3 * (w->cell_nodes[(d + 1) * (ic - 1) + k] - 1) + a];
                    for (int ip = 0; ip < np_init; ip++) {
                        double S = po_rf(w), T = po_rf(w), U = po_rf(w), xp, yp, zp;
                        if (S + T > 1) { S = 1 - S; T = 1 - T; }
                        if (d == 1) {
                            xp = V[0][0] + (V[1][0] - V[0][0]) * S;
                            yp = w->c.ymin + (w->c.ymax - w->c.ymin) * T;
                            zp = w->c.zmin + (w->c.zmax - w->c.zmin) * U;
                        } else if (d == 2) {
                            xp = V[0][0] + (V[1][0] - V[0][0]) * S + (V[2][0] - V[0][0]) * T;
                            yp = V[0][1] + (V[1][1] - V[0][1]) * S + (V[2][1] - V[0][1]) * T;
                            zp = w->c.axi ? 0.0 : w->c.zmin + (w->c.zmax - w->c.zmin) * U;
                        } else {
                            double P, Q, R;
                            if (S + T + U <= 1) { P = S; Q = T; R = U; }
                            else if (T + U > 1) { P = S; Q = 1 - U; R = 1 - S - T; }
                            else { P = 1 - T - U; Q = T; R = S + T + U - 1; }
                            xp = V[0][0] + (V[1][0] - V[0][0]) * P + (V[2][0] - V[0][0]) * Q + (V[3][0] - V[0][0]) * R;
                            yp = V[0][1] + (V[1][1] - V[0][1]) * P + (V[2][1] - V[0][1]) * Q + (V[3][1] - V[0][1]) * R;
                            zp = V[0][2] + (V[1][2] - V[0][2]) * P + (V[2][2] - V[0][2]) * Q + (V[3][2] - V[0][2]) * R;
                        }
                        double v3[3];
                        po_maxwell(w, tp[1], tp[2], tp[3], tp[4], tp[5], tp[6], m, v3);
                        double erot = po_internal_energy(w, w->rotdof[s_id - 1], tp[7]);
                        double evib = po_internal_energy(w, w->vibdof[s_id - 1], tp[8]);
                        add_particle(w, xp, yp, zp, v3[0], v3[1], v3[2], erot, evib, s_id, ic, w->c.dt);
                    }
                }
            } else {
                double domain_volume;
                if (w->c.axi) domain_volume = 0.5 * (w->c.xmax - w->c.xmin) * (w->c.ymax * w->c.ymax - w->c.ymin * w->c.ymin) * (w->c.zmax - w->c.zmin);
                else domain_volume = (w->c.xmax - w->c.xmin) * (w->c.ymax - w->c.ymin) * (w->c.zmax - w->c.zmin);
                po_rng_stream(w, kind, tID, 0ULL, (uint64_t)(it * 256 + icomp));
                int np_init = randint(w, rate / (w->c.fnum * spwt) * domain_volume * frac / nranks);
                for (int ip = 0; ip < np_init; ip++) {
                    po_rng_stream(w, kind, tID, (uint64_t)(1 + ip), (uint64_t)(it * 256 + icomp));
                    double xp = w->c.xmin + (w->c.xmax - w->c.xmin) * po_rf(w), yp, zp;
                    if (w->c.axi && !w->cell_fnum) { /* IF (AXI .AND. (.NOT. BOOL_RADIAL_WEIGHTING)) */
                        yp = sqrt(w->c.ymin * w->c.ymin + po_rf(w) * (w->c.ymax * w->c.ymax - w->c.ymin * w->c.ymin));
                        zp = 0;
                    } else {
                        yp = w->c.ymin + (w->c.ymax - w->c.ymin) * po_rf(w);
                        zp = w->c.zmin + (w->c.zmax - w->c.zmin) * po_rf(w);
                    }
                    double v3[3];
                    po_maxwell(w, tp[1], tp[2], tp[3], tp[4], tp[5], tp[6], m, v3);
                    double erot = po_internal_energy(w, w->rotdof[s_id - 1], tp[7]);
                    double evib = po_internal_energy(w, w->vibdof[s_id - 1], tp[8]);
                    int cid = cell_from_position(w, xp, yp);
                    add_particle(w, xp, yp, zp, v3[0], v3[1], v3[2], erot, evib, s_id, cid, w->c.dt);
                }
            }
        }
    }
    return PGPU_OK;
}
int po_initial_seed(po_ctx* w) {
    if (!w->mix_off) FAIL(w, PGPU_ERR_STATE, "po_initial_seed: mixtures not set");
    return cell_source(w, 5, 0, w->n_seed, w->seed_par, w->seed_mix);
}
int po_volume_inject(po_ctx* w, int32_t tID) {
    if (!w->mix_off) FAIL(w, PGPU_ERR_STATE, "po_volume_inject: mixtures not set");
    return cell_source(w, 6, tID, w->n_vinj, w->vinj_par, w->vinj_mix);
}
/* ref: src/timecycle.f90:368-547 */
int po_boundaries_emit(po_ctx* w, int32_t tID) {
    if (w->n_emit && !w->mix_off) FAIL(w, PGPU_ERR_STATE, "po_boundaries_emit: mixtures not set");
    const int d = w->dims_mesh;
    for (int it = 0; it < w->n_emit; it++) {
        const double* tp = w->emit_par + 7 * it;
        const int* ti = w->emit_int + 6 * it;
        const int mix = ti[0], ic = ti[1], iv1 = ti[2], iv2 = ti[3], iface = ti[4];
        double fn[3] = {0, 0, 0}, t1[3], t2[3], x1 = 0, y1 = 0, x2 = 0, y2 = 0, V1[3] = {0}, V2[3] = {0}, V3[3] = {0};
        const size_t f = (size_t)(d + 1) * (ic - 1) + (iface - 1);
        for (int a = 0; a < 3; a++) fn[a] = w->face_normal[3 * f + a];
        if (d == 1) {
            x1 = w->node_coords[3 * (iv1 - 1)]; x2 = w->node_coords[3 * (iv2 - 1)];
            t1[0] = 0; t1[1] = fn[0]; t1[2] = 0; t2[0] = 0; t2[1] = 0; t2[2] = 1;
        } else if (d == 2) {
            x1 = w->node_coords[3 * (iv1 - 1)]; y1 = w->node_coords[3 * (iv1 - 1) + 1];
            x2 = w->node_coords[3 * (iv2 - 1)]; y2 = w->node_coords[3 * (iv2 - 1) + 1];
            t1[0] = -fn[1]; t1[1] = fn[0]; t1[2] = 0; t2[0] = 0; t2[1] = 0; t2[2] = 1;
        } else {
            for (int a = 0; a < 3; a++) {
                t1[a] = w->face_tang1[3 * f + a];
                t2[a] = w->face_tang2[3 * f + a];
                V1[a] = w->node_coords[3 * (w->cell_nodes[4 * (ic - 1) + FACE_IND[iface - 1][0] - 1] - 1) + a];
                V2[a] = w->node_coords[3 * (w->cell_nodes[4 * (ic - 1) + FACE_IND[iface - 1][1] - 1] - 1) + a];
                V3[a] = w->node_coords[3 * (w->cell_nodes[4 * (ic - 1) + FACE_IND[iface - 1][2] - 1] - 1) + a];
            }
        }
        for (int i = w->mix_off[mix]; i < w->mix_off[mix + 1]; i++) {
            const int icomp = i - w->mix_off[mix];
            const int s_id = w->mix_sp[i];
            const double m = w->mass[s_id - 1];
            po_rng_stream(w, 7, tID, (uint64_t)it, (uint64_t)icomp);
            const double nfs_x = w->emit_nfs[ti[5] + icomp];
            int nfs = (int)floor(nfs_x);
            if (nfs_x - (double)nfs >= po_rf(w)) nfs = nfs + 1;
            double beta = 0;
            if (tp[4] != 0) beta = 1. / sqrt(2. * KB / m * tp[4]);
            for (int ip = 0; ip < nfs; ip++) {
                po_rng_stream(w, 7, tID, (uint64_t)it, (uint64_t)(icomp + ((ip + 1) << 8))); /* (device contract: a stream per particle) */
                double v3[3];
                po_maxwell(w, 0., 0., 0., tp[4], tp[4], tp[4], m, v3);
                double erot = po_internal_energy(w, w->rotdof[s_id - 1], tp[5]);
                double evib = po_internal_energy(w, w->vibdof[s_id - 1], tp[6]);
                double v_norm = 0;
                if (tp[4] != 0) v_norm = po_flx(w, tp[3] * beta, tp[4], m);
                double vx = tp[0] - v_norm * fn[0] - v3[1] * t1[0] - v3[2] * t2[0];
                double vy = tp[1] - v_norm * fn[1] - v3[1] * t1[1] - v3[2] * t2[1];
                double vz = tp[2] - v_norm * fn[2] - v3[1] * t1[2] - v3[2] * t2[2];
                double dtfrac = po_rf(w) * w->c.dt;
                double x, y, z;
                if (d == 1) {
                    double R = po_rf(w);
                    x = x1 + R * (x2 - x1);
                    y = w->c.ymin + (w->c.ymax - w->c.ymin) * po_rf(w);
                    z = w->c.zmin + (w->c.zmax - w->c.zmin) * po_rf(w);
                } else if (d == 2) {
                    double R = po_rf(w);
                    if (w->c.axi) {
                        if (y1 == y2) { y = y1; x = x1 + R * (x2 - x1); z = 0; }
                        else if (y1 < y2) { y = sqrt(y1 * y1 + R * (y2 * y2 - y1 * y1)); x = x1 + (y - y1) / (y2 - y1) * (x2 - x1); z = 0; }
                        else { y = sqrt(y2 * y2 + R * (y1 * y1 - y2 * y2)); x = x2 + (y - y2) / (y1 - y2) * (x1 - x2); z = 0; }
                    } else {
                        x = x1 + R * (x2 - x1);
                        y = y1 + R * (y2 - y1);
                        z = w->c.zmin + (w->c.zmax - w->c.zmin) * po_rf(w);
                    }
                } else {
                    double S = po_rf(w), T = po_rf(w), P, Q;
                    if (S > 1 - T) { Q = 1 - T; P = 1 - S; }
                    else { Q = T; P = S; }
                    x = V1[0] + (V2[0] - V1[0]) * P + (V3[0] - V1[0]) * Q;
                    y = V1[1] + (V2[1] - V1[1]) * P + (V3[1] - V1[1]) * Q;
                    z = V1[2] + (V2[2] - V1[2]) * P + (V3[2] - V1[2]) * Q;
                }
                add_particle(w, x, y, z, vx, vy, vz, erot, evib, s_id, ic, dtfrac);
            }
        }
    }
    return PGPU_OK;
}
int po_set_species_weights(po_ctx* w, int ns, const double* spwt) { REPL(w->spwt, spwt, ns); return PGPU_OK; }
int po_set_mixtures(po_ctx* w, int nmix, const int32_t* comp_off, const int32_t* comp_sp, const double* comp_molfrac) {
    REPL(w->mix_off, comp_off, nmix + 1);
    REPL(w->mix_sp, comp_sp, comp_off[nmix]);
    REPL(w->mix_frac, comp_molfrac, comp_off[nmix]);
    w->nmix = nmix;
    return PGPU_OK;
}
static int set_source_tasks(double** par, int** mix, int* count, int n, const int32_t* mix_id, const double* rate, const double* ux,
                            const double* uy, const double* uz, const double* tx, const double* ty, const double* tz,
                            const double* trot, const double* tvib) {
    free(*par); free(*mix);
    *par = (double*)calloc((size_t)(9 * (n > 0 ? n : 1)), sizeof(double));
    *mix = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++) {
        double* t = *par + 9 * i;
        t[0] = rate[i]; t[1] = ux[i]; t[2] = uy[i]; t[3] = uz[i]; t[4] = tx[i]; t[5] = ty[i]; t[6] = tz[i]; t[7] = trot[i]; t[8] = tvib[i];
        (*mix)[i] = mix_id[i];
    }
    *count = n;
    return PGPU_OK;
}
int po_set_initial_particles_tasks(po_ctx* w, int n, const int32_t* mix_id, const double* nrho, const double* ux, const double* uy,
                                   const double* uz, const double* ttrax, const double* ttray, const double* ttraz,
                                   const double* trot, const double* tvib) {
    return set_source_tasks(&w->seed_par, &w->seed_mix, &w->n_seed, n, mix_id, nrho, ux, uy, uz, ttrax, ttray, ttraz, trot, tvib);
}
int po_set_volume_inject_tasks(po_ctx* w, int n, const int32_t* mix_id, const double* nrhodot, const double* ux, const double* uy,
                               const double* uz, const double* ttrax, const double* ttray, const double* ttraz,
                               const double* trot, const double* tvib) {
    return set_source_tasks(&w->vinj_par, &w->vinj_mix, &w->n_vinj, n, mix_id, nrhodot, ux, uy, uz, ttrax, ttray, ttraz, trot, tvib);
}
int po_set_emit_tasks(po_ctx* w, int n, const int32_t* mix_id, const int32_t* ic, const int32_t* iv1, const int32_t* iv2,
                      const int32_t* iface, const int32_t* nfs_off, const double* nfs, const double* ux, const double* uy,
                      const double* uz, const double* u_norm, const double* ttra, const double* trot, const double* tvib) {
    free(w->emit_par); free(w->emit_int); free(w->emit_nfs);
    w->emit_par = (double*)calloc((size_t)(7 * (n > 0 ? n : 1)), sizeof(double));
    w->emit_int = (int*)calloc((size_t)(6 * (n > 0 ? n : 1)), sizeof(int));
    int total = n > 0 ? nfs_off[n] : 0;
    w->emit_nfs = (double*)dupmem(nfs, sizeof(double) * (size_t)(total > 0 ? total : 1));
    for (int i = 0; i < n; i++) {
        double* t = w->emit_par + 7 * i;
        t[0] = ux[i]; t[1] = uy[i]; t[2] = uz[i]; t[3] = u_norm[i]; t[4] = ttra[i]; t[5] = trot[i]; t[6] = tvib[i];
        int* k = w->emit_int + 6 * i;
        k[0] = mix_id[i]; k[1] = ic[i]; k[2] = iv1[i]; k[3] = iv2[i]; k[4] = iface[i]; k[5] = nfs_off[i];
    }
    w->n_emit = n;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * Field step: ASSEMBLE_POISSON for 1-D line meshes (src/fields.f90:137-253, 758-760), SET_WALL_POTENTIAL (:5566-5607),
 * the right-hand side rows of DEPOSIT_CHARGE (:5538-5561), COMPUTE_E_FIELD (:5237-5296).  The reference solves with
 * PETSc's KSP; here the assembled matrix (dense, MatSetValue semantics) is solved by Gaussian elimination with partial
 * pivoting -- deliberately not the device's tridiagonal sweep.
 * ---------------------------------------------------------------------------------------- */
int po_set_field_bcs(po_ctx* w, int nbc, const double* wall_potential, const double* wall_rf_potential, const double* rf_frequency,
                     const double* wall_efield, const double* eps_rel) {
    double** dst[5] = {&w->fb_pot, &w->fb_rfpot, &w->fb_freq, &w->fb_efield, &w->fb_eps};
    const double* src[5] = {wall_potential, wall_rf_potential, rf_frequency, wall_efield, eps_rel};
    for (int k = 0; k < 5; k++) {
        free(*dst[k]);
        *dst[k] = (double*)calloc((size_t)(nbc > 0 ? nbc : 1), sizeof(double));
        for (int i = 0; i < nbc; i++) (*dst[k])[i] = src[k] ? src[k][i] : (k == 4 ? 1.0 : 0.0);
    }
    return PGPU_OK;
}
static void compute_e_field(const po_ctx* w, const double* phi, double* e) {
    const int d = w->dims_mesh, nf = d + 1;
    for (int i = 0; i < w->nc; i++) {
        double ev[3] = {0, 0, 0};
        for (int id = 0; id < d; id++)
            for (int p = 0; p < nf; p++) {
                int vp = w->cell_nodes[nf * i + p];
                ev[id] = ev[id] - phi[vp - 1] * w->basis_coeffs[(size_t)nf * nf * i + nf * p + id];
            }
        e[3 * i] = ev[0]; e[3 * i + 1] = ev[1]; e[3 * i + 2] = ev[2];
    }
}
int po_compute_e_field(po_ctx* w, const double* phi, int64_t nnodes, double* e_out) {
    if (w->c.grid_type != PGPU_UNSTRUCTURED || !w->basis_coeffs) FAIL(w, PGPU_ERR_STATE, "po_compute_e_field: unstructured mesh with BASIS_COEFFS needed");
    if (nnodes < w->nn) FAIL(w, PGPU_ERR_INVALID, "po_compute_e_field: phi too short");
    double* e = (double*)calloc((size_t)(3 * w->nc), sizeof(double));
    compute_e_field(w, phi, e);
    REPL(w->e_field, e, 3 * w->nc);
    w->ne = 3 * w->nc;
    if (e_out) memcpy(e_out, e, sizeof(double) * (size_t)(3 * w->nc));
    free(e);
    return PGPU_OK;
}
/* The whole field step on any unstructured mesh.  One body for the three branches of ASSEMBLE_POISSON
 * (src/fields.f90:137-253 segments, :256-424 triangles, :425-637 tetrahedra; unused nodes :639-646; identity rows
 * :758-760), SET_WALL_POTENTIAL (:5566-5717) and the right-hand side rows of DEPOSIT_CHARGE (:5538-5547).  The matrix
 * is dense with MatSetValue(ADD_VALUES) semantics -- cells in ascending order -- and solved by Gaussian elimination with
 * partial pivoting: deliberately neither the device's tridiagonal sweep nor its conjugate gradient. */
static const int FACE_LOCAL_1D[2][3] = {{1, 0, 0}, {2, 0, 0}};
static const int FACE_LOCAL_2D[3][3] = {{1, 2, 0}, {2, 3, 0}, {3, 1, 0}};
static const int FACE_LOCAL_3D[4][3] = {{1, 3, 2}, {1, 2, 4}, {2, 3, 4}, {1, 4, 3}};
static int face_local(int d, int j, int t) { /* local node t (0-based) of face j (1-based), 1-based local number */
    return d == 1 ? FACE_LOCAL_1D[j - 1][t] : d == 2 ? FACE_LOCAL_2D[j - 1][t] : FACE_LOCAL_3D[j - 1][t];
}
static int field_solve(po_ctx* w, int32_t tID, const double* rhs, double* phi_out, double* e_out) {
    if (w->c.grid_type != PGPU_UNSTRUCTURED) FAIL(w, PGPU_ERR_UNSUPPORTED, "po_field_solve: unstructured meshes only");
    if (!rhs) FAIL(w, PGPU_ERR_INVALID, "po_field_solve: the oracle needs the RHS passed in");
    const int nn = w->nn, nc = w->nc, d = w->dims_mesh, nf = d + 1;
    double* A = (double*)calloc((size_t)nn * nn, sizeof(double));
    double* b = (double*)calloc((size_t)nn, sizeof(double));
    double* dir = (double*)calloc((size_t)nn, sizeof(double));
    double* neu = (double*)calloc((size_t)nn, sizeof(double));
    char* isd = (char*)calloc((size_t)nn, 1);
    char* isn = (char*)calloc((size_t)nn, 1);
    char* unused = (char*)malloc((size_t)nn);
    memset(unused, 1, (size_t)nn);
#define FB(arr, pg) ((arr) ? (arr)[(pg)-1] : 0.0)
    for (int i = 1; i <= nc; i++) { /* flags, with SET_WALL_POTENTIAL's time-dependent value */
        for (int p = 1; p <= nf; p++) unused[CELL_NODE(w, p, i) - 1] = 0;
        for (int j = 1; j <= nf; j++) {
            int pg = FACE_PG(w, j, i);
            if (pg == -1) continue;
            int fbc = w->bc_field[pg - 1];
            for (int t = 0; t < d; t++) {
                int v = CELL_NODE(w, face_local(d, j, t), i) - 1;
                if (fbc == PGPU_DIRICHLET_BC || fbc == PGPU_RF_VOLTAGE_BC || fbc == PGPU_DECOUPLED_RF_VOLTAGE_BC) {
                    double pot = FB(w->fb_pot, pg);
                    if (fbc != PGPU_DIRICHLET_BC) pot = FB(w->fb_pot, pg) + 0.5 * FB(w->fb_rfpot, pg) * cos(2 * PI * FB(w->fb_freq, pg) * tID * w->c.dt);
                    dir[v] = pot;
                    isd[v] = 1;
                }
                if (fbc == PGPU_NEUMANN_BC) isn[v] = 1;
            }
        }
    }
    for (int i = 1; i <= nc; i++) {
        double eps_rel = 1.0;
        if (w->cell_pg && w->cell_pg[i - 1] != -1 && w->fb_eps) eps_rel = w->fb_eps[w->cell_pg[i - 1] - 1];
        double K[4][4];
        if (d == 1) { /* :204-212 */
            double length = fabs(NODE_X(w, 1, CELL_NODE(w, 2, i)) - NODE_X(w, 1, CELL_NODE(w, 1, i))); /* SEGMENT_LENGTHS */
            K[0][0] = 1.0 / length * eps_rel; K[1][1] = 1.0 / length * eps_rel; K[0][1] = K[1][0] = -1.0 / length * eps_rel;
        } else if (d == 2) { /* :332-361 */
            int v1 = CELL_NODE(w, 1, i), v2 = CELL_NODE(w, 2, i), v3 = CELL_NODE(w, 3, i);
            double x1 = NODE_X(w, 1, v1), x2 = NODE_X(w, 1, v2), x3 = NODE_X(w, 1, v3);
            double y1 = NODE_X(w, 2, v1), y2 = NODE_X(w, 2, v2), y3 = NODE_X(w, 2, v3);
            double area = 0.5 * fabs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)); /* CELL_AREAS, grid_and_partition.f90:1315 */
            double k11 = 0.25 * ((y2 - y3) * (y2 - y3) + (x2 - x3) * (x2 - x3)) / area * eps_rel;
            double k22 = 0.25 * ((y1 - y3) * (y1 - y3) + (x1 - x3) * (x1 - x3)) / area * eps_rel;
            double k33 = 0.25 * ((y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1)) / area * eps_rel;
            double k12 = -(0.25 * ((y2 - y3) * (y1 - y3) + (x2 - x3) * (x1 - x3)) / area * eps_rel);
            double k23 = 0.25 * ((y1 - y3) * (y2 - y1) + (x1 - x3) * (x2 - x1)) / area * eps_rel;
            double k13 = -(0.25 * ((y2 - y3) * (y2 - y1) + (x2 - x3) * (x2 - x1)) / area * eps_rel);
            if (w->c.axi) {
                k11 = k11 * (y1 + y2 + y3) / 3.; k22 = k22 * (y1 + y2 + y3) / 3.; k33 = k33 * (y1 + y2 + y3) / 3.;
                k12 = k12 * (y1 + y2 + y3) / 3.; k23 = k23 * (y1 + y2 + y3) / 3.; k13 = k13 * (y1 + y2 + y3) / 3.;
            }
            K[0][0] = k11; K[1][1] = k22; K[2][2] = k33;
            K[0][1] = K[1][0] = k12; K[1][2] = K[2][1] = k23; K[0][2] = K[2][0] = k13;
        } else { /* :560-577 */
            double volume = w->cell_volumes[i - 1];
            for (int p = 1; p <= 4; p++)
                for (int q = 1; q <= 4; q++)
                    K[p - 1][q - 1] = volume * (BASIS(w, 1, p, i) * BASIS(w, 1, q, i) + BASIS(w, 2, p, i) * BASIS(w, 2, q, i) + BASIS(w, 3, p, i) * BASIS(w, 3, q, i)) * eps_rel;
        }
        for (int p = 1; p <= nf; p++) {
            int vp = CELL_NODE(w, p, i) - 1;
            if (isd[vp]) continue;
            for (int q = 1; q <= nf; q++) A[(size_t)vp * nn + (CELL_NODE(w, q, i) - 1)] += K[p - 1][q - 1];
        }
        for (int j = 1; j <= nf; j++) { /* Neumann faces: :236-251, :391-421, :579-633 */
            int pg = FACE_PG(w, j, i);
            if (pg == -1 || w->bc_field[pg - 1] != PGPU_NEUMANN_BC) continue;
            double add;
            if (d == 1) add = FB(w->fb_efield, pg);
            else if (d == 2) {
                int va = CELL_NODE(w, face_local(2, j, 0), i), vb = CELL_NODE(w, face_local(2, j, 1), i);
                double dx = NODE_X(w, 1, vb) - NODE_X(w, 1, va), dy = NODE_X(w, 2, vb) - NODE_X(w, 2, va);
                double edgelength = sqrt(dx * dx + dy * dy);
                add = FB(w->fb_efield, pg) * 0.5 * edgelength;
            } else {
                int va = CELL_NODE(w, face_local(3, j, 0), i), vb = CELL_NODE(w, face_local(3, j, 1), i), vc = CELL_NODE(w, face_local(3, j, 2), i);
                double u[3], t[3], cr[3];
                for (int k = 1; k <= 3; k++) { u[k - 1] = NODE_X(w, k, vb) - NODE_X(w, k, va); t[k - 1] = NODE_X(w, k, vc) - NODE_X(w, k, va); }
                cross3(u, t, cr);
                double facearea = 0.5 * norm2_3(cr); /* FACE_AREA, grid_and_partition.f90:1957 */
                add = FB(w->fb_efield, pg) * facearea / 3.;
            }
            for (int t = 0; t < d; t++) {
                int v = CELL_NODE(w, face_local(d, j, t), i) - 1;
                if (!isd[v]) neu[v] = neu[v] + add;
            }
        }
    }
#undef FB
    for (int v = 0; v < nn; v++)
        if (unused[v]) { isd[v] = 1; dir[v] = 0.0; } /* :639-646 */
    for (int v = 0; v < nn; v++) { /* :758-760 and :5538-5547 */
        if (isd[v]) { for (int k = 0; k < nn; k++) A[(size_t)v * nn + k] = 0; A[(size_t)v * nn + v] = 1.0; b[v] = dir[v]; }
        else if (isn[v]) b[v] = rhs[v] + neu[v];
        else b[v] = rhs[v];
    }
    int singular = 0;
    for (int k = 0; k < nn && !singular; k++) { /* Gaussian elimination, partial pivoting */
        int piv = k;
        for (int r = k + 1; r < nn; r++)
            if (fabs(A[(size_t)r * nn + k]) > fabs(A[(size_t)piv * nn + k])) piv = r;
        if (A[(size_t)piv * nn + k] == 0.0) { singular = 1; break; }
        if (piv != k) {
            for (int cidx = 0; cidx < nn; cidx++) { double t = A[(size_t)k * nn + cidx]; A[(size_t)k * nn + cidx] = A[(size_t)piv * nn + cidx]; A[(size_t)piv * nn + cidx] = t; }
            double t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        for (int r = k + 1; r < nn; r++) {
            double m = A[(size_t)r * nn + k] / A[(size_t)k * nn + k];
            if (m == 0.0) continue;
            for (int cidx = k; cidx < nn; cidx++) A[(size_t)r * nn + cidx] -= m * A[(size_t)k * nn + cidx];
            b[r] -= m * b[k];
        }
    }
    int rc = PGPU_OK;
    if (!singular) {
        for (int k = nn - 1; k >= 0; k--) {
            double sum = b[k];
            for (int cidx = k + 1; cidx < nn; cidx++) sum -= A[(size_t)k * nn + cidx] * b[cidx];
            b[k] = sum / A[(size_t)k * nn + k];
        }
        if (phi_out) memcpy(phi_out, b, sizeof(double) * (size_t)nn);
        rc = po_compute_e_field(w, b, nn, e_out);
    }
    free(A); free(b); free(dir); free(neu); free(isd); free(isn); free(unused);
    if (singular) FAIL(w, PGPU_ERR_INVALID, "po_field_solve: singular system");
    return rc;
}
int po_field_solve_1d(po_ctx* w, int32_t tID, const double* rhs, double* phi_out, double* e_out) {
    if (w->c.grid_type != PGPU_UNSTRUCTURED || w->dims_mesh != 1) FAIL(w, PGPU_ERR_UNSUPPORTED, "po_field_solve_1d: 1-D meshes only");
    return field_solve(w, tID, rhs, phi_out, e_out);
}
/* (rtol, max_iter: the oracle's solve is direct) */
int po_field_solve(po_ctx* w, int32_t tID, const double* rhs, double* phi_out, double* e_out, double rtol, int32_t max_iter,
                   int32_t* iters_out, double* relres_out) {
    (void)rtol; (void)max_iter;
    if (iters_out) *iters_out = 0;
    if (relres_out) *relres_out = 0.0;
    return field_solve(w, tID, rhs, phi_out, e_out);
}

/* ------------------------------------------------------------------------------------------
 * GRID_AVG / GRID_RESET / CHECKS. ref: src/postprocess.f90:34-416 (BOOL_DUMP_MOMENTS = .FALSE.), :1070-1094, :1722-1800
 * ---------------------------------------------------------------------------------------- */
static int64_t ncells_of(const po_ctx* w) {
    return (w->c.grid_type == PGPU_UNSTRUCTURED) ? w->nc : (int64_t)w->c.nx * w->c.ny;
}
static int gavg_alloc(po_ctx* w) {
    int64_t nb = ncells_of(w) * w->ns;
    if (w->gavg && w->gbins == nb) return PGPU_OK;
    free(w->gavg); free(w->gcum1); free(w->gcum2);
    w->gavg = (double*)calloc((size_t)(9 * nb), sizeof(double));
    w->gcum1 = (int*)calloc((size_t)nb, sizeof(int));
    w->gcum2 = (int*)calloc((size_t)nb, sizeof(int));
    w->gbins = nb;
    w->avg_cumulated = 0;
    return PGPU_OK;
}
int po_grid_reset(po_ctx* w) {
    gavg_alloc(w);
    memset(w->gavg, 0, sizeof(double) * (size_t)(9 * w->gbins));
    memset(w->gcum1, 0, sizeof(int) * (size_t)w->gbins);
    memset(w->gcum2, 0, sizeof(int) * (size_t)w->gbins);
    w->avg_cumulated = 0;
    return PGPU_OK;
}
int po_grid_avg(po_ctx* w) {
    gavg_alloc(w);
    const int64_t nb = w->gbins, ncells = ncells_of(w);
    int* tnp = (int*)calloc((size_t)nb, sizeof(int));
    double* t = (double*)calloc((size_t)(12 * nb), sizeof(double));
    double *vx = t, *vy = t + nb, *vz = t + 2 * nb, *vx2 = t + 3 * nb, *vy2 = t + 4 * nb, *vz2 = t + 5 * nb, *erot = t + 6 * nb,
           *evib = t + 7 * nb, *ttrx = t + 8 * nb, *ttry = t + 9 * nb, *ttrz = t + 10 * nb;
    double* trotvib = (double*)calloc((size_t)(2 * nb), sizeof(double));
    double *trot = trotvib, *tvib = trotvib + nb;
    char* one = (char*)calloc((size_t)nb, 1);
    char* two = (char*)calloc((size_t)nb, 1);
    for (int64_t jp = 0; jp < w->np; jp++) { /* :133-138 */
        int64_t idx = (w->p[jp].IC - 1) + ncells * (w->p[jp].S_ID - 1);
        tnp[idx]++;
    }
    for (int64_t jp = 0; jp < w->np; jp++) { /* :141-148 */
        int64_t idx = (w->p[jp].IC - 1) + ncells * (w->p[jp].S_ID - 1);
        vx[idx] += w->p[jp].VX; vy[idx] += w->p[jp].VY; vz[idx] += w->p[jp].VZ;
    }
    for (int64_t i = 0; i < nb; i++) /* :151-158 */
        if (tnp[i] > 0) {
            one[i] = 1;
            vx[i] = vx[i] / (double)tnp[i]; vy[i] = vy[i] / (double)tnp[i]; vz[i] = vz[i] / (double)tnp[i];
        }
    for (int64_t jp = 0; jp < w->np; jp++) { /* :161-171 */
        const part_t* q = &w->p[jp];
        int64_t idx = (q->IC - 1) + ncells * (q->S_ID - 1);
        vx2[idx] += q->VX * q->VX; vy2[idx] += q->VY * q->VY; vz2[idx] += q->VZ * q->VZ;
        erot[idx] += q->EROT; evib[idx] += q->EVIB;
    }
    for (int64_t jc = 0; jc < ncells; jc++) /* :262-297 */
        for (int js = 0; js < w->ns; js++) {
            int64_t i = jc + ncells * js;
            if (tnp[i] > 0) {
                if (w->rotdof[js] == 0) trot[i] = 0.0;
                else trot[i] = (double)(2.f / (float)w->rotdof[js]) / KB * (erot[i] / (double)tnp[i]);
                if (w->vibdof[js] == 0) tvib[i] = 0.0;
                else tvib[i] = (double)(2.f / (float)w->vibdof[js]) / KB * (evib[i] / (double)tnp[i]);
                if (tnp[i] > 1) {
                    two[i] = 1;
                    double numpart = (double)tnp[i], sampledof = numpart;
                    ttrx[i] = w->mass[js] / KB * (vx2[i] - numpart * vx[i] * vx[i]) / sampledof;
                    ttry[i] = w->mass[js] / KB * (vy2[i] - numpart * vy[i] * vy[i]) / sampledof;
                    ttrz[i] = w->mass[js] / KB * (vz2[i] - numpart * vz[i] * vz[i]) / sampledof;
                }
            }
        }
    /* :338-381 running averages */
    double* A = w->gavg;
    double dcum = (double)w->avg_cumulated;
    for (int64_t i = 0; i < nb; i++) A[i] = (A[i] * dcum + (double)tnp[i]) / (double)((float)w->avg_cumulated + 1.f);
    w->avg_cumulated++;
    for (int64_t i = 0; i < nb; i++) {
        if (one[i]) {
            double k = (double)w->gcum1[i], den = (double)((float)w->gcum1[i] + 1.f);
            A[1 * nb + i] = (A[1 * nb + i] * k + vx[i]) / den;
            A[2 * nb + i] = (A[2 * nb + i] * k + vy[i]) / den;
            A[3 * nb + i] = (A[3 * nb + i] * k + vz[i]) / den;
            A[7 * nb + i] = (A[7 * nb + i] * k + trot[i]) / den;
            A[8 * nb + i] = (A[8 * nb + i] * k + tvib[i]) / den;
            w->gcum1[i]++;
        }
        if (two[i]) {
            double k = (double)w->gcum2[i], den = (double)((float)w->gcum2[i] + 1.f);
            A[4 * nb + i] = (A[4 * nb + i] * k + ttrx[i]) / den;
            A[5 * nb + i] = (A[5 * nb + i] * k + ttry[i]) / den;
            A[6 * nb + i] = (A[6 * nb + i] * k + ttrz[i]) / den;
            w->gcum2[i]++;
        }
    }
    free(tnp); free(t); free(trotvib); free(one); free(two);
    return PGPU_OK;
}
int po_get_grid_avg(po_ctx* w, int64_t n, double* avg_np, double* avg_vx, double* avg_vy, double* avg_vz, double* avg_ttrx,
                    double* avg_ttry, double* avg_ttrz, double* avg_ttr, double* avg_trot, double* avg_tvib, int32_t* avg_cumulated) {
    if (!w->gavg) FAIL(w, PGPU_ERR_STATE, "po_get_grid_avg: po_grid_avg has not been called");
    const int64_t nb = w->gbins;
    if (n < nb) FAIL(w, PGPU_ERR_INVALID, "po_get_grid_avg: arrays too small");
    double* dst[9] = {avg_np, avg_vx, avg_vy, avg_vz, avg_ttrx, avg_ttry, avg_ttrz, avg_trot, avg_tvib};
    for (int k = 0; k < 9; k++)
        if (dst[k]) memcpy(dst[k], w->gavg + (size_t)k * nb, sizeof(double) * (size_t)nb);
    if (avg_ttr)
        for (int64_t i = 0; i < nb; i++) avg_ttr[i] = (w->gavg[4 * nb + i] + w->gavg[5 * nb + i] + w->gavg[6 * nb + i]) / (double)3.0f;
    if (avg_cumulated) *avg_cumulated = w->avg_cumulated;
    return PGPU_OK;
}
/* CHECKS particle loop, ref: src/postprocess.f90:1752-1786 */
int po_checks(po_ctx* w, int64_t* tot_num, double* tot_momentum, double* tot_ke_part, double* tot_ie) {
    for (int js = 0; js < w->ns; js++) {
        if (tot_num) tot_num[js] = 0;
        if (tot_ke_part) tot_ke_part[js] = 0;
        if (tot_momentum) tot_momentum[3 * js] = tot_momentum[3 * js + 1] = tot_momentum[3 * js + 2] = 0;
    }
    double ie = 0;
    for (int64_t jp = 0; jp < w->np; jp++) {
        const part_t* q = &w->p[jp];
        int js = q->S_ID - 1;
        double cfnum = w->cell_fnum ? w->cell_fnum[q->IC - 1] : w->c.fnum, m = w->mass[js]; /* ref: :1764-1768 */
        if (tot_num) tot_num[js]++;
        if (tot_momentum) {
            tot_momentum[0 + 3 * js] += m * q->VX * cfnum;
            tot_momentum[1 + 3 * js] += m * q->VY * cfnum;
            tot_momentum[2 + 3 * js] += m * q->VZ * cfnum;
        }
        if (tot_ke_part) tot_ke_part[js] += 0.5 * m * (q->VX * q->VX + q->VY * q->VY + q->VZ * q->VZ) * cfnum;
        ie += (q->EROT + q->EVIB) * cfnum;
    }
    if (tot_ie) *tot_ie = ie;
    return PGPU_OK;
}

/* ref: src/timecycle.f90:271-299 (PIC_TYPE==NONE branch of TIME_LOOP), single rank */
int po_run_steps(po_ctx* w, int32_t tID0, int32_t nsteps, int64_t* total_coll, int64_t* total_reac) {
    int64_t tc = 0, tr = 0;
    for (int s = 0; s < nsteps; s++) {
        int rc = po_advect(w, tID0 + s);
        if (rc) return rc;
        if (w->c.collision_type == PGPU_DSMC) {
            int64_t a = 0, b = 0;
            rc = po_collide(w, tID0 + s, &a, &b);
            if (rc) return rc;
            tc += a; tr += b;
        } else if (w->c.collision_type == PGPU_MCC_VAHEDI) {
            int64_t a = 0, b = 0;
            rc = po_mcc_vahedi(w, tID0 + s, &a, &b);
            if (rc) return rc;
            tc += a; tr += b;
        } else if (w->c.collision_type == PGPU_MCC) {
            int64_t a = 0, b = 0;
            rc = po_mcc_collisions(w, tID0 + s, &a, &b);
            if (rc) return rc;
            tc += a; tr += b;
        }
    }
    if (total_coll) *total_coll = tc;
    if (total_reac) *total_reac = tr;
    return PGPU_OK;
}

/* ------------------------------------------------------------------------------------------
 * setters / container plumbing
 * ---------------------------------------------------------------------------------------- */
int po_create(po_ctx** ctx, int device_ignored) {
    (void)device_ignored;
    po_ctx* w = (po_ctx*)calloc(1, sizeof(po_ctx));
    if (!w) return PGPU_ERR_INVALID;
    w->c.nranks = 1;
    w->c.eps_scaling = 1.0;
    *ctx = w;
    return PGPU_OK;
}

int po_destroy(po_ctx* w) {
    if (!w) return PGPU_OK;
    free(w->mass); free(w->charge); free(w->rotrel); free(w->vibrel); free(w->rotdof); free(w->vibdof);
    free(w->sigmas); free(w->omegas); free(w->alphas); free(w->grefs);
    free(w->reac); free(w->tab_off); free(w->tab_e); free(w->tab_cs);
    free(w->wr_r); free(w->wr_p1); free(w->wr_nprod); free(w->wr_prob);
    free(w->mcc_sp); free(w->mcc_frac); free(w->mcc_cell_nrho);
    free(w->node_coords); free(w->cell_nodes); free(w->cell_neighbors); free(w->face_normal);
    free(w->face_tang1); free(w->face_tang2); free(w->cell_edges_len); free(w->cell_faces_pg);
    free(w->cell_pg); free(w->basis_coeffs); free(w->cell_volumes); free(w->cell_procs);
    free(w->face_bidx); free(w->bound_area);
    free(w->bc_particle); free(w->bc_field); free(w->bc_volume); free(w->bc_react); free(w->bc_dump);
    free(w->bc_temp); free(w->bc_accn); free(w->bc_acct);
    free(w->w_x1); free(w->w_y1); free(w->w_x2); free(w->w_y2); free(w->w_nx); free(w->w_ny);
    free(w->w_temp); free(w->w_transm); free(w->w_spec); free(w->w_diff); free(w->w_porous); free(w->w_react);
    free(w->xcoord); free(w->ycoord);
    free(w->e_field); free(w->b_field); free(w->surface_charge); free(w->spwt); free(w->mix_off); free(w->mix_sp); free(w->mix_frac); free(w->seed_par); free(w->seed_mix);
    free(w->vinj_par); free(w->vinj_mix); free(w->emit_par); free(w->emit_int); free(w->emit_nfs);
    free(w->cell_fnum); free(w->fb_pot); free(w->fb_rfpot); free(w->fb_freq); free(w->fb_efield); free(w->fb_eps);
    free(w->script); free(w->gavg); free(w->gcum1); free(w->gcum2);
    free(w->p); free(w->bcount); free(w->wcount); free(w->tallies); free(w->npcall);
    free(w);
    return PGPU_OK;
}

const char* po_last_error(const po_ctx* w) { return w ? w->err : "null context"; }

int po_set_rng_mode(po_ctx* w, int mode) {
    if (mode < 0 || mode > 2) FAIL(w, PGPU_ERR_INVALID, "rng mode must be 0 (MT19937-64), 1 (Philox) or 2 (scripted draws)");
    w->rng_mode = mode;
    w->seeded = 0;
    return PGPU_OK;
}

/* rng_mode 2: the sequence rf() will return; po_rng_script_used() = draws consumed so far (may exceed n) */
int po_set_rng_script(po_ctx* w, const double* draws, int64_t n) {
    REPL(w->script, draws, n);
    w->script_n = n;
    w->script_pos = 0;
    return PGPU_OK;
}
int64_t po_rng_script_used(const po_ctx* w) { return w->script_pos; }
int po_collision_type(const po_ctx* w) { return w->c.collision_type; }
int po_set_cell_fnum(po_ctx* w, int nc, const double* cell_fnum) {
    if (!cell_fnum || nc <= 0) { free(w->cell_fnum); w->cell_fnum = NULL; return PGPU_OK; }
    REPL(w->cell_fnum, cell_fnum, nc);
    return PGPU_OK;
}
int po_set_deterministic(po_ctx* w, int32_t on) { (void)w; (void)on; return PGPU_OK; } /* the oracle is sequential */

int po_set_config(po_ctx* w, const pgpu_config* cfg) {
    w->c = *cfg;
    if (w->c.nranks < 1) w->c.nranks = 1;
    if (w->c.eps_scaling == 0.0) w->c.eps_scaling = 1.0;
    w->seeded = 0;
    return PGPU_OK;
}

static void alloc_counters(po_ctx* w) {
    free(w->bcount);
    w->bcount = (int64_t*)calloc((size_t)4 * (w->ns > 0 ? w->ns : 1), sizeof(int64_t));
    free(w->wcount);
    w->wcount = (int64_t*)calloc((size_t)(w->nw > 0 ? w->nw : 1) * (w->ns > 0 ? w->ns : 1), sizeof(int64_t));
    free(w->tallies);
    w->tallies = NULL;
    if (w->nboundcells > 0 && w->ns > 0)
        w->tallies = (double*)calloc((size_t)10 * w->nboundcells * w->ns, sizeof(double));
}

int po_set_species(po_ctx* w, int ns, const double* mass, const double* charge, const int32_t* rotdof,
                   const int32_t* vibdof, const double* rotrel, const double* vibrel) {
    if (ns < 1) FAIL(w, PGPU_ERR_INVALID, "ns < 1");
    w->ns = ns;
    REPL(w->mass, mass, ns); REPL(w->charge, charge, ns);
    REPL(w->rotdof, rotdof, ns); REPL(w->vibdof, vibdof, ns);
    REPL(w->rotrel, rotrel, ns); REPL(w->vibrel, vibrel, ns);
    alloc_counters(w);
    return PGPU_OK;
}

int po_set_vss(po_ctx* w, int ns, const double* sigmas, const double* omegas, const double* alphas,
               const double* grefs, double sigmamax) {
    if (ns != w->ns) FAIL(w, PGPU_ERR_INVALID, "set_vss: ns mismatch");
    REPL(w->sigmas, sigmas, ns * ns); REPL(w->omegas, omegas, ns * ns);
    REPL(w->alphas, alphas, ns * ns); REPL(w->grefs, grefs, ns * ns);
    w->sigmamax = sigmamax;
    return PGPU_OK;
}

int po_set_reactions(po_ctx* w, int nr, const int32_t* type, const int32_t* r1, const int32_t* r2,
                     const int32_t* p1, const int32_t* p2, const int32_t* p3, const int32_t* p4,
                     const int32_t* nprod, const int32_t* is_cex, const double* ea, const double* c1,
                     const double* c2, const double* c3, const int32_t* tab_off, const double* tab_e,
                     const double* tab_cs) {
    free(w->reac); w->reac = NULL;
    w->nr = nr;
    if (nr == 0) return PGPU_OK;
    int ntab = tab_off ? tab_off[nr] : 0;
    REPL(w->tab_off, tab_off, nr + 1);
    REPL(w->tab_e, tab_e, ntab);
    REPL(w->tab_cs, tab_cs, ntab);
    w->reac = (reaction_t*)calloc((size_t)nr, sizeof(reaction_t));
    for (int i = 0; i < nr; i++) {
        reaction_t* r = &w->reac[i];
        r->type = type[i]; r->r1 = r1[i]; r->r2 = r2[i];
        r->p1 = p1[i]; r->p2 = p2[i]; r->p3 = p3 ? p3[i] : -1; r->p4 = p4 ? p4[i] : -1;
        r->nprod = nprod[i]; r->is_cex = is_cex ? is_cex[i] : 0;
        r->ea = ea[i]; r->c1 = c1 ? c1[i] : 0; r->c2 = c2 ? c2[i] : 0; r->c3 = c3 ? c3[i] : 0;
        if (tab_off) {
            r->tab_n = tab_off[i + 1] - tab_off[i];
            r->tab_e = w->tab_e + tab_off[i];
            r->tab_cs = w->tab_cs + tab_off[i];
        }
    }
    return PGPU_OK;
}

int po_set_wall_reactions(po_ctx* w, int n, const int32_t* r_sp, const int32_t* p1_sp,
                          const int32_t* nprod, const double* prob) {
    w->nwr = n;
    REPL(w->wr_r, r_sp, n); REPL(w->wr_p1, p1_sp, n); REPL(w->wr_nprod, nprod, n); REPL(w->wr_prob, prob, n);
    return PGPU_OK;
}

int po_set_mcc_background(po_ctx* w, int ncomp, const int32_t* sp_id, const double* molfrac) {
    w->mcc_ncomp = ncomp;
    REPL(w->mcc_sp, sp_id, ncomp); REPL(w->mcc_frac, molfrac, ncomp);
    return PGPU_OK;
}

/* MCC_BG_CELL_NRHO(SP_ID, IC), ref: src/initialization.f90:2752-2788; n = 0 clears it */
int po_set_mcc_cell_density(po_ctx* w, int ns, int64_t nc, const double* nrho) {
    if (!nrho || ns <= 0 || nc <= 0) {
        free(w->mcc_cell_nrho);
        w->mcc_cell_nrho = NULL;
        w->mcc_cell_nrho_n = 0;
        return PGPU_OK;
    }
    if (ns != w->ns) FAIL(w, PGPU_ERR_INVALID, "po_set_mcc_cell_density: ns %d != %d", ns, w->ns);
    REPL(w->mcc_cell_nrho, nrho, (size_t)ns * nc);
    w->mcc_cell_nrho_n = (int64_t)ns * nc;
    return PGPU_OK;
}
int po_set_mesh(po_ctx* w, int dims, int nn, int nc, const double* node_coords, const int32_t* cell_nodes,
                const int32_t* cell_neighbors, const double* face_normal, const double* face_tang1,
                const double* face_tang2, const double* cell_edges_len, const int32_t* cell_faces_pg,
                const int32_t* cell_pg, const double* basis_coeffs, const double* cell_volumes,
                const int32_t* cell_procs, const int32_t* face_boundary_index, int nboundcells,
                const double* boundary_area) {
    if (dims < 1 || dims > 3) FAIL(w, PGPU_ERR_INVALID, "set_mesh: dims must be 1..3");
    int nf = dims + 1;
    w->dims_mesh = dims; w->nn = nn; w->nc = nc;
    REPL(w->node_coords, node_coords, 3 * (size_t)nn);
    REPL(w->cell_nodes, cell_nodes, (size_t)nf * nc);
    REPL(w->cell_neighbors, cell_neighbors, (size_t)nf * nc);
    REPL(w->face_normal, face_normal, 3 * (size_t)nf * nc);
    REPL(w->face_tang1, face_tang1, face_tang1 ? 12 * (size_t)nc : 0);
    REPL(w->face_tang2, face_tang2, face_tang2 ? 12 * (size_t)nc : 0);
    REPL(w->cell_edges_len, cell_edges_len, cell_edges_len ? 3 * (size_t)nc : 0);
    REPL(w->cell_faces_pg, cell_faces_pg, (size_t)nf * nc);
    REPL(w->cell_pg, cell_pg, cell_pg ? (size_t)nc : 0);
    REPL(w->basis_coeffs, basis_coeffs, basis_coeffs ? (size_t)nf * nf * nc : 0);
    REPL(w->cell_volumes, cell_volumes, (size_t)nc);
    REPL(w->cell_procs, cell_procs, cell_procs ? (size_t)nc : 0);
    REPL(w->face_bidx, face_boundary_index, face_boundary_index ? (size_t)nf * nc : 0);
    w->nboundcells = face_boundary_index ? nboundcells : 0;
    REPL(w->bound_area, boundary_area, w->nboundcells);
    free(w->surface_charge);
    w->surface_charge = (double*)calloc((size_t)nn, sizeof(double));
    w->nsc = nn;
    alloc_counters(w);
    return PGPU_OK;
}

int po_set_cell_procs(po_ctx* w, int nc, const int32_t* cell_procs) {
    if (nc != ncells(w)) FAIL(w, PGPU_ERR_INVALID, "set_cell_procs: nc mismatch");
    REPL(w->cell_procs, cell_procs, nc);
    return PGPU_OK;
}

int po_set_bcs(po_ctx* w, int nbc, const int32_t* particle_bc, const int32_t* field_bc,
               const int32_t* volume_bc, const double* wall_temp, const double* acc_n,
               const double* acc_t, const int32_t* react, const int32_t* dump_fluxes) {
    w->nbc = nbc;
    REPL(w->bc_particle, particle_bc, nbc); REPL(w->bc_field, field_bc, nbc);
    REPL(w->bc_volume, volume_bc, nbc); REPL(w->bc_temp, wall_temp, nbc);
    REPL(w->bc_accn, acc_n, nbc); REPL(w->bc_acct, acc_t, nbc);
    REPL(w->bc_react, react, nbc); REPL(w->bc_dump, dump_fluxes, nbc);
    return PGPU_OK;
}

int po_set_walls(po_ctx* w, int nw, const double* x1, const double* y1, const double* x2, const double* y2,
                 const double* normx, const double* normy, const double* temp, const int32_t* specular,
                 const int32_t* diffuse, const int32_t* porous, const double* transmissivity,
                 const int32_t* react) {
    w->nw = nw;
    REPL(w->w_x1, x1, nw); REPL(w->w_y1, y1, nw); REPL(w->w_x2, x2, nw); REPL(w->w_y2, y2, nw);
    REPL(w->w_nx, normx, nw); REPL(w->w_ny, normy, nw); REPL(w->w_temp, temp, nw);
    REPL(w->w_spec, specular, nw); REPL(w->w_diff, diffuse, nw); REPL(w->w_porous, porous, nw);
    REPL(w->w_transm, transmissivity, nw); REPL(w->w_react, react, nw);
    alloc_counters(w);
    return PGPU_OK;
}

int po_set_grid_coords(po_ctx* w, int nxp, const double* xcoord, int nyp, const double* ycoord) {
    w->nxp = nxp; w->nyp = nyp;
    REPL(w->xcoord, xcoord, nxp); REPL(w->ycoord, ycoord, nyp);
    return PGPU_OK;
}

int po_upload_particles(po_ctx* w, const void* aos, int64_t n, int32_t stride) {
    if (stride < (int32_t)92) FAIL(w, PGPU_ERR_INVALID, "stride too small");
    reserve(w, n);
    for (int64_t i = 0; i < n; i++) memcpy(&w->p[i], (const char*)aos + (size_t)i * stride, sizeof(part_t) < (size_t)stride ? sizeof(part_t) : (size_t)stride);
    w->np = n;
    w->id_counter = n;
    return PGPU_OK;
}

int po_append_particles(po_ctx* w, const void* aos, int64_t n, int32_t stride) {
    reserve(w, w->np + n);
    for (int64_t i = 0; i < n; i++) memcpy(&w->p[w->np + i], (const char*)aos + (size_t)i * stride, sizeof(part_t) < (size_t)stride ? sizeof(part_t) : (size_t)stride);
    w->np += n;
    return PGPU_OK;
}

int po_download_particles(po_ctx* w, void* aos, int64_t capacity, int64_t* n, int32_t stride) {
    if (capacity < w->np) FAIL(w, PGPU_ERR_INVALID, "download: capacity %lld < %lld", (long long)capacity, (long long)w->np);
    for (int64_t i = 0; i < w->np; i++) memcpy((char*)aos + (size_t)i * stride, &w->p[i], sizeof(part_t) < (size_t)stride ? sizeof(part_t) : (size_t)stride);
    if (n) *n = w->np;
    return PGPU_OK;
}

int po_num_particles(po_ctx* w, int64_t* n) { *n = w->np; return PGPU_OK; }
/* record-plane transfers (see pantera_gpu.h): the oracle keeps AoS records, so these are plain transpositions */
int po_upload_planes(po_ctx* w, const void* pos, const void* vel, const void* ie, int64_t n) {
    const pgpu_pos_record* P = (const pgpu_pos_record*)pos;
    const pgpu_vel_record* V = (const pgpu_vel_record*)vel;
    const pgpu_ie_record* E = (const pgpu_ie_record*)ie;
    reserve(w, n);
    for (int64_t i = 0; i < n; i++) {
        part_t* q = &w->p[i];
        memset(q, 0, sizeof(*q));
        q->X = P[i].X; q->Y = P[i].Y; q->Z = P[i].Z; q->IC = P[i].IC; q->S_ID = P[i].S_ID;
        q->VX = V[i].VX; q->VY = V[i].VY; q->VZ = V[i].VZ;
        if (E) { q->EROT = E[i].EROT; q->EVIB = E[i].EVIB; }
        q->DTRIM = w->c.dt;
        q->ID = (int64_t)w->c.rank + (i << 8);
    }
    w->np = n;
    w->id_counter = n;
    return PGPU_OK;
}
int po_download_planes(po_ctx* w, void* pos, void* vel, void* ie, int64_t capacity, int64_t* n) {
    if (capacity < w->np) FAIL(w, PGPU_ERR_INVALID, "po_download_planes: capacity too small");
    pgpu_pos_record* P = (pgpu_pos_record*)pos;
    pgpu_vel_record* V = (pgpu_vel_record*)vel;
    pgpu_ie_record* E = (pgpu_ie_record*)ie;
    for (int64_t i = 0; i < w->np; i++) {
        const part_t* q = &w->p[i];
        P[i].X = q->X; P[i].Y = q->Y; P[i].Z = q->Z; P[i].IC = q->IC; P[i].S_ID = q->S_ID;
        V[i].VX = q->VX; V[i].VY = q->VY; V[i].VZ = q->VZ; V[i].S_ID = q->S_ID; V[i].pad_ = 0;
        if (E) { E[i].EROT = q->EROT; E[i].EVIB = q->EVIB; }
    }
    if (n) *n = w->np;
    return PGPU_OK;
}

int po_set_efield(po_ctx* w, const double* e, int64_t n) { REPL(w->e_field, e, n); w->ne = n; return PGPU_OK; }
int po_set_bfield(po_ctx* w, const double* b, int64_t n) { REPL(w->b_field, b, n); w->nb = n; return PGPU_OK; }
int po_get_surface_charge(po_ctx* w, double* sc, int64_t n) {
    for (int64_t i = 0; i < n && i < w->nsc; i++) sc[i] = w->surface_charge[i];
    return PGPU_OK;
}
int po_set_surface_charge(po_ctx* w, const double* sc, int64_t n) {
    REPL(w->surface_charge, sc, n); w->nsc = n; return PGPU_OK;
}

int po_get_boundary_counts(po_ctx* w, int64_t* b, int64_t nb, int64_t* wc, int64_t nw) {
    for (int64_t i = 0; i < nb && i < 4 * (int64_t)w->ns; i++) b[i] = w->bcount[i];
    for (int64_t i = 0; wc && i < nw && i < (int64_t)w->nw * w->ns; i++) wc[i] = w->wcount[i];
    return PGPU_OK;
}
int po_get_boundary_tallies(po_ctx* w, double* t, int64_t n, int32_t reset) {
    int64_t tot = (int64_t)10 * w->nboundcells * w->ns;
    if (!w->tallies) return PGPU_OK;
    for (int64_t i = 0; i < n && i < tot; i++) t[i] = w->tallies[i];
    if (reset) memset(w->tallies, 0, sizeof(double) * (size_t)tot);
    return PGPU_OK;
}
int po_get_cell_counts(po_ctx* w, int32_t* npc, int64_t nc) {
    if (!w->npcall) FAIL(w, PGPU_ERR_STATE, "no sort has run yet");
    for (int64_t i = 0; i < nc && i < ncells(w); i++) npc[i] = w->npcall[i];
    return PGPU_OK;
}
int po_get_timings(po_ctx* w, double* sec, int32_t n) {
    for (int i = 0; i < n && i < 3; i++) sec[i] = w->tsec[i];
    return PGPU_OK;
}
