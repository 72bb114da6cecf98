/*
 * pgpu_collide.cu -- DSMC collisions on the cell-sorted record store.
 *
 * Replaces: the cell loop of DSMC_COLLISIONS (src/collisions.f90:126-138), VSS_COLLIS (:155-558),
 * FUNCTION_I (:892-908), COLL_INTERNAL_ENERGY (:915-937), HS_SCATTER (:956-992).
 *
 * The reference walks the NCOLL candidate pairs of a cell one after another, and a later candidate sees
 * the velocities an earlier accepted one wrote (:259-551).  That order is kept EXACTLY, in parallel:
 *   prep    velocity bounding box of the cell -> VRMAX, NCOLL, FCORR (:202-253)
 *   pick    one thread per candidate: the pair (JP1, JP2) depends only on the candidate's own counter-based
 *           Philox stream keyed by (cell, candidate) (:263-271), never on particle state
 *   rounds  every waiting candidate claims both of its particles with its sequence number; a candidate that
 *           holds both claims is the earliest one touching either particle, so evaluating it now is what the
 *           sequential loop would do.  Evaluate, repeat until none waits.  Cells, and disjoint pairs inside
 *           a cell, proceed in parallel; chains of candidates sharing a particle serialise, as they must.
 *
 * Two executions:
 *   CELL    (the hot path) one LANE per cell walks that cell's candidates in the reference's order, so the
 *           sequential semantics hold by construction and there is no bookkeeping at all; a warp covers 32
 *           cells (~1300 particles) and re-converges for the expensive deflection math.  Measured against a
 *           shared-memory tile kernel running the dependency rounds below per CTA (r01 profiles), it executes
 *           ~3x fewer instructions, which is what bounds this phase: it is FP64/Philox issue-bound, not
 *           HBM-bound.  32 B read + at most 32 B written per particle, no atomics, no host round trip.
 *   GLOBAL  cells with more than PG_TILE_BIG particles (e.g. the one-cell heat bath of config C1) would
 *           serialise a lane; their candidates go through dependency rounds: every waiting candidate claims
 *           both of its particles with its sequence number; a candidate that holds both claims is the
 *           earliest one touching either particle, so evaluating it now is what the sequential loop would do.
 *           Evaluate, repeat until none waits (claims in global memory, one kernel pair per round).
 * With the oracle run on the same Philox streams both agree with it value for value (tests/test_gpu_collide.py).
 */
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "pgpu_internal.cuh"
#include "pgpu_physics.cuh"

namespace {

#define CAND_DONE 0x80000000u
#define OWNER_FREE 0xFFFFFFFFu

struct MaxOp {
    __device__ __forceinline__ unsigned int operator()(unsigned int a, unsigned int b) const { return a > b ? a : b; }
};

/* ------------------------------------------------------------------------------------------
 * per-cell majorant, ref: src/collisions.f90:224-253
 * ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ int cell_majorant(const PgDev& d, int cell0, int n, double vxmin, double vxmax, double vymin,
                                             double vymax, double vzmin, double vzmax, double& vrmax, double& fcorr) {
    vrmax = sqrt((vxmax - vxmin) * (vxmax - vxmin) + (vymax - vymin) * (vymax - vymin) + (vzmax - vzmin) * (vzmax - vzmin));
    /* ref: :224-230 */
    double vol;
    if (d.grid_type == PGPU_UNSTRUCTURED) vol = __ldg(&d.cell_vol[cell0]);
    else vol = (d.xmax - d.xmin) * (d.ymax - d.ymin) * (d.zmax - d.zmin) / (d.nx * d.ny);
    /* ref: :232  0.5*NPC*(NPC-1) is evaluated in single precision */
    float pairs_f = (0.5f * (float)n) * (float)(n - 1);
    const double cfnum = d.cell_fnum ? __ldg(&d.cell_fnum[cell0]) : d.fnum; /* ref: :192-196 CFNUM = CELL_FNUM(JC) */
    double ncollmax = (double)pairs_f * d.sigmamax * vrmax * cfnum * d.dt / vol;
    double r = floor(ncollmax + 0.5);
    long long half = (long long)floorf(0.5f * (float)n);
    long long ncoll;
    if (!(r >= 1.)) ncoll = 1;
    else if (r > (double)half) ncoll = half;
    else ncoll = (long long)r;
    fcorr = ncollmax / (double)ncoll;
    return (int)ncoll;
}

/* ref: src/collisions.f90:263-271 pair selection from the candidate's own stream; returns the draws used */
__device__ __forceinline__ unsigned int pick_pair(const PgDev& d, unsigned int cell0, unsigned int j, int npcj, int& k1, int& k2) {
    PgRng rng;
    rng.init(d.seed_local, 2, d.tID, (unsigned long long)cell0, (unsigned long long)j);
    k1 = rng.pick(npcj);
    k2 = k1;
    while (k2 == k1) k2 = rng.pick(npcj);
    return rng.ndraw;
}

/* ------------------------------------------------------------------------------------------
 * one candidate pair of VSS_COLLIS, ref: src/collisions.f90:274-546, on register copies of the pair
 * ---------------------------------------------------------------------------------------- */
struct PairIO {
    double c1[3], c2[3];
    double erot1, evib1, erot2, evib2;
    int s1, s2;
};
enum { PR_ACCEPTED = 1, PR_REACTED = 2, PR_V = 4, PR_E1 = 8, PR_E2 = 16, PR_S = 32 };

#define VSS2(arr, s1, s2) __ldg(&(arr)[((s1)-1) + (size_t)d.ns * ((s2)-1)])

/* acceptance test of a candidate, ref: src/collisions.f90:274-325; rfp is the candidate's acceptance draw.
 * The cross-section power law is first bracketed in single precision (MUFU lg2/ex2, relative error
 * < 1e-5): only when the draw falls inside the bracket is the FP64 pow() evaluated, so the decision is
 * the one the full-precision expression gives while almost no candidate pays for an FP64 pow. */
__device__ __forceinline__ bool pair_accept(const PgDev& d, double rfp, const double* c1, const double* c2, int s1, int s2,
                                            double vrmax, double fcorr) {
    double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
    double vr = sqrt(vr2);
    double cref = VSS2(d.grefs, s1, s2), sigma = VSS2(d.sigmas, s1, s2);
    double omega = VSS2(d.omegas, s1, s2);
    double collsigma;
    if (fabs(omega - 0.5) < 1.e-6) {
        collsigma = sigma;
    } else {
        if (vr < (double)1e-20f) return false; /* ref: :314 (single-precision literal) */
        double est = fcorr / (d.sigmamax * vrmax) * vr * (sigma * (double)__powf((float)(vr / cref), (float)(1. - 2. * omega)));
        if (rfp < est * (1. - 1.e-4)) return true;
        if (rfp > est * (1. + 1.e-4)) return false;
        collsigma = sigma * pow(vr / cref, 1. - 2. * omega);
    }
    double collprob = fcorr / (d.sigmamax * vrmax) * vr * collsigma;
    return rfp < collprob;
}

/* An ACCEPTED candidate, ref: src/collisions.f90:327-546, in three pieces so that the lane-per-cell kernel can keep its
 * warp together: (1) the reaction test, (2) the redistribution after a reaction -- rare and long, batched there --,
 * (3) Borgnakke-Larsen + VSS deflection.  g continues the candidate's stream right after the acceptance draw.
 * gate_all: MCC_COLLISIONS skips a reaction whose activation energy is not exceeded whatever its type (ref: :1383-1384);
 * VSS_COLLIS applies that test to TCE reactions only (ref: :349-350). */

/* (1) ref: :331-352: returns the index of the reaction that fires (ecoll_out = ECOLL - EA) or -1 */
__device__ __forceinline__ int pair_react_test(const PgDev& d, PgRng& g, const PairIO& p, bool gate_all, double& ecoll_out) {
    if (d.nr == 0) return -1;
    const double* c1 = p.c1;
    const double* c2 = p.c2;
    const int s1 = p.s1, s2 = p.s2;
    const double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
    const double vr = sqrt(vr2);
    const double m1 = __ldg(&d.mass[s1 - 1]), m2 = __ldg(&d.mass[s2 - 1]);
    const double mred = m1 * m2 / (m1 + m2);
    const double etr = 0.5 * mred * vr2;
    const double ecoll = etr + p.erot1 + p.erot2 + p.evib1 + p.evib2;
    for (int jr = 0; jr < d.nr; jr++) {
        const PgReaction* r = &d.reac[jr];
        const int rr1 = __ldg(&r->r1), rr2 = __ldg(&r->r2);
        if (rr1 != s1) {
            if (rr1 != s2) continue;
            else if (rr2 != s1) continue;
        } else {
            if (rr2 != s2) continue;
        }
        const double ea = __ldg(&r->ea);
        const int type = __ldg(&r->type);
        double ptce;
        if (gate_all && ecoll <= ea) continue;
        if (type == PGPU_TCE) {
            if (ecoll <= ea) continue;
            ptce = __ldg(&r->c1) * pow(ecoll - ea, __ldg(&r->c2)) * pow(1. - ea / ecoll, __ldg(&r->c3));
        } else if (type == PGPU_LXCAT) {
            const double cref = VSS2(d.grefs, s1, s2), sigma = VSS2(d.sigmas, s1, s2), omega = VSS2(d.omegas, s1, s2);
            double sigma_r = pg_interp_cs(etr, r->tab_n, d.tab_e + r->tab_off, d.tab_cs + r->tab_off);
            ptce = sigma_r / (sigma * pow(vr / cref, 1. - 2. * omega));
        } else {
            ptce = 0;
        }
        const double rfp = g.rf();
        if (rfp < ptce) {
            ecoll_out = ecoll - ea;
            return jr;
        }
    }
    return -1;
}

/* (2) ref: :353-447.  Roles are NOT swapped when the pair order is opposite to the reaction. */
__device__ int pair_react_apply(const PgDev& d, PgRng& g, PairIO& p, int jr, double ecoll, long long gslot2, int cell0) {
    const PgReaction r = d.reac[jr];
    double erot1 = p.erot1, evib1 = p.evib1, erot2 = p.erot2, evib2 = p.evib2;
    const PgPos P2 = d.p.pos[gslot2];
    pg_react_redistribute(d, g, r, ecoll, p.c1, p.c2, erot1, evib1, erot2, evib2, P2.x, P2.y, P2.z, cell0 + 1);
    p.s1 = r.p1;
    p.s2 = r.p2;
    p.erot1 = erot1; p.evib1 = evib1; p.erot2 = erot2; p.evib2 = evib2;
    return PR_ACCEPTED | PR_REACTED | PR_V | PR_E1 | PR_E2 | PR_S;
}

/* (3) ref: :454-546 Borgnakke-Larsen (at most one mode) + VSS deflection.  The four branches of the reference's
 * IF / ELSE IF chain differ only in WHICH energy is redrawn: the mode is selected first and COLL_INTERNAL_ENERGY is
 * called from one place, so that the lanes of a warp share its rejection loop instead of running it four times over. */
__device__ int pair_inelastic_vss(const PgDev& d, PgRng& g, PairIO& p) {
    const double PI2 = 2. * PG_PI;
    double* c1 = p.c1;
    double* c2 = p.c2;
    const int s1 = p.s1, s2 = p.s2;
    const double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
    const double omega = VSS2(d.omegas, s1, s2), alpha = VSS2(d.alphas, s1, s2);
    const double m1 = __ldg(&d.mass[s1 - 1]), m2 = __ldg(&d.mass[s2 - 1]);
    const double mred = m1 * m2 / (m1 + m2);
    double etr = 0.5 * mred * vr2;
    int flags = PR_ACCEPTED;
    const int rd1 = __ldg(&d.rotdof[s1 - 1]), rd2 = __ldg(&d.rotdof[s2 - 1]);
    const int vd1 = __ldg(&d.vibdof[s1 - 1]), vd2 = __ldg(&d.vibdof[s2 - 1]);
    const double rr1 = __ldg(&d.rotrel[s1 - 1]), rr2 = __ldg(&d.rotrel[s2 - 1]);
    const double vr1 = __ldg(&d.vibrel[s1 - 1]), vr2_ = __ldg(&d.vibrel[s2 - 1]);
    double trdof = 3.;
    const double prot1 = (trdof + rd1) / trdof * rr1;
    const double prot2 = (trdof + rd2) / trdof * rr2;
    const double pvib1 = (trdof + vd1) / trdof * vr1;
    const double pvib2 = (trdof + vd2) / trdof * vr2_;
    trdof = 5. - 2. * omega;
    /* when neither species can exchange internal energy the selector cannot fall in any branch and its draw is not
     * taken from the stream (stream contract shared with the oracle's rng_mode 1) */
    double rfp;
    if (rr1 + rr2 + vr1 + vr2_ == 0.) rfp = 1.;
    else rfp = g.rf();
    int mode = 0; /* 1 rot of particle 1, 2 rot of 2, 3 vib of 1, 4 vib of 2 */
    if (rfp < prot1) mode = 1;
    else if (rfp < prot1 + prot2) mode = 2;
    else if (rfp < prot1 + prot2 + pvib1) mode = 3;
    else if (rfp < prot1 + prot2 + pvib1 + pvib2) mode = 4;
    if (mode) {
        const double eold = (mode == 1) ? p.erot1 : (mode == 2) ? p.erot2 : (mode == 3) ? p.evib1 : p.evib2;
        const int dof = (mode == 1) ? rd1 : (mode == 2) ? rd2 : (mode == 3) ? vd1 : vd2;
        const double ecoll = etr + eold;
        const double enew = coll_internal_energy(g, ecoll, trdof, dof);
        etr = ecoll - enew;
        if (mode == 1) p.erot1 = enew;
        else if (mode == 2) p.erot2 = enew;
        else if (mode == 3) p.evib1 = enew;
        else p.evib2 = enew;
        flags |= (mode == 1 || mode == 3) ? PR_E1 : PR_E2;
    }

    /* ref: :495-546 VSS deflection */
    double coschi = 2. * pow(g.rf(), 1. / alpha) - 1.;
    double sinchi = sqrt(1. - coschi * coschi);
    double theta = PI2 * g.rf();
    double costheta = cos(theta), sintheta = sin(theta);
    double gx, gy, gz, gg;
    if (rfp < prot1 + prot2 + pvib1 + pvib2) {
        gg = sqrt(2. * etr / mred);
        double cosa = 2. * g.rf() - 1.0;
        double sina = sqrt(1 - cosa * cosa);
        double bb = PI2 * g.rf();
        gx = gg * sina * cos(bb);
        gy = gg * sina * sin(bb);
        gz = gg * cosa;
    } else {
        gx = c1[0] - c2[0];
        gy = c1[1] - c2[1];
        gz = c1[2] - c2[2];
        gg = sqrt(gx * gx + gy * gy + gz * gz);
    }
    double grel[3], wc[3];
    double gyz = sqrt(gy * gy + gz * gz);
    grel[0] = gx * coschi + gyz * sintheta * sinchi;
    grel[1] = gy * coschi + (gg * gz * costheta - gx * gy * sintheta) / gyz * sinchi;
    grel[2] = gz * coschi - (gg * gy * costheta + gx * gz * sintheta) / gyz * sinchi;
    for (int i = 0; i < 3; i++) wc[i] = m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i];
    for (int i = 0; i < 3; i++) {
        c1[i] = wc[i] + m2 / (m1 + m2) * grel[i];
        c2[i] = wc[i] - m1 / (m1 + m2) * grel[i];
    }
    return flags | PR_V;
}

/* the three pieces in sequence (GLOBAL path, MCC_COLLISIONS) */
__device__ int pair_collide(const PgDev& d, PgRng& g, PairIO& p, long long gslot2, int cell0, bool gate_all = false) {
    double ecoll = 0.;
    const int jr = pair_react_test(d, g, p, gate_all, ecoll);
    if (jr >= 0) return pair_react_apply(d, g, p, jr, ecoll, gslot2, cell0);
    return pair_inelastic_vss(d, g, p);
}

/* one candidate pair of VSS_COLLIS start to end (GLOBAL path) */
__device__ int collide_pair_core(const PgDev& d, PgRng& g, PairIO& p, long long gslot2, int cell0, double vrmax, double fcorr) {
    if (!pair_accept(d, g.rf(), p.c1, p.c2, p.s1, p.s2, vrmax, fcorr)) return 0;
    return pair_collide(d, g, p, gslot2, cell0);
}

/* ==========================================================================================
 * CELL path: one lane per cell
 * ======================================================================================== */

/* the deflection of an accepted pair without internal energy or reactions; same expressions as pair_collide.
 * g continues the candidate's stream right after the acceptance draw. */
__device__ __forceinline__ void pair_collide_fast(const PgDev& d, PgRng& g, double* c1, double* c2, int s1, int s2) {
    const double PI2 = 2. * PG_PI;
    double m1 = __ldg(&d.mass[s1 - 1]), m2 = __ldg(&d.mass[s2 - 1]);
    /* ref: :454-490 with no internal modes: the selector only decides whether the post-collision relative
     * velocity is redirected isotropically (:505-514; ROTREL/VIBREL != 0 on a species without modes) */
    const double psum = __ldg(&d.rotrel[s1 - 1]) + __ldg(&d.rotrel[s2 - 1]) + __ldg(&d.vibrel[s1 - 1]) + __ldg(&d.vibrel[s2 - 1]);
    /* (psum == 0: the selector cannot fall in any branch and its draw is not taken from the stream; an accepted
     *  elastic pair then needs exactly one more Philox block: deflection angle and azimuth) */
    const double rbl = (psum == 0.) ? 1. : g.rf();
    /* ref: :495-499 */
    const double alpha = VSS2(d.alphas, s1, s2);
    double coschi = 2. * pow(g.rf(), 1. / alpha) - 1.;
    double sinchi = sqrt(1. - coschi * coschi);
    double theta = PI2 * g.rf();
    double costheta = cos(theta), sintheta = sin(theta);
    double gx, gy, gz, gg;
    if (rbl < psum) {
        double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
        double mred = m1 * m2 / (m1 + m2);
        double etr = 0.5 * mred * vr2;
        double ecoll = etr + 0.;
        etr = ecoll - 0.;
        gg = sqrt(2. * etr / mred);
        double cosa = 2. * g.rf() - 1.0;
        double sina = sqrt(1 - cosa * cosa);
        double bb = PI2 * g.rf();
        gx = gg * sina * cos(bb);
        gy = gg * sina * sin(bb);
        gz = gg * cosa;
    } else {
        gx = c1[0] - c2[0];
        gy = c1[1] - c2[1];
        gz = c1[2] - c2[2];
        gg = sqrt(gx * gx + gy * gy + gz * gz);
    }
    double grel[3], wc[3];
    double gyz = sqrt(gy * gy + gz * gz);
    grel[0] = gx * coschi + gyz * sintheta * sinchi;
    grel[1] = gy * coschi + (gg * gz * costheta - gx * gy * sintheta) / gyz * sinchi;
    grel[2] = gz * coschi - (gg * gy * costheta + gx * gz * sintheta) / gyz * sinchi;
    for (int i = 0; i < 3; i++) wc[i] = m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i];
    for (int i = 0; i < 3; i++) {
        c1[i] = wc[i] + m2 / (m1 + m2) * grel[i];
        c2[i] = wc[i] - m1 / (m1 + m2) * grel[i];
    }
}

/* pair -> store, by what the collision touched */
template <bool IE>
__device__ __forceinline__ void write_pair(const PgDev& d, const PairIO& p, unsigned int j1, unsigned int j2, int fl) {
    if (fl & (PR_V | PR_S)) {
        PgVel W1, W2;
        W1.vx = p.c1[0]; W1.vy = p.c1[1]; W1.vz = p.c1[2]; W1.sid = p.s1; W1.pad = 0;
        W2.vx = p.c2[0]; W2.vy = p.c2[1]; W2.vz = p.c2[2]; W2.sid = p.s2; W2.pad = 0;
        d.p.vel[j1] = W1;
        d.p.vel[j2] = W2;
    }
    if (IE) {
        if (fl & PR_E1) { PgIe E; E.erot = p.erot1; E.evib = p.evib1; d.p.ie[j1] = E; }
        if (fl & PR_E2) { PgIe E; E.erot = p.erot2; E.evib = p.evib2; d.p.ie[j2] = E; }
        if (fl & PR_S) { /* the species mirror in the pos records */
            d.p.pos[j1].sid = p.s1;
            d.p.pos[j2].sid = p.s2;
        }
    }
}

/*
 * Every lane walks the candidates of ITS cell in the reference's order (so no dependency bookkeeping at all),
 * directly on the sorted store: a cell's records are a contiguous 32-byte-aligned run that stays in L1/L2
 * between the bounding-box sweep and the pair accesses.  The lanes of a warp are kept together where it pays:
 * each lane first runs the cheap part (pair selection + acceptance test, most candidates are rejected) until it
 * holds an ACCEPTED candidate, then the warp executes the expensive deflection (pow, sincos) for all of them at
 * once.  Cells with more than PG_TILE_BIG particles are left to the GLOBAL path.
 */
template <bool IE, bool GATHER>
__global__ void __launch_bounds__(128, IE ? PG_CELL_MINBLOCKS_IE : PG_CELL_MINBLOCKS) cell_collide_kernel(const __grid_constant__ PgDev d) {
    const unsigned int FULL = 0xffffffffu;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    /* cells are handed out grouped by particle count (pg_sort_impl): the lanes of a warp then walk similar numbers
     * of candidates instead of idling behind the fullest cell */
    const long long c = (tid < d.order_n) ? (d.cell_order ? (long long)__ldg(&d.cell_order[tid]) : tid + d.cells_lo) : (long long)d.ncells;
    unsigned int o0 = 0;
    int n = 0;
    if (c < d.ncells) {
        o0 = d.cell_off[c];
        n = (int)(d.cell_off[c + 1] - o0);
    }
    if (n > PG_TILE_BIG) n = 0; /* oversized cells: GLOBAL path (which also finishes their vel reorder) */
    if (!GATHER && n < 2) n = 0; /* ref: src/collisions.f90:129 */

    /* velocity bounding box, ref: :202-222.  GATHER: the sort left the velocity records where they were; this sweep
     * fetches them through the source index and writes the sorted plane (pg_sort_impl defer_vel), so a velocity
     * record crosses HBM once on its way to its cell instead of twice. */
    double vxmax = -1.e+38, vxmin = 1.e+38, vymax = -1.e+38, vymin = 1.e+38, vzmax = -1.e+38, vzmin = 1.e+38;
    {
        PgVel* __restrict__ vp = d.p.vel + o0;
        const unsigned int* __restrict__ so = d.src_of + o0;
        int i = 0;
        for (; i + 4 <= n; i += 4) {
            PgVel A, B, C, D;
            if (GATHER) {
                const unsigned int ia = __ldcs(so + i), ib = __ldcs(so + i + 1), ic = __ldcs(so + i + 2), id = __ldcs(so + i + 3);
                A = d.vel_src[ia]; B = d.vel_src[ib]; C = d.vel_src[ic]; D = d.vel_src[id];
                vp[i] = A; vp[i + 1] = B; vp[i + 2] = C; vp[i + 3] = D;
            } else {
                A = vp[i]; B = vp[i + 1]; C = vp[i + 2]; D = vp[i + 3];
            }
            vxmin = fmin(fmin(vxmin, A.vx), fmin(B.vx, fmin(C.vx, D.vx)));
            vxmax = fmax(fmax(vxmax, A.vx), fmax(B.vx, fmax(C.vx, D.vx)));
            vymin = fmin(fmin(vymin, A.vy), fmin(B.vy, fmin(C.vy, D.vy)));
            vymax = fmax(fmax(vymax, A.vy), fmax(B.vy, fmax(C.vy, D.vy)));
            vzmin = fmin(fmin(vzmin, A.vz), fmin(B.vz, fmin(C.vz, D.vz)));
            vzmax = fmax(fmax(vzmax, A.vz), fmax(B.vz, fmax(C.vz, D.vz)));
        }
        for (; i < n; i++) {
            PgVel A;
            if (GATHER) { A = d.vel_src[__ldcs(so + i)]; vp[i] = A; }
            else A = vp[i];
            vxmin = fmin(vxmin, A.vx); vxmax = fmax(vxmax, A.vx);
            vymin = fmin(vymin, A.vy); vymax = fmax(vymax, A.vy);
            vzmin = fmin(vzmin, A.vz); vzmax = fmax(vzmax, A.vz);
        }
    }
    if (GATHER && n < 2) n = 0; /* ref: src/collisions.f90:129 (a lone particle was only moved to its slot) */
    int ncoll = 0;
    double vrmax = 0., fcorr = 0.;
    if (n) ncoll = cell_majorant(d, (int)c, n, vxmin, vxmax, vymin, vymax, vzmin, vzmax, vrmax, fcorr);

    int nacc = 0, nreac = 0;
    int j = 0;
    /* state that survives a loop trip only for a lane parked on a reaction (IE) */
    int react_jr = -1;
    double react_ecoll = 0.;
    unsigned int j1 = 0, j2 = 0, jhit = 0, used = 4;
    PairIO p;
    PgRng rng;
    rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, 0ULL);
    while (__any_sync(FULL, j < ncoll || react_jr >= 0)) {
        /* cheap part: this lane's next candidates, up to and including its next accepted one; bounded, so that
         * the lanes that already hold an accepted candidate do not idle behind the unluckiest lane */
        bool hit = false;
        for (int tries = 0; tries < PG_CELL_TRIES && j < ncoll && !hit && react_jr < 0; tries++) {
            /* ref: :263-271.  The first block of the candidate's stream (PgRng contract: kind 2, cell, candidate)
             * holds everything a candidate needs: word 0 -> first index, word 1 -> second index, words 2,3 -> the
             * acceptance draw (n <= PG_TILE_BIG here, so the indices are one-word picks).  Only a repeated index
             * (probability 1/n) goes on through the general stream. */
            rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, (unsigned long long)j);
            uint32_t o[4];
            pg_philox(0u, rng.e0, rng.e1, rng.e2, rng.k0, rng.k1, o);
            const int k1 = (int)__umulhi(o[0], (uint32_t)n);
            int k2 = (int)__umulhi(o[1], (uint32_t)n);
            double racc = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * (1.0 / 9007199254740992.0);
            used = 4;
            if (k2 == k1) {
                rng.seek(2);
                while (k2 == k1) k2 = rng.pick(n);
                racc = rng.rf();
                used = rng.ndraw;
            }
            j1 = o0 + (unsigned int)k1;
            j2 = o0 + (unsigned int)k2;
            const PgVel V1 = d.p.vel[j1], V2 = d.p.vel[j2];
            p.c1[0] = V1.vx; p.c1[1] = V1.vy; p.c1[2] = V1.vz; p.s1 = V1.sid;
            p.c2[0] = V2.vx; p.c2[1] = V2.vy; p.c2[2] = V2.vz; p.s2 = V2.sid;
            hit = pair_accept(d, racc, p.c1, p.c2, p.s1, p.s2, vrmax, fcorr);
            jhit = (unsigned int)j;
            j++;
        }
        /* expensive part, for all lanes that hold an accepted candidate */
        if (hit) {
            nacc++;
            rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, (unsigned long long)jhit); /* the candidate's stream, */
            rng.seek(used);                                                                     /* after its acceptance draw */
            if (!IE) {
                pair_collide_fast(d, rng, p.c1, p.c2, p.s1, p.s2);
                PgVel W1, W2;
                W1.vx = p.c1[0]; W1.vy = p.c1[1]; W1.vz = p.c1[2]; W1.sid = p.s1; W1.pad = 0;
                W2.vx = p.c2[0]; W2.vy = p.c2[1]; W2.vz = p.c2[2]; W2.sid = p.s2; W2.pad = 0;
                d.p.vel[j1] = W1;
                d.p.vel[j2] = W2;
            } else {
                const PgIe E1 = d.p.ie[j1], E2 = d.p.ie[j2];
                p.erot1 = E1.erot; p.evib1 = E1.evib; p.erot2 = E2.erot; p.evib2 = E2.evib;
                react_jr = pair_react_test(d, rng, p, false, react_ecoll);
                if (react_jr < 0) {
                    const int fl = pair_inelastic_vss(d, rng, p);
                    write_pair<IE>(d, p, j1, j2, fl);
                }
            }
        }
        if (IE) {
            /* A reaction (energy redistribution over the products' modes, two isotropic scatterings, a third product)
             * is rare -- about one accepted pair in a hundred in hot air -- and about ten times longer than an inelastic
             * collision; run as it comes it would leave one or two lanes working while thirty wait (ncu r02f: 6 of 32
             * lanes active).  A lane whose pair reacts PARKS (its cell's later candidates must see the products, so
             * the lane does not go on) until a quarter of the warp is parked or nobody else has work; then the parked
             * lanes run the redistribution together. */
            const unsigned int parked = __ballot_sync(FULL, react_jr >= 0);
            if (parked) {
                const unsigned int working = __ballot_sync(FULL, react_jr < 0 && j < ncoll);
                if (__popc(parked) >= 8 || !working) {
                    if (react_jr >= 0) {
                        nreac++;
                        const int fl = pair_react_apply(d, rng, p, react_jr, react_ecoll, (long long)j2, (int)c);
                        write_pair<IE>(d, p, j1, j2, fl);
                        react_jr = -1;
                    }
                }
            }
        }
    }

#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        nacc += __shfl_xor_sync(FULL, nacc, s);
        nreac += __shfl_xor_sync(FULL, nreac, s);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nacc) atomicAdd(&d.status->ncoll, (unsigned long long)nacc);
        if (nreac) atomicAdd(&d.status->nreac, (unsigned long long)nreac);
    }
}

/* ==========================================================================================
 * MCC_COLLISIONS (VSS cross-section MCC against a background gas), ref: src/collisions.f90:1261-1595.
 * One thread per particle present at entry: per background component a partner is sampled from the background
 * Maxwellian (:1314-1316), the collision is accepted with P = 1 - exp(-DT n VR sigma (VR/CREF)^(1-2w)) (:1327),
 * and an accepted partner becomes a real particle (:1344-1351) that goes with the particle through the same
 * reaction / Borgnakke-Larsen / VSS code as a DSMC pair.  Everything is per particle, so there is no ordering
 * to preserve beyond the particle's own draw sequence: Philox stream kind 4 keyed by the slot.
 * ======================================================================================== */
__global__ void __launch_bounds__(128) mcc_vss_kernel(const __grid_constant__ PgDev d) {
    long long jp1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (jp1 >= (long long)d.status->np_cur) return; /* the particles present at entry (ref: :1646-1647) */
    const PgPos P1 = d.p.pos[jp1];
    const int ic = P1.ic;
    if (ic < 1) return;
    PgRng g;
    g.init(d.seed_local, 4, d.tID, (unsigned long long)jp1, 0ULL);
    const int sp1 = P1.sid; /* SP_ID1 is read once (:1291) and NOT refreshed after a reaction */
    int sid_now = sp1;
    const PgVel V1 = d.p.vel[jp1];
    const PgIe E1 = pg_ie_load(d.p.ie, jp1);
    PairIO p;
    p.c1[0] = V1.vx; p.c1[1] = V1.vy; p.c1[2] = V1.vz;
    p.erot1 = E1.erot; p.evib1 = E1.evib;
    int dirty = 0, ncoll = 0, nreac = 0;
    for (int j = 0; j < d.mcc_ncomp; j++) {
        const int s2 = d.mcc_sp[j];
        double bg_nrho;
        if (d.mcc_cell_nrho) bg_nrho = __ldg(&d.mcc_cell_nrho[(s2 - 1) + (size_t)d.ns * (size_t)(ic - 1)]); /* :1296-1297 */
        else bg_nrho = d.mcc_frac[j] * d.mcc_dens;                                                           /* :1299-1300 */
        if (bg_nrho == 0) continue;
        const double cref = VSS2(d.grefs, sp1, s2), sigma = VSS2(d.sigmas, sp1, s2), omega = VSS2(d.omegas, sp1, s2);
        pg_maxwell0(g, d.mcc_ttra, d.mass[s2 - 1], p.c2);
        const double vr2 = (p.c2[0] - p.c1[0]) * (p.c2[0] - p.c1[0]) + (p.c2[1] - p.c1[1]) * (p.c2[1] - p.c1[1]) +
                           (p.c2[2] - p.c1[2]) * (p.c2[2] - p.c1[2]);
        const double vr = sqrt(vr2);
        const double p_coll = 1 - exp(-d.dt * bg_nrho * vr * sigma * pow(vr / cref, 1. - 2. * omega));
        if (!(g.rf() < p_coll)) continue;
        ncoll++;
        p.erot2 = pg_internal_energy(g, d.rotdof[s2 - 1], d.mcc_ttra);
        p.evib2 = pg_internal_energy(g, d.vibdof[s2 - 1], d.mcc_ttra);
        p.s1 = sp1;
        p.s2 = s2;
        /* the partner sits at the particle's own position and cell (:1346-1347), which is where a third
         * product is created (:1478-1479) */
        const int fl = pair_collide(d, g, p, jp1, ic - 1, true);
        if (fl & PR_REACTED) nreac++;
        if (fl & PR_S) sid_now = p.s1;
        dirty |= fl;
        /* the partner is appended with its final state; the reference appends it first and then edits it */
        pg_append_particle(d, P1.x, P1.y, P1.z, p.c2[0], p.c2[1], p.c2[2], p.erot2, p.evib2, p.s2, ic);
    }
    if (dirty & (PR_V | PR_S)) {
        PgVel W;
        W.vx = p.c1[0]; W.vy = p.c1[1]; W.vz = p.c1[2]; W.sid = sid_now; W.pad = 0;
        d.p.vel[jp1] = W;
    }
    if (dirty & PR_E1) pg_ie_store(d.p.ie, jp1, p.erot1, p.evib1, d.status);
    if (dirty & PR_S) d.p.pos[jp1].sid = sid_now;
    if (ncoll) atomicAdd(&d.status->ncoll, (unsigned long long)ncoll);
    if (nreac) atomicAdd(&d.status->nreac, (unsigned long long)nreac);
}

/* ==========================================================================================
 * GLOBAL path: the cells of big_list (more than PG_TILE_BIG particles)
 * ======================================================================================== */
__device__ __forceinline__ unsigned long long enc_f64(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_f64(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
    return __longlong_as_double((long long)u);
}

__global__ void big_init_kernel(unsigned long long* bbox, unsigned int nbig) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbig * 6) return;
    bbox[i] = ((i % 6) < 3) ? ~0ULL : 0ULL; /* [0..2] minima, [3..5] maxima */
}

/* grid (nbig, Y): the Y CTAs of a cell stride over its particles */
__global__ void __launch_bounds__(256) big_bbox_kernel(const __grid_constant__ PgDev d, const unsigned int* __restrict__ big_list,
                                                       unsigned long long* __restrict__ bbox) {
    const unsigned int b = blockIdx.x, c = big_list[b];
    const unsigned int o0 = d.cell_off[c], o1 = d.cell_off[c + 1];
    double vxmax = -1.e+38, vxmin = 1.e+38, vymax = -1.e+38, vymin = 1.e+38, vzmax = -1.e+38, vzmin = 1.e+38;
    bool any = false;
    for (unsigned long long i = (unsigned long long)o0 + blockIdx.y * 256u + threadIdx.x; i < o1; i += (unsigned long long)gridDim.y * 256u) {
        PgVel V;
        if (d.src_of) { V = d.vel_src[d.src_of[i]]; d.p.vel[i] = V; } /* deferred vel reorder, see cell_collide_kernel */
        else V = d.p.vel[i];
        vxmin = fmin(vxmin, V.vx); vxmax = fmax(vxmax, V.vx);
        vymin = fmin(vymin, V.vy); vymax = fmax(vymax, V.vy);
        vzmin = fmin(vzmin, V.vz); vzmax = fmax(vzmax, V.vz);
        any = true;
    }
    if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        vxmin = fmin(vxmin, __shfl_xor_sync(0xffffffffu, vxmin, s));
        vxmax = fmax(vxmax, __shfl_xor_sync(0xffffffffu, vxmax, s));
        vymin = fmin(vymin, __shfl_xor_sync(0xffffffffu, vymin, s));
        vymax = fmax(vymax, __shfl_xor_sync(0xffffffffu, vymax, s));
        vzmin = fmin(vzmin, __shfl_xor_sync(0xffffffffu, vzmin, s));
        vzmax = fmax(vzmax, __shfl_xor_sync(0xffffffffu, vzmax, s));
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* q = bbox + 6 * (size_t)b;
        atomicMin(&q[0], enc_f64(vxmin)); atomicMin(&q[1], enc_f64(vymin)); atomicMin(&q[2], enc_f64(vzmin));
        atomicMax(&q[3], enc_f64(vxmax)); atomicMax(&q[4], enc_f64(vymax)); atomicMax(&q[5], enc_f64(vzmax));
    }
}

__global__ void big_prep_kernel(const __grid_constant__ PgDev d, const unsigned int* __restrict__ big_list, unsigned int nbig,
                                const unsigned long long* __restrict__ bbox, int* __restrict__ big_ncoll,
                                double* __restrict__ big_vrmax, double* __restrict__ big_fcorr) {
    unsigned int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nbig) return;
    if (b == nbig) { big_ncoll[b] = 0; return; } /* terminator of the exclusive scan */
    const unsigned int c = big_list[b];
    const int n = (int)(d.cell_off[c + 1] - d.cell_off[c]);
    const unsigned long long* q = bbox + 6 * (size_t)b;
    double vrmax, fcorr;
    big_ncoll[b] = cell_majorant(d, (int)c, n, dec_f64(q[0]), dec_f64(q[3]), dec_f64(q[1]), dec_f64(q[4]), dec_f64(q[2]), dec_f64(q[5]),
                                 vrmax, fcorr);
    big_vrmax[b] = vrmax;
    big_fcorr[b] = fcorr;
}

__global__ void __launch_bounds__(256) cand_head_kernel(unsigned int nbig, const int* __restrict__ big_ncoll,
                                                        const unsigned int* __restrict__ cand_off,
                                                        unsigned int* __restrict__ cand_big) {
    unsigned int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbig) return;
    if (big_ncoll[b] > 0) cand_big[cand_off[b]] = b;
}

/* pair selection; also the first claim round */
__global__ void __launch_bounds__(256) cand_pick_kernel(const __grid_constant__ PgDev d, long long ncand,
                                                        const unsigned int* __restrict__ big_list,
                                                        const unsigned int* __restrict__ cand_big,
                                                        const unsigned int* __restrict__ cand_off,
                                                        uint4* __restrict__ cand, unsigned int* __restrict__ owner) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncand) return;
    unsigned int b = cand_big[g];
    unsigned int c = big_list[b];
    unsigned int j = (unsigned int)(g - cand_off[b]);
    unsigned int o0 = d.cell_off[c];
    int npcj = (int)(d.cell_off[c + 1] - o0);
    int k1, k2;
    uint4 rec;
    rec.w = pick_pair(d, c, j, npcj, k1, k2);
    rec.x = o0 + (unsigned int)k1;
    rec.y = o0 + (unsigned int)k2;
    rec.z = b;
    cand[g] = rec;
    atomicMin(&owner[rec.x], (unsigned int)g);
    atomicMin(&owner[rec.y], (unsigned int)g);
}

__global__ void __launch_bounds__(256) cand_mark_kernel(long long ncand, const uint4* __restrict__ cand,
                                                        unsigned int* __restrict__ owner) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncand) return;
    uint4 rec = cand[g];
    if (rec.z & CAND_DONE) return;
    atomicMin(&owner[rec.x], (unsigned int)g);
    atomicMin(&owner[rec.y], (unsigned int)g);
}

__global__ void __launch_bounds__(256) cand_process_kernel(const __grid_constant__ PgDev d, long long ncand,
                                                           uint4* __restrict__ cand, unsigned int* __restrict__ owner,
                                                           const unsigned int* __restrict__ big_list,
                                                           const unsigned int* __restrict__ cand_off,
                                                           const double* __restrict__ big_vrmax,
                                                           const double* __restrict__ big_fcorr) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncand) return;
    uint4 rec = cand[g];
    if (rec.z & CAND_DONE) return;
    /* the claims were written by the previous kernel; releases made by this kernel race harmlessly
     * (a released entry is OWNER_FREE, never equal to g) */
    if (owner[rec.x] != (unsigned int)g || owner[rec.y] != (unsigned int)g) {
        atomicAdd(&d.status->pending, 1ULL);
        return;
    }
    const unsigned int b = rec.z, c = big_list[b];
    PgRng rng;
    rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, (unsigned long long)(g - cand_off[b]));
    rng.seek(rec.w); /* the pick consumed rec.w draws of this candidate's stream */
    PairIO p;
    const PgVel V1 = d.p.vel[rec.x], V2 = d.p.vel[rec.y];
    p.c1[0] = V1.vx; p.c1[1] = V1.vy; p.c1[2] = V1.vz; p.s1 = V1.sid;
    p.c2[0] = V2.vx; p.c2[1] = V2.vy; p.c2[2] = V2.vz; p.s2 = V2.sid;
    if (d.use_ie) {
        const PgIe E1 = d.p.ie[rec.x], E2 = d.p.ie[rec.y];
        p.erot1 = E1.erot; p.evib1 = E1.evib; p.erot2 = E2.erot; p.evib2 = E2.evib;
    } else {
        p.erot1 = p.evib1 = p.erot2 = p.evib2 = 0.;
    }
    const int fl = collide_pair_core(d, rng, p, (long long)rec.y, (int)c, big_vrmax[b], big_fcorr[b]);
    if (fl & PR_ACCEPTED) atomicAdd(&d.status->ncoll, 1ULL);
    if (fl & PR_REACTED) atomicAdd(&d.status->nreac, 1ULL);
    if (fl & (PR_V | PR_S)) {
        PgVel W1, W2;
        W1.vx = p.c1[0]; W1.vy = p.c1[1]; W1.vz = p.c1[2]; W1.sid = p.s1; W1.pad = 0;
        W2.vx = p.c2[0]; W2.vy = p.c2[1]; W2.vz = p.c2[2]; W2.sid = p.s2; W2.pad = 0;
        d.p.vel[rec.x] = W1;
        d.p.vel[rec.y] = W2;
    }
    if (d.use_ie) {
        if (fl & PR_E1) { PgIe E; E.erot = p.erot1; E.evib = p.evib1; d.p.ie[rec.x] = E; }
        if (fl & PR_E2) { PgIe E; E.erot = p.erot2; E.evib = p.evib2; d.p.ie[rec.y] = E; }
    }
    if (fl & PR_S) {
        d.p.pos[rec.x].sid = p.s1;
        d.p.pos[rec.y].sid = p.s2;
    }
    cand[g].z = rec.z | CAND_DONE;
    owner[rec.x] = OWNER_FREE;
    owner[rec.y] = OWNER_FREE;
}

} /* namespace */

static int ensure_big_scratch(pgpu_ctx* c) {
    long long nb = c->big_alloc + 1;
    if (nb > c->cellscratch_alloc) {
        cudaFree(c->d_cell_ncoll); cudaFree(c->d_cand_off); cudaFree(c->d_cell_vrmax); cudaFree(c->d_cell_fcorr);
        c->d_cell_ncoll = nullptr; c->d_cand_off = nullptr; c->d_cell_vrmax = nullptr; c->d_cell_fcorr = nullptr;
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_ncoll, sizeof(int) * (size_t)nb));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cand_off, sizeof(unsigned int) * (size_t)nb));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_vrmax, sizeof(double) * (size_t)nb));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_fcorr, sizeof(double) * (size_t)nb));
        c->cellscratch_alloc = (int)nb;
    }
    if (c->capacity > c->owner_capacity) {
        cudaFree(c->d_owner);
        c->d_owner = nullptr;
        PG_CUDA(c, cudaMalloc((void**)&c->d_owner, sizeof(unsigned int) * (size_t)c->capacity));
        PG_CUDA(c, cudaMemset(c->d_owner, 0xFF, sizeof(unsigned int) * (size_t)c->capacity));
        c->owner_capacity = c->capacity;
    }
    return PGPU_OK;
}

/* the GLOBAL path over big_list; synchronises with the host once per dependency round */
static int collide_big_cells(pgpu_ctx* c) {
    int rc;
    if ((rc = ensure_big_scratch(c))) return rc;
    const PgDev& d = c->dev;
    const unsigned int nbig = c->n_big;
    big_init_kernel<<<(nbig * 6 + 255) / 256, 256, 0, c->stream>>>(c->d_big_bbox, nbig);
    big_bbox_kernel<<<dim3(nbig, 32), 256, 0, c->stream>>>(d, c->d_big_list, c->d_big_bbox);
    big_prep_kernel<<<(nbig + 1 + 255) / 256, 256, 0, c->stream>>>(d, c->d_big_list, nbig, c->d_big_bbox, c->d_cell_ncoll,
                                                                 c->d_cell_vrmax, c->d_cell_fcorr);
    c->launches += 3;
    /* candidate offsets */
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const unsigned int*)c->d_cell_ncoll, c->d_cand_off, (int)nbig + 1, c->stream);
    if (tmp > c->scan_tmp_bytes) PG_FAIL(c, PGPU_ERR_STATE, "collide: scan workspace too small");
    tmp = c->scan_tmp_bytes;
    PG_CUDA(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, tmp, (const unsigned int*)c->d_cell_ncoll, c->d_cand_off,
                                             (int)nbig + 1, c->stream));
    c->launches++;
    unsigned int* h_total = &c->h_status->scratch[2];
    PG_CUDA(c, cudaMemcpyAsync(h_total, c->d_cand_off + nbig, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    long long ncand = (long long)*h_total;
    if (ncand == 0) return PGPU_OK;

    /* scratch: one uint4 record + one cell label per candidate, plus the max-scan workspace */
    if (ncand > c->cand_capacity) {
        cudaFree(c->d_cand);
        c->d_cand = nullptr;
        long long cap = std::max<long long>(ncand + ncand / 4, 1024);
        size_t scan_bytes = 0;
        cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (unsigned int*)nullptr, (unsigned int*)nullptr, MaxOp(), (int)cap, c->stream);
        size_t bytes = sizeof(uint4) * (size_t)cap + sizeof(unsigned int) * (size_t)cap + scan_bytes + 256;
        PG_CUDA(c, cudaMalloc(&c->d_cand, bytes));
        c->cand_capacity = cap;
    }
    uint4* cand = (uint4*)c->d_cand;
    unsigned int* cand_big = (unsigned int*)(cand + c->cand_capacity);
    void* scan_ws = (void*)(cand_big + c->cand_capacity);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, cand_big, cand_big, MaxOp(), (int)ncand, c->stream);
    scan_ws = (void*)(((uintptr_t)scan_ws + 255) & ~(uintptr_t)255);

    PG_CUDA(c, cudaMemsetAsync(cand_big, 0, sizeof(unsigned int) * (size_t)ncand, c->stream));
    cand_head_kernel<<<(nbig + 255) / 256, 256, 0, c->stream>>>(nbig, c->d_cell_ncoll, c->d_cand_off, cand_big);
    PG_CUDA(c, cub::DeviceScan::InclusiveScan(scan_ws, scan_bytes, cand_big, cand_big, MaxOp(), (int)ncand, c->stream));
    unsigned int cblocks = (unsigned int)((ncand + 255) / 256);
    cand_pick_kernel<<<cblocks, 256, 0, c->stream>>>(d, ncand, c->d_big_list, cand_big, c->d_cand_off, cand, c->d_owner);
    c->launches += 3;

    /* dependency rounds */
    for (int round = 0;; round++) {
        PG_CUDA(c, cudaMemsetAsync(&c->d_status->pending, 0, sizeof(unsigned long long), c->stream));
        cand_process_kernel<<<cblocks, 256, 0, c->stream>>>(d, ncand, cand, c->d_owner, c->d_big_list, c->d_cand_off,
                                                           c->d_cell_vrmax, c->d_cell_fcorr);
        c->launches++;
        PG_CUDA(c, cudaMemcpyAsync(&c->h_status->pending, &c->d_status->pending, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        PG_CUDA(c, cudaGetLastError());
        if (c->h_status->pending == 0) break;
        if (round > 10000000) PG_FAIL(c, PGPU_ERR_STATE, "collide: dependency rounds do not converge");
        cand_mark_kernel<<<cblocks, 256, 0, c->stream>>>(ncand, cand, c->d_owner);
        c->launches++;
    }
    return PGPU_OK;
}

/* DSMC_COLLISIONS on the sorted store; leaves TIMESTEP_COLL/REAC in the status block */
int pg_collide_big(pgpu_ctx* c) {
    int rc;
    pg_refresh_dev(c);
    if (c->n_big > 0 && (rc = collide_big_cells(c))) return rc;
    if (c->vel_deferred) { /* every cell has been swept: the sorted vel plane is complete */
        c->vel_deferred = false;
        pg_refresh_dev(c);
    }
    return PGPU_OK;
}

/* big_now: the number of oversized cells is known on the host (the sort synchronised); otherwise the caller runs
 * pg_collide_big after the step's read-back */
int pg_collide_impl(pgpu_ctx* c, int tID, bool big_now) {
    if (!c->sorted) PG_FAIL(c, PGPU_ERR_STATE, "collide: store is not sorted");
    if (c->np_ub < 2) return pg_resolve_vel(c);
    c->dev.tID = tID;
    pg_refresh_dev(c);
    const PgDev& d = c->dev;

    {
        unsigned int blocks = (unsigned int)((d.order_n + 127) / 128); /* (the cells the sort covered: pg_cell_range) */
        const bool gather = c->vel_deferred;
        if (d.use_ie) {
            if (gather) cell_collide_kernel<true, true><<<blocks, 128, 0, c->stream>>>(d);
            else cell_collide_kernel<true, false><<<blocks, 128, 0, c->stream>>>(d);
        } else {
            if (gather) cell_collide_kernel<false, true><<<blocks, 128, 0, c->stream>>>(d);
            else cell_collide_kernel<false, false><<<blocks, 128, 0, c->stream>>>(d);
        }
        c->launches++;
        PG_CUDA(c, cudaGetLastError());
    }
    c->hist_valid = false;
    if (big_now) return pg_collide_big(c);
    return PGPU_OK;
}

/* MCC_COLLISIONS for all slots [0, np); products and partners go to the append region */
int pg_mcc_vss_impl(pgpu_ctx* c, int tID) {
    if (!c->dev.mcc_ncomp) PG_FAIL(c, PGPU_ERR_STATE, "mcc_collisions: background mixture not set");
    if (!c->dev.sigmas) PG_FAIL(c, PGPU_ERR_STATE, "mcc_collisions: VSS tables not set");
    if (c->dev.mcc_cell_nrho && c->mcc_cell_nrho_cells < (int64_t)pg_ncells(c))
        PG_FAIL(c, PGPU_ERR_STATE, "mcc_collisions: background density table covers %lld of %d cells",
                (long long)c->mcc_cell_nrho_cells, pg_ncells(c));
    if (c->np_ub == 0) return PGPU_OK;
    { int rc0; if ((rc0 = pg_resolve_vel(c))) return rc0; }
    c->dev.tID = tID;
    pg_refresh_dev(c);
    mcc_vss_kernel<<<(unsigned int)((c->np_ub + 127) / 128), 128, 0, c->stream>>>(c->dev);
    c->launches++;
    PG_CUDA(c, cudaGetLastError());
    c->sorted = false;
    c->hist_valid = false;
    return PGPU_OK;
}

extern "C" int pgpu_mcc_collisions(pgpu_ctx* c, int32_t tID, int64_t* timestep_coll, int64_t* timestep_reac) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    memset(c->ms, 0, sizeof(c->ms));
    c->launches = 0;
    if ((rc = pg_reset_status(c))) return rc;
    cudaEventRecord(c->ev[4], c->stream);
    if ((rc = pg_mcc_vss_impl(c, tID))) return rc;
    cudaEventRecord(c->ev[5], c->stream);
    rc = pg_finish(c, "pgpu_mcc_collisions");
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->ms[2] = ms;
    c->ms[5] = c->ms[2];
    if (timestep_coll) *timestep_coll = (int64_t)c->h_status->ncoll;
    if (timestep_reac) *timestep_reac = (int64_t)c->h_status->nreac;
    return rc;
}
