/*
 * pgpu_collide.cu -- DSMC collisions on the cell-sorted SoA store.
 *
 * Replaces: the cell loop of DSMC_COLLISIONS (src/collisions.f90:126-138), VSS_COLLIS (:155-558),
 * FUNCTION_I (:892-908), COLL_INTERNAL_ENERGY (:915-937), HS_SCATTER (:956-992).
 *
 * The reference walks the NCOLL candidate pairs of a cell one after another, and a later candidate sees
 * the velocities an earlier accepted one wrote (:259-551).  That order is kept EXACTLY, in parallel:
 *   1. cell_prep     8 lanes per cell: velocity bounding box -> VRMAX, NCOLL, FCORR (:202-253)
 *   2. scan          NCOLL -> first candidate of each cell; a max-scan labels every candidate with its cell
 *   3. pick          one thread per candidate: the pair (JP1, JP2) depends only on the candidate's own
 *                    counter-based Philox stream (:263-271), never on particle state
 *   4. rounds        every waiting candidate atomically claims both of its particles with its sequence
 *                    number (atomicMin); a candidate that holds both claims is the earliest one touching
 *                    either particle, so evaluating it now is what the sequential loop would do.  Evaluate,
 *                    release, repeat until none waits.  Cells, and disjoint pairs inside a cell, proceed in
 *                    parallel; chains of candidates sharing a particle serialise, as they must.
 * With the oracle run on the same Philox streams the results agree value for value (tests/test_collide_*).
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

/* ref: src/collisions.f90:202-253 */
__global__ void __launch_bounds__(256) cell_prep_kernel(const __grid_constant__ PgDev d, int* __restrict__ cell_ncoll,
                                                        double* __restrict__ cell_vrmax,
                                                        double* __restrict__ cell_fcorr) {
    const int G = 8;
    long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int l = threadIdx.x % G;
    bool valid = gid < d.ncells;
    unsigned int o0 = 0, o1 = 0;
    if (valid) {
        o0 = d.cell_off[gid];
        o1 = d.cell_off[gid + 1];
    }
    int n = (int)(o1 - o0);
    double vxmax = -1.e+38, vxmin = 1.e+38, vymax = -1.e+38, vymin = 1.e+38, vzmax = -1.e+38, vzmin = 1.e+38;
    if (n > 1) {
        for (unsigned int i = o0 + l; i < o1; i += G) {
            double vx = d.p.vx[i], vy = d.p.vy[i], vz = d.p.vz[i];
            vxmin = fmin(vxmin, vx); vxmax = fmax(vxmax, vx);
            vymin = fmin(vymin, vy); vymax = fmax(vymax, vy);
            vzmin = fmin(vzmin, vz); vzmax = fmax(vzmax, vz);
        }
    }
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) {
        vxmin = fmin(vxmin, __shfl_xor_sync(0xffffffffu, vxmin, s, G));
        vxmax = fmax(vxmax, __shfl_xor_sync(0xffffffffu, vxmax, s, G));
        vymin = fmin(vymin, __shfl_xor_sync(0xffffffffu, vymin, s, G));
        vymax = fmax(vymax, __shfl_xor_sync(0xffffffffu, vymax, s, G));
        vzmin = fmin(vzmin, __shfl_xor_sync(0xffffffffu, vzmin, s, G));
        vzmax = fmax(vzmax, __shfl_xor_sync(0xffffffffu, vzmax, s, G));
    }
    if (!valid || l != 0) return;
    if (n <= 1) { /* ref: src/collisions.f90:129 only cells with more than one particle collide */
        cell_ncoll[gid] = 0;
        return;
    }
    double vrmax = sqrt((vxmax - vxmin) * (vxmax - vxmin) + (vymax - vymin) * (vymax - vymin) +
                        (vzmax - vzmin) * (vzmax - vzmin));
    /* ref: :224-230 */
    double vol;
    if (d.grid_type == PGPU_UNSTRUCTURED) vol = d.cell_vol[gid];
    else vol = (d.xmax - d.xmin) * (d.ymax - d.ymin) * (d.zmax - d.zmin) / (d.nx * d.ny);
    /* ref: :232  0.5*NPC*(NPC-1) is evaluated in single precision */
    float pairs_f = (0.5f * (float)n) * (float)(n - 1);
    double ncollmax = (double)pairs_f * d.sigmamax * vrmax * d.fnum * d.dt / vol;
    double r = floor(ncollmax + 0.5);
    long long half = (long long)floorf(0.5f * (float)n);
    long long ncoll;
    if (!(r >= 1.)) ncoll = 1;
    else if (r > (double)half) ncoll = half;
    else ncoll = (long long)r;
    cell_ncoll[gid] = (int)ncoll;
    cell_vrmax[gid] = vrmax;
    cell_fcorr[gid] = ncollmax / (double)ncoll;
}

__global__ void __launch_bounds__(256) cand_head_kernel(int ncells, const int* __restrict__ cell_ncoll,
                                                        const unsigned int* __restrict__ cand_off,
                                                        unsigned int* __restrict__ cand_cell) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    if (cell_ncoll[c] > 0) cand_cell[cand_off[c]] = (unsigned int)c;
}

/* ref: src/collisions.f90:263-271 pair selection; also the first claim round */
__global__ void __launch_bounds__(256) cand_pick_kernel(const __grid_constant__ PgDev d, long long ncand,
                                                        const unsigned int* __restrict__ cand_cell,
                                                        const unsigned int* __restrict__ cand_off,
                                                        uint4* __restrict__ cand, unsigned int* __restrict__ owner) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncand) return;
    unsigned int c = cand_cell[g];
    unsigned int j = (unsigned int)(g - cand_off[c]);
    unsigned int o0 = d.cell_off[c];
    int npcj = (int)(d.cell_off[c + 1] - o0);
    PgRng rng;
    rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, (unsigned long long)j);
    int k1 = (int)(npcj * rng.rf());
    if (k1 >= npcj) k1 = npcj - 1;
    int k2 = k1;
    while (k2 == k1) {
        k2 = (int)(npcj * rng.rf());
        if (k2 >= npcj) k2 = npcj - 1;
    }
    uint4 rec;
    rec.x = o0 + (unsigned int)k1;
    rec.y = o0 + (unsigned int)k2;
    rec.z = c;
    rec.w = rng.ndraw;
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

#define VSS2(arr, s1, s2) __ldg(&(arr)[((s1)-1) + (size_t)d.ns * ((s2)-1)])

/* one candidate pair of VSS_COLLIS, ref: src/collisions.f90:274-546 */
__device__ void collide_pair(const PgDev& d, PgRng& g, unsigned int jp1, unsigned int jp2, int cell0, double vrmax,
                             double fcorr, int& accepted, int& reacted) {
    const double PI2 = 2. * PG_PI;
    double c1[3] = {d.p.vx[jp1], d.p.vy[jp1], d.p.vz[jp1]};
    double c2[3] = {d.p.vx[jp2], d.p.vy[jp2], d.p.vz[jp2]};
    int s1 = d.p.sid[jp1], s2 = d.p.sid[jp2];
    double vr2 = (c2[0] - c1[0]) * (c2[0] - c1[0]) + (c2[1] - c1[1]) * (c2[1] - c1[1]) + (c2[2] - c1[2]) * (c2[2] - c1[2]);
    double vr = sqrt(vr2);
    double cref = VSS2(d.grefs, s1, s2), sigma = VSS2(d.sigmas, s1, s2);
    double omega = VSS2(d.omegas, s1, s2), alpha = VSS2(d.alphas, s1, s2);
    double m1 = __ldg(&d.mass[s1 - 1]), m2 = __ldg(&d.mass[s2 - 1]);
    double mred = m1 * m2 / (m1 + m2);
    double etr = 0.5 * mred * vr2;

    double rfp = g.rf();
    double collsigma;
    if (fabs(omega - 0.5) < 1.e-6) {
        collsigma = sigma;
    } else {
        if (vr < (double)1e-20f) return; /* ref: :314 (single-precision literal) */
        collsigma = sigma * pow(vr / cref, 1. - 2. * omega);
    }
    double collprob = fcorr / (d.sigmamax * vrmax) * vr * collsigma;
    if (!(rfp < collprob)) return;
    accepted = 1;

    int rd1 = __ldg(&d.rotdof[s1 - 1]), rd2 = __ldg(&d.rotdof[s2 - 1]);
    int vd1 = __ldg(&d.vibdof[s1 - 1]), vd2 = __ldg(&d.vibdof[s2 - 1]);
    bool internal = (d.nr > 0) || rd1 || rd2 || vd1 || vd2;
    double erot1 = 0., evib1 = 0., erot2 = 0., evib2 = 0.;
    if (internal) {
        erot1 = d.p.erot[jp1]; evib1 = d.p.evib[jp1];
        erot2 = d.p.erot[jp2]; evib2 = d.p.evib[jp2];
    }

    /* ref: :331-449 TCE / LXCat reaction test */
    double ecoll = etr + erot1 + erot2 + evib1 + evib2;
    for (int jr = 0; jr < d.nr; jr++) {
        const PgReaction r = d.reac[jr];
        if (r.r1 != s1) {
            if (r.r1 != s2) continue;
            else if (r.r2 != s1) continue;
        } else {
            if (r.r2 != s2) continue;
        }
        double ea = r.ea, ptce;
        if (r.type == PGPU_TCE) {
            if (ecoll <= ea) continue;
            ptce = r.c1 * pow(ecoll - ea, r.c2) * pow(1. - ea / ecoll, r.c3);
        } else if (r.type == PGPU_LXCAT) {
            double sigma_r = pg_interp_cs(etr, r.tab_n, d.tab_e + r.tab_off, d.tab_cs + r.tab_off);
            ptce = sigma_r / (sigma * pow(vr / cref, 1. - 2. * omega));
        } else {
            ptce = 0;
        }
        rfp = g.rf();
        if (rfp < ptce) {
            reacted = 1;
            ecoll = ecoll - ea;
            /* roles are NOT swapped when the pair order is opposite to the reaction */
            pg_react_redistribute(d, g, r, ecoll, c1, c2, erot1, evib1, erot2, evib2, d.p.x[jp2], d.p.y[jp2],
                                  d.p.z[jp2], cell0 + 1);
            d.p.sid[jp1] = r.p1;
            d.p.sid[jp2] = r.p2;
            d.p.erot[jp1] = erot1; d.p.evib[jp1] = evib1;
            d.p.erot[jp2] = erot2; d.p.evib[jp2] = evib2;
            d.p.vx[jp1] = c1[0]; d.p.vy[jp1] = c1[1]; d.p.vz[jp1] = c1[2];
            d.p.vx[jp2] = c2[0]; d.p.vy[jp2] = c2[1]; d.p.vz[jp2] = c2[2];
            return;
        }
    }

    /* ref: :454-490 Borgnakke-Larsen, at most one mode */
    double trdof = 3.;
    double prot1 = (trdof + rd1) / trdof * __ldg(&d.rotrel[s1 - 1]);
    double prot2 = (trdof + rd2) / trdof * __ldg(&d.rotrel[s2 - 1]);
    double pvib1 = (trdof + vd1) / trdof * __ldg(&d.vibrel[s1 - 1]);
    double pvib2 = (trdof + vd2) / trdof * __ldg(&d.vibrel[s2 - 1]);
    trdof = 5. - 2. * omega;
    rfp = g.rf();
    if (rfp < prot1) {
        ecoll = etr + erot1;
        erot1 = coll_internal_energy(g, ecoll, trdof, rd1);
        etr = ecoll - erot1;
        d.p.erot[jp1] = erot1;
    } else if (rfp < prot1 + prot2) {
        ecoll = etr + erot2;
        erot2 = coll_internal_energy(g, ecoll, trdof, rd2);
        etr = ecoll - erot2;
        d.p.erot[jp2] = erot2;
    } else if (rfp < prot1 + prot2 + pvib1) {
        ecoll = etr + evib1;
        evib1 = coll_internal_energy(g, ecoll, trdof, vd1);
        etr = ecoll - evib1;
        d.p.evib[jp1] = evib1;
    } else if (rfp < prot1 + prot2 + pvib1 + pvib2) {
        ecoll = etr + evib2;
        evib2 = coll_internal_energy(g, ecoll, trdof, vd2);
        etr = ecoll - evib2;
        d.p.evib[jp2] = evib2;
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
    d.p.vx[jp1] = c1[0]; d.p.vy[jp1] = c1[1]; d.p.vz[jp1] = c1[2];
    d.p.vx[jp2] = c2[0]; d.p.vy[jp2] = c2[1]; d.p.vz[jp2] = c2[2];
}

__global__ void __launch_bounds__(256) cand_process_kernel(const __grid_constant__ PgDev d, long long ncand,
                                                           uint4* __restrict__ cand, unsigned int* __restrict__ owner,
                                                           const unsigned int* __restrict__ cand_off,
                                                           const double* __restrict__ cell_vrmax,
                                                           const double* __restrict__ cell_fcorr) {
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
    unsigned int c = rec.z;
    PgRng rng;
    rng.init(d.seed_local, 2, d.tID, (unsigned long long)c, (unsigned long long)(g - cand_off[c]));
    rng.seek(rec.w); /* the pick consumed rec.w draws of this candidate's stream */
    int accepted = 0, reacted = 0;
    collide_pair(d, rng, rec.x, rec.y, (int)c, cell_vrmax[c], cell_fcorr[c], accepted, reacted);
    if (accepted) atomicAdd(&d.status->ncoll, 1ULL);
    if (reacted) atomicAdd(&d.status->nreac, 1ULL);
    cand[g].z = rec.z | CAND_DONE;
    owner[rec.x] = OWNER_FREE;
    owner[rec.y] = OWNER_FREE;
}

} /* namespace */

static int ensure_collide_scratch(pgpu_ctx* c) {
    int ncells = pg_ncells(c);
    if (ncells > c->cellscratch_alloc) {
        cudaFree(c->d_cell_ncoll); cudaFree(c->d_cand_off); cudaFree(c->d_cell_vrmax); cudaFree(c->d_cell_fcorr);
        c->d_cell_ncoll = nullptr; c->d_cand_off = nullptr; c->d_cell_vrmax = nullptr; c->d_cell_fcorr = nullptr;
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_ncoll, sizeof(int) * ((size_t)ncells + 1)));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cand_off, sizeof(unsigned int) * ((size_t)ncells + 1)));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_vrmax, sizeof(double) * (size_t)ncells));
        PG_CUDA(c, cudaMalloc((void**)&c->d_cell_fcorr, sizeof(double) * (size_t)ncells));
        PG_CUDA(c, cudaMemset(c->d_cell_ncoll, 0, sizeof(int) * ((size_t)ncells + 1)));
        c->cellscratch_alloc = ncells;
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

/* DSMC_COLLISIONS on the sorted store; leaves TIMESTEP_COLL/REAC in the status block */
int pg_collide_impl(pgpu_ctx* c, int tID) {
    int rc;
    if (!c->sorted) PG_FAIL(c, PGPU_ERR_STATE, "collide: store is not sorted");
    if (c->np < 2) return PGPU_OK;
    if ((rc = ensure_collide_scratch(c))) return rc;
    c->dev.tID = tID;
    pg_refresh_dev(c);
    const PgDev& d = c->dev;
    int ncells = d.ncells;
    const int G = 8;
    unsigned int blocks = (unsigned int)(((long long)ncells * G + 255) / 256);
    cell_prep_kernel<<<blocks, 256, 0, c->stream>>>(d, c->d_cell_ncoll, c->d_cell_vrmax, c->d_cell_fcorr);
    c->launches++;
    /* candidate offsets; the scan workspace sized for ncells+1 unsigned ints is reused */
    size_t tmp = c->scan_tmp_bytes;
    PG_CUDA(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp, tmp, (const unsigned int*)c->d_cell_ncoll, c->d_cand_off,
                                             ncells + 1, c->stream));
    c->launches++;
    unsigned int* h_total = reinterpret_cast<unsigned int*>(&c->h_status->pad);
    PG_CUDA(c, cudaMemcpyAsync(h_total, c->d_cand_off + ncells, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
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
    unsigned int* cand_cell = (unsigned int*)(cand + c->cand_capacity);
    void* scan_ws = (void*)(cand_cell + c->cand_capacity);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, cand_cell, cand_cell, MaxOp(), (int)ncand, c->stream);
    scan_ws = (void*)(((uintptr_t)scan_ws + 255) & ~(uintptr_t)255);

    PG_CUDA(c, cudaMemsetAsync(cand_cell, 0, sizeof(unsigned int) * (size_t)ncand, c->stream));
    cand_head_kernel<<<(unsigned int)((ncells + 255) / 256), 256, 0, c->stream>>>(ncells, c->d_cell_ncoll, c->d_cand_off, cand_cell);
    PG_CUDA(c, cub::DeviceScan::InclusiveScan(scan_ws, scan_bytes, cand_cell, cand_cell, MaxOp(), (int)ncand, c->stream));
    unsigned int cblocks = (unsigned int)((ncand + 255) / 256);
    cand_pick_kernel<<<cblocks, 256, 0, c->stream>>>(d, ncand, cand_cell, c->d_cand_off, cand, c->d_owner);
    c->launches += 3;

    /* dependency rounds */
    for (int round = 0;; round++) {
        PG_CUDA(c, cudaMemsetAsync(&c->d_status->pending, 0, sizeof(unsigned long long), c->stream));
        cand_process_kernel<<<cblocks, 256, 0, c->stream>>>(d, ncand, cand, c->d_owner, c->d_cand_off, c->d_cell_vrmax,
                                                           c->d_cell_fcorr);
        c->launches++;
        PG_CUDA(c, cudaMemcpyAsync(&c->h_status->pending, &c->d_status->pending, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        PG_CUDA(c, cudaGetLastError());
        if (c->h_status->pending == 0) break;
        if (round > 100000) PG_FAIL(c, PGPU_ERR_STATE, "collide: dependency rounds do not converge");
        cand_mark_kernel<<<cblocks, 256, 0, c->stream>>>(ncand, cand, c->d_owner);
        c->launches++;
    }
    c->hist_valid = false;
    return PGPU_OK;
}
