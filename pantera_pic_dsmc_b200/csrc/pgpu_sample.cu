/*
 * pgpu_sample.cu -- flow-field sampling and conservation checks on the device.
 *
 * Replaces: GRID_AVG (src/postprocess.f90:34-416; BOOL_DUMP_MOMENTS = .FALSE., the reference's default), GRID_RESET
 * (:1070-1094) and the particle loop of CHECKS (:1722-1800).  Without these every sample costs a download of the whole
 * particle store (12 GB at the bench size); here a sample is one pass over the record planes and the running averages
 * AVG_NP, AVG_VX/VY/VZ, AVG_TTRX/Y/Z, AVG_TROT, AVG_TVIB stay on the device until the host asks for them.
 *
 * Per (cell, species) bin INDEX = JC + NCELLS*(JS-1) the reference accumulates NP, sum V, sum V^2, sum EROT, sum EVIB
 * sequentially over the particles; here the sums are FP64 atomics (order differs -> compare at 1e-12 relative of the
 * summed magnitudes, like DEPOSIT_CHARGE).  The per-bin arithmetic that follows (means, temperatures, running
 * averages with their per-bin sample counters) is evaluated with the reference's expressions in the reference's order.
 * Several ranks: a cell is sampled by the rank that owns it; pgpu_get_grid_avg sums the ranks' arrays (ncclAllReduce),
 * which is what the reference's MPI_REDUCE(SUM) of the per-step arrays amounts to for a cell-partitioned domain.
 */
#include <algorithm>

#include "pgpu_internal.cuh"

int pg_allreduce_sum_f64(pgpu_ctx* c, double* buf, long long n); /* pgpu_exchange.cu */

namespace {

enum { GA_VX = 0, GA_VY, GA_VZ, GA_VX2, GA_VY2, GA_VZ2, GA_EROT, GA_EVIB, GA_NSUM };
enum { AV_NP = 0, AV_VX, AV_VY, AV_VZ, AV_TTRX, AV_TTRY, AV_TTRZ, AV_TROT, AV_TVIB, AV_N };

__global__ void __launch_bounds__(256) gavg_accum_kernel(const __grid_constant__ PgDev d, long long nbins, unsigned int* __restrict__ np,
                                                         double* __restrict__ sums) {
    const long long n = (long long)d.status->np_cur;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const PgPos P = d.p.pos[i];
        if (P.ic < 1) continue;
        const PgVel V = d.p.vel[i];
        const long long idx = (long long)(P.ic - 1) + (long long)d.ncells * (long long)(P.sid - 1);
        atomicAdd(&np[idx], 1u);
        atomicAdd(&sums[GA_VX * nbins + idx], V.vx);
        atomicAdd(&sums[GA_VY * nbins + idx], V.vy);
        atomicAdd(&sums[GA_VZ * nbins + idx], V.vz);
        atomicAdd(&sums[GA_VX2 * nbins + idx], V.vx * V.vx);
        atomicAdd(&sums[GA_VY2 * nbins + idx], V.vy * V.vy);
        atomicAdd(&sums[GA_VZ2 * nbins + idx], V.vz * V.vz);
        if (d.use_ie) {
            const PgIe E = d.p.ie[i];
            atomicAdd(&sums[GA_EROT * nbins + idx], E.erot);
            atomicAdd(&sums[GA_EVIB * nbins + idx], E.evib);
        }
    }
}

/* ref: src/postprocess.f90:147-155 (means), :262-297 (temperatures), :344-381 (running averages) */
__global__ void __launch_bounds__(256) gavg_update_kernel(const __grid_constant__ PgDev d, long long nbins, const unsigned int* __restrict__ np,
                                                          const double* __restrict__ sums, double* __restrict__ avg,
                                                          int* __restrict__ cum1, int* __restrict__ cum2, int avg_cumulated) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nbins) return;
    const int js = (int)(idx / d.ncells);
    const unsigned int tnp = np[idx];
    /* AVG_NP = (AVG_NP*DBLE(AVG_CUMULATED) + DBLE(TIMESTEP_NP))/(AVG_CUMULATED + 1.)   (the divisor is single precision) */
    const double dcum = (double)avg_cumulated;
    avg[AV_NP * nbins + idx] = (avg[AV_NP * nbins + idx] * dcum + (double)tnp) / (double)((float)avg_cumulated + 1.f);
    if (tnp == 0) return;
    const double numpart = (double)tnp;
    const double vx = sums[GA_VX * nbins + idx] / numpart, vy = sums[GA_VY * nbins + idx] / numpart, vz = sums[GA_VZ * nbins + idx] / numpart;
    const int rotdof = d.rotdof[js], vibdof = d.vibdof[js];
    double trot = 0.0, tvib = 0.0;
    if (rotdof != 0) trot = (double)(2.f / (float)rotdof) / PG_KB * (sums[GA_EROT * nbins + idx] / numpart);
    if (vibdof != 0) tvib = (double)(2.f / (float)vibdof) / PG_KB * (sums[GA_EVIB * nbins + idx] / numpart);
    {
        const int k = cum1[idx];
        const double dk = (double)k, den = (double)((float)k + 1.f);
        avg[AV_VX * nbins + idx] = (avg[AV_VX * nbins + idx] * dk + vx) / den;
        avg[AV_VY * nbins + idx] = (avg[AV_VY * nbins + idx] * dk + vy) / den;
        avg[AV_VZ * nbins + idx] = (avg[AV_VZ * nbins + idx] * dk + vz) / den;
        avg[AV_TROT * nbins + idx] = (avg[AV_TROT * nbins + idx] * dk + trot) / den;
        avg[AV_TVIB * nbins + idx] = (avg[AV_TVIB * nbins + idx] * dk + tvib) / den;
        cum1[idx] = k + 1;
    }
    if (tnp > 1) {
        const double m_kb = d.mass[js] / PG_KB;
        const double sampledof = numpart;
        const double ttrx = m_kb * (sums[GA_VX2 * nbins + idx] - numpart * vx * vx) / sampledof;
        const double ttry = m_kb * (sums[GA_VY2 * nbins + idx] - numpart * vy * vy) / sampledof;
        const double ttrz = m_kb * (sums[GA_VZ2 * nbins + idx] - numpart * vz * vz) / sampledof;
        const int k = cum2[idx];
        const double dk = (double)k, den = (double)((float)k + 1.f);
        avg[AV_TTRX * nbins + idx] = (avg[AV_TTRX * nbins + idx] * dk + ttrx) / den;
        avg[AV_TTRY * nbins + idx] = (avg[AV_TTRY * nbins + idx] * dk + ttry) / den;
        avg[AV_TTRZ * nbins + idx] = (avg[AV_TTRZ * nbins + idx] * dk + ttrz) / den;
        cum2[idx] = k + 1;
    }
}

/* CHECKS particle loop, ref: src/postprocess.f90:1752-1786: per species count, momentum, kinetic energy; total internal
 * energy.  out: [ns] counts as double | [3*ns] momentum (k + 3*(js-1)) | [ns] kinetic energy | [1] internal energy */
__global__ void __launch_bounds__(256) checks_kernel(const __grid_constant__ PgDev d, double* __restrict__ out) {
    extern __shared__ double sh[]; /* [ns*5 + 1] */
    const int ns = d.ns, nacc = ns * 5 + 1;
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) sh[i] = 0.;
    __syncthreads();
    const long long n = (long long)d.status->np_cur;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const PgPos P = d.p.pos[i];
        if (P.ic < 1) continue;
        const PgVel V = d.p.vel[i];
        const int js = P.sid - 1;
        const double m = d.mass[js], cf = d.cell_fnum ? d.cell_fnum[P.ic - 1] : d.fnum; /* ref: :1764-1768 */
        atomicAdd(&sh[js], 1.0);
        atomicAdd(&sh[ns + 3 * js + 0], m * V.vx * cf);
        atomicAdd(&sh[ns + 3 * js + 1], m * V.vy * cf);
        atomicAdd(&sh[ns + 3 * js + 2], m * V.vz * cf);
        atomicAdd(&sh[4 * ns + js], 0.5 * m * (V.vx * V.vx + V.vy * V.vy + V.vz * V.vz) * cf);
        if (d.use_ie) {
            const PgIe E = d.p.ie[i];
            atomicAdd(&sh[5 * ns], (E.erot + E.evib) * cf);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += blockDim.x)
        if (sh[i] != 0.) atomicAdd(&out[i], sh[i]);
}

} /* namespace */

static int gavg_alloc(pgpu_ctx* c) {
    const long long nbins = (long long)pg_ncells(c) * c->ns;
    if (c->gavg_bins == nbins && c->d_gavg) return PGPU_OK;
    cudaFree(c->d_gavg); cudaFree(c->d_gavg_np); cudaFree(c->d_gavg_sums); cudaFree(c->d_gavg_cum);
    c->d_gavg = nullptr; c->d_gavg_np = nullptr; c->d_gavg_sums = nullptr; c->d_gavg_cum = nullptr;
    PG_CUDA(c, cudaMalloc((void**)&c->d_gavg, sizeof(double) * AV_N * (size_t)nbins));
    PG_CUDA(c, cudaMalloc((void**)&c->d_gavg_np, sizeof(unsigned int) * (size_t)nbins));
    PG_CUDA(c, cudaMalloc((void**)&c->d_gavg_sums, sizeof(double) * GA_NSUM * (size_t)nbins));
    PG_CUDA(c, cudaMalloc((void**)&c->d_gavg_cum, sizeof(int) * 2 * (size_t)nbins));
    PG_CUDA(c, cudaMemset(c->d_gavg, 0, sizeof(double) * AV_N * (size_t)nbins));
    PG_CUDA(c, cudaMemset(c->d_gavg_cum, 0, sizeof(int) * 2 * (size_t)nbins));
    c->gavg_bins = nbins;
    c->avg_cumulated = 0;
    return PGPU_OK;
}

/* GRID_RESET, ref: src/postprocess.f90:1070-1094 */
extern "C" int pgpu_grid_reset(pgpu_ctx* c) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if ((rc = gavg_alloc(c))) return rc;
    PG_CUDA(c, cudaMemsetAsync(c->d_gavg, 0, sizeof(double) * AV_N * (size_t)c->gavg_bins, c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_gavg_cum, 0, sizeof(int) * 2 * (size_t)c->gavg_bins, c->stream));
    c->avg_cumulated = 0;
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    return PGPU_OK;
}

/* GRID_AVG: one more sample into the running averages */
extern "C" int pgpu_grid_avg(pgpu_ctx* c) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if ((rc = gavg_alloc(c))) return rc;
    if ((rc = pg_reset_status(c))) return rc;
    if ((rc = pg_resolve_vel(c))) return rc;
    pg_refresh_dev(c);
    const long long nbins = c->gavg_bins;
    PG_CUDA(c, cudaMemsetAsync(c->d_gavg_np, 0, sizeof(unsigned int) * (size_t)nbins, c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_gavg_sums, 0, sizeof(double) * GA_NSUM * (size_t)nbins, c->stream));
    if (c->np > 0) {
        const int blocks = (int)std::min<long long>((c->np + 255) / 256, 148LL * 16);
        gavg_accum_kernel<<<blocks, 256, 0, c->stream>>>(c->dev, nbins, c->d_gavg_np, c->d_gavg_sums);
        c->launches++;
    }
    gavg_update_kernel<<<(unsigned int)((nbins + 255) / 256), 256, 0, c->stream>>>(c->dev, nbins, c->d_gavg_np, c->d_gavg_sums, c->d_gavg,
                                                                                  c->d_gavg_cum, c->d_gavg_cum + nbins, c->avg_cumulated);
    c->launches++;
    c->avg_cumulated++;
    PG_CUDA(c, cudaGetLastError());
    return pg_finish(c, "pgpu_grid_avg");
}

/* the running averages, each an array of NCELLS*N_SPECIES (index JC + NCELLS*(JS-1), 0-based here); null pointers are
 * skipped; AVG_TTR = (AVG_TTRX + AVG_TTRY + AVG_TTRZ) / 3.0 (ref: :383) */
extern "C" int pgpu_get_grid_avg(pgpu_ctx* c, int64_t n, double* avg_np, double* avg_vx, double* avg_vy, double* avg_vz, double* avg_ttrx,
                                 double* avg_ttry, double* avg_ttrz, double* avg_ttr, double* avg_trot, double* avg_tvib,
                                 int32_t* avg_cumulated) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->d_gavg) PG_FAIL(c, PGPU_ERR_STATE, "get_grid_avg: pgpu_grid_avg has not been called");
    const long long nbins = c->gavg_bins;
    if (n < nbins) PG_FAIL(c, PGPU_ERR_INVALID, "get_grid_avg: arrays of %lld entries, %lld bins", (long long)n, nbins);
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    const double* src = c->d_gavg;
    double* tmp = nullptr;
    if (c->cfg.nranks > 1) { /* a cell is sampled by its owner only: the sum over ranks is the owner's value */
        int rc;
        PG_CUDA(c, cudaMalloc((void**)&tmp, sizeof(double) * AV_N * (size_t)nbins));
        PG_CUDA(c, cudaMemcpyAsync(tmp, c->d_gavg, sizeof(double) * AV_N * (size_t)nbins, cudaMemcpyDeviceToDevice, c->stream));
        if ((rc = pg_allreduce_sum_f64(c, tmp, AV_N * nbins))) { cudaFree(tmp); return rc; }
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        src = tmp;
    }
    double* dst[AV_N] = {avg_np, avg_vx, avg_vy, avg_vz, avg_ttrx, avg_ttry, avg_ttrz, avg_trot, avg_tvib};
    std::vector<double> tx, ty, tz;
    if (avg_ttr) {
        tx.resize((size_t)nbins); ty.resize((size_t)nbins); tz.resize((size_t)nbins);
        if (!avg_ttrx) dst[AV_TTRX] = tx.data();
        if (!avg_ttry) dst[AV_TTRY] = ty.data();
        if (!avg_ttrz) dst[AV_TTRZ] = tz.data();
    }
    for (int k = 0; k < AV_N; k++)
        if (dst[k]) cudaMemcpy(dst[k], src + (size_t)k * nbins, sizeof(double) * (size_t)nbins, cudaMemcpyDeviceToHost);
    if (avg_ttr)
        for (long long i = 0; i < nbins; i++) avg_ttr[i] = (dst[AV_TTRX][i] + dst[AV_TTRY][i] + dst[AV_TTRZ][i]) / (double)3.0f;
    if (tmp) cudaFree(tmp);
    if (avg_cumulated) *avg_cumulated = c->avg_cumulated;
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}

/* CHECKS, particle part (ref: src/postprocess.f90:1752-1786 + the MPI_REDUCEs of :1790-1796): tot_num(NS),
 * tot_momentum(3,NS), tot_ke_part(NS), tot_ie; summed over the ranks.  The electrostatic terms (APPLY_POTENTIAL, E
 * field energy) need PHI/E on the host's side of the solver boundary and stay there. */
extern "C" int pgpu_checks(pgpu_ctx* c, int64_t* tot_num, double* tot_momentum, double* tot_ke_part, double* tot_ie) {
    if (!c) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    if ((rc = pg_reset_status(c))) return rc;
    if ((rc = pg_resolve_vel(c))) return rc;
    pg_refresh_dev(c);
    const int ns = (int)c->ns, nacc = ns * 5 + 1;
    double* d_out = nullptr;
    PG_CUDA(c, cudaMalloc((void**)&d_out, sizeof(double) * (size_t)nacc));
    PG_CUDA(c, cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)nacc, c->stream));
    if (c->np > 0) {
        const int blocks = (int)std::min<long long>((c->np + 255) / 256, 148LL * 8);
        checks_kernel<<<blocks, 256, sizeof(double) * (size_t)nacc, c->stream>>>(c->dev, d_out);
        c->launches++;
    }
    if (c->cfg.nranks > 1 && (rc = pg_allreduce_sum_f64(c, d_out, nacc))) { cudaFree(d_out); return rc; }
    std::vector<double> h((size_t)nacc);
    PG_CUDA(c, cudaMemcpyAsync(h.data(), d_out, sizeof(double) * (size_t)nacc, cudaMemcpyDeviceToHost, c->stream));
    rc = pg_finish(c, "pgpu_checks");
    cudaFree(d_out);
    if (rc) return rc;
    for (int js = 0; js < ns; js++) {
        if (tot_num) tot_num[js] = (int64_t)(h[js] + 0.5);
        if (tot_ke_part) tot_ke_part[js] = h[4 * ns + js];
        if (tot_momentum)
            for (int k = 0; k < 3; k++) tot_momentum[k + 3 * js] = h[ns + 3 * js + k];
    }
    if (tot_ie) *tot_ie = h[5 * ns];
    return PGPU_OK;
}
