/*
 * pgpu_io.cu -- particle dump / restart files written and read straight from the device record planes
 * (SURVEY 8(f) row 3).
 *
 * Replaces: DUMP_PARTICLES_FILE (src/tools.f90:327-365) and READ_PARTICLES_FILE (src/tools.f90:465-517), binary
 * flavour (BOOL_BINARY_OUTPUT = .TRUE., the reference's default, src/global.f90:94).  The file is a Fortran
 * SEQUENTIAL UNFORMATTED stream opened with CONVERT='BIG_ENDIAN': per particle one record
 *     [int32 80] X Y Z VX VY VZ EROT EVIB (8 x real64) S_ID IC (2 x int32) DTRIM (real64) [int32 80]
 * all big-endian = 88 bytes on disk, named <path>proc_<PROC_ID:05d>_time_<TIMESTEP:08d> (the WRITE(filename, ...)
 * format "(A,A,I0.5,A6,I0.8)").  The layout is therefore unchanged for PANTERA's own restart reader and for the
 * post-processing scripts; what changes is that no 96-byte AoS copy of the store is made on the host: a kernel
 * byte-swaps the records out of the planes into a staging buffer, which is copied and written in chunks.
 * Reading: the reference ignores the file's DTRIM and gives every restarted particle DTRIM = DT (:497).
 */
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>

#include <sys/stat.h>

#include "pgpu_internal.cuh"

namespace {

#define DUMP_REC 88

__device__ __forceinline__ unsigned long long bswap64(unsigned long long v) {
    return ((unsigned long long)__byte_perm((unsigned int)(v & 0xffffffffULL), 0, 0x0123) << 32) |
           (unsigned long long)__byte_perm((unsigned int)(v >> 32), 0, 0x0123);
}
__device__ __forceinline__ unsigned int bswap32(unsigned int v) { return __byte_perm(v, 0, 0x0123); }

/* one thread per particle: 22 big-endian 32-bit words per record (records are 4-byte aligned, not 8) */
__global__ void __launch_bounds__(256) dump_pack_kernel(PgPlanes p, long long base, long long n, double dt, int has_dtrim,
                                                        unsigned int* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long o = base + i;
    const PgPos P = p.pos[o];
    const PgVel V = p.vel[o];
    const PgIe E = pg_ie_load(p.ie, o);
    const double f[8] = {P.x, P.y, P.z, V.vx, V.vy, V.vz, E.erot, E.evib};
    unsigned int* w = out + (size_t)i * (DUMP_REC / 4);
    w[0] = bswap32(80u);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned long long b = bswap64((unsigned long long)__double_as_longlong(f[k]));
        w[1 + 2 * k] = (unsigned int)(b & 0xffffffffULL); /* little-endian host: low word first = first 4 bytes in memory */
        w[2 + 2 * k] = (unsigned int)(b >> 32);
    }
    w[17] = bswap32((unsigned int)P.sid);
    w[18] = bswap32((unsigned int)P.ic);
    const unsigned long long b = bswap64((unsigned long long)__double_as_longlong((has_dtrim && p.dtrim) ? p.dtrim[o] : dt));
    w[19] = (unsigned int)(b & 0xffffffffULL);
    w[20] = (unsigned int)(b >> 32);
    w[21] = bswap32(80u);
}

__global__ void __launch_bounds__(256) dump_unpack_kernel(const unsigned int* __restrict__ in, long long n, PgPlanes p, long long base,
                                                          double dt, int rank, long long id_base, int ncells, int relax, PgStatus* st) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int* w = in + (size_t)i * (DUMP_REC / 4);
    double f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned long long b = ((unsigned long long)w[2 + 2 * k] << 32) | (unsigned long long)w[1 + 2 * k];
        f[k] = __longlong_as_double((long long)bswap64(b));
    }
    const int sid = (int)bswap32(w[17]);
    int ic = (int)bswap32(w[18]);
    /* PGPU_RESTART_UNASSIGNED: the dump's IC is not trusted; what is no cell of this grid waits as IC = 0 for the re-binning */
    if (relax && ncells > 0 && (ic < 1 || ic > ncells)) ic = 0;
    const long long o = base + i;
    PgPos P;
    P.x = f[0]; P.y = f[1]; P.z = f[2]; P.ic = ic; P.sid = sid;
    PgVel V;
    V.vx = f[3]; V.vy = f[4]; V.vz = f[5]; V.sid = sid; V.pad = 0;
    PgIe E;
    E.erot = f[6]; E.evib = f[7];
    p.pos[o] = P;
    p.vel[o] = V;
    pg_ie_store(p.ie, o, E.erot, E.evib, st);
    if (p.dtrim) p.dtrim[o] = dt; /* INIT_PARTICLE(..., DT, ...), ref: src/tools.f90:497 */
    if (p.id) p.id[o] = (long long)rank + ((id_base + i) << 8);
    if (E.erot != 0. || E.evib != 0.) st->flag_ie = 1;
    if (bswap32(w[0]) != 80u || bswap32(w[21]) != 80u) atomicCAS(&st->err, 0, PG_DEVERR_CELL_RANGE);
    if (!relax && ncells > 0 && (ic < 1 || ic > ncells)) atomicCAS(&st->err, 0, PG_DEVERR_CELL_RANGE);
}

const long long IO_CHUNK = 1 << 20; /* particles per staging chunk (88 MiB) */

void dump_name(char* out, size_t n, const char* prefix, int rank, int timestep) {
    snprintf(out, n, "%sproc_%05d_time_%08d", prefix ? prefix : "", rank, timestep);
}

} /* namespace */

extern "C" int pgpu_dump_particles_file(pgpu_ctx* c, const char* path_prefix, int32_t timestep, int32_t binary) {
    if (!c) return PGPU_ERR_INVALID;
    if (!binary) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "dump_particles_file: only the binary flavour (Binary_output: .TRUE.) is written from the device");
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    /* the reference dumps the live particles in store order: drop tombstones first (as pgpu_download_particles does) */
    if ((rc = pg_reset_status(c))) return rc;
    if (c->maybe_tombstones && c->np > 0) {
        if ((rc = pg_sort_impl(c))) return rc;
        if ((rc = pg_finish(c, "pgpu_dump_particles_file"))) return rc;
    }
    if ((rc = pg_resolve_vel(c))) return rc;
    char name[1024];
    dump_name(name, sizeof(name), path_prefix, c->cfg.rank, timestep);
    struct stat sb;
    if (stat(name, &sb) == 0) PG_FAIL(c, PGPU_ERR_INVALID, "dump_particles_file: %s exists (the reference opens it with STATUS='NEW')", name);
    FILE* fp = fopen(name, "wb");
    if (!fp) PG_FAIL(c, PGPU_ERR_INVALID, "dump_particles_file: cannot create %s (%s)", name, strerror(errno));
    const long long chunk = std::min<long long>(IO_CHUNK, std::max<long long>(c->np, 1));
    void *d_buf[2] = {nullptr, nullptr}, *h_buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    rc = PGPU_OK;
    for (int k = 0; k < 2; k++) {
        if (cudaMalloc(&d_buf[k], (size_t)chunk * DUMP_REC) != cudaSuccess || cudaMallocHost(&h_buf[k], (size_t)chunk * DUMP_REC) != cudaSuccess)
            rc = PGPU_ERR_CUDA;
        cudaEventCreate(&ev[k]);
    }
    PgPlanes p = c->planes[c->cur];
    /* two staging buffers: the kernel + copy of chunk k+1 overlap the fwrite of chunk k */
    long long pending_n[2] = {0, 0};
    int k = 0;
    for (long long done = 0; rc == PGPU_OK && done < c->np; done += chunk, k ^= 1) {
        const long long m = std::min<long long>(chunk, c->np - done);
        if (pending_n[k]) {
            cudaEventSynchronize(ev[k]);
            if (fwrite(h_buf[k], DUMP_REC, (size_t)pending_n[k], fp) != (size_t)pending_n[k]) rc = PGPU_ERR_INVALID;
        }
        dump_pack_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, c->stream>>>(p, done, m, c->cfg.dt, c->has_dtrim ? 1 : 0, (unsigned int*)d_buf[k]);
        cudaMemcpyAsync(h_buf[k], d_buf[k], (size_t)m * DUMP_REC, cudaMemcpyDeviceToHost, c->stream);
        cudaEventRecord(ev[k], c->stream);
        pending_n[k] = m;
        c->launches++;
    }
    for (int j = 0; j < 2 && rc == PGPU_OK; j++, k ^= 1) { /* drain in issue order */
        if (pending_n[k]) {
            cudaEventSynchronize(ev[k]);
            if (fwrite(h_buf[k], DUMP_REC, (size_t)pending_n[k], fp) != (size_t)pending_n[k]) rc = PGPU_ERR_INVALID;
            pending_n[k] = 0;
        }
    }
    if (fclose(fp) != 0 && rc == PGPU_OK) rc = PGPU_ERR_INVALID;
    for (int j = 0; j < 2; j++) {
        cudaFree(d_buf[j]);
        cudaFreeHost(h_buf[j]);
        cudaEventDestroy(ev[j]);
    }
    if (rc == PGPU_ERR_CUDA) PG_FAIL(c, rc, "dump_particles_file: staging buffers (%lld bytes) could not be allocated", (long long)chunk * DUMP_REC);
    if (rc) PG_FAIL(c, rc, "dump_particles_file: short write to %s", name);
    return PGPU_OK;
}

extern "C" int pgpu_read_particles_file(pgpu_ctx* c, const char* path_prefix, int32_t timestep, int32_t binary) {
    if (!c) return PGPU_ERR_INVALID;
    if (!(binary & 1)) PG_FAIL(c, PGPU_ERR_UNSUPPORTED, "read_particles_file: only the binary flavour is read on the device");
    const int relax = (binary & PGPU_RESTART_UNASSIGNED) ? 1 : 0;
    int rc;
    c->reassigning = true; /* (a second file may follow a first one read with PGPU_RESTART_UNASSIGNED) */
    rc = pg_prepare(c);
    c->reassigning = false;
    if (rc) return rc;
    char name[1024];
    dump_name(name, sizeof(name), path_prefix, c->cfg.rank, timestep);
    FILE* fp = fopen(name, "rb");
    if (!fp) PG_FAIL(c, PGPU_ERR_INVALID, "read_particles_file: particle restart file %s not found", name);
    struct stat sb;
    if (stat(name, &sb) != 0 || (sb.st_size % DUMP_REC) != 0) {
        fclose(fp);
        PG_FAIL(c, PGPU_ERR_INVALID, "read_particles_file: %s is not a sequence of 88-byte records", name);
    }
    const long long n = (long long)(sb.st_size / DUMP_REC);
    if (c->capacity == 0) {
        long long cap = c->cfg.capacity > 0 ? (long long)c->cfg.capacity : (long long)(1.5 * (double)n);
        rc = pg_ensure_capacity(c, std::max<long long>(cap, c->np + n));
    } else {
        rc = pg_ensure_capacity(c, c->np + n);
    }
    if (rc) { fclose(fp); return rc; }
    if ((rc = pg_resolve_vel(c))) { fclose(fp); return rc; }
    if ((rc = pg_reset_status(c))) { fclose(fp); return rc; }
    const long long chunk = std::min<long long>(IO_CHUNK, std::max<long long>(n, 1));
    void *d_buf = nullptr, *h_buf = nullptr;
    if (cudaMalloc(&d_buf, (size_t)chunk * DUMP_REC) != cudaSuccess || cudaMallocHost(&h_buf, (size_t)chunk * DUMP_REC) != cudaSuccess) {
        cudaFree(d_buf);
        fclose(fp);
        PG_FAIL(c, PGPU_ERR_CUDA, "read_particles_file: staging buffers could not be allocated");
    }
    if (c->np > 0 && !c->has_dtrim) { /* (every restarted particle has DTRIM = DT, so the plane stays uniform) */ }
    PgPlanes p = c->planes[c->cur];
    const int ncells = (c->cfg.grid_type == PGPU_UNSTRUCTURED) ? (c->mesh_set ? c->nc : 0) : c->cfg.nx * c->cfg.ny;
    rc = PGPU_OK;
    for (long long done = 0; done < n; done += chunk) {
        const long long m = std::min<long long>(chunk, n - done);
        if (fread(h_buf, DUMP_REC, (size_t)m, fp) != (size_t)m) { rc = PGPU_ERR_INVALID; break; }
        cudaMemcpyAsync(d_buf, h_buf, (size_t)m * DUMP_REC, cudaMemcpyHostToDevice, c->stream);
        dump_unpack_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, c->stream>>>((const unsigned int*)d_buf, m, p, c->np + done, c->cfg.dt,
                                                                                  c->cfg.rank, c->id_counter + done, ncells, relax, c->d_status);
        cudaStreamSynchronize(c->stream); /* one staging buffer: the host refills it next */
        c->launches++;
    }
    fclose(fp);
    cudaFree(d_buf);
    cudaFreeHost(h_buf);
    if (rc) PG_FAIL(c, rc, "read_particles_file: short read from %s", name);
    const long long np_old = c->np;
    c->np += n;
    c->sorted = false;
    c->hist_valid = false;
    c->owned_clean = false;
    if ((rc = pg_finish(c, "pgpu_read_particles_file"))) {
        c->np = np_old;
        return rc;
    }
    if (c->h_status->flag_ie) c->ie_seen = true;
    c->id_counter = std::max(c->id_counter, c->np);
    if (relax && n > 0) c->cells_unassigned = true;
    return PGPU_OK;
}
