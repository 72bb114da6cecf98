/*
 * po_mpi_sim.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): the reference's MPI execution model on host
 * threads, used by `bench.py --impl reference` and by the exchange tests.
 *
 * PANTERA runs one MPI rank per spatial strip (ASSIGN_CELLS_TO_PROCS, src/grid_and_partition.f90:123-353);
 * every timestep each rank does ADVECT, then EXCHANGE ships the particles whose cell now belongs to another
 * rank (MPI_ALLTOALL + MPI_ALLTOALLV, :525-622), then the collisions run between two MPI_BARRIERs
 * (src/timecycle.f90:271-299).  Here a rank is a po_ctx driven by one pthread and the all-to-all is a
 * shared-memory hand-over of the packed send buffers between two barriers.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pantera_oracle.h"

typedef struct {
    po_ctx** ranks;
    int nranks, rank, tID0, nsteps;
    int collide; /* 0: NO_COLL runs (ADVECT + EXCHANGE only, src/timecycle.f90:288) */
    pthread_barrier_t* bar;
    pgpu_particle** sendbuf; /* [nranks] */
    int64_t* sendcap;        /* [nranks] */
    int64_t* sendcount;      /* [nranks][nranks] */
    int64_t ncoll, nreac, nsent;
    int rc;
} worker_t;

static void* worker(void* arg) {
    worker_t* w = (worker_t*)arg;
    po_ctx* me = w->ranks[w->rank];
    const int nr = w->nranks, r = w->rank;
    for (int s = 0; s < w->nsteps && !w->rc; s++) {
        int tID = w->tID0 + s;
        if ((w->rc = po_advect(me, tID))) break;
        if (nr > 1) {
            int64_t np;
            po_num_particles(me, &np);
            if (np > w->sendcap[r]) {
                free(w->sendbuf[r]);
                w->sendcap[r] = np + np / 4 + 16;
                w->sendbuf[r] = (pgpu_particle*)malloc(sizeof(pgpu_particle) * (size_t)w->sendcap[r]);
            }
            w->rc = po_exchange_pack(me, w->sendbuf[r], w->sendcap[r], &w->sendcount[(size_t)r * nr]);
            pthread_barrier_wait(w->bar);
            for (int src = 0; src < nr; src++) {
                if (src == r) continue;
                int64_t off = 0;
                for (int d = 0; d < r; d++) off += w->sendcount[(size_t)src * nr + d];
                int64_t cnt = w->sendcount[(size_t)src * nr + r];
                if (cnt) po_append_particles(me, w->sendbuf[src] + off, cnt, (int32_t)sizeof(pgpu_particle));
                w->nsent += cnt;
            }
            pthread_barrier_wait(w->bar);
        }
        int64_t a = 0, b = 0;
        if (w->collide && (w->rc = po_collide(me, tID, &a, &b))) break;
        w->ncoll += a;
        w->nreac += b;
    }
    return NULL;
}

/* Runs nsteps of [ADVECT -> EXCHANGE -> DSMC_COLLISIONS] on nranks contexts, one thread each.
 * Returns wall-clock seconds in *seconds, summed counters in totals[3] = {ncoll, nreac, particles exchanged}. */
int po_sim_run(po_ctx** ranks, int nranks, int tID0, int nsteps, double* seconds, int64_t* totals) {
    pthread_t* th = (pthread_t*)calloc((size_t)nranks, sizeof(pthread_t));
    worker_t* ws = (worker_t*)calloc((size_t)nranks, sizeof(worker_t));
    pgpu_particle** sendbuf = (pgpu_particle**)calloc((size_t)nranks, sizeof(void*));
    int64_t* sendcap = (int64_t*)calloc((size_t)nranks, sizeof(int64_t));
    int64_t* sendcount = (int64_t*)calloc((size_t)nranks * nranks, sizeof(int64_t));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)nranks);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < nranks; r++) {
        ws[r].ranks = ranks; ws[r].nranks = nranks; ws[r].rank = r; ws[r].tID0 = tID0; ws[r].nsteps = nsteps;
        ws[r].collide = po_collision_type(ranks[r]) != PGPU_NO_COLL; ws[r].bar = &bar; ws[r].sendbuf = sendbuf; ws[r].sendcap = sendcap; ws[r].sendcount = sendcount;
        pthread_create(&th[r], NULL, worker, &ws[r]);
    }
    int rc = 0;
    totals[0] = totals[1] = totals[2] = 0;
    for (int r = 0; r < nranks; r++) {
        pthread_join(th[r], NULL);
        if (ws[r].rc) rc = ws[r].rc;
        totals[0] += ws[r].ncoll; totals[1] += ws[r].nreac; totals[2] += ws[r].nsent;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    pthread_barrier_destroy(&bar);
    for (int r = 0; r < nranks; r++) free(sendbuf[r]);
    free(sendbuf); free(sendcap); free(sendcount); free(ws); free(th);
    return rc;
}
