/*
 * pgpu_physics.cuh -- collision-physics device functions shared by the DSMC and MCC kernels:
 * FUNCTION_I, COLL_INTERNAL_ENERGY, HS_SCATTER (src/collisions.f90:892-992), particle creation
 * (ADD_PARTICLE_ARRAY, src/particle.f90:86-136) and the post-reaction Borgnakke-Larsen split.
 */
#pragma once
#include "pgpu_internal.cuh"

/* ref: src/collisions.f90:892-908 FUNCTION_I */
static __device__ __forceinline__ double function_i(double ecoll, double ei, double trdof, int idof) {
    if (idof == 2) return pow(1. - ei / ecoll, trdof / 2. - 1.);
    return pow((trdof + idof - 4.) / (idof - 2.), idof / 2. - 1.) * pow((trdof + idof - 4.) / (trdof - 2.), trdof / 2. - 1.) *
           pow(ei / ecoll, idof / 2. - 1.) * pow(1. - ei / ecoll, trdof / 2. - 1.);
}

/* ref: src/collisions.f90:915-937 COLL_INTERNAL_ENERGY */
static __device__ double coll_internal_energy(PgRng& g, double ecoll, double trdof, int intdof) {
    if (trdof < 1.e-6) return ecoll;
    if (intdof == 0) return 0.0;
    for (;;) {
        double r1 = g.rf();
        double ei = r1 * ecoll;
        double r2 = g.rf();
        if (r2 <= function_i(ecoll, ei, trdof, intdof)) return ei;
    }
}

/* ref: src/collisions.f90:956-992 HS_SCATTER */
static __device__ void hs_scatter(PgRng& g, double ecoll, double m1, double m2, double* c1, double* c2) {
    const double PI2 = 2. * PG_PI;
    double mred = m1 * m2 / (m1 + m2);
    double gg = sqrt(2. * ecoll / mred);
    double cosphi = 2. * g.rf() - 1.;
    double sinphi = sqrt(1 - cosphi * cosphi);
    double theta = PI2 * g.rf();
    double grel[3], wc[3];
    grel[0] = gg * sinphi * cos(theta);
    grel[1] = gg * sinphi * sin(theta);
    grel[2] = gg * cosphi;
    for (int i = 0; i < 3; i++) wc[i] = m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i];
    for (int i = 0; i < 3; i++) {
        c1[i] = wc[i] + m2 / (m1 + m2) * grel[i];
        c2[i] = wc[i] - m1 / (m1 + m2) * grel[i];
    }
}

/* append one particle created by a reaction (ADD_PARTICLE_ARRAY, ref: src/particle.f90:86-136) */
static __device__ long long pg_append_particle(const PgDev& d, double x, double y, double z, double vx, double vy, double vz,
                                        double erot, double evib, int sid, int ic) {
    unsigned long long k = atomicAdd(&d.status->n_appended, 1ULL);
    long long o = (long long)d.status->np_cur + (long long)k; /* (np_cur does not move while a collision kernel runs) */
    if (o >= d.capacity) {
        pg_set_err(d, PG_DEVERR_CAPACITY);
        return -1;
    }
    PgPos P;
    P.x = x; P.y = y; P.z = z; P.ic = ic; P.sid = sid;
    PgVel V;
    V.vx = vx; V.vy = vy; V.vz = vz; V.sid = sid; V.pad = 0;
    d.p.pos[o] = P;
    d.p.vel[o] = V;
    pg_ie_store(d.p.ie, o, erot, evib, d.status);
    if (d.p.dtrim) d.p.dtrim[o] = d.dt;
    if (d.p.id) d.p.id[o] = (long long)d.rank + ((d.id_base + (long long)k) << 8);
    return o;
}

/*
 * Borgnakke-Larsen redistribution after a reaction; shared by VSS_COLLIS (ref: src/collisions.f90:374-444)
 * and MCC_COLLISIONS_VAHEDI (:1749-1879).  e1/e2 = (erot, evib) of the two reactant slots, written back
 * by the caller.  Returns 0, or -1 if a product could not be appended.
 */
static __device__ __noinline__ int pg_react_redistribute(const PgDev& d, PgRng& g, const PgReaction& r, double ecoll, double* c1,
                                     double* c2, double& erot1, double& evib1, double& erot2, double& evib2,
                                     double x2, double y2, double z2, int ic2) {
    int p1 = r.p1, p2 = r.p2, p3 = r.p3, p4 = r.p4;
    double totdof = 3. + d.rotdof[p2 - 1] + d.vibdof[p2 - 1] + d.rotdof[p1 - 1];
    if (r.nprod >= 3) totdof = totdof + 3. + d.rotdof[p3 - 1] + d.vibdof[p3 - 1];
    if (r.nprod >= 4) totdof = totdof + 3. + d.rotdof[p4 - 1] + d.vibdof[p4 - 1];

    double ei = coll_internal_energy(g, ecoll, totdof, d.vibdof[p1 - 1]);
    evib1 = ei;
    ecoll = ecoll - ei;
    totdof = totdof - d.rotdof[p1 - 1];
    ei = coll_internal_energy(g, ecoll, totdof, d.rotdof[p1 - 1]);
    erot1 = ei;
    ecoll = ecoll - ei;
    totdof = totdof - d.vibdof[p2 - 1];
    ei = coll_internal_energy(g, ecoll, totdof, d.vibdof[p2 - 1]);
    evib2 = ei;
    ecoll = ecoll - ei;
    totdof = totdof - d.rotdof[p2 - 1];
    ei = coll_internal_energy(g, ecoll, totdof, d.rotdof[p2 - 1]);
    erot2 = ei;
    ecoll = ecoll - ei;

    double m1 = d.mass[p1 - 1], m2 = 0;
    if (r.nprod == 2) m2 = d.mass[p2 - 1];
    else if (r.nprod == 3) m2 = d.mass[p2 - 1] + d.mass[p3 - 1];
    else if (r.nprod == 4) m2 = d.mass[p2 - 1] + d.mass[p3 - 1] + d.mass[p4 - 1];
    totdof = totdof - 3.;
    ei = coll_internal_energy(g, ecoll, totdof, 3);
    hs_scatter(g, ei, m1, m2, c1, c2);
    ecoll = ecoll - ei;
    /* c1 is final for product 1 */
    if (r.nprod == 2) return 0;

    double c1f[3] = {c1[0], c1[1], c1[2]};
    if (r.nprod == 3) {
        totdof = totdof - d.vibdof[p3 - 1];
        double evib = coll_internal_energy(g, ecoll, totdof, d.vibdof[p3 - 1]);
        ecoll = ecoll - evib;
        totdof = totdof - d.rotdof[p3 - 1];
        double erot = coll_internal_energy(g, ecoll, totdof, d.rotdof[p3 - 1]);
        ecoll = ecoll - erot;
        m1 = d.mass[p2 - 1];
        m2 = d.mass[p3 - 1];
        double a[3] = {c2[0], c2[1], c2[2]}, b[3] = {c2[0], c2[1], c2[2]};
        hs_scatter(g, ecoll, m1, m2, a, b);
        /* product 2 keeps slot JP2, product 3 is a new particle at JP2's position, DTRIM = DT (:420-444) */
        c2[0] = a[0]; c2[1] = a[1]; c2[2] = a[2];
        c1[0] = c1f[0]; c1[1] = c1f[1]; c1[2] = c1f[2];
        return pg_append_particle(d, x2, y2, z2, b[0], b[1], b[2], erot, evib, p3, ic2) < 0 ? -1 : 0;
    }
    /* nprod == 4, ref: src/collisions.f90:1826-1879 (MCC only) */
    totdof = totdof - d.vibdof[p3 - 1];
    double evib = coll_internal_energy(g, ecoll, totdof, d.vibdof[p3 - 1]);
    ecoll = ecoll - evib;
    totdof = totdof - d.rotdof[p3 - 1];
    double erot = coll_internal_energy(g, ecoll, totdof, d.rotdof[p3 - 1]);
    ecoll = ecoll - erot;
    m1 = d.mass[p2 - 1];
    m2 = d.mass[p3 - 1] + d.mass[p4 - 1];
    double a[3] = {c2[0], c2[1], c2[2]}, b[3] = {c2[0], c2[1], c2[2]};
    hs_scatter(g, ecoll, m1, m2, a, b);
    c2[0] = a[0]; c2[1] = a[1]; c2[2] = a[2];
    /* third product: velocity b for now, re-scattered against the fourth below */
    totdof = totdof - d.vibdof[p4 - 1];
    double evib4 = coll_internal_energy(g, ecoll, totdof, d.vibdof[p4 - 1]);
    ecoll = ecoll - evib4;
    totdof = totdof - d.rotdof[p4 - 1];
    double erot4 = coll_internal_energy(g, ecoll, totdof, d.rotdof[p4 - 1]);
    ecoll = ecoll - erot4;
    m1 = d.mass[p3 - 1];
    m2 = d.mass[p4 - 1];
    double e3[3] = {b[0], b[1], b[2]}, e4[3] = {b[0], b[1], b[2]};
    hs_scatter(g, ecoll, m1, m2, e3, e4);
    int bad = 0;
    if (pg_append_particle(d, x2, y2, z2, e3[0], e3[1], e3[2], erot, evib, p3, ic2) < 0) bad = 1;
    /* the reference passes P3_SP_ID (not P4) for the fourth product (ref: :1875-1876); kept */
    if (pg_append_particle(d, x2, y2, z2, e4[0], e4[1], e4[2], erot4, evib4, p3, ic2) < 0) bad = 1;
    c1[0] = c1f[0]; c1[1] = c1f[1]; c1[2] = c1f[2];
    return bad ? -1 : 0;
}

