#!/usr/bin/env python
"""Independent pins, second set (round 2, after VERDICT r01 asked for more of them): single events of the hot path
evaluated in PLAIN PYTHON from the Fortran source of the reference (file:line at every routine), NOT from
oracle/pantera_oracle.c.  Same conventions as tests/golden/fortran_pins.py: IEEE doubles, no FMA, left-to-right like
gfortran without -ffast-math, `x*x` for `X**2`; random draws supplied as a script.

    python tests/golden/fortran_pins2.py        # regenerates tests/golden/fortran_pins2.json

Covered reference lines:
  axi_walk       src/timecycle.f90:1173-1253 (the axisymmetric edge test of unstructured grids: vertical edges, the
                 hyperboloid quadratic with its root filters, the trial MOVE_PARTICLE forth and back -- which is not an
                 exact inverse and leaves its rounding in the particle --, the missing ordering check between the two
                 roots), :1308-1343 (move, neighbour), :1433-1443 (specular), :1648-1657 (hop / final flight), :2297-2324
                 (MOVE_PARTICLE, axisymmetric); edge geometry src/grid_and_partition.f90:1401-1418
  structured_walk  src/timecycle.f90:1072-1081 (NORMX, NORMY, XW, YW), :1659-1661, :1735-1752 (box sides, non-axisymmetric: a
                 strict '>' here, unlike the '>=' of the unstructured faces), :1841-1863 (interior walls: one-sided, hit
                 point inside the segment), :1868-1935 (periodic shift, specular flip, diffuse re-emission with BOUNDTEMP),
                 :1937-1965 (specular wall: rf() is drawn even when the wall is not porous; porous removal), :2015-2019,
                 :2044-2054 (Z wrap), :2104-2106 + src/grid_and_partition.f90:46-72 (re-index on the uniform grid)
  mcc_collisions  src/collisions.f90:1261-1595 (MCC_COLLISIONS: partner drawn from the background Maxwellian BEFORE the test,
                 P = 1 - exp(-DT n VR sigma (VR/CREF)^(1-2 omega)), partner created in the store, TCE test over the reaction list
                 with the ORIGINAL species ids, 3-product redistribution, Borgnakke-Larsen selector, VSS deflection; loop
                 over the NP_PROC present at entry and over the components of the background mixture)
  structured_axi  src/timecycle.f90:1662-1729 (box sides of an axisymmetric rectilinear grid: vertical sides by division, the
                 horizontal ones by the quadratic in the radius, roots >= 0 here -- no TOL --, trial moves), :1762-1838 (interior
                 walls, axisymmetric: the cone quadratic, hit parameter along the segment), :1885-1935, :1944-1957
  tri_walk_solid  src/timecycle.f90:1270-1285 (plane 2-D face test), :1308-1343 (a neighbour cell of a SOLID volume group is a
                 boundary: FLUIDBOUNDARY, the wall model is that of the face's own physical group), :1433-1443, :1648-1657
  emit_2d        src/timecycle.f90:368-547 (BOUNDARIES_EMIT, DIMS == 2: FACE_TANG1 = (-n_y, n_x, 0), the point on the edge -- plane:
                 linear in R and a Z draw; axisymmetric: Y = SQRT(Y1^2 + R (Y2^2 - Y1^2)) with the three orderings of Y1, Y2)
  seed_2d_1d     src/initialization.f90:2062-2141 (INITIAL_SEED on triangles -- plane and axisymmetric: the (S, T) fold, Z drawn
                 or zero, U always drawn -- and on segments: X from S, Y from T, Z from U), CELL_VOLUMES src/grid_and_partition.f90
                 :1315-1323 (triangles) / :870-871 (segments)
  vinject_tri    src/timecycle.f90:553-670 (VOLUME_INJECT, unstructured branch: DT*NRHODOT/(FNUM*SPWT)*VOL*MOLFRAC/N_MPI_THREADS
                 per cell, every cell -- no ownership test --, the same point / velocity / energy draws as INITIAL_SEED, DTRIM = DT)
  tri_diffuse    src/timecycle.f90:1327-1343 (2-D: FACE_NORMAL = -EDGE_NORMAL, FACE_TANG1 = (-FACE_NORMAL(2), FACE_NORMAL(1), 0),
                 FACE_TANG2 = (0, 0, 1)), :1445-1476 (diffuse re-emission from a wall of the plane triangle mesh)
  dsmc_order     src/collisions.f90:37-149 (DSMC_COLLISIONS: INDALL is a stable counting sort by cell, cells are visited in
                 ascending order), :259-271 (INT(NPC*rf()) picks into that order)
  cell_nonuniform  src/grid_and_partition.f90:34-96 (CELL_FROM_POSITION on RECTILINEAR_NONUNIFORM grids), src/tools.f90:1499-1535
                 (BINARY_SEARCH; a value equal to the last array element never leaves the reference's loop: not pinned)
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_pins as fp1  # noqa: E402  (species tables, VSS pair tables, vss_collis_cell of the first set)

hx = fp1.hx
Script = fp1.Script
TOL = 1.0e-15  # src/timecycle.f90:1068


# ------------------------------------------------------------------------------------------ axisymmetric triangles
def tri_mesh():
    """3 x 2 quads cut into 12 triangles over [0, 0.09] x [0.01, 0.05] (y is the radius; no cell touches the axis),
    interior nodes displaced so that no edge is vertical or horizontal except on the boundary.  EDGE_NORMAL and
    CELL_EDGES_LEN as src/grid_and_partition.f90:1401-1418 (edge J joins local nodes J and J mod 3 + 1).
    Physical groups: 1 x = 0, 2 x = 0.09, 3 y = 0.01, 4 y = 0.05."""
    nx, ny = 3, 2
    xs, ys = [0.0, 0.03, 0.06, 0.09], [0.01, 0.03, 0.05]
    nodes = []
    for i in range(nx + 1):
        for j in range(ny + 1):
            x, y = xs[i], ys[j]
            if 0 < i < nx and 0 < j < ny:
                x += 0.004 * (1 if i % 2 else -1)
                y += 0.003
            nodes.append([x, y])
    nid = lambda i, j: i * (ny + 1) + j + 1
    cells = []
    for i in range(nx):
        for j in range(ny):
            a, b, c, d = nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)
            cells.append([a, b, c])
            cells.append([a, c, d])
    normal, length = [], []
    for cn in cells:
        nrm, ln = [], []
        for j in range(3):
            x1, y1 = nodes[cn[j] - 1]
            x2, y2 = nodes[cn[(j + 1) % 3] - 1]
            le = math.sqrt((y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1))
            ln.append(le)
            nrm.append([(y2 - y1) / le, (x1 - x2) / le])
        normal.append(nrm)
        length.append(ln)
    edge_owner = {}
    for c, cn in enumerate(cells):
        for j in range(3):
            key = tuple(sorted((cn[j], cn[(j + 1) % 3])))
            edge_owner.setdefault(key, []).append((c, j))
    neighbors = [[-1] * 3 for _ in cells]
    faces_pg = [[-1] * 3 for _ in cells]
    for key, own in edge_owner.items():
        if len(own) == 2:
            (c1, j1), (c2, j2) = own
            neighbors[c1][j1] = c2 + 1
            neighbors[c2][j2] = c1 + 1
        else:
            (c1, j1), = own
            (xa, ya), (xb, yb) = nodes[key[0] - 1], nodes[key[1] - 1]
            if xa == 0.0 and xb == 0.0:
                pg = 1
            elif xa == 0.09 and xb == 0.09:
                pg = 2
            elif ya == 0.01 and yb == 0.01:
                pg = 3
            else:
                pg = 4
            faces_pg[c1][j1] = pg
    return dict(nodes=nodes, cells=cells, normal=normal, edges_len=length, neighbors=neighbors, faces_pg=faces_pg)


def move_axi(q, time):
    """MOVE_PARTICLE, axisymmetric branch, src/timecycle.f90:2311-2324 (in place)."""
    q["x"] = q["x"] + q["vx"] * time
    q["z"] = 0.0
    r = math.sqrt((q["y"] + q["vy"] * time) * (q["y"] + q["vy"] * time) + (q["vz"] * time) * (q["vz"] * time))
    sintheta = q["vz"] * time / r
    costheta = math.copysign(math.sqrt(1.0 - sintheta * sintheta), q["y"] + q["vy"] * time)
    q["y"] = r
    vz, vy = q["vz"], q["vy"]
    q["vz"] = costheta * vz - sintheta * vy
    q["vy"] = sintheta * vz + costheta * vy


def advect_axi(mesh, p, dt):
    """One particle through ADVECT on the axisymmetric triangle mesh, every boundary specular.
    src/timecycle.f90:1166-1253, 1308-1343, 1433-1443, 1648-1657."""
    q = dict(x=p["x"], y=p["y"], z=p["z"], vx=p["vx"], vy=p["vy"], vz=p["vz"])
    ic = p["ic"]
    dtrim = dt
    trace = []
    ntrial = 0
    while dtrim > 0.0:
        dtcoll = dtrim
        boundcoll = -1
        for i in range(1, 4):
            edge_x1, edge_y1 = mesh["nodes"][mesh["cells"][ic - 1][i - 1] - 1]
            n1, n2 = mesh["normal"][ic - 1][i - 1]
            test = []
            if abs(n2) < 1.0e-10:                                   # vertical wall, :1181-1189
                sol1 = (edge_x1 - q["x"]) / q["vx"]
                if sol1 >= 0 and sol1 < dtcoll:
                    test = [sol1]
            else:                                                   # :1191-1227
                alpha = -n1 / n2
                beta = (q["x"] - edge_x1) * alpha
                coefa = q["vy"] * q["vy"] + q["vz"] * q["vz"] - alpha * alpha * (q["vx"] * q["vx"])
                coefb = q["y"] * q["vy"] - (beta + edge_y1) * alpha * q["vx"]
                coefc = q["y"] * q["y"] - (beta + edge_y1) * (beta + edge_y1)
                delta = coefb * coefb - coefa * coefc
                if delta >= 0:
                    sol1 = (-coefb - math.sqrt(delta)) / coefa
                    sol2 = (-coefb + math.sqrt(delta)) / coefa
                    ok1 = sol1 >= -TOL and sol1 < dtcoll and beta + alpha * q["vx"] * sol1 + edge_y1 >= 0
                    ok2 = sol2 >= -TOL and sol2 < dtcoll and beta + alpha * q["vx"] * sol2 + edge_y1 >= 0
                    if ok1:
                        if ok2:
                            test = [sol1, sol2] if sol1 <= sol2 else [sol2, sol1]
                        else:
                            test = [sol1]
                    elif ok2:
                        test = [sol2]
            for t in test:                                          # :1233-1251 (no comparison of TEST(SOL) with DTCOLL here)
                move_axi(q, t)
                ntrial += 1
                if (q["vx"] * n1 + q["vy"] * n2) > 0:
                    le = mesh["edges_len"][ic - 1][i - 1]
                    colldist = -n2 * (q["x"] - edge_x1) / le + n1 * (q["y"] - edge_y1) / le
                    if colldist >= 0 and colldist <= 1:
                        dtcoll = t
                        boundcoll = i
                move_axi(q, -t)
        if boundcoll != -1:
            move_axi(q, dtcoll)
            dtrim = dtrim - dtcoll
            trace.append([ic, boundcoll])
            neighbor = mesh["neighbors"][ic - 1][boundcoll - 1]
            if neighbor == -1:
                fn = [-mesh["normal"][ic - 1][boundcoll - 1][0], -mesh["normal"][ic - 1][boundcoll - 1][1], -0.0]
                vdotn = q["vx"] * fn[0] + q["vy"] * fn[1] + q["vz"] * fn[2]
                q["vx"] = q["vx"] - 2.0 * vdotn * fn[0]
                q["vy"] = q["vy"] - 2.0 * vdotn * fn[1]
                q["vz"] = q["vz"] - 2.0 * vdotn * fn[2]
            else:
                ic = neighbor
        else:
            move_axi(q, dtrim)
            dtrim = 0.0
    q["ic"] = ic
    return q, trace, ntrial


def advect_tri_plane(mesh, p, dt, cell_pg, solid_groups, diffuse=None, species=None, rf=None):
    """One particle through ADVECT on the plane (not axisymmetric) triangle mesh, every tagged face specular.  cell_pg[c] is
    CELL_PG (-1 none); a neighbour whose group is in solid_groups is a wall (FLUIDBOUNDARY)."""
    x, y, z, vx, vy, vz = p["x"], p["y"], p["z"], p["vx"], p["vy"], p["vz"]
    ic = p["ic"]
    dtrim = dt
    trace = []
    while dtrim > 0.0:
        dtcoll = dtrim
        boundcoll = -1
        for i in range(1, 4):
            n1, n2 = mesh["normal"][ic - 1][i - 1]
            px, py = mesh["nodes"][mesh["cells"][ic - 1][i - 1] - 1]
            vn = vx * n1 + vy * n2
            dx = (px - x) * n1 + (py - y) * n2
            if vn >= 0.0 and vn * dtcoll >= dx:
                dtcoll = dx / vn
                boundcoll = i
        if boundcoll == -1:
            x, y, z = x + vx * dtrim, y + vy * dtrim, z + vz * dtrim
            dtrim = 0.0
            continue
        x, y, z = x + vx * dtcoll, y + vy * dtcoll, z + vz * dtcoll
        dtrim = dtrim - dtcoll
        neighbor = mesh["neighbors"][ic - 1][boundcoll - 1]
        fluidboundary = neighbor == -1
        if not fluidboundary and cell_pg[neighbor - 1] != -1 and cell_pg[neighbor - 1] in solid_groups:
            fluidboundary = True
        trace.append([ic, boundcoll, int(fluidboundary)])
        if fluidboundary and diffuse and mesh["faces_pg"][ic - 1][boundcoll - 1] in diffuse:
            wall_temp = diffuse[mesh["faces_pg"][ic - 1][boundcoll - 1]]
            fn = [-mesh["normal"][ic - 1][boundcoll - 1][0], -mesh["normal"][ic - 1][boundcoll - 1][1], -0.0]
            t1 = [-fn[1], fn[0], 0.0]
            t2 = [0.0, 0.0, 1.0]
            sp = species[p["sid"]]
            _, v_tang1, v_tang2 = fp1.maxwell(rf, 0.0, 0.0, 0.0, wall_temp, wall_temp, wall_temp, sp["mass"])
            p["erot"] = fp1.internal_energy(rf, sp["rotdof"], wall_temp)
            p["evib"] = fp1.internal_energy(rf, sp["vibdof"], wall_temp)
            v_perp = fp1.flx(rf, 0.0, wall_temp, sp["mass"])
            vx = v_perp * fn[0] + v_tang1 * t1[0] + v_tang2 * t2[0]
            vy = v_perp * fn[1] + v_tang1 * t1[1] + v_tang2 * t2[1]
            vz = v_perp * fn[2] + v_tang1 * t1[2] + v_tang2 * t2[2]
        elif fluidboundary:
            assert mesh["faces_pg"][ic - 1][boundcoll - 1] != -1
            fn = [-mesh["normal"][ic - 1][boundcoll - 1][0], -mesh["normal"][ic - 1][boundcoll - 1][1]]
            vdotn = vx * fn[0] + vy * fn[1] + vz * 0.0
            vx = vx - 2.0 * vdotn * fn[0]
            vy = vy - 2.0 * vdotn * fn[1]
            vz = vz - 2.0 * vdotn * (-0.0)
        else:
            ic = neighbor
    return dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, ic=ic), trace


# ------------------------------------------------------------------------------------------ rectilinear grids
def advect_structured(g, walls, species, parts, dt, rf):
    """ADVECT on a uniform rectilinear 2-D grid, not axisymmetric, particles in store order sharing ONE rf() stream.
    g: xmin, xmax, ymin, ymax, zmin, zmax, nx, ny, periodic[4], specular[4], diffuse[4], boundtemp, z_periodic.
    walls: dicts x1, y1, x2, y2, normx, normy, specular, diffuse, porous, transm, temp.  Returns per particle the final
    state (or None when it was removed) and its events."""
    normx, normy = [1.0, -1.0, 0.0, 0.0], [0.0, 0.0, 1.0, -1.0]
    xw = [g["xmin"], g["xmax"], g["xmin"], g["xmax"]]
    yw = [g["ymax"], g["ymin"], g["ymin"], g["ymax"]]
    out = []
    for p in parts:
        q = dict(p)
        dtrim = dt
        removed = False
        events = []
        while dtrim > 0.0:
            dtcoll = dtrim
            boundcoll = -1
            for i in range(4):
                vn = -(q["vx"] * normx[i] + q["vy"] * normy[i])
                dx = (q["x"] - xw[i]) * normx[i] + (q["y"] - yw[i]) * normy[i]
                if vn >= 0.0 and vn * dtcoll > dx:
                    dtcoll = dx / vn
                    boundcoll = i + 1
            wallcoll = -1
            for k, w in enumerate(walls):
                vn = -(q["vx"] * w["normx"] + q["vy"] * w["normy"])
                dx = (q["x"] - w["x1"]) * w["normx"] + (q["y"] - w["y1"]) * w["normy"]
                if dx >= 0.0 and vn * dtcoll >= dx:
                    cand = dx / vn
                    xcoll = q["x"] + q["vx"] * cand
                    ycoll = q["y"] + q["vy"] * cand
                    colldist = ((xcoll - w["x1"]) * (w["x2"] - w["x1"]) + (ycoll - w["y1"]) * (w["y2"] - w["y1"])) / \
                               ((w["x2"] - w["x1"]) * (w["x2"] - w["x1"]) + (w["y2"] - w["y1"]) * (w["y2"] - w["y1"]))
                    if colldist >= 0 and colldist <= 1:
                        boundcoll = -1
                        wallcoll = k + 1
                        dtcoll = cand

            def move(t):
                q["x"] = q["x"] + q["vx"] * t
                q["y"] = q["y"] + q["vy"] * t
                q["z"] = q["z"] + q["vz"] * t

            if boundcoll != -1:
                events.append(["side", boundcoll])
                if g["periodic"][boundcoll - 1]:
                    move(dtcoll)
                    if boundcoll == 1:
                        q["x"] = q["x"] + g["xmax"] - g["xmin"]
                    elif boundcoll == 2:
                        q["x"] = q["x"] - g["xmax"] + g["xmin"]
                    elif boundcoll == 3:
                        q["y"] = q["y"] + g["ymax"] - g["ymin"]
                    else:
                        q["y"] = q["y"] - g["ymax"] + g["ymin"]
                    dtrim = dtrim - dtcoll
                elif g["specular"][boundcoll - 1]:
                    move(dtcoll)
                    if boundcoll <= 2:
                        q["vx"] = -q["vx"]
                    else:
                        q["vy"] = -q["vy"]
                    dtrim = dtrim - dtcoll
                elif g["diffuse"][boundcoll - 1]:
                    move(dtcoll)
                    dtrim = dtrim - dtcoll
                    sp = species[q["sid"]]
                    _, v_perp, vz = fp1.maxwell(rf, 0.0, 0.0, 0.0, g["boundtemp"], g["boundtemp"], g["boundtemp"], sp["mass"])
                    erot = fp1.internal_energy(rf, sp["rotdof"], g["boundtemp"])
                    evib = fp1.internal_energy(rf, sp["vibdof"], g["boundtemp"])
                    v_norm = fp1.flx(rf, 0.0, g["boundtemp"], sp["mass"])
                    q["vx"] = v_norm * normx[boundcoll - 1] - v_perp * normy[boundcoll - 1]
                    q["vy"] = v_perp * normx[boundcoll - 1] + v_norm * normy[boundcoll - 1]
                    q["vz"] = vz
                    q["erot"], q["evib"] = erot, evib
                else:
                    removed = True
                    dtrim = 0.0
            elif wallcoll != -1:
                w = walls[wallcoll - 1]
                events.append(["wall", wallcoll])
                if w["specular"]:
                    rfp = rf()
                    if w["porous"] and rfp <= w["transm"]:
                        removed = True
                        dtrim = 0.0
                    else:
                        move(dtcoll)
                        vdotn = q["vx"] * w["normx"] + q["vy"] * w["normy"]
                        q["vx"] = q["vx"] - 2.0 * vdotn * w["normx"]
                        q["vy"] = q["vy"] - 2.0 * vdotn * w["normy"]
                        dtrim = dtrim - dtcoll
                elif w["diffuse"]:
                    rfp = rf()
                    if w["porous"] and rfp <= w["transm"]:
                        removed = True
                        dtrim = 0.0
                    else:
                        move(dtcoll)
                        dtrim = dtrim - dtcoll
                        sp = species[q["sid"]]
                        _, v_perp, vz = fp1.maxwell(rf, 0.0, 0.0, 0.0, w["temp"], w["temp"], w["temp"], sp["mass"])
                        erot = fp1.internal_energy(rf, sp["rotdof"], w["temp"])
                        evib = fp1.internal_energy(rf, sp["vibdof"], w["temp"])
                        v_norm = fp1.flx(rf, 0.0, w["temp"], sp["mass"])
                        q["vx"] = v_norm * w["normx"] - v_perp * w["normy"]
                        q["vy"] = v_perp * w["normx"] + v_norm * w["normy"]
                        q["vz"] = vz
                        q["erot"], q["evib"] = erot, evib
                else:
                    removed = True
                    dtrim = 0.0
            else:
                move(dtrim)
                dtrim = 0.0
        if g["z_periodic"]:
            while q["z"] > g["zmax"]:
                q["z"] = g["zmin"] + (q["z"] - g["zmax"])
            while q["z"] < g["zmin"]:
                q["z"] = g["zmax"] + (q["z"] - g["zmin"])
        if removed:
            out.append((None, events))
        else:
            ddx = (g["xmax"] - g["xmin"]) / g["nx"]
            ddy = (g["ymax"] - g["ymin"]) / g["ny"]
            q["ic"] = int((q["x"] - g["xmin"]) / ddx) + g["nx"] * int((q["y"] - g["ymin"]) / ddy) + 1
            out.append((q, events))
    return out


def advect_structured_axi(g, walls, parts, dt):
    """ADVECT on a uniform rectilinear AXISYMMETRIC grid, specular sides and specular (non-porous) walls only: no draw
    changes the outcome (the rf() of a specular wall is consumed and ignored).  Same containers as advect_structured."""
    normx, normy = [1.0, -1.0, 0.0, 0.0], [0.0, 0.0, 1.0, -1.0]
    boundpos = [g["xmin"], g["xmax"], g["ymin"], g["ymax"]]
    out = []
    for p in parts:
        q = dict(x=p["x"], y=p["y"], z=p["z"], vx=p["vx"], vy=p["vy"], vz=p["vz"])
        dtrim = dt
        events = []
        while dtrim > 0.0:
            dtcoll = dtrim
            boundcoll = -1
            for i in range(1, 5):
                test = []
                if i <= 2:
                    sol1 = (boundpos[i - 1] - q["x"]) / q["vx"]
                    if sol1 >= 0 and sol1 < dtcoll:
                        test = [sol1]
                else:
                    coefa = q["vy"] * q["vy"] + q["vz"] * q["vz"]
                    coefb = q["y"] * q["vy"]
                    coefc = q["y"] * q["y"] - boundpos[i - 1] * boundpos[i - 1]
                    delta = coefb * coefb - coefa * coefc
                    if delta >= 0:
                        sol1 = (-coefb - math.sqrt(delta)) / coefa
                        sol2 = (-coefb + math.sqrt(delta)) / coefa
                        ok1 = sol1 >= 0 and sol1 < dtcoll
                        ok2 = sol2 >= 0 and sol2 < dtcoll
                        if ok1:
                            test = ([sol1, sol2] if sol1 <= sol2 else [sol2, sol1]) if ok2 else [sol1]
                        elif ok2:
                            test = [sol2]
                for t in test:
                    move_axi(q, t)
                    if (q["vx"] * normx[i - 1] + q["vy"] * normy[i - 1]) < 0:
                        dtcoll = t
                        boundcoll = i
                    move_axi(q, -t)
            wallcoll = -1
            for k, w in enumerate(walls):
                test = []
                if w["x1"] == w["x2"]:
                    sol1 = (w["x1"] - q["x"]) / q["vx"]
                    if sol1 >= 0 and sol1 < dtcoll:
                        test = [sol1]
                else:
                    alpha = (w["y2"] - w["y1"]) / (w["x2"] - w["x1"])
                    beta = (q["x"] - w["x1"]) * alpha
                    coefa = q["vy"] * q["vy"] + q["vz"] * q["vz"] - alpha * alpha * (q["vx"] * q["vx"])
                    coefb = q["y"] * q["vy"] - (beta + w["y1"]) * alpha * q["vx"]
                    coefc = q["y"] * q["y"] - (beta + w["y1"]) * (beta + w["y1"])
                    delta = coefb * coefb - coefa * coefc
                    if delta >= 0:
                        sol1 = (-coefb - math.sqrt(delta)) / coefa
                        sol2 = (-coefb + math.sqrt(delta)) / coefa
                        ok1 = sol1 >= 0 and sol1 < dtcoll and beta + alpha * q["vx"] * sol1 + w["y1"] >= 0
                        ok2 = sol2 >= 0 and sol2 < dtcoll and beta + alpha * q["vx"] * sol2 + w["y1"] >= 0
                        if ok1:
                            test = ([sol1, sol2] if sol1 <= sol2 else [sol2, sol1]) if ok2 else [sol1]
                        elif ok2:
                            test = [sol2]
                for t in test:
                    move_axi(q, t)
                    if (q["vx"] * w["normx"] + q["vy"] * w["normy"]) < 0:
                        colldist = ((q["x"] - w["x1"]) * (w["x2"] - w["x1"]) + (q["y"] - w["y1"]) * (w["y2"] - w["y1"])) / \
                                   ((w["x2"] - w["x1"]) * (w["x2"] - w["x1"]) + (w["y2"] - w["y1"]) * (w["y2"] - w["y1"]))
                        if colldist >= 0 and colldist <= 1:
                            dtcoll = t
                            boundcoll = -1
                            wallcoll = k + 1
                    move_axi(q, -t)
            if boundcoll != -1:
                events.append(["side", boundcoll])
                assert g["specular"][boundcoll - 1]
                move_axi(q, dtcoll)
                if boundcoll <= 2:
                    q["vx"] = -q["vx"]
                else:
                    q["vy"] = -q["vy"]
                dtrim = dtrim - dtcoll
            elif wallcoll != -1:
                w = walls[wallcoll - 1]
                events.append(["wall", wallcoll])
                move_axi(q, dtcoll)
                vdotn = q["vx"] * w["normx"] + q["vy"] * w["normy"]
                q["vx"] = q["vx"] - 2.0 * vdotn * w["normx"]
                q["vy"] = q["vy"] - 2.0 * vdotn * w["normy"]
                dtrim = dtrim - dtcoll
            else:
                move_axi(q, dtrim)
                dtrim = 0.0
        ddx = (g["xmax"] - g["xmin"]) / g["nx"]
        ddy = (g["ymax"] - g["ymin"]) / g["ny"]
        q["ic"] = int((q["x"] - g["xmin"]) / ddx) + g["nx"] * int((q["y"] - g["ymin"]) / ddy) + 1
        out.append((q, events))
    return out


def boundaries_emit_2d(mesh, species, mix, tasks, dt, rf, axi, zmin, zmax):
    """src/timecycle.f90:368-547, DIMS == 2, no colocated electrons; tasks in order, one rf() stream.  task: dict ic, iface,
    iv1, iv2, nfs[], ux, uy, uz, u_norm, ttra, trot, tvib.  Returns the particles in creation order."""
    made = []
    for task in tasks:
        ic = task["ic"]
        x1, y1 = mesh["nodes"][task["iv1"] - 1]
        x2, y2 = mesh["nodes"][task["iv2"] - 1]
        n = mesh["normal"][ic - 1][task["iface"] - 1]
        fn = [n[0], n[1], 0.0]
        t1 = [-fn[1], fn[0], 0.0]
        t2 = [0.0, 0.0, 1.0]
        for k, (name, _) in enumerate(mix):
            m = species[name]["mass"]
            nfs = math.floor(task["nfs"][k])
            if task["nfs"][k] - float(nfs) >= rf():
                nfs = nfs + 1
            beta = 0.0 if task["ttra"] == 0 else 1.0 / math.sqrt(2.0 * fp1.KB / m * task["ttra"])
            for _ip in range(nfs):
                _, v_tang1, v_tang2 = fp1.maxwell(rf, 0.0, 0.0, 0.0, task["ttra"], task["ttra"], task["ttra"], m)
                erot = fp1.internal_energy(rf, species[name]["rotdof"], task["trot"])
                evib = fp1.internal_energy(rf, species[name]["vibdof"], task["tvib"])
                v_norm = 0.0 if task["ttra"] == 0 else fp1.flx(rf, task["u_norm"] * beta, task["ttra"], m)
                vx = task["ux"] - v_norm * fn[0] - v_tang1 * t1[0] - v_tang2 * t2[0]
                vy = task["uy"] - v_norm * fn[1] - v_tang1 * t1[1] - v_tang2 * t2[1]
                vz = task["uz"] - v_norm * fn[2] - v_tang1 * t1[2] - v_tang2 * t2[2]
                dtfrac = rf() * dt
                r = rf()
                if axi:
                    if y1 == y2:
                        y = y1
                        x = x1 + r * (x2 - x1)
                    elif y1 < y2:
                        y = math.sqrt(y1 * y1 + r * (y2 * y2 - y1 * y1))
                        x = x1 + (y - y1) / (y2 - y1) * (x2 - x1)
                    else:
                        y = math.sqrt(y2 * y2 + r * (y1 * y1 - y2 * y2))
                        x = x2 + (y - y2) / (y1 - y2) * (x1 - x2)
                    z = 0.0
                else:
                    x = x1 + r * (x2 - x1)
                    y = y1 + r * (y2 - y1)
                    z = zmin + (zmax - zmin) * rf()
                made.append(dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, erot=erot, evib=evib, sid=name, ic=ic, dtrim=dtfrac))
    return made


def initial_seed_12d(d, nodes, cells, vols, species, mixtures, spwt, tasks, fnum, rf, axi, lim):
    """src/initialization.f90:2017-2233, GRID_TYPE == UNSTRUCTURED, DIMS == 1 or 2, one rank, no solid cells.  lim = (ymin,
    ymax, zmin, zmax).  Loops: task, mixture component, cell."""
    ymin, ymax, zmin, zmax = lim
    made = []
    for task in tasks:
        for name, molfrac in mixtures[task["mix_id"]]:
            for ic in range(1, len(cells) + 1):
                np_init = fp1.randint(rf, task["nrho"] / (fnum * spwt[name]) * vols[ic - 1] * molfrac)
                if np_init == 0:
                    continue
                v = [nodes[k - 1] for k in cells[ic - 1]]
                for _ip in range(np_init):
                    sd, td, ud = rf(), rf(), rf()
                    if sd + td > 1:
                        sd, td = 1 - sd, 1 - td
                    if d == 1:
                        xp = v[0][0] + (v[1][0] - v[0][0]) * sd
                        yp = ymin + (ymax - ymin) * td
                        zp = zmin + (zmax - zmin) * ud
                    else:
                        xp = v[0][0] + (v[1][0] - v[0][0]) * sd + (v[2][0] - v[0][0]) * td
                        yp = v[0][1] + (v[1][1] - v[0][1]) * sd + (v[2][1] - v[0][1]) * td
                        zp = 0.0 if axi else zmin + (zmax - zmin) * ud
                    m = species[name]["mass"]
                    vx, vy, vz = fp1.maxwell(rf, task["ux"], task["uy"], task["uz"], task["ttrax"], task["ttray"], task["ttraz"], m)
                    erot = fp1.internal_energy(rf, species[name]["rotdof"], task["trot"])
                    evib = fp1.internal_energy(rf, species[name]["vibdof"], task["tvib"])
                    made.append(dict(x=xp, y=yp, z=zp, vx=vx, vy=vy, vz=vz, erot=erot, evib=evib, sid=name, ic=ic))
    return made


# ------------------------------------------------------------------------------------------ MCC_COLLISIONS
def mcc_collisions(species, parts, reactions, bg_mix, bg_dens, bg_ttra, dt, rf):
    """src/collisions.f90:1261-1595 with BOOL_BG_DENSITY_FILE = .FALSE.  parts: dicts pos, v, erot, evib, sid (name), in store
    order; new particles are appended to `parts` (ADD_PARTICLE_ARRAY).  bg_mix: [(name, molfrac)].  reactions: dicts r1, r2,
    p1, p2, p3, nprod, ea, c1, c2, c3 (TCE).  Returns (TIMESTEP_COLL, TIMESTEP_REAC)."""
    pi2 = 2 * fp1.PI
    ncoll = nreac = 0
    np_initial = len(parts)
    for jp1 in range(np_initial):
        a = parts[jp1]
        sp_id1 = a["sid"]                                     # read once per particle, :1292
        for sp_id2, frac in bg_mix:
            bg_nrho = frac * bg_dens
            if bg_nrho == 0:
                continue
            s1, s2 = species[sp_id1], species[sp_id2]
            pr = fp1.vss_pair(s1, s2)
            cref, sigma, omega, alpha = pr["cref"], pr["sigma"], pr["omega"], pr["alpha"]
            c2 = list(fp1.maxwell(rf, 0.0, 0.0, 0.0, bg_ttra, bg_ttra, bg_ttra, s2["mass"]))
            vr2 = (c2[0] - a["v"][0]) * (c2[0] - a["v"][0]) + (c2[1] - a["v"][1]) * (c2[1] - a["v"][1]) + (c2[2] - a["v"][2]) * (c2[2] - a["v"][2])
            vr = math.sqrt(vr2)
            p_coll = 1 - math.exp(-dt * bg_nrho * vr * sigma * (vr / cref) ** (1.0 - 2.0 * omega))
            if not (rf() < p_coll):
                continue
            ncoll += 1
            c1 = list(a["v"])
            erot = fp1.internal_energy(rf, s2["rotdof"], bg_ttra)
            evib = fp1.internal_energy(rf, s2["vibdof"], bg_ttra)
            b = dict(pos=list(a["pos"]), v=list(c2), erot=erot, evib=evib, sid=sp_id2, ic=a["ic"])
            parts.append(b)
            m1, m2 = s1["mass"], s2["mass"]
            mred = m1 * m2 / (m1 + m2)
            etr = 0.5 * mred * vr2
            ecoll = etr + a["erot"] + b["erot"] + a["evib"] + b["evib"]
            skip = False
            for r in reactions:
                if r["r1"] != sp_id1:
                    if r["r1"] != sp_id2:
                        continue
                    elif r["r2"] != sp_id1:
                        continue
                elif r["r2"] != sp_id2:
                    continue
                ea = r["ea"]
                if ecoll <= ea:
                    continue
                ptce = r["c1"] * (ecoll - ea) ** r["c2"] * (1.0 - ea / ecoll) ** r["c3"]
                rfp = rf()
                if rfp < ptce:
                    skip = True
                    nreac += 1
                    ecoll = ecoll - ea
                    p1, p2 = species[r["p1"]], species[r["p2"]]
                    a["sid"], b["sid"] = r["p1"], r["p2"]
                    totdof = 3.0 + p2["rotdof"] + p2["vibdof"] + p1["rotdof"]
                    if r["nprod"] == 3:
                        p3 = species[r["p3"]]
                        totdof = totdof + 3.0 + p3["rotdof"] + p3["vibdof"]
                    ei = fp1.coll_internal_energy(rf, ecoll, totdof, p1["vibdof"])
                    a["evib"] = ei
                    ecoll = ecoll - ei
                    totdof = totdof - p1["rotdof"]
                    ei = fp1.coll_internal_energy(rf, ecoll, totdof, p1["rotdof"])
                    a["erot"] = ei
                    ecoll = ecoll - ei
                    totdof = totdof - p2["vibdof"]
                    ei = fp1.coll_internal_energy(rf, ecoll, totdof, p2["vibdof"])
                    b["evib"] = ei
                    ecoll = ecoll - ei
                    totdof = totdof - p2["rotdof"]
                    ei = fp1.coll_internal_energy(rf, ecoll, totdof, p2["rotdof"])
                    b["erot"] = ei
                    ecoll = ecoll - ei
                    mm1 = p1["mass"]
                    mm2 = p2["mass"] if r["nprod"] == 2 else p2["mass"] + p3["mass"]
                    totdof = totdof - 3.0
                    ei = fp1.coll_internal_energy(rf, ecoll, totdof, 3)
                    c1, c2 = fp1.hs_scatter(rf, ei, mm1, mm2, c1, c2)
                    ecoll = ecoll - ei
                    a["v"] = list(c1)
                    if r["nprod"] == 2:
                        b["v"] = list(c2)
                    else:
                        totdof = totdof - p3["vibdof"]
                        evib3 = fp1.coll_internal_energy(rf, ecoll, totdof, p3["vibdof"])
                        ecoll = ecoll - evib3
                        totdof = totdof - p3["rotdof"]
                        erot3 = fp1.coll_internal_energy(rf, ecoll, totdof, p3["rotdof"])
                        ecoll = ecoll - erot3
                        c1 = list(c2)
                        c1, c2 = fp1.hs_scatter(rf, ecoll, p2["mass"], p3["mass"], c1, c2)
                        b["v"] = list(c1)
                        parts.append(dict(pos=list(b["pos"]), v=list(c2), erot=erot3, evib=evib3, sid=r["p3"], ic=b["ic"]))
                    break
            if skip:
                continue
            trdof = 3.0
            prot1 = (trdof + s1["rotdof"]) / trdof * s1["rotrel"]
            prot2 = (trdof + s2["rotdof"]) / trdof * s2["rotrel"]
            pvib1 = (trdof + s1["vibdof"]) / trdof * s1["vibrel"]
            pvib2 = (trdof + s2["vibdof"]) / trdof * s2["vibrel"]
            trdof = 5.0 - 2.0 * omega
            rfp = rf()
            if rfp < prot1:
                ecoll = etr + a["erot"]
                a["erot"] = fp1.coll_internal_energy(rf, ecoll, trdof, s1["rotdof"])
                etr = ecoll - a["erot"]
            elif rfp < prot1 + prot2:
                ecoll = etr + b["erot"]
                b["erot"] = fp1.coll_internal_energy(rf, ecoll, trdof, s2["rotdof"])
                etr = ecoll - b["erot"]
            elif rfp < prot1 + prot2 + pvib1:
                ecoll = etr + a["evib"]
                a["evib"] = fp1.coll_internal_energy(rf, ecoll, trdof, s1["vibdof"])
                etr = ecoll - a["evib"]
            elif rfp < prot1 + prot2 + pvib1 + pvib2:
                ecoll = etr + b["evib"]
                b["evib"] = fp1.coll_internal_energy(rf, ecoll, trdof, s2["vibdof"])
                etr = ecoll - b["evib"]
            coschi = 2.0 * rf() ** (1.0 / alpha) - 1.0
            sinchi = math.sqrt(1.0 - coschi * coschi)
            theta = pi2 * rf()
            costheta, sintheta = math.cos(theta), math.sin(theta)
            if rfp < prot1 + prot2 + pvib1 + pvib2:
                g = math.sqrt(2.0 * etr / mred)
                cosa = 2.0 * rf() - 1.0
                sina = math.sqrt(1 - cosa * cosa)
                bb = pi2 * rf()
                gx, gy, gz = g * sina * math.cos(bb), g * sina * math.sin(bb), g * cosa
            else:
                gx, gy, gz = c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]
                g = math.sqrt(gx * gx + gy * gy + gz * gz)
            grel = [gx * coschi + math.sqrt(gy * gy + gz * gz) * sintheta * sinchi,
                    gy * coschi + (g * gz * costheta - gx * gy * sintheta) / math.sqrt(gy * gy + gz * gz) * sinchi,
                    gz * coschi - (g * gy * costheta + gx * gz * sintheta) / math.sqrt(gy * gy + gz * gz) * sinchi]
            w = [m1 / (m1 + m2) * c1[i] + m2 / (m1 + m2) * c2[i] for i in range(3)]
            a["v"] = [w[i] + m2 / (m1 + m2) * grel[i] for i in range(3)]
            b["v"] = [w[i] - m1 / (m1 + m2) * grel[i] for i in range(3)]
    return ncoll, nreac


# ------------------------------------------------------------------------------------------ DSMC_COLLISIONS order
def dsmc_collisions(species, names, parts, cells, ncells, sigmamax, fnum, dt, vol, rf):
    """src/collisions.f90:37-149: INDALL(IOFALL(JC) + NPCALL(JC)) = JP for JP = 1..NP_PROC (stable in store order), then
    VSS_COLLIS cell by cell in ascending JC for the cells holding more than one particle.  `parts` in store order,
    cells[k] = IC of particle k.  Returns TIMESTEP_COLL."""
    indall = {jc: [] for jc in range(1, ncells + 1)}
    for jp, jc in enumerate(cells):
        indall[jc].append(jp)
    total = 0
    for jc in range(1, ncells + 1):
        if len(indall[jc]) > 1:
            local = [parts[jp] for jp in indall[jc]]          # VSS_COLLIS indexes INDALL(IOFALL(JC) + INT(NPC*rf()))
            nc, nr, created = fp1.vss_collis_cell(species, names, local, [], sigmamax, fnum, dt, vol, rf)
            assert not created
            total += nc
    return total


# ------------------------------------------------------------------------------------------ CELL_FROM_POSITION
def binary_search(value, array):
    """BINARY_SEARCH, src/tools.f90:1499-1535 (1-based INDEX; -1 outside the array)."""
    lo, hi = 1, len(array)
    if value < array[lo - 1] or value > array[hi - 1]:
        return -1
    if hi == lo + 1:
        return lo
    while True:
        index = (lo + hi) // 2
        if array[index - 1] <= value:
            if array[index] > value:
                return index
            lo = index
        else:
            if array[index - 2] <= value:
                return index - 1
            hi = index


def cell_from_position_nonuniform(xp, yp, xcoord, ycoord, nx):
    """src/grid_and_partition.f90:62-90 (RECTILINEAR_NONUNIFORM): XCELL = BINARY_SEARCH(XP, XCOORD), likewise Y;
    IDCELL = XCELL + NX*(YCELL-1)."""
    return binary_search(xp, xcoord) + nx * (binary_search(yp, ycoord) - 1)


def pdict(q):
    return {k: (hx(v) if isinstance(v, float) else v) for k, v in q.items()}


def main():
    import random
    out = {}
    air = fp1.read_species("air")
    # ---- axi_walk
    mesh = tri_mesh()
    out["tri_mesh"] = mesh
    dt = 6.0e-5

    def locate(x, y):
        for c, cn in enumerate(mesh["cells"]):
            ok = True
            for j in range(3):
                x1, y1 = mesh["nodes"][cn[j] - 1]
                n1, n2 = mesh["normal"][c][j]
                if (x - x1) * n1 + (y - y1) * n2 > 0:
                    ok = False
            if ok:
                return c + 1
        raise AssertionError((x, y))

    starts = [
        dict(x=0.012, y=0.018, z=0.0, vx=900.0, vy=350.0, vz=-420.0),      # several hops, azimuthal velocity folds into the radius
        dict(x=0.071, y=0.041, z=0.0, vx=1400.0, vy=600.0, vz=150.0),      # reflects at x = 0.09 (vertical edge) and y = 0.05
        dict(x=0.044, y=0.016, z=0.0, vx=-300.0, vy=-900.0, vz=260.0),     # reflects at the inner radius y = 0.01
        dict(x=0.008, y=0.046, z=0.0, vx=-1100.0, vy=-200.0, vz=-800.0),   # reflects at x = 0, large swirl
        dict(x=0.05, y=0.03, z=0.0, vx=12.0, vy=-7.0, vz=5.0),             # slow: stays in its cell, trial moves only
    ]
    walk = []
    for p in starts:
        p["ic"] = locate(p["x"], p["y"])
        q, trace, ntrial = advect_axi(mesh, p, dt)
        walk.append(dict(inp=pdict(p), out=pdict(q), trace=trace, trial_moves=ntrial))
    assert any(len(w["trace"]) >= 3 for w in walk) and walk[-1]["trace"] == []
    pgs = {mesh["faces_pg"][c - 1][f - 1] for w in walk for c, f in w["trace"] if mesh["neighbors"][c - 1][f - 1] == -1}
    assert {1, 2, 3, 4} <= pgs, pgs                                          # every wall of the box was hit by some case
    out["axi_walk"] = dict(dt=hx(dt), cases=walk)

    # ---- tri_walk_solid: the same triangles, plane geometry; cells 7 and 8 (the quad in the middle of the upper row) belong to
    # a SOLID volume group 5 and their outer edges carry the specular group 6
    solid_cells = [7, 8]
    cell_pg = [-1] * len(mesh["cells"])
    for c in solid_cells:
        cell_pg[c - 1] = 5
    faces_pg2 = [list(r) for r in mesh["faces_pg"]]
    for c in range(len(mesh["cells"])):
        for j in range(3):
            nbc = mesh["neighbors"][c][j]
            if nbc != -1 and ((c + 1 in solid_cells) != (nbc in solid_cells)):
                faces_pg2[c][j] = 6
    mesh2 = dict(mesh, faces_pg=faces_pg2)
    out["tri_solid"] = dict(cell_pg=cell_pg, faces_pg=faces_pg2, solid_group=5, wall_group=6)
    dtp = 8.0e-5
    tri_cases = []
    for p in [dict(x=0.012, y=0.018, z=0.1, vx=900.0, vy=150.0, vz=-420.0),
              dict(x=0.085, y=0.03, z=-0.2, vx=-1100.0, vy=-250.0, vz=300.0),
              dict(x=0.045, y=0.015, z=0.0, vx=60.0, vy=1300.0, vz=50.0),
              dict(x=0.02, y=0.04, z=0.3, vx=700.0, vy=100.0, vz=10.0)]:
        p["ic"] = locate(p["x"], p["y"])
        assert p["ic"] not in solid_cells
        q, trace = advect_tri_plane(mesh2, p, dtp, cell_pg, {5})
        tri_cases.append(dict(inp=pdict(p), out=pdict(q), trace=trace))
    hits_solid = sum(1 for c in tri_cases for (cc, f, fb) in c["trace"] if fb and mesh["neighbors"][cc - 1][f - 1] != -1)
    assert hits_solid >= 2 and all(c["out"]["ic"] not in solid_cells for c in tri_cases), (hits_solid, [c["trace"] for c in tri_cases])
    out["tri_walk_solid"] = dict(dt=hx(dtp), cases=tri_cases)

    # ---- emit_2d: three tasks on boundary edges of the pin triangles -- x = 0 traversed with Y1 > Y2, x = 0.09 with Y1 < Y2,
    # y = 0.01 with Y1 == Y2 -- N2/O2 mixture, plane and axisymmetric
    def edge_task(cell, face, **kw):
        cn = mesh["cells"][cell - 1]
        return dict(ic=cell, iface=face, iv1=cn[face - 1], iv2=cn[face % 3], **kw)
    etasks = []
    for cell, faces in enumerate(mesh["faces_pg"], start=1):
        for face, pg in enumerate(faces, start=1):
            if pg in (1, 2, 3) and not any(mesh["faces_pg"][t["ic"] - 1][t["iface"] - 1] == pg for t in etasks):
                n = mesh["normal"][cell - 1][face - 1]
                u = (120.0, -40.0, 15.0)
                etasks.append(edge_task(cell, face, nfs=[2.3, 1.6], ux=u[0], uy=u[1], uz=u[2], u_norm=-(u[0] * n[0] + u[1] * n[1]),
                                        ttra=700.0, trot=500.0, tvib=900.0))
    ys = [(mesh["nodes"][t["iv1"] - 1][1], mesh["nodes"][t["iv2"] - 1][1]) for t in etasks]
    assert len(etasks) == 3 and any(a > b for a, b in ys) and any(a < b for a, b in ys) and any(a == b for a, b in ys), ys
    mixe = [("N2", 0.79), ("O2", 0.21)]
    out["emit_2d"] = {}
    for label, axi_flag, (zlo, zhi) in (("plane", False, (-0.5, 0.5)), ("axi", True, (0.0, 1.0))):
        rr = random.Random(4242 + int(axi_flag))
        draws = [rr.random() for _ in range(400)]
        s = Script(draws)
        made = boundaries_emit_2d(mesh, air, mixe, etasks, 3.0e-6, s, axi_flag, zlo, zhi)
        assert len(made) >= 9
        out["emit_2d"][label] = dict(tasks=etasks, mix=mixe, dt=hx(3.0e-6), zmin=zlo, zmax=zhi, draws=[hx(v) for v in draws[:s.k]], used=s.k,
                                     out=[dict(pdict(q), dtrim=hx(q["dtrim"])) for q in made])

    # ---- seed_2d_1d: INITIAL_SEED on the pin triangles (plane, axisymmetric) and on a 5-segment line
    def tri_vols(axi_flag, depth):
        v = []
        for cn in mesh["cells"]:
            (ax_, ay), (bx, by), (cx, cy) = (mesh["nodes"][k - 1] for k in cn)
            area = 0.5 * abs(ax_ * (by - cy) + bx * (cy - ay) + cx * (ay - by))
            v.append(area * depth * ((ay + by + cy) / 3.0) if axi_flag else area * depth)     # grid_and_partition.f90:1315-1323
        return v
    seed_tasks = [dict(mix_id=0, nrho=8.0e3, ux=30.0, uy=-20.0, uz=5.0, ttrax=400.0, ttray=0.0, ttraz=650.0, trot=500.0, tvib=300.0),
                  dict(mix_id=1, nrho=3.0e3, ux=0.0, uy=0.0, uz=0.0, ttrax=900.0, ttray=900.0, ttraz=900.0, trot=0.0, tvib=1200.0)]
    seed_mix = [[("N2", 0.79), ("O2", 0.21)], [("O2", 1.0)]]
    seed_spwt = {"N2": 1.0, "O2": 2.0}
    out["seed_2d_1d"] = {}
    for label, dsz, axi_flag, lim in (("plane", 2, False, (0.01, 0.05, -0.5, 0.5)), ("axi", 2, True, (0.01, 0.05, 0.0, 1.0)), ("line", 1, False, (0.0, 0.2, -0.1, 0.3))):
        rr = random.Random(99 + len(label))
        draws = [rr.random() for _ in range(4000)]
        if dsz == 2:
            nds, cls, vls = mesh["nodes"], mesh["cells"], tri_vols(axi_flag, lim[3] - lim[2])
            fn = 1.0 if not axi_flag else 0.03
        else:
            xs = [0.0, 0.011, 0.019, 0.034, 0.041, 0.05]
            nds = [[x, 0.0] for x in xs]
            cls = [[k + 1, k + 2] for k in range(5)]
            vls = [abs(xs[k + 1] - xs[k]) * (lim[1] - lim[0]) * (lim[3] - lim[2]) for k in range(5)]     # LENGTH*(YMAX-YMIN)*(ZMAX-ZMIN), :870-871
            fn = 2.0
        s = Script(draws)
        made = initial_seed_12d(dsz, nds, cls, vls, air, seed_mix, seed_spwt, seed_tasks, fn, s, axi_flag, lim)
        assert 10 <= len(made) <= 200, (label, len(made))
        out["seed_2d_1d"][label] = dict(dims=dsz, axi=axi_flag, lim=lim, fnum=hx(fn), nodes=nds, cells=cls, vols=[hx(v) for v in vls],
                                        tasks=seed_tasks, mixtures=seed_mix, spwt=seed_spwt, draws=[hx(v) for v in draws[:s.k]], used=s.k,
                                        out=[pdict(q) for q in made])

    # ---- vinject_tri: VOLUME_INJECT on the plane pin triangles (the loop body is INITIAL_SEED's with the rate DT*NRHODOT)
    dtv = 2.0e-6
    vtasks = [dict(mix_id=0, nrhodot=5.0e9, ux=-15.0, uy=25.0, uz=0.0, ttrax=300.0, ttray=450.0, ttraz=600.0, trot=350.0, tvib=0.0)]
    rr = random.Random(5150)
    draws = [rr.random() for _ in range(3000)]
    s = Script(draws)
    lim = (0.01, 0.05, -0.5, 0.5)
    made = initial_seed_12d(2, mesh["nodes"], mesh["cells"], tri_vols(False, 1.0), air, seed_mix, seed_spwt,
                            [dict(t, nrho=dtv * t["nrhodot"]) for t in vtasks], 1.0, s, False, lim)
    assert 10 <= len(made) <= 150, len(made)
    out["vinject_tri"] = dict(dt=hx(dtv), lim=lim, tasks=vtasks, mixtures=seed_mix, spwt=seed_spwt, vols=[hx(v) for v in tri_vols(False, 1.0)],
                              draws=[hx(v) for v in draws[:s.k]], used=s.k, out=[pdict(q) for q in made])

    # ---- tri_diffuse: plane triangles, the walls x = 0 (group 1, 800 K) and y = 0.01 (group 3, 450 K) diffuse, the others specular
    dtd = 6.0e-5
    rr = random.Random(31337)
    draws = [rr.random() for _ in range(200)]
    s = Script(draws)
    dcases = []
    for p in [dict(x=0.012, y=0.02, z=0.1, vx=-900.0, vy=150.0, vz=-420.0, erot=1.0e-21, evib=2.0e-21, sid="N2"),
              dict(x=0.05, y=0.016, z=-0.2, vx=200.0, vy=-800.0, vz=300.0, erot=3.0e-21, evib=0.5e-21, sid="O2")]:
        p["ic"] = locate(p["x"], p["y"])
        pin = dict(p)
        q, trace = advect_tri_plane(mesh, p, dtd, [-1] * len(mesh["cells"]), set(), diffuse={1: 800.0, 3: 450.0}, species=air, rf=s)
        q["erot"], q["evib"] = p["erot"], p["evib"]
        dcases.append(dict(inp=pdict(pin), out=pdict(q), trace=trace))
    assert s.k >= 16 and all(any(fb for _, _, fb in c["trace"]) for c in dcases), (s.k, [c["trace"] for c in dcases])
    out["tri_diffuse"] = dict(dt=hx(dtd), walls={"1": 800.0, "3": 450.0}, draws=[hx(v) for v in draws[:s.k]], used=s.k, cases=dcases)

    # ---- structured_walk: 4 x 3 cells on [0, 0.4] x [0, 0.3]; x periodic, ymin specular, ymax diffuse at 450 K; wall 1 a plain
    # specular segment at x = 0.2 facing -x, wall 2 a porous specular segment at y = 0.15 facing +y, wall 3 diffuse at x = 0.3
    g = dict(xmin=0.0, xmax=0.4, ymin=0.0, ymax=0.3, zmin=-0.5, zmax=0.5, nx=4, ny=3, periodic=[1, 1, 0, 0], specular=[0, 0, 1, 0],
             diffuse=[0, 0, 0, 1], boundtemp=450.0, z_periodic=1)
    walls = [dict(x1=0.2, y1=0.05, x2=0.2, y2=0.2, normx=-1.0, normy=0.0, specular=1, diffuse=0, porous=0, transm=0.0, temp=300.0),
             dict(x1=0.22, y1=0.15, x2=0.38, y2=0.15, normx=0.0, normy=1.0, specular=1, diffuse=0, porous=1, transm=0.5, temp=300.0),
             dict(x1=0.3, y1=0.2, x2=0.3, y2=0.29, normx=1.0, normy=0.0, specular=0, diffuse=1, porous=0, transm=0.0, temp=700.0)]
    dts = 1.0e-4
    base = dict(z=0.1, erot=1.0e-21, evib=2.0e-21, sid="N2")
    det = [dict(base, x=0.35, y=0.04, vx=900.0, vy=-700.0, vz=3000.0),       # xmax periodic wrap, ymin specular, z wraps
           dict(base, x=0.12, y=0.1, vx=1500.0, vy=300.0, vz=-200.0),         # wall 1 (draw consumed, not porous), then xmin wrap
           dict(base, x=0.15, y=0.21, vx=1000.0, vy=400.0, vz=-9000.0),       # passes above wall 1 (COLLDIST > 1): free flight, z wraps
           dict(base, x=0.05, y=0.02, vx=-2600.0, vy=-900.0, vz=50.0),        # ymin specular, xmin wrap, far end
           dict(base, x=0.25, y=0.1, vx=-800.0, vy=100.0, vz=0.0)]            # behind wall 1 (DX < 0): one-sided, flies through
    s = Script([0.6, 0.6, 0.6, 0.6, 0.6, 0.6])
    res = advect_structured(g, walls, air, det, dts, s)
    assert [e for q, ev in res for e in ev if e[0] == "wall"] == [["wall", 1]] and s.k == 1 and all(q is not None for q, _ in res)
    out["structured_walk"] = dict(grid=g, walls=walls, dt=hx(dts),
                                  det=[dict(inp=pdict(p), out=pdict(q), events=ev) for p, (q, ev) in zip(det, res)])
    sto = [dict(base, x=0.3, y=0.2, vx=100.0, vy=-1200.0, vz=10.0, sid="O2"),     # wall 2, draw 0.3 <= 0.5: removed
           dict(base, x=0.3, y=0.2, vx=100.0, vy=-1200.0, vz=10.0, sid="O2"),     # wall 2, draw 0.8: reflected
           dict(base, x=0.1, y=0.25, vx=300.0, vy=1400.0, vz=-20.0),              # ymax diffuse with BOUNDTEMP
           dict(base, x=0.36, y=0.25, vx=-1300.0, vy=20.0, vz=5.0, sid="O2")]     # wall 3 diffuse at 700 K (draw for the porous test first)
    draws = [0.3, 0.8,
             0.11, 0.52, 0.73, 0.29, 0.91, 0.64, 0.35, 0.48, 0.95, 0.99, 0.62, 0.17,
             0.44, 0.21, 0.77, 0.58, 0.39, 0.83, 0.26, 0.67, 0.09, 0.5, 0.4, 0.55, 0.3, 0.61, 0.2, 0.72, 0.1, 0.5, 0.05]
    s = Script(draws)
    res = advect_structured(g, walls, air, sto, dts, s)
    assert res[0][0] is None and res[1][0] is not None and res[2][1][0] == ["side", 4] and res[3][1][0] == ["wall", 3], [ev for _, ev in res]
    out["structured_walk"]["sto"] = dict(draws=draws[:s.k], used=s.k,
                                          cases=[dict(inp=pdict(p), out=(pdict(q) if q else None), events=ev) for p, (q, ev) in zip(sto, res)])

    # ---- structured_axi: 4 x 3 cells on [0, 0.4] x [0.02, 0.17] (y = radius), every side specular; wall 1 a cone segment,
    # wall 2 a vertical disc
    ga = dict(xmin=0.0, xmax=0.4, ymin=0.02, ymax=0.17, zmin=0.0, zmax=1.0, nx=4, ny=3, periodic=[0, 0, 0, 0], specular=[1, 1, 1, 1],
              diffuse=[0, 0, 0, 0], boundtemp=300.0, z_periodic=0)
    s45 = math.sqrt(0.5)
    wa = [dict(x1=0.1, y1=0.05, x2=0.2, y2=0.15, normx=-s45, normy=s45, specular=1, diffuse=0, porous=0, transm=0.0, temp=300.0),
          dict(x1=0.3, y1=0.03, x2=0.3, y2=0.1, normx=-1.0, normy=0.0, specular=1, diffuse=0, porous=0, transm=0.0, temp=300.0)]
    dta = 2.0e-4
    ax_cases = [dict(x=0.05, y=0.12, z=0.0, vx=300.0, vy=500.0, vz=420.0),        # outer radius (side 4), swirl
                dict(x=0.35, y=0.05, z=0.0, vx=700.0, vy=-250.0, vz=90.0),         # inner radius (side 3) and x = 0.4
                dict(x=0.1, y=0.14, z=0.0, vx=450.0, vy=-350.0, vz=60.0),          # the cone wall from above
                dict(x=0.22, y=0.06, z=0.0, vx=600.0, vy=40.0, vz=-30.0),          # the disc at x = 0.3
                dict(x=0.02, y=0.09, z=0.0, vx=-500.0, vy=15.0, vz=800.0)]         # x = 0 with a large azimuthal velocity
    resa = advect_structured_axi(ga, wa, ax_cases, dta)
    seen = {tuple(e) for _, ev in resa for e in ev}
    assert {("side", 1), ("side", 2), ("side", 3), ("side", 4), ("wall", 1), ("wall", 2)} <= seen, seen
    out["structured_axi"] = dict(grid=ga, walls=wa, dt=hx(dta), cases=[dict(inp=pdict(p), out=pdict(q), events=ev) for p, (q, ev) in zip(ax_cases, resa)])

    # ---- mcc_collisions: three test particles (a slow O2, a slow N2, a 14 km/s N2) in an O2 0.21 / N2 0.79 background at 800 K;
    # random script; the fast particle's collision energy is above the O2 dissociation threshold: N2 + O2 --> N2 + O + O can fire
    names5 = ["O2", "N2", "O", "N", "NO"]
    tce = fp1.read_tce(air)
    reactions = []
    for r in tce:
        if all(x in names5 for x in r["react"] + r["prod"]):
            c1, c2, c3 = fp1.tce_coeffs(air, r["react"][0], r["react"][1], r["A"], r["n"], r["ea"])
            reactions.append(dict(r1=r["react"][0], r2=r["react"][1], p1=r["prod"][0], p2=r["prod"][1],
                                  p3=r["prod"][2] if len(r["prod"]) > 2 else None, nprod=len(r["prod"]), ea=r["ea"], c1=c1, c2=c2, c3=c3))
    bg_mix, bg_dens, bg_ttra, dtm = [("O2", 0.21), ("N2", 0.79)], 4.0e21, 800.0, 1.0e-6
    best = None
    for seed in range(400):
        rr = random.Random(seed)
        draws = [rr.random() for _ in range(400)]
        mparts = [dict(pos=[0.1, 0.2, 0.0], v=[900.0, -300.0, 150.0], erot=2.0e-21, evib=1.0e-21, sid="O2", ic=1),
                  dict(pos=[0.6, 0.3, 0.1], v=[-400.0, 700.0, 250.0], erot=3.0e-21, evib=0.5e-21, sid="N2", ic=1),
                  dict(pos=[0.4, 0.8, -0.2], v=[14000.0, -2000.0, 1000.0], erot=4.0e-21, evib=6.0e-21, sid="N2", ic=1)]
        m_in = json.loads(json.dumps(mparts))
        s = Script(draws)
        nc, nr = mcc_collisions(air, mparts, reactions, bg_mix, bg_dens, bg_ttra, dtm, s)
        if nr >= 1 and nc >= 3 and any(q["sid"] == "O" for q in mparts[3:]):
            best = (seed, draws[:s.k], m_in, mparts, nc, nr)
            break
    assert best, "no script with a reaction found"
    seed, draws, m_in, mparts, nc, nr = best
    pm = lambda q: dict(pos=q["pos"], v=[hx(c) for c in q["v"]], erot=hx(q["erot"]), evib=hx(q["evib"]), sid=q["sid"])
    out["mcc_collisions"] = dict(species=names5, bg_mix=bg_mix, bg_dens=hx(bg_dens), bg_ttra=hx(bg_ttra), dt=hx(dtm), draws=[hx(v) for v in draws],
                                 used=len(draws), ncoll=nc, nreac=nr, inp=[pm(q) for q in m_in], out=[pm(q) for q in mparts])

    # ---- dsmc_order: 7 argon particles in a 2 x 1 structured grid, store order interleaved between the cells; one candidate
    # per cell (NCOLLMAX < 1 -> NCOLL = 1); draws per cell: JP1, JP2, acceptance, the Borgnakke-Larsen selector (no internal
    # mode for argon: nothing exchanged), then COSCHI and THETA of the VSS deflection
    ar = fp1.read_species("ar")
    names = ["Ar"]
    sigmamax = fp1.vss_pair(ar["Ar"], ar["Ar"])["sigma"]
    fnum, dtc, vol = 1.0e10, 1.0e-7, 0.5
    cells = [2, 1, 2, 1, 1, 2, 1]
    vel = [[310.0, -120.0, 45.0], [-250.0, 90.0, 400.0], [120.0, 330.0, -210.0], [-80.0, -410.0, 150.0],
           [505.0, 60.0, -95.0], [-330.0, 270.0, 15.0], [40.0, -75.0, -380.0]]
    parts = [dict(v=list(v), erot=0.0, evib=0.0, sid="Ar", pos=None) for v in vel]
    # cell 1 holds store indices (0-based) 1, 3, 4, 6: INT(4*0.55) = 2 -> particle 4, INT(4*0.05) = 0 -> particle 1
    # cell 2 holds 0, 2, 5: INT(3*0.7) = 2 -> particle 5, INT(3*0.7) = 2 again -> redraw, INT(3*0.34) = 1 -> particle 2
    draws = [0.55, 0.05, 1.0e-12, 0.9, 0.31, 0.77,
             0.7, 0.7, 0.34, 1.0e-12, 0.5, 0.64, 0.12]
    s = Script(draws)
    before = [list(q["v"]) for q in parts]
    total = dsmc_collisions(ar, names, parts, cells, 2, sigmamax, fnum, dtc, vol, s)
    changed = [k for k in range(7) if parts[k]["v"] != before[k]]
    assert total == 2 and s.k == len(draws) and changed == [1, 2, 4, 5], (total, s.k, changed)
    out["dsmc_order"] = dict(cells=cells, sigmamax=hx(sigmamax), fnum=hx(fnum), dt=hx(dtc), draws=draws,
                             inp=[[hx(c) for c in v] for v in before], out=[[hx(c) for c in q["v"]] for q in parts], changed=changed)

    # ---- cell_nonuniform
    xcoord = [0.0, 0.1, 0.15, 0.4, 0.45, 1.0]
    ycoord = [-0.5, -0.2, 0.3, 0.35, 2.0]
    pts = [(0.0, -0.5), (0.05, -0.3), (0.1, -0.2), (0.149999999, 0.29), (0.15, 0.3), (0.42, 0.34), (0.45, 0.35), (0.999, 1.99),
           (0.3, 0.0), (0.0999999999999, 0.3499999999)]
    out["cell_nonuniform"] = dict(xcoord=xcoord, ycoord=ycoord,
                                  cases=[dict(x=hx(x), y=hx(y), ic=cell_from_position_nonuniform(x, y, xcoord, ycoord, len(xcoord) - 1)) for x, y in pts])
    with open(os.path.join(HERE, "fortran_pins2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mcc_collisions: script seed", seed, "->", nc, "collisions,", nr, "reactions,", len(mparts), "particles,", len(draws), "draws")
    print("wrote fortran_pins2.json:", {k: (len(v["cases"]) if isinstance(v, dict) and "cases" in v else "-") for k, v in out.items()})


if __name__ == "__main__":
    main()
