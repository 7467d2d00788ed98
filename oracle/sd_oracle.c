/* TEST INFRASTRUCTURE -- CPU oracle (see sd_oracle.h).  Scalar, one particle at a time, written
 * for clarity: every function names the reference code it restates.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off: the reference is numpy, which never fuses a*b+c).
 */
#include "sd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NOFACE 999
static const double PI = 3.141592653589793;
static const double DEG2M = 6271e3 * 3.141592653589793 / 180; /* sic: lagrangian.py:22, utils.py:306 */
/* DIRECTIONS, topology.py:30 */
static const double DIRECTIONS[4] = {3.141592653589793 / 2, -3.141592653589793 / 2,
                                     3.141592653589793, 0.0};

/* ------------------------------------------------------------------------------------------ */
/* topology.py                                                                                  */
/* ------------------------------------------------------------------------------------------ */
static int fc(const ora_grid *g, int face, int edge) { return g->face_connect[face * 4 + edge]; }

static int where_first(const ora_grid *g, int face, int target) {
    for (int e = 0; e < 4; ++e)
        if (fc(g, face, e) == target) return e;
    return -1;
}

/* _llc_get_the_other_edge, topology.py:84 */
static int other_edge(const ora_grid *g, int face, int edge, int *nf, int *ne) {
    int n = fc(g, face, edge);
    if (n == NOFACE) return ORA_INDEX_ERROR;
    *nf = n;
    *ne = where_first(g, n, face);
    return ORA_OK;
}

/* _llc_mutual_direction, topology.py:34 */
static int mutual_direction(const ora_grid *g, int face, int nface, int transitive, int *edge,
                            int *new_edge) {
    int a = where_first(g, face, nface);
    int b = where_first(g, nface, face);
    if (a >= 0 && b >= 0) {
        *edge = a;
        *new_edge = b;
        return ORA_OK;
    }
    if (!transitive) return ORA_VALUE_ERROR;
    int common = -1;
    for (int e = 0; e < 4 && common < 0; ++e) {
        int cand = fc(g, face, e);
        if (where_first(g, nface, cand) >= 0) common = cand;
    }
    if (common < 0) return ORA_VALUE_ERROR;
    if (common == NOFACE) return ORA_INDEX_ERROR; /* reference indexes face_connect[999] here */
    int edge_1 = where_first(g, face, common);
    int new_edge_1 = where_first(g, common, face);
    int edge_2 = where_first(g, common, nface);
    int new_edge_2 = where_first(g, nface, common);
    if ((edge_1 <= 1 && new_edge_1 <= 1) || (edge_1 >= 2 && new_edge_1 >= 2)) {
        *edge = edge_2;
        *new_edge = new_edge_2;
        return ORA_OK;
    }
    if ((edge_2 <= 1 && new_edge_2 <= 1) || (edge_2 >= 2 && new_edge_2 >= 2)) {
        *edge = edge_1;
        *new_edge = new_edge_1;
        return ORA_OK;
    }
    return ORA_VALUE_ERROR; /* NotImplementedError in the reference */
}

/* _box_ind_tend topology.py:100, _x_per_ind_tend :124 ; ind = (unused, iy, ix) */
static int flat_ind_tend(const ora_grid *g, const int32_t ind[3], int tend, int32_t out[3]) {
    int iy = ind[1], ix = ind[2];
    int iymax = g->ny - 1, ixmax = g->nx - 1;
    if (tend == 0) iy += 1;
    else if (tend == 1) iy -= 1;
    else if (tend == 2) ix -= 1;
    else if (tend == 3) ix += 1;
    out[0] = ind[0];
    if (iy > iymax || iy < 0) {
        out[1] = -1;
        out[2] = -1;
        return ORA_OK;
    }
    if (g->typ == ORA_BOX) {
        if (ix > ixmax || ix < 0) {
            out[1] = -1;
            out[2] = -1;
            return ORA_OK;
        }
    } else {
        if (ix > ixmax) ix = ix - ixmax; /* sic: topology.py:141-142 */
        else if (ix < 0) ix = ixmax + ix + 1;
    }
    out[1] = iy;
    out[2] = ix;
    return ORA_OK;
}

/* _llc_ind_tend, topology.py:149 */
static int llc_ind_tend(const ora_grid *g, const int32_t ind[3], int tend, int32_t out[3]) {
    int face = ind[0], iy = ind[1], ix = ind[2];
    int iymax = g->ny - 1, ixmax = g->nx - 1;
    int nf, ne, err;
    if (tend == 3) {
        if (ix != ixmax) ix += 1;
        else {
            if ((err = other_edge(g, face, 3, &nf, &ne))) return err;
            int oy = iy;
            face = nf;
            if (ne == 1) { iy = 0; ix = ixmax - oy; }
            else if (ne == 0) { iy = iymax; ix = oy; }
            else if (ne == 2) { iy = oy; ix = 0; }
            else { iy = iymax - oy; ix = ixmax; }
        }
    } else if (tend == 2) {
        if (ix != 0) ix -= 1;
        else {
            if ((err = other_edge(g, face, 2, &nf, &ne))) return err;
            int oy = iy;
            face = nf;
            if (ne == 1) { iy = 0; ix = oy; }
            else if (ne == 0) { iy = iymax; ix = ixmax - oy; }
            else if (ne == 2) { iy = iymax - oy; ix = 0; }
            else { iy = oy; ix = ixmax; }
        }
    } else if (tend == 0) {
        if (iy != iymax) iy += 1;
        else {
            if ((err = other_edge(g, face, 0, &nf, &ne))) return err;
            int ox = ix;
            face = nf;
            if (ne == 1) { iy = 0; ix = ox; }
            else if (ne == 0) { iy = iymax; ix = ixmax - ox; }
            else if (ne == 2) { iy = iymax - ox; ix = 0; }
            else { iy = ox; ix = ixmax; }
        }
    } else if (tend == 1) {
        if (iy != 0) iy -= 1;
        else {
            if ((err = other_edge(g, face, 1, &nf, &ne))) return err;
            int ox = ix;
            face = nf;
            if (ne == 1) { iy = 0; ix = ixmax - ox; }
            else if (ne == 0) { iy = iymax; ix = ox; }
            else if (ne == 2) { iy = ox; ix = 0; }
            else { iy = iymax - ox; ix = ixmax; }
        }
    }
    out[0] = face;
    out[1] = iy;
    out[2] = ix;
    return ORA_OK;
}

static int has_minus_one(const ora_grid *g, const int32_t ind[3]) {
    return (g->has_face && ind[0] == -1) || ind[1] == -1 || ind[2] == -1;
}

/* Topology.check_illegal, topology.py:491 */
static int illegal(const ora_grid *g, const int32_t ind[3], char cuvwg) {
    int ny = cuvwg == 'C' ? g->ny : g->gny;
    int nx = cuvwg == 'C' ? g->nx : g->gnx;
    if (g->has_face && (ind[0] < 0 || ind[0] > g->nface - 1)) return 1;
    return ind[1] < 0 || ind[1] > ny - 1 || ind[2] < 0 || ind[2] > nx - 1;
}

static int ind_tend_U(const ora_grid *g, const int32_t ind[3], int tend, int *is_v, int32_t out[3]);
static int ind_tend_V(const ora_grid *g, const int32_t ind[3], int tend, int *is_v, int32_t out[3]);

/* Topology.ind_tend, topology.py:389 */
int ora_ind_tend(const ora_grid *g, const int32_t ind[3], int tend, char cuvwg, int32_t out[3]) {
    if (has_minus_one(g, ind)) {
        out[0] = g->has_face ? -1 : ind[0];
        out[1] = -1;
        out[2] = -1;
        return ORA_OK;
    }
    if (g->typ == ORA_LLC) {
        int is_v;
        if (cuvwg == 'C') return llc_ind_tend(g, ind, tend, out);
        if (cuvwg == 'U') return ind_tend_U(g, ind, tend, &is_v, out);
        if (cuvwg == 'V') return ind_tend_V(g, ind, tend, &is_v, out);
        if (cuvwg == 'G') { /* _ind_tend_G, topology.py:649 */
            if (tend <= 1) return ind_tend_U(g, ind, tend, &is_v, out);
            return ind_tend_V(g, ind, tend, &is_v, out);
        }
        return ORA_VALUE_ERROR;
    }
    return flat_ind_tend(g, ind, tend, out);
}

/* Topology.ind_moves, topology.py:436 */
int ora_ind_moves(const ora_grid *g, const int32_t ind_in[3], const int *moves_in, int nmoves,
                  char cuvwg, int32_t out[3]) {
    int32_t ind[3] = {ind_in[0], ind_in[1], ind_in[2]};
    if (illegal(g, ind, 'C')) {
        out[0] = g->has_face ? -1 : ind[0];
        out[1] = -1;
        out[2] = -1;
        return ORA_OK;
    }
    int moves[16];
    if (nmoves > 16) return ORA_VALUE_ERROR;
    for (int k = 0; k < nmoves; ++k) moves[k] = moves_in[k];
    if (g->typ == ORA_LLC) {
        int face = ind[0];
        for (int k = 0; k < nmoves; ++k) {
            int32_t nxt[3];
            int err = ora_ind_tend(g, ind, moves[k], cuvwg, nxt);
            if (err) return err;
            memcpy(ind, nxt, sizeof nxt);
            if (ind[0] != face) {
                int edge, new_edge;
                if ((err = mutual_direction(g, face, ind[0], 1, &edge, &new_edge))) return err;
                double rot = fmod(PI - DIRECTIONS[edge] + DIRECTIONS[new_edge], 2 * PI);
                if (rot < 0) rot += 2 * PI;
                static const int r90[4] = {2, 3, 1, 0};
                static const int r270[4] = {3, 2, 0, 1};
                if (fabs(rot - PI / 2) < 1e-6)
                    for (int j = k + 1; j < nmoves; ++j) moves[j] = r90[moves[j]];
                else if (fabs(rot - 3 * PI / 2) < 1e-6)
                    for (int j = k + 1; j < nmoves; ++j) moves[j] = r270[moves[j]];
                face = ind[0];
            }
        }
    } else {
        for (int k = 0; k < nmoves; ++k) {
            int32_t nxt[3];
            ora_ind_tend(g, ind, moves[k], 'C', nxt);
            memcpy(ind, nxt, sizeof nxt);
        }
    }
    memcpy(out, ind, sizeof ind);
    return ORA_OK;
}

static int same_ind(const int32_t a[3], const int32_t b[3]) {
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2];
}

/* Topology._find_wall_between, topology.py:569 */
static int find_wall_between(const ora_grid *g, const int32_t ind1[3], const int32_t ind2[3],
                             int *is_v, int32_t out[3]) {
    if (ind1[0] == ind2[0]) {
        int32_t ret[3];
        for (int k = 0; k < 3; ++k) ret[k] = (int32_t)ceil((ind1[k] + ind2[k]) / 2.0);
        const int32_t *other = same_ind(ind2, ret) ? ind1 : ind2;
        memcpy(out, ret, sizeof ret);
        if (ret[2] > other[2]) { *is_v = 0; return ORA_OK; }
        if (ret[1] > other[1]) { *is_v = 1; return ORA_OK; }
        return ORA_INDEX_ERROR;
    }
    int d1to2, d2to1;
    int err = mutual_direction(g, ind1[0], ind2[0], 0, &d1to2, &d2to1);
    if (err) return err;
    if (d1to2 == 0 || d1to2 == 3) {
        memcpy(out, ind2, 3 * sizeof(int32_t));
        if (d2to1 == 1) { *is_v = 1; return ORA_OK; }
        if (d2to1 == 2) { *is_v = 0; return ORA_OK; }
        return ORA_VALUE_ERROR;
    }
    if (d2to1 == 0 || d2to1 == 3) {
        memcpy(out, ind1, 3 * sizeof(int32_t));
        if (d1to2 == 1) { *is_v = 1; return ORA_OK; }
        if (d1to2 == 2) { *is_v = 0; return ORA_OK; }
        return ORA_VALUE_ERROR;
    }
    return ORA_VALUE_ERROR;
}

/* Topology._ind_tend_U, topology.py:618 */
static int ind_tend_U(const ora_grid *g, const int32_t ind[3], int tend, int *is_v, int32_t out[3]) {
    int32_t first[3], alter[3];
    int err = ora_ind_tend(g, ind, tend, 'C', first);
    if (err) return err;
    if (first[0] == ind[0]) {
        *is_v = 0;
        memcpy(out, first, sizeof first);
        return ORA_OK;
    }
    int moves[2] = {2, tend};
    if ((err = ora_ind_moves(g, ind, moves, 2, 'C', alter))) return err;
    if (same_ind(alter, first)) {
        *is_v = 0;
        memcpy(out, alter, sizeof alter);
        return ORA_OK;
    }
    return find_wall_between(g, first, alter, is_v, out);
}

/* Topology._ind_tend_V, topology.py:635 */
static int ind_tend_V(const ora_grid *g, const int32_t ind[3], int tend, int *is_v, int32_t out[3]) {
    int32_t first[3], alter[3];
    int err = ora_ind_tend(g, ind, tend, 'C', first);
    if (err) return err;
    if (first[0] == ind[0]) {
        *is_v = 1;
        memcpy(out, first, sizeof first);
        return ORA_OK;
    }
    int moves[2] = {1, tend};
    if ((err = ora_ind_moves(g, ind, moves, 2, 'C', alter))) return err;
    return find_wall_between(g, first, alter, is_v, out);
}

/* one element of Topology.ind_tend_vec, topology.py:526 */
static int tend_vec1(const ora_grid *g, const int32_t ind[3], int tend, char cuvwg, int32_t out[3]) {
    static const int dy[4] = {1, -1, 0, 0}, dx[4] = {0, 0, -1, 1};
    int32_t naive[3] = {ind[0], ind[1] + dy[tend], ind[2] + dx[tend]};
    if (!illegal(g, naive, cuvwg)) {
        memcpy(out, naive, sizeof naive);
        return ORA_OK;
    }
    int err = ora_ind_tend(g, ind, tend, cuvwg, out);
    if (err == ORA_INDEX_ERROR) { /* "Some points are on the edge": keep the old index */
        memcpy(out, ind, 3 * sizeof(int32_t));
        return ORA_OK;
    }
    return err;
}

/* find_px_py, utils.py:495: corners LL, LR, UR, UL */
int ora_corners(const ora_grid *g, const int32_t ind[3], int32_t out[4][3]) {
    int err;
    memcpy(out[0], ind, 3 * sizeof(int32_t));
    if ((err = tend_vec1(g, ind, 3, 'G', out[1]))) return err;
    if ((err = tend_vec1(g, out[1], 0, 'G', out[2]))) return err;
    if ((err = tend_vec1(g, ind, 0, 'G', out[3]))) return err;
    return ORA_OK;
}

/* one node of Position._fatten_h, eulerian.py:386 (moves from _translate_to_tendency,
 * kernel_weight.py:46) */
static int fatten_node(const ora_grid *g, const int32_t ind[3], int ddx, int ddy, char cuvwg,
                       int32_t out[3]) {
    int32_t naive[3] = {ind[0], ind[1] + ddy, ind[2] + ddx};
    if (!illegal(g, naive, cuvwg)) {
        memcpy(out, naive, sizeof naive);
        return ORA_OK;
    }
    int moves[16], n = 0;
    if (ddy > 0) for (int j = 0; j < ddy; ++j) moves[n++] = 0;
    else for (int j = 0; j < -ddy; ++j) moves[n++] = 1;
    if (ddx < 0) for (int j = 0; j < -ddx; ++j) moves[n++] = 2;
    else for (int j = 0; j < ddx; ++j) moves[n++] = 3;
    return ora_ind_moves(g, ind, moves, n, cuvwg, out);
}

/* ------------------------------------------------------------------------------------------ */
/* utils.py numerics                                                                            */
/* ------------------------------------------------------------------------------------------ */
static double py_mod(double a, double b) { /* numpy / python float % */
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

static double py_floordiv(double a, double b) { /* numpy floor_divide for floats */
    double m = fmod(a, b);
    double div = (a - m) / b;
    if (m != 0.0 && ((b < 0) != (m < 0))) div -= 1.0;
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
    } else {
        fl = copysign(0.0, a / b);
    }
    return fl;
}

/* to_180, utils.py:199 */
static double to_180(double x) {
    x = py_mod(py_mod(x, 360.0), 360.0);
    return x + (-1) * py_floordiv(x, 360.0 / 2) * 360.0;
}

static double nan_to_num(double x) {
    if (isnan(x)) return 0.0;
    if (isinf(x)) return x > 0 ? DBL_MAX : -DBL_MAX;
    return x;
}

static double sign(double x) { return (x > 0) - (x < 0); }

/* _time2wall utils.py:859 + _stationary_time :821 + _uleftright_from_udu :852 + _which_early :875
 * + _move_within_cell lagrangian.py:528 (+ _increment utils.py:777) */
void ora_analytic(const double x0[3], const double u[3], const double du[3], double tf, int *tend,
                  double *t_event, double newx[3], double newu[3]) {
    double ts[7];
    double sg = sign(tf);
    for (int i = 0; i < 3; ++i) {
        double tl = log(1 - du[i] / u[i] * (0.5 + x0[i])) / du[i];
        double tr = log(1 + du[i] / u[i] * (0.5 - x0[i])) / du[i];
        if (fabs(du[i]) < 1e-12) {
            tl = (-x0[i] - 0.5) / u[i];
            tr = (0.5 - x0[i]) / u[i];
        }
        double ul = u[i] - (x0[i] + 0.5) * du[i];
        double ur = u[i] + (0.5 - x0[i]) * du[i];
        if (-ul * sg <= 1e-12) tl = -sg;
        if (ur * sg <= 1e-12) tr = -sg;
        ts[2 * i] = tl;
        ts[2 * i + 1] = tr;
    }
    ts[6] = tf;
    int best = 0;
    double bestv = INFINITY;
    for (int k = 0; k < 7; ++k) {
        double d = ts[k] * sg;
        if (isnan(d) || d < 0) d = INFINITY;
        if (k == 0 || d < bestv) { /* first minimum wins (np.argmin) */
            if (k == 0) { best = 0; bestv = d; }
            else { best = k; bestv = d; }
        }
    }
    *tend = best;
    *t_event = ts[best];
    double t = ts[best];
    for (int i = 0; i < 3; ++i) {
        double incr = u[i] / du[i] * (exp(du[i] * t) - 1);
        if (fabs(du[i]) < 1e-12) incr = u[i] * t;
        newu[i] = u[i] + du[i] * incr;
        newx[i] = incr + x0[i];
    }
}

/* find_rx_ry_oceanparcel, utils.py:532 */
static void find_rx_ry_oceanparcel(double x, double y, const double px_in[4], const double py[4],
                                   double *rx_out, double *ry_out) {
    double x0 = px_in[0];
    double px[4];
    x = to_180(x - x0);
    for (int j = 0; j < 4; ++j) px[j] = to_180(px_in[j] - x0);
    /* a = invA . px (np.dot -> BLAS, accumulation over k in order) */
    double a[4], b[4];
    a[0] = px[0];
    a[1] = -px[0] + px[1];
    a[2] = -px[0] + px[3];
    a[3] = ((px[0] - px[1]) + px[2]) - px[3];
    b[0] = py[0];
    b[1] = -py[0] + py[1];
    b[2] = -py[0] + py[3];
    b[3] = ((py[0] - py[1]) + py[2]) - py[3];
    double aa = a[3] * b[2] - a[2] * b[3];
    double bb = a[3] * b[0] - a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + x * b[3] - y * a[3];
    double cc = a[1] * b[0] - a[0] * b[1] + x * b[1] - y * a[1];
    double det2 = bb * bb - 4 * aa * cc;
    int order1 = fabs(aa) < 1e-12;
    int order2 = !order1 && det2 >= 0;
    double ry = -(cc / bb);
    if (order2) ry = (-bb + sqrt(det2)) / (2 * aa);
    int rot_rectilinear = fabs(a[1] + a[3] * ry) < 1e-12;
    double rx;
    if (rot_rectilinear)
        rx = ((y - py[0]) / (py[1] - py[0]) + (y - py[3]) / (py[2] - py[3])) * 0.5;
    else
        rx = (x - a[0] - a[2] * ry) / (a[1] + a[3] * ry);
    *rx_out = rx - 1.0 / 2;
    *ry_out = ry - 1.0 / 2;
}

/* ------------------------------------------------------------------------------------------ */
/* field access                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static double rd(const void *a, int dtype, int64_t i) {
    return dtype == 0 ? ((const double *)a)[i] : (double)((const float *)a)[i];
}

static int wrap(int i, int n) { return i < 0 ? i + n : i; } /* numpy negative indexing */

static double read_uv(const ora_grid *g, const ora_vel *vel, int which, int it, int iz, int f,
                      int iy, int ix) {
    int ny = which == 0 ? vel->nyU : vel->nyV;
    int nx = which == 0 ? vel->nxU : vel->nxV;
    int64_t idx = 0;
    if (vel->has_time) idx = wrap(it - vel->it0, vel->nt);
    if (vel->has_z) idx = idx * vel->nzU + wrap(iz, vel->nzU);
    if (g->has_face) idx = idx * g->nface + wrap(f, g->nface);
    idx = (idx * ny + wrap(iy, ny)) * nx + wrap(ix, nx);
    return nan_to_num(rd(which == 0 ? vel->U : vel->V, vel->dtype, idx));
}

static double read_w(const ora_grid *g, const ora_vel *vel, int it, int izl, int f, int iy, int ix) {
    int64_t idx = 0;
    if (vel->has_time) idx = wrap(it - vel->it0, vel->nt);
    idx = idx * vel->nzW + wrap(izl, vel->nzW);
    if (g->has_face) idx = idx * g->nface + wrap(f, g->nface);
    idx = (idx * g->ny + wrap(iy, g->ny)) * g->nx + wrap(ix, g->nx);
    return nan_to_num(rd(vel->W, vel->dtype, idx));
}

/* rotation coefficients of a kernel node on another face: _llc_get_uv_mask_from_face,
 * topology.py:222, rounded as in eulerian.py:934 */
static int uv_rotation(const ora_grid *g, int f0, int f1, double m[4]) {
    m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1; /* UfromU UfromV VfromU VfromV */
    if (f0 == f1) return ORA_OK;
    int edge, new_edge;
    int err = mutual_direction(g, f0, f1, 1, &edge, &new_edge);
    if (err) return err;
    double rot = PI - DIRECTIONS[edge] + DIRECTIONS[new_edge];
    m[0] = nearbyint(cos(rot));
    m[1] = nearbyint(sin(rot));
    m[2] = nearbyint(-sin(rot));
    m[3] = nearbyint(cos(rot));
    return ORA_OK;
}

/* Particle.get_vol lagrangian.py:202 */
static double get_vol(const ora_grid *g, const ora_vel *vel, int izl, int f, int iy, int ix) {
    int64_t idx = 0;
    if (vel->vol_has_z) idx = wrap(izl - 1, g->n_zl > 0 ? g->n_z : 1);
    if (g->has_face) idx = idx * g->nface + f;
    idx = (idx * g->ny + wrap(iy, g->ny)) * g->nx + wrap(ix, g->nx);
    return rd(vel->Vol, vel->vol_dtype, idx);
}

/* Particle.get_u_du lagrangian.py:218 -> Position.interpolate with the module level kernels
 * (lagrangian.py:25-53): u from the left/right wall, du their difference; same for v; w from
 * W[izl_lin] and W[izl_lin-1]. */
void ora_get_u_du(const ora_grid *g, const ora_vel *vel, ora_particles *p, int64_t i, int has_dep) {
    int f = g->has_face ? p->face[i] : 0, iy = p->iy[i], ix = p->ix[i];
    int it = p->it ? p->it[i] : 0;
    int izl = has_dep ? p->izl[i] : 0;
    int iz = izl - 1;
    double rx = p->rx[i], ry = p->ry[i];
    int32_t ind[3] = {f, iy, ix};
    double ul, ur, vl, vr;
    int bad = 0;
    if (g->has_face) {
        /* vector kernels, cuvwg 'U' / 'V' moves and face rotation */
        int32_t n1[3];
        double m[4];
        ul = read_uv(g, vel, 0, it, iz, f, iy, ix) * 1.0 + read_uv(g, vel, 0, it, iz, f, iy, ix) * 0.0;
        if (fatten_node(g, ind, 1, 0, 'U', n1) || uv_rotation(g, f, n1[0], m)) {
            bad = 1;
            ur = 0;
        } else {
            double val = NAN;
            if (fabs(m[0]) != 0) val = read_uv(g, vel, 0, it, iz, n1[0], n1[1], n1[2]);
            if (fabs(m[1]) != 0) val = read_uv(g, vel, 1, it, iz, n1[0], n1[1], n1[2]);
            ur = val * m[0] + val * m[1];
        }
        vl = read_uv(g, vel, 1, it, iz, f, iy, ix) * 0.0 + read_uv(g, vel, 1, it, iz, f, iy, ix) * 1.0;
        if (fatten_node(g, ind, 0, 1, 'V', n1) || uv_rotation(g, f, n1[0], m)) {
            bad = 1;
            vr = 0;
        } else {
            double val = NAN;
            if (fabs(m[2]) != 0) val = read_uv(g, vel, 0, it, iz, n1[0], n1[1], n1[2]);
            if (fabs(m[3]) != 0) val = read_uv(g, vel, 1, it, iz, n1[0], n1[1], n1[2]);
            vr = val * m[2] + val * m[3];
        }
    } else {
        /* scalar kernels: cuvwg = 'G' when the variable lives on Xp1/Yp1, else 'C' */
        int32_t n1[3];
        ul = read_uv(g, vel, 0, it, iz, 0, iy, ix);
        fatten_node(g, ind, 1, 0, vel->u_stag ? 'G' : 'C', n1);
        ur = read_uv(g, vel, 0, it, iz, 0, n1[1], n1[2]);
        vl = read_uv(g, vel, 1, it, iz, 0, iy, ix);
        fatten_node(g, ind, 0, 1, vel->v_stag ? 'G' : 'C', n1);
        vr = read_uv(g, vel, 1, it, iz, 0, n1[1], n1[2]);
    }
    /* weights: kernel_weight_x "x"/"y" functions, kernel_weight.py:211/250, on rx + 1/2 when
     * staggered (eulerian.py:1229-1236, 1259-1264) */
    double rxs = (g->has_face || vel->u_stag) ? rx + 0.5 : rx;
    double rys = (g->has_face || vel->v_stag) ? ry + 0.5 : ry;
    double wx0 = (rxs - 1.0) / (0.0 - 1.0), wx1 = (rxs - 0.0) / (1.0 - 0.0);
    double wy0 = (rys - 1.0) / (0.0 - 1.0), wy1 = (rys - 0.0) / (1.0 - 0.0);
    double u = ul * wx0 + ur * wx1;
    double du = ul * (1.0 / (0.0 - 1.0)) + ur * (1.0 / (1.0 - 0.0));
    double v = vl * wy0 + vr * wy1;
    double dv = vl * (1.0 / (0.0 - 1.0)) + vr * (1.0 / (1.0 - 0.0));
    double w = 0, dw = 0;
    if (vel->has_w) {
        double rz = p->rzl[i];
        double w0 = read_w(g, vel, it, izl, f, iy, ix), w1 = read_w(g, vel, it, izl - 1, f, iy, ix);
        w = w0 * (1 - rz) + w1 * rz;
        dw = w0 * -1.0 + w1 * 1.0;
    }
    if (!vel->transport) {
        double dx = p->dx[i], dy = p->dy[i];
        p->u[i] = u / dx;
        p->v[i] = v / dy;
        p->du[i] = du / dx;
        p->dv[i] = dv / dy;
    } else {
        double vol = p->vol[i];
        p->u[i] = u / vol;
        p->v[i] = v / vol;
        p->du[i] = du / vol;
        p->dv[i] = dv / vol;
    }
    if (vel->has_w) {
        double s = vel->transport ? p->vol[i] : p->dzl[i];
        p->w[i] = w / s;
        p->dw[i] = dw / s;
    }
    p->u[i] = nan_to_num(p->u[i]);
    p->v[i] = nan_to_num(p->v[i]);
    p->w[i] = nan_to_num(p->w[i]);
    p->du[i] = nan_to_num(p->du[i]);
    p->dv[i] = nan_to_num(p->dv[i]);
    p->dw[i] = nan_to_num(p->dw[i]);
    if (bad && p->status) p->status[i] |= 2;
}

/* ------------------------------------------------------------------------------------------ */
/* lagrangian.py                                                                                */
/* ------------------------------------------------------------------------------------------ */
static void note(ora_log *log, const ora_particles *p, int64_t i, int stamp, int has_dep,
                 int64_t *cursor) {
    if (!log) return;
    if (log->offset) {
        int64_t r = log->offset[i] + *cursor;
        if (r < log->capacity) {
            log->fc[r] = p->face ? p->face[i] : 0;
            log->it[r] = p->it ? p->it[i] : 0;
            log->izl[r] = has_dep ? p->izl[i] : 0;
            log->rzl[r] = has_dep ? p->rzl[i] : 0.0;
            log->iy[r] = p->iy[i];
            log->ix[r] = p->ix[i];
            log->rx[r] = p->rx[i];
            log->ry[r] = p->ry[i];
            log->tt[r] = p->t[i];
            log->uu[r] = p->u[i];
            log->vv[r] = p->v[i];
            log->ww[r] = p->w[i];
            log->du[r] = p->du[i];
            log->dv[r] = p->dv[i];
            log->dw[r] = p->dw[i];
            log->xx[r] = p->lon[i];
            log->yy[r] = p->lat[i];
            log->zz[r] = p->dep ? p->dep[i] : 0.0;
            log->vs[r] = stamp;
        }
    }
    *cursor += 1;
}

/* Particle.trim lagrangian.py:400 */
static void trim1(double *x, double *u, double du, double lo, double hi) {
    if (*x >= hi) {
        double d = hi - *x;
        *x += d;
        *u += du * d;
    }
    if (*x <= lo) {
        double d = lo - *x;
        *x += d;
        *u += du * d;
    }
}

/* Position.get_px_py eulerian.py:592 */
static int read_corners(const ora_grid *g, ora_particles *p, int64_t i) {
    int32_t ind[3] = {g->has_face ? p->face[i] : 0, p->iy[i], p->ix[i]};
    int32_t c[4][3];
    int err = ora_corners(g, ind, c);
    if (err) return err;
    for (int j = 0; j < 4; ++j) {
        int64_t idx = ((int64_t)(g->has_face ? wrap(c[j][0], g->nface) : 0) * g->gny +
                       wrap(c[j][1], g->gny)) * g->gnx + wrap(c[j][2], g->gnx);
        double x = (double)g->XG[idx];
        p->px[j * p->n + i] = p->lon[i] + to_180(x - p->lon[i]);
        p->py[j * p->n + i] = (double)g->YG[idx];
    }
    return ORA_OK;
}

static void advect_one(const ora_grid *g, const ora_vel *vel, ora_particles *p, int64_t i,
                       double t_stop, const ora_params *par, ora_log *log) {
    const int has_dep = par->has_dep;
    const int64_t n = p->n;
    int64_t cursor = 0;
    int niter = 0, ncross = 0;
    if (log && log->emit_start) note(log, p, i, 15, has_dep, &cursor);
    double tf = t_stop - p->t[i];
    int live = fabs(tf) > par->tol;
    while (live && niter < par->max_iteration) {
        /* trim, lagrangian.py:400 */
        double lo = -0.5 + par->trim_tol, hi = 0.5 - par->trim_tol;
        trim1(&p->rx[i], &p->u[i], p->du[i], lo, hi);
        trim1(&p->ry[i], &p->v[i], p->dv[i], lo, hi);
        if (has_dep) trim1(&p->rzl[i], &p->w[i], p->dw[i], -0.0 + par->trim_tol, 1.0 - par->trim_tol);
        /* analytical_step, lagrangian.py:561 */
        double x0[3] = {p->rx[i], p->ry[i], has_dep ? p->rzl[i] - 1.0 / 2 : 0.0};
        double uu[3] = {p->u[i], p->v[i], p->w[i]};
        double dd[3] = {p->du[i], p->dv[i], p->dw[i]};
        double nx[3], nu[3], t_event;
        int tend;
        ora_analytic(x0, uu, dd, tf, &tend, &t_event, nx, nu);
        p->t[i] += t_event;
        p->rx[i] = nx[0];
        p->ry[i] = nx[1];
        if (has_dep) p->rzl[i] = nx[2] + 1.0 / 2;
        p->u[i] = nu[0];
        p->v[i] = nu[1];
        p->w[i] = nu[2];
        /* _sync_latlondep_before_cross, lagrangian.py:539 */
        if (has_dep) p->dep[i] = p->bzl[i] + p->dzl[i] * p->rzl[i];
        if (g->hmode == ORA_H_OCEANPARCEL) {
            double rx = p->rx[i], ry = p->ry[i];
            double w4[4] = {(0.5 - rx) * (0.5 - ry), (0.5 + rx) * (0.5 - ry), (0.5 + rx) * (0.5 + ry),
                            (0.5 - rx) * (0.5 + ry)};
            double lon = 0, lat = 0;
            for (int j = 0; j < 4; ++j) {
                lon += w4[j] * p->px[j * n + i];
                lat += w4[j] * p->py[j * n + i];
            }
            p->lon[i] = lon;
            p->lat[i] = lat;
        } else {
            /* rel2latlon, utils.py:214 (no px / py on the particle: lagrangian.py:549); compiled by numba in the
             * reference, so the cosine is taken in double precision */
            double temp_x = p->rx[i] * p->dx[i] / DEG2M;
            double temp_y = p->ry[i] * p->dy[i] / DEG2M;
            double dlon = (temp_x * p->cs[i] - temp_y * p->sn[i]) / cos(p->by[i] * PI / 180);
            double dlat = temp_x * p->sn[i] + temp_y * p->cs[i];
            p->lon[i] = dlon + p->bx[i];
            p->lat[i] = dlat + p->by[i];
        }
        note(log, p, i, tend, has_dep, &cursor);
        if (tend < 6) ncross += 1;
        /* _cross_cell_wall_index, lagrangian.py:598 */
        if (tend <= 3) {
            static const int translate[4] = {2, 3, 1, 0};
            int32_t ind[3] = {g->has_face ? p->face[i] : 0, p->iy[i], p->ix[i]}, out[3];
            int err = tend_vec1(g, ind, translate[tend], 'C', out);
            if (err) p->status[i] |= 2;
            else {
                if (g->has_face) p->face[i] = out[0];
                p->iy[i] = out[1];
                p->ix[i] = out[2];
            }
        } else if (tend == 4) {
            p->izl[i] += 1;
        } else if (tend == 5) {
            if (!par->kick_back) p->izl[i] -= 1;
            else if (p->izl[i] == 1) p->dep[i] = p->bzl[i] + p->dzl[i] / 2;
            else p->izl[i] -= 1;
        }
        /* _cross_cell_wall_read, lagrangian.py:640 */
        double cos_by32 = 1.0;
        if (g->hmode == ORA_H_OCEANPARCEL) {
            int64_t hidx = ((int64_t)(g->has_face ? p->face[i] : 0) * g->ny + wrap(p->iy[i], g->ny)) *
                               g->nx + wrap(p->ix[i], g->nx);
            p->bx[i] = (double)g->XC[hidx];
            p->by[i] = (double)g->YC[hidx];
            if (read_corners(g, p, i)) p->status[i] |= 2;
        } else if (g->hmode == ORA_H_LOCAL_CARTESIAN) { /* :647-659 */
            int64_t hidx = ((int64_t)(g->has_face ? p->face[i] : 0) * g->ny + wrap(p->iy[i], g->ny)) *
                               g->nx + wrap(p->ix[i], g->nx);
            p->bx[i] = (double)g->XC[hidx];
            p->by[i] = (double)g->YC[hidx];
            p->cs[i] = (double)g->CS[hidx];
            p->sn[i] = (double)g->SN[hidx];
            p->dx[i] = (double)g->dX[hidx];
            p->dy[i] = (double)g->dY[hidx];
            cos_by32 = (double)g->coslat32[hidx]; /* per cell here: np.cos(YC * np.pi / 180) in float32 */
        } else if (g->hmode == ORA_H_RECTILINEAR) { /* :667-672 */
            int ix = wrap(p->ix[i], g->nx), iy = wrap(p->iy[i], g->ny);
            p->bx[i] = (double)g->lon[ix];
            p->by[i] = (double)g->lat[iy];
            p->cs[i] = 1.0;
            p->sn[i] = 0.0;
            cos_by32 = (double)g->coslat32[iy]; /* np.cos(float32 lat * np.pi / 180): float32 arithmetic */
            p->dx[i] = g->dlon[ix] * cos_by32;
            p->dy[i] = g->dlat[iy];
        }
        if (has_dep) {
            p->bzl[i] = (double)g->Zl[wrap(p->izl[i], g->n_zl)];
            p->dzl[i] = (double)g->dZl[wrap(p->izl[i] - 1, g->n_z)];
        }
        /* _cross_cell_wall_rel, lagrangian.py:681 */
        if (has_dep) p->rzl[i] = (p->dep[i] - p->bzl[i]) / p->dzl[i];
        if (g->hmode == ORA_H_OCEANPARCEL) {
            double px[4], py[4];
            for (int j = 0; j < 4; ++j) {
                px[j] = p->px[j * n + i];
                py[j] = p->py[j * n + i];
            }
            find_rx_ry_oceanparcel(p->lon[i], p->lat[i], px, py, &p->rx[i], &p->ry[i]);
        } else { /* :693-703; by is the float32 array just read, so is its cosine */
            double dlon = to_180(p->lon[i] - p->bx[i]);
            double dlat = to_180(p->lat[i] - p->by[i]);
            p->rx[i] = (dlon * cos_by32 * p->cs[i] + dlat * p->sn[i]) * DEG2M / p->dx[i];
            p->ry[i] = (dlat * p->cs[i] - dlon * p->sn[i] * cos_by32) * DEG2M / p->dy[i];
        }
        if (vel->transport)
            p->vol[i] = get_vol(g, vel, has_dep ? p->izl[i] : 0, g->has_face ? p->face[i] : 0,
                                p->iy[i], p->ix[i]);
        ora_get_u_du(g, vel, p, i, has_dep);
        niter += 1;
        tf = t_stop - p->t[i];
        live = fabs(tf) > par->tol;
        if (live) note(log, p, i, 7, has_dep, &cursor);
    }
    if (live && p->status) p->status[i] |= 1;
    if (p->niter) p->niter[i] = niter;
    if (p->ncross) p->ncross[i] = ncross;
    if (log && log->count) log->count[i] = (int32_t)cursor;
}

/* Particle.to_next_stop, lagrangian.py:741 -- every live particle takes one analytic step per
 * iteration and particles never interact, so the loop is done particle by particle (and, for the
 * all-cores CPU baseline, by several threads over disjoint chunks of particles). */
typedef struct {
    const ora_grid *g;
    const ora_vel *vel;
    ora_particles *p;
    double t_stop;
    const ora_params *par;
    ora_log *log;
    int64_t *next;
    pthread_mutex_t *lock;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    const int64_t chunk = 256;
    for (;;) {
        pthread_mutex_lock(j->lock);
        int64_t lo = *j->next;
        *j->next = lo + chunk;
        pthread_mutex_unlock(j->lock);
        if (lo >= j->p->n) break;
        int64_t hi = lo + chunk < j->p->n ? lo + chunk : j->p->n;
        for (int64_t i = lo; i < hi; ++i) advect_one(j->g, j->vel, j->p, i, j->t_stop, j->par, j->log);
    }
    return NULL;
}

/* Position._fatten_h, eulerian.py:386: the index of every kernel node (kx, ky) of every point --
 * naive offset, and a topology walk (ind_moves with the moves of _translate_to_tendency,
 * kernel_weight.py:46) where the naive index is illegal.  Output arrays are n*m; oerr[j*m+i] is the
 * ORA_* code of the walk (the reference raises there and aborts the whole call). */
int ora_fatten_h(const ora_grid *g, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix, int m,
                 const int32_t *kx, const int32_t *ky, char cuvwg, int32_t *of, int32_t *oy, int32_t *ox,
                 int32_t *oerr) {
    for (int64_t j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) {
            const int64_t o = j * m + i;
            int32_t ind[3] = {g->has_face ? face[j] : 0, iy[j] + ky[i], ix[j] + kx[i]};
            oerr[o] = ORA_OK;
            if (!illegal(g, ind, cuvwg == 'C' ? 'C' : 'G')) {
                of[o] = ind[0];
                oy[o] = ind[1];
                ox[o] = ind[2];
                continue;
            }
            int moves[32], nm = 0;
            const int x = kx[i], y = ky[i];
            for (int k = 0; k < (y > 0 ? y : -y) && nm < 32; ++k) moves[nm++] = y > 0 ? 0 : 1;
            for (int k = 0; k < (x > 0 ? x : -x) && nm < 32; ++k) moves[nm++] = x < 0 ? 2 : 3;
            const int32_t start[3] = {g->has_face ? face[j] : 0, iy[j], ix[j]};
            int32_t out[3] = {start[0], start[1], start[2]};
            oerr[o] = ora_ind_moves(g, start, moves, nm, cuvwg, out);
            of[o] = out[0];
            oy[o] = out[1];
            ox[o] = out[2];
        }
    }
    return ORA_OK;
}

/* ind_tend_uv, lagrangian_budget.py:163: the U wall to the right (iw = 0) or the V wall above (iw = 1)
 * of a cell, as (new iw, new index), through Topology._ind_tend_U / _ind_tend_V */
int ora_ind_tend_uv(const ora_grid *g, int iw, const int32_t ind[3], int *new_iw, int32_t out[3]) {
    int is_v = 0, err;
    if (iw == 1) err = ind_tend_V(g, ind, 0, &is_v, out);
    else if (iw == 0) err = ind_tend_U(g, ind, 3, &is_v, out);
    else return ORA_VALUE_ERROR;
    *new_iw = is_v;
    return err;
}

/* Topology.four_matrix_for_uv, topology.py:689, rounded as in eulerian.py:934-937: for an (n, m) array
 * of faces (node 0 is the reference) the four coefficients UfromU, UfromV, VfromU, VfromV. */
/* _llc_get_uv_mask_from_face topology.py:222 as it stands (cos / sin of the rotation, not rounded: what
 * lagrangian_budget.read_wall_list :406 multiplies with): m = UfromU UfromV VfromU VfromV of face f1 seen from f0 */
int ora_uv_rotation_raw(const ora_grid *g, int f0, int f1, double m[4]) {
    m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1;
    if (f0 == f1) return ORA_OK;
    int edge, new_edge;
    int err = mutual_direction(g, f0, f1, 1, &edge, &new_edge);
    if (err) return err;
    double rot = PI - DIRECTIONS[edge] + DIRECTIONS[new_edge];
    m[0] = cos(rot);
    m[1] = sin(rot);
    m[2] = -sin(rot);
    m[3] = cos(rot);
    return ORA_OK;
}

int ora_four_matrix_for_uv(const ora_grid *g, int64_t n, int m, const int32_t *faces, double *ufu, double *ufv,
                           double *vfu, double *vfv, int32_t *oerr) {
    for (int64_t j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) {
            double c[4];
            const int64_t o = j * m + i;
            oerr[o] = uv_rotation(g, faces[j * m], faces[o], c);
            ufu[o] = c[0];
            ufv[o] = c[1];
            vfu[o] = c[2];
            vfv[o] = c[3];
        }
    }
    return ORA_OK;
}

int ora_advect(const ora_grid *g, const ora_vel *vel, ora_particles *p, double t_stop,
               const ora_params *par, ora_log *log) {
    int nthreads = par->nthreads > 0 ? par->nthreads : 1;
    if (nthreads == 1) {
        for (int64_t i = 0; i < p->n; ++i) advect_one(g, vel, p, i, t_stop, par, log);
        return 0;
    }
    if (nthreads > 256) nthreads = 256;
    pthread_t tid[256];
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    int64_t next = 0;
    job_t job = {g, vel, p, t_stop, par, log, &next, &lock};
    for (int k = 0; k < nthreads; ++k) pthread_create(&tid[k], NULL, worker, &job);
    for (int k = 0; k < nthreads; ++k) pthread_join(tid[k], NULL);
    return 0;
}
