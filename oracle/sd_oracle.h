/* TEST INFRASTRUCTURE -- CPU oracle for the seaduck Lagrangian hot path.
 *
 * A plain C restatement of the reference's numpy algorithm (MaceKuailv/seaduck, files cited per
 * function in sd_oracle.c).  It exists to check the CUDA path and to time a CPU baseline; the
 * product never links, imports or calls it.  Parity status: PINNED -- tests/test_oracle_golden.py
 * checks it against golden vectors produced by running the unmodified reference
 * (oracle/make_golden.py) and against the reference's own known-answer tests.
 */
#ifndef SD_ORACLE_H
#define SD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORA_BOX = 0, ORA_XPERIODIC = 1, ORA_LLC = 2 };
enum { ORA_H_OCEANPARCEL = 0, ORA_H_LOCAL_CARTESIAN = 1, ORA_H_RECTILINEAR = 2 };
enum { ORA_OK = 0, ORA_INDEX_ERROR = 1, ORA_VALUE_ERROR = 2 };

typedef struct {
    int32_t typ;
    int32_t has_face;
    int32_t nface; /* 1 when there is no face dimension */
    int32_t ny, nx;   /* h_shape */
    int32_t gny, gnx; /* g_shape */
    int32_t hmode;
    int32_t n_zl; /* len(Zl) (= nz + 1 when built from Zp1), 0 if no vertical */
    int32_t n_z;
    const int32_t *face_connect; /* nface x 4, 999 = none */
    const float *XC, *YC, *XG, *YG, *dX, *dY, *CS, *SN;
    const float *Zl, *dZl, *Z, *dZ;
    const float *lon, *lat; /* rectilinear */
    const double *dlon, *dlat;
    /* np.cos(lat * np.pi / 180) of the float32 latitudes as the host's numpy evaluates it in float32: per row of
     * `lat` (rectilinear) or per cell of `YC` (local_cartesian).  lagrangian.py:671, :693-703 take the cosine of a
     * float32 latitude; its last bit is numpy's, so it is data */
    const float *coslat32;
} ora_grid;

typedef struct {
    const void *U, *V, *W;
    int32_t dtype;    /* 0 = float64, 1 = float32 */
    int32_t has_time; /* leading time axis */
    int32_t it0;      /* prefetch prefix (itmin) */
    int32_t nt;
    int32_t has_z;  /* U, V have a Z axis */
    int32_t nzU;    /* levels of U,V */
    int32_t nzW;    /* levels of W (nz + 1 with the zero bottom buffer) */
    int32_t nyU, nxU, nyV, nxV;
    int32_t u_stag; /* 1 if U is on Xp1 (rx shifted by 1/2) */
    int32_t v_stag;
    int32_t has_w;
    int32_t transport;
    const void *Vol; /* [nz][nface][ny][nx] or [nface][ny][nx] */
    int32_t vol_dtype;
    int32_t vol_has_z;
} ora_vel;

typedef struct {
    int64_t n;
    double *lon, *lat, *dep, *t;
    int32_t *face, *iy, *ix, *izl, *it;
    double *rx, *ry, *rzl;
    double *u, *v, *w, *du, *dv, *dw;
    double *px, *py; /* [4][n] */
    double *dx, *dy; /* float32 values widened */
    double *bzl, *dzl;
    double *vol;
    double *bx, *by, *cs, *sn;
    int32_t *niter;  /* out: iterations done for this particle in this call */
    int32_t *ncross; /* out: iterations with tend < 6 */
    int32_t *status; /* out: bit 0 max_iteration reached, bit 1 topology error */
} ora_particles;

typedef struct {
    double tol;      /* 1e-4 */
    double trim_tol; /* 1e-14 */
    int32_t max_iteration;
    int32_t kick_back; /* free_surface == "kick_back" */
    int32_t has_dep;   /* rzl_lin is not None */
    int32_t nthreads;
} ora_params;

/* crossing log (reference Particle.note_taking, lagrangian.py:301) */
typedef struct {
    int64_t capacity;       /* records available */
    const int64_t *offset;  /* per particle start, or NULL -> count only */
    int32_t *count;         /* per particle records written */
    int32_t *fc, *it, *izl, *iy, *ix, *vs;
    double *rzl, *rx, *ry, *tt, *uu, *vv, *ww, *du, *dv, *dw, *xx, *yy, *zz;
    int32_t emit_start; /* write the stamp-15 record first */
} ora_log;

int ora_advect(const ora_grid *g, const ora_vel *vel, ora_particles *p, double t_stop,
               const ora_params *par, ora_log *log);

/* topology restatement, exposed for tests */
int ora_ind_tend(const ora_grid *g, const int32_t ind[3], int tend, char cuvwg, int32_t out[3]);
int ora_ind_moves(const ora_grid *g, const int32_t ind[3], const int *moves, int nmoves, char cuvwg,
                  int32_t out[3]);
int ora_fatten_h(const ora_grid *g, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix, int m,
                 const int32_t *kx, const int32_t *ky, char cuvwg, int32_t *of, int32_t *oy, int32_t *ox,
                 int32_t *oerr);
int ora_ind_tend_uv(const ora_grid *g, int iw, const int32_t ind[3], int *new_iw, int32_t out[3]);
int ora_four_matrix_for_uv(const ora_grid *g, int64_t n, int m, const int32_t *faces, double *ufu, double *ufv,
                           double *vfu, double *vfv, int32_t *oerr);
int ora_uv_rotation_raw(const ora_grid *g, int f0, int f1, double m[4]);
int ora_corners(const ora_grid *g, const int32_t ind[3], int32_t out[4][3]);
void ora_get_u_du(const ora_grid *g, const ora_vel *vel, ora_particles *p, int64_t i, int has_dep);
void ora_analytic(const double x0[3], const double u[3], const double du[3], double tf, int *tend,
                  double *t_event, double newx[3], double newu[3]);

#ifdef __cplusplus
}
#endif
#endif
