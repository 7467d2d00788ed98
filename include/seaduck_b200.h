/* seaduck_b200 -- C ABI of the B200 (sm_100a) library behind seaduck's Lagrangian / interpolation
 * hot path.  Plain pointers and sizes only; no torch types.
 *
 * The reference (MaceKuailv/seaduck) has no FFI: its seams are Python methods.
 * Each entry point below names the reference method whose body it replaces (file:line into the
 * reference tree).  INTEGRATION.md shows the ctypes stub a seaduck maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SDK_E* code on failure; sdk_last_error()
 *    returns a thread-local message.
 *  - pointers documented "device" must be device memory on the current CUDA device; the library
 *    borrows them for the duration of the call (grids: for the lifetime of the handle).
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device-pointer entry
 *    points are asynchronous with respect to the host; *_host entry points synchronise.
 */
#ifndef SEADUCK_B200_H
#define SEADUCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDK_ABI_VERSION 2

enum { SDK_OK = 0, SDK_EINVAL = -1, SDK_ECUDA = -2, SDK_EUNSUPPORTED = -3, SDK_ENOMEM = -4 };

/* grid kinds: reference Topology.typ, seaduck/topology.py:303-338 */
enum { SDK_TYP_BOX = 0, SDK_TYP_XPERIODIC = 1, SDK_TYP_LLC = 2 };
/* horizontal scheme: reference OceData.readiness["h"], seaduck/ocedata.py:231 */
enum { SDK_H_OCEANPARCEL = 0, SDK_H_LOCAL_CARTESIAN = 1, SDK_H_RECTILINEAR = 2 };
/* storage type of velocity / volume fields (math is always fp64) */
enum { SDK_F64 = 0, SDK_F32 = 1 };

/* per-particle status bits written by sdk_advect */
enum {
    SDK_ST_MAX_ITER = 1,   /* still live after max_iteration steps (reference warns, lagrangian.py:794) */
    SDK_ST_EDGE = 2,       /* tried to leave through an edge without neighbour ("Some points are on the edge") */
    SDK_ST_BAD_CELL = 4,   /* cell whose walls/corners the reference cannot index (it raises there) */
    SDK_ST_OUT = 8,        /* left a closed (box) domain; particle frozen */
    /* sdk_locate: value outside the table, the reference raises "Value out of bound." (utils.py:301) */
    SDK_ST_OOB = 16,       /* depth below the last Zl level (linear Zl lookup) */
    SDK_ST_OOB_Z = 32,     /* depth below the last Z level (linear Z lookup, lazily used by interpolation) */
    SDK_ST_OOB_T = 64,     /* sdk_locate: time before the first level (linear time lookup, lazily used);
                              sdk_advect: the particle's time level is outside the staged velocity slab */
    SDK_ST_OOB_LEVEL = 128 /* sdk_advect: the particle left through the deepest staged level (non-zero vertical
                              velocity at the bottom); the reference raises IndexError at Zl[izl_lin]
                              (seaduck/lagrangian.py:675).  The particle is frozen on that wall. */
};

#define SDK_MAX_FACES 16

typedef struct sdk_grid sdk_grid; /* opaque */

/* Replaces the in-memory grid of OceData._grid2array (seaduck/ocedata.py:411) and the Topology
 * tables (seaduck/topology.py:8-30).  All array pointers are DEVICE pointers except nbr_face /
 * nbr_edge which are HOST pointers copied into the handle (they end up in constant memory). */
typedef struct sdk_grid_desc {
    int32_t typ;       /* SDK_TYP_* */
    int32_t has_face;  /* 1 if the grid has a face dimension */
    int32_t nface;     /* 1 when has_face == 0 */
    int32_t ny, nx;    /* shape of XC (per face) */
    int32_t gny, gnx;  /* shape of XG (per face) */
    int32_t hmode;     /* SDK_H_* */
    int32_t n_zl;      /* len(Zl) (nz + 1), 0 = no vertical staggered grid */
    int32_t n_z;       /* len(dZl) == number of cells in the vertical, 0 = none */
    int32_t n_zc;      /* len(Z), 0 = none */
    int32_t reserved;
    const int32_t *nbr_face; /* host [nface*4]: face across edge (up,down,left,right), -1 = none */
    const int32_t *nbr_edge; /* host [nface*4]: which edge of that face touches us */
    const float *XC, *YC;    /* device [nface*ny*nx] */
    const float *XG, *YG;    /* device [nface*gny*gnx] */
    const float *dX, *dY;    /* device [nface*ny*nx] or NULL */
    const float *CS, *SN;    /* device or NULL */
    const float *Zl, *dZl;   /* device [n_zl], [n_z] or NULL */
    const float *Z, *dZ;     /* device [n_zc] or NULL */
    /* device [nface*(nx+ny)*3] flat XG indices of the LR, UR, UL corners of the cells on the top
     * row (slot ix) and right column (slot nx+iy) of each face, -1 where the reference raises;
     * required when gny == ny or gnx == nx (reference find_px_py, seaduck/utils.py:495). */
    const int32_t *corner_tab;
    /* rectilinear scheme (OceData.lon / lat / dlon / dlat, seaduck/ocedata.py:255-259, 368-370): the 1-D axes
     * lon [nx], lat [ny] (float32, device) and their metric spacing dlon [nx], dlat [ny] (float64 values of
     * np.gradient(axis) * 6371e3 * pi / 180, device).  XC / YC / XG / YG are not needed for this scheme. */
    const float *lon, *lat;
    const double *dlon, *dlat;
    /* local_cartesian and rectilinear schemes: np.cos(by * np.pi / 180) as the HOST's numpy evaluates it for the
     * float32 latitude of a cell (seaduck/lagrangian.py:671, :693-703 -- float32 arithmetic and numpy's float32
     * cosine, which is not correctly rounded, so it is data, not something to recompute on the device):
     * [nface*ny*nx] from YC (local_cartesian) or [ny] from lat (rectilinear), float32, device. */
    const float *coslat32;
    int32_t dlon_is_f32; /* the dataset's lon / lat are float32: dlon[ix] * cos is a float32 product (numpy promotion) */
    int32_t reserved2;
} sdk_grid_desc;

/* Velocity (or transport) snapshot slab; replaces Particle.update_uvw_array's uarray/varray/warray
 * (seaduck/lagrangian.py:175) and OceData["Vol"] (seaduck/ocedata.py:303).  Device pointers.
 * Layout: U [nt][nzU][nface][nyU][nxU], V [nt][nzU][nface][nyV][nxV], W [nt][nzW][nface][ny][nx],
 * Vol [nz][nface][ny][nx] (axes of size 1 when absent). */
typedef struct sdk_velocity {
    const void *U, *V, *W; /* W may be NULL (wname=None) */
    const void *Vol;       /* required when transport != 0 */
    int32_t dtype;         /* SDK_F64 / SDK_F32 for U, V, W */
    int32_t vol_dtype;
    int32_t has_time, it0, nt; /* it0 = first snapshot held (reference prefetch_prefix / itmin) */
    int32_t has_z, nzU, nzW;
    int32_t nyU, nxU, nyV, nxV;
    int32_t u_stag, v_stag;    /* U lives on Xp1 / V on Yp1 */
    int32_t vol_has_z;
    int32_t transport;         /* divide by Vol instead of dx|dy|dzl (lagrangian.py:268-286) */
} sdk_velocity;

/* Particle state, structure of arrays, DEVICE pointers, n entries each (px, py: 4*n, corner-major).
 * Replaces the attributes of Particle / RelCoord that the time loop touches
 * (seaduck/lagrangian.py:56, seaduck/ocedata.py:119-143).
 * px / py hold the horizontal description of the particle's cell: in the oceanparcel scheme the longitudes /
 * latitudes of its four corners (LL, LR, UR, UL); in the local_cartesian and rectilinear schemes, which have
 * no corners, rows 0..2 of px are bx, cs, dx and rows 0..2 of py are by, sn, dy as float64 (row 3 unused) --
 * the attributes the reference refreshes in _cross_cell_wall_read (seaduck/lagrangian.py:647-673). */
typedef struct sdk_particles {
    int64_t n;
    double *lon, *lat, *dep, *t;
    int32_t *face, *iy, *ix, *izl, *it; /* izl = izl_lin */
    double *rx, *ry, *rzl;              /* rzl = rzl_lin */
    double *u, *v, *w, *du, *dv, *dw;
    double *px, *py;
    float *dx, *dy;      /* kept as state: the reference never refreshes them in oceanparcel mode */
    double *bzl, *dzl;   /* bzl_lin, dzl_lin */
    double *vol;
    int32_t *status;     /* out, SDK_ST_* bits */
    int32_t *niter;      /* out, analytic steps taken in this call */
    int32_t *ncross;     /* out, steps that ended on a wall (tend < 6) */
    const int32_t *active; /* optional: particles with active[i] == 0 are left untouched (callback) */
} sdk_particles;

struct sdk_exchange; /* declared with the record exchange further down */

typedef struct sdk_advect_params {
    double tol;            /* 1e-4  (lagrangian.py:754) */
    double trim_tol;       /* 1e-14 (lagrangian.py:765) */
    int32_t max_iteration; /* Particle(max_iteration=200) */
    int32_t kick_back;     /* free_surface == "kick_back" (lagrangian.py:629) */
    int32_t has_dep;       /* particles carry a depth (rzl_lin is not None) */
    int32_t refill_threshold; /* idle lanes per warp that trigger a refill, 0 = library default */
    int32_t blocks_per_sm;    /* 0 = library default */
    int32_t threads_per_block;/* 0 = library default */
    /* leave this many SMs without a CTA of the (persistent) kernel, so that kernels of other streams
     * -- the NCCL all-gather of the previous stop's snapshot -- run beside it; needs one CTA per SM
     * (blocks_per_sm = 1, threads_per_block = 512) to be exact */
    int32_t reserve_sms;
    /* != 0: the status array holds flags from an earlier stage (sdk_locate) that the call ORs its own into
     * instead of overwriting them */
    int32_t keep_status;
    /* optional DEVICE array of 4 counters, ADDED to by the call: particles live at the start,
     * wall crossings, particles that hit max_iteration, particles with a non-zero status */
    uint64_t *stats;
    /* Fused output exchange (optional, NULL = off): when a particle is written back at the end of its leg, the
     * kernel also stores its 32-byte output record into block [ship_step % n_slots][rank] of every rank's receive
     * buffer (position ship_index ? ship_index[i] : i) -- the advection kernel is its own sdk_pack_ship, the transfer
     * rides on the compute instead of following it.  Same protocol as sdk_pack_ship: waits for the slot to be free
     * before the first store, posts arrived[rank] = ship_step after the last.  Not with a log or an `active` mask. */
    const struct sdk_exchange *ship;
    const int32_t *ship_index;
    uint64_t ship_step;
} sdk_advect_params;

/* One record of the crossing log: the 19 fields Particle.note_taking appends (seaduck/lagrangian.py:301),
 * 128 bytes, so a lane writes four full 32-byte sectors per record. */
typedef struct sdk_log_record {
    double rzl, rx, ry, tt, uu, vv, ww, du, dv, dw, xx, yy, zz;
    int32_t fc, it, izl, iy, ix, vs; /* vs = stamp: 0..6 tend, 7 after a crossing, 15 start of a stop */
} sdk_log_record;

/* Crossing log, replaces Particle.note_taking (seaduck/lagrangian.py:301).  Device pointers.
 * Records of particle i go to records[offset[i] .. offset[i] + count[i]).  With offset == NULL only
 * count[] is written (first pass of count -> scan -> fill). */
typedef struct sdk_log {
    int64_t capacity;
    const int64_t *offset;
    int32_t *count;
    sdk_log_record *records;
    int32_t emit_start; /* write the stamp-15 record first (lagrangian.py:887) */
    /* bit 0: do NOT write the stamp-7 record after a crossing (the host decides who continues: Python
     * callback, lagrangian.py:787); bit 1: write a stamp-7 record for every active live particle when
     * it is loaded (the record deferred from the previous one-iteration launch) */
    int32_t continue_mode;
} sdk_log;

int sdk_abi_version(void);
/* sizeof() of the structures above as the library was compiled (binding self-check): 0 sdk_grid_desc,
 * 1 sdk_velocity, 2 sdk_particles, 3 sdk_advect_params, 4 sdk_log, 5 sdk_located, 6 sdk_track_io,
 * 7 sdk_kernel_desc, 8 sdk_field, 9 sdk_points, 10 sdk_budget_fields, 11 sdk_wall_fields, 12 sdk_snapshot_record,
 * 13 sdk_exchange, 14 sdk_interp_job, 15 sdk_interp_io; -1 for anything else */
int64_t sdk_struct_size(int which);
const char *sdk_last_error(void);
/* number of SMs of the current device (grid sizing; no reference counterpart), <0 on error */
int sdk_sm_count(void);

/* OceData(...) grid staging, seaduck/ocedata.py:164,411 */
int sdk_grid_create(const sdk_grid_desc *desc, sdk_grid **out);
int sdk_grid_destroy(sdk_grid *grid);

/* Particle.to_next_stop, seaduck/lagrangian.py:741 (trim :400, analytical_step :561,
 * cross_cell_wall :706, get_vol :202, get_u_du :218).  In place on the device state.
 * `commit` = 0 runs the loop without writing the state back (counting pass for the log). */
int sdk_advect(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, double t_stop,
               const sdk_advect_params *par, const sdk_log *log, int commit, void *stream);

/* Which compile-time variant of the advection kernel sdk_advect would launch for these arguments (no launch, no device
 * access; no reference counterpart -- the reference has one numpy code path): 0 = every switch read from the arguments,
 * 1 = LLC faces + depth + vertical velocity + volume transports (BASELINE.json configs[1]), 2 = the horizontal schemes
 * without corner points (local_cartesian, rectilinear), 3 = LLC surface run with 2-D transports (configs[3]).  All variants
 * give bit-identical results (tests/test_gpu_parity.py); negative = SDK_E* error. */
int sdk_advect_profile(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, const sdk_advect_params *par,
                       const sdk_log *log);

/* Particle.get_u_du alone, seaduck/lagrangian.py:218 (after a velocity-slab swap). */
int sdk_get_u_du(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, int has_dep,
                 void *stream);

/* Same as sdk_advect but every pointer in `p` is HOST memory (pinned for full speed): copies the
 * state to the device, advects, copies it back, synchronises.  `scratch` is a device buffer of at
 * least sdk_particles_bytes(n) bytes. */
int64_t sdk_particles_bytes(int64_t n);
int sdk_advect_host(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *host_p,
                    double t_stop, const sdk_advect_params *par, void *scratch, void *stream);


/* (sdk_snapshot_record is declared further down with the output-record entry points) */
struct sdk_snapshot_record;
/* Positions in, positions out, all HOST memory (pinned for full speed): what a user of
 * `Particle(x=..,y=..,z=..,t=..).to_next_stop(t_stop)` followed by reading lon/lat/dep back gets
 * (seaduck/lagrangian.py:93,741; seaduck/oceinterp.py:113-126).  Copies the inputs to the device,
 * locates them (sdk_locate), reads the velocities (sdk_get_u_du), advects (sdk_advect) and copies
 * the results back; synchronises.  Optional outputs may be NULL.  `scratch` is a device buffer of
 * sdk_track_scratch_bytes(n) bytes; `ts_dev` the device array of model time levels (or NULL). */
typedef struct sdk_track_io {
    int64_t n;
    const double *lon, *lat, *dep, *t;
    /* results, one 32-byte record per particle in the caller's order (or NULL) */
    struct sdk_snapshot_record *out_records;
    /* the same as separate arrays, each optional; not available together with `sort` */
    double *out_lon, *out_lat, *out_dep;
    int32_t *out_face, *out_iy, *out_ix, *out_izl;
    int32_t *out_ncross, *out_status;
    /* optional, 4 counters: particles live at the start, wall crossings, particles that hit max_iteration,
     * particles with a non-zero status */
    uint64_t *out_stats;
    /* != 0: put the particles of every chunk into cell order on the device before advecting (sdk_cell_sort);
     * the records come back in the caller's order all the same */
    int32_t sort;
    int32_t reserved;
} sdk_track_io;
int64_t sdk_track_scratch_bytes(int64_t n);
int sdk_track_host(const sdk_grid *grid, const sdk_velocity *vel, const sdk_track_io *io,
                   const double *ts_dev, int32_t n_t, double t_stop, const sdk_advect_params *par,
                   void *scratch, void *stream);

/* Output of sdk_locate: DEVICE pointers, n entries each (px, py: 4*n).  NULL members are skipped.
 * Mirrors the RelCoord families of the reference (seaduck/ocedata.py:119-143).  px / py as in
 * sdk_particles (corners, or bx, cs, dx / by, sn, dy in float64 for the schemes without corners: the
 * rectilinear dx, dy = cos(by) * deg2m * spacing are float64 in the reference, seaduck/utils.py:602). */
typedef struct sdk_located {
    int32_t *face, *iy, *ix;
    double *rx, *ry;
    float *cs, *sn, *dx, *dy, *bx, *by;
    double *px, *py;
    int32_t *iz;       double *rz, *dz, *bz;                 /* nearest Z      (_find_rel_v)      */
    int32_t *izl_near; double *rzl_near, *dzl_near, *bzl_near; /* nearest Zl   (_find_rel_vl)     */
    int32_t *izl;      double *rzl, *dzl, *bzl;              /* linear Zl      (_find_rel_vl_lin) */
    int32_t *it;       double *rt, *dt, *bt;                 /* nearest time   (_find_rel_t)      */
    int32_t *iz_lin;   double *rz_lin, *dz_lin, *bz_lin;     /* linear Z       (_find_rel_v_lin)  */
    int32_t *it_lin;   double *rt_lin, *dt_lin, *bt_lin;     /* linear time    (_find_rel_t_lin)  */
    int32_t *status;   /* SDK_ST_BAD_CELL, SDK_ST_OOB* */
} sdk_located;

/* Build the lat/lon bucket grid used by sdk_locate (replaces create_tree, seaduck/utils.py:225).
 * Buckets cover [lon0, lon0+span_lon] x [lat0, lat0+span_lat] degrees with nlon x nlat buckets;
 * span_lon >= 360 means periodic in longitude. */
int sdk_grid_build_locator(sdk_grid *grid, double lon0, double lat0, double span_lon, double span_lat,
                           int32_t nlon, int32_t nlat, void *stream);

/* Position.from_latlon, seaduck/eulerian.py:138 -> OceData._find_rel_h/_v/_vl/_vl_lin/_t
 * (seaduck/ocedata.py:422-488).  lon/lat, dep, t: device arrays of n doubles or NULL.
 * ts: device array of the n_t model time levels (seconds) or NULL. */
int sdk_locate(const sdk_grid *grid, int64_t n, const double *lon, const double *lat, const double *dep,
               const double *t, const double *ts, int32_t n_t, const sdk_located *out, void *stream);

/* ---- masked kernel-weight interpolation (Position.interpolate, seaduck/eulerian.py:1268) ---------- */
#define SDK_MAX_NODES 32
#define SDK_MAX_LEVELS 8
#define SDK_MAX_COEF 9
/* neighbour index of a stencil node in the `rim` tables: 5 bits face, 13 bits iy, 13 bits ix */
#define SDK_PACK(f, iy, ix) (((f) << 26) | ((iy) << 13) | (ix))
#define SDK_NODE_RAISES (-1) /* the reference raises for this node (no such neighbour): result NaN */
#define SDK_NODE_LAST (-2)   /* the reference indexes [-1, -1] (stepped out of a closed box) */
enum { SDK_MODE_NEAREST = 0, SDK_MODE_LINEAR = 1, SDK_MODE_DERIV = 2 };
/* how the weight of a node is put together from its two polynomials A(rx), C(ry) */
enum { SDK_W_A = 0, SDK_W_C = 1, SDK_W_A_PLUS_C = 2, SDK_W_A_TIMES_C = 3 };

/* One level of a KnW inheritance (seaduck/kernel_weight.py:645): which nodes it uses and how each
 * node's Lagrange weight (kernel_weight_x :84, kernel_weight_s :313) is formed. */
typedef struct sdk_kernel_level {
    uint32_t node_mask;
    uint32_t reserved;
    uint8_t mode[SDK_MAX_NODES]; /* SDK_W_* per node */
} sdk_kernel_level;

/* A KnW: horizontal nodes, inheritance levels, vertical / temporal mode.
 * `coef` is the DEVICE table of the weight polynomials, [n_levels][n_nodes][2][n_coef] doubles: for
 * every level and node the coefficients (ascending powers) of A(rx) and of C(ry); the host expands the
 * reference's products of (r - node) factors, derivative order included.
 * `rim` is the DEVICE table of neighbour indices for the cells whose stencil leaves the (per face)
 * array: for every rim cell (slot order documented in seaduck_b200/interp.py) and node,
 * SDK_PACK(face, iy, ix) or SDK_NODE_*; `rim_code` (vectors only) bit0 = value comes from the OTHER
 * component, bit1 = negative sign (four_matrix_for_uv, seaduck/topology.py:689). */
typedef struct sdk_kernel_desc {
    int32_t n_nodes, n_levels, vmode, tmode, ignore_mask, bottom_no_flux;
    int32_t reach;     /* max(|dx|, |dy|): cells closer than this to a face edge use `rim` */
    int32_t rim_dense; /* 1: `rim` covers every cell (tiny grids), slot = iy*nx + ix */
    int32_t n_coef, reserved;
    int8_t dx[SDK_MAX_NODES], dy[SDK_MAX_NODES];
    sdk_kernel_level level[SDK_MAX_LEVELS];
    const double *coef;
    const int32_t *rim;
    const uint8_t *rim_code;
} sdk_kernel_desc;

/* A gridded variable on the device: data [t][z][face][ny][nx] (axes of size 1 when absent) and the
 * wet mask that goes with its grid position (uint8, [mask_nz][face][mask_ny][mask_nx], or NULL). */
typedef struct sdk_field {
    const void *data;
    int32_t dtype, has_time, nt, it0, has_z, nz, ny, nx, xstag, ystag;
    const uint8_t *mask;
    int32_t mask_nz, mask_ny, mask_nx, reserved;
} sdk_field;

/* Points: indices and rel-coords (DEVICE pointers).  kz/rz and kt/rt are the vertical / temporal
 * index and rel-coordinate that match the variable's grid and the kernel's mode (iz|iz_lin|izl|
 * izl_lin, it|it_lin); the second vertical node is kz-1, the second time node kt+1. */
typedef struct sdk_points {
    int64_t n;
    const int32_t *face, *iy, *ix;
    const double *rx, *ry;
    const int32_t *kz;
    const double *rz;
    const int32_t *kt;
    const double *rt;
    const float *cs, *sn;
    /* optional: result of point i goes to out[out_index[i]] (the caller runs the kernel over a
     * cell-sorted copy of the points and wants the results in its own order); NULL = out[i] */
    const int32_t *out_index;
} sdk_points;

/* Position.interpolate for one scalar variable, seaduck/eulerian.py:1268 (str var_name; the register / mask / read /
 * weight stages :887-1266 and KnW.get_weight kernel_weight.py:772 in one launch) */
int sdk_interp_scalar(const sdk_grid *grid, const sdk_points *pts, const sdk_field *field,
                      const sdk_kernel_desc *knw, double *out, void *stream);
/* several scalar variables in one launch (at most SDK_MAX_MULTI_FIELDS): they must share the kernel, the grid position
 * (dimensions, staggering), the mask and the time / vertical layout -- T and S of an ocean model, say.  Node indices, wet
 * flags, the choice of the stencil and the weights are worked out once per point; every field adds its gathers and a dot
 * product.  fields: HOST array of n_fields descriptors, outs: HOST array of n_fields DEVICE pointers. */
#define SDK_MAX_MULTI_FIELDS 4
int sdk_interp_scalar_multi(const sdk_grid *grid, const sdk_points *pts, const sdk_field *fields, int32_t n_fields,
                            const sdk_kernel_desc *knw, double *const *outs, void *stream);
/* horizontal vector on a grid with faces (components may swap / change sign across faces);
 * vec_transform != 0 rotates the result to east/north with cs, sn (utils.py:206). */
int sdk_interp_vector(const sdk_grid *grid, const sdk_points *pts, const sdk_field *fu, const sdk_field *fv,
                      const sdk_kernel_desc *ku, const sdk_kernel_desc *kv, int vec_transform, double *out_u,
                      double *out_v, void *stream);
/* maskU / maskV / maskWvel from maskC (seaduck/get_masks.py:12,46,79), all [nz][face][ny][nx] uint8 */
int sdk_build_masks(const sdk_grid *grid, const uint8_t *maskC, int32_t nz, uint8_t *maskU, uint8_t *maskV,
                    uint8_t *maskW, void *stream);

/* Coordinates in, interpolated values out, all HOST memory (pinned for full speed): what
 * `OceInterp(od, var_list, x, y, z, t)` does for Eulerian points (seaduck/oceinterp.py:107-112 ->
 * Position.from_latlon eulerian.py:138 + Position.interpolate :1268).  The points are cut into chunks that go
 * round three internal streams: upload, (level, bucket) sort, sdk_locate, one launch per job with the results
 * scattered back to the caller's order, read-back -- the copies of one chunk overlap the kernels of the others.
 * A job is what one launch computes: up to SDK_MAX_MULTI_FIELDS scalars that share a kernel and a grid position,
 * or one horizontal vector.  `zsel` / `tsel` say which of sdk_locate's index families the variable and the
 * kernel's vkernel / tkernel ask for (eulerian.py:451-518, 1196-1224). */
enum { SDK_SEL_NONE = 0, SDK_SEL_Z = 1, SDK_SEL_Z_LIN = 2, SDK_SEL_ZL = 3, SDK_SEL_ZL_LIN = 4 };
enum { SDK_SEL_T = 1, SDK_SEL_T_LIN = 2 };
enum { SDK_JOB_SCALARS = 0, SDK_JOB_VECTOR = 1 };
typedef struct sdk_interp_job {
    int32_t kind;     /* SDK_JOB_* */
    int32_t n_fields; /* scalars: 1 .. SDK_MAX_MULTI_FIELDS; vector: 2 (u, v) */
    int32_t zsel, tsel;
    int32_t vec_transform, reserved;
    sdk_field fields[SDK_MAX_MULTI_FIELDS];
    sdk_kernel_desc knw[2]; /* scalars: knw[0]; vector: kernels of u and v */
    double *out[SDK_MAX_MULTI_FIELDS]; /* HOST arrays of n doubles, one per field */
} sdk_interp_job;
typedef struct sdk_interp_io {
    int64_t n;
    const double *lon, *lat, *dep, *t; /* HOST; dep and t may be NULL */
    int32_t sort;       /* != 0: order the points of a chunk by (level, bucket) before locating them */
    int32_t max_chunks; /* 0 = the library's default */
    /* optional HOST word: OR of sdk_locate's status bits over all points (SDK_ST_OOB* = the reference's
     * "Value out of bound.", SDK_ST_BAD_CELL) */
    uint32_t *out_flags;
} sdk_interp_io;
int64_t sdk_interp_host_scratch_bytes(int64_t n, int32_t n_outputs);
int sdk_interp_host(const sdk_grid *grid, const sdk_interp_io *io, const double *ts_dev, int32_t n_t,
                    const sdk_interp_job *jobs, int32_t n_jobs, void *scratch, void *stream);
/* OR of n status words (device) into *flags (device, not cleared first): what decides whether the reference's
 * "Value out of bound." (find_rel, seaduck/utils.py:321) is raised for a batch of located points */
int sdk_status_or(const int32_t *status, int64_t n, uint32_t *flags, void *stream);

/* End of Particle.to_next_stop without a host round trip (seaduck/lagrangian.py:793-802): if the
 * advect call counted live particles (stats[0] > 0; the reference returns early otherwise) set
 * t := t_stop for every particle and, when time_midp is given, it := time index of t_stop; then add
 * stats[0..3] to total[0..3] (running totals read lazily by the host) -- all on the device.
 * stats, total: DEVICE arrays of 4 uint64 (total may be NULL); time_midp / it may be NULL. */
int sdk_finish_stop(double *t, int32_t *it, const double *time_midp, int32_t n_midp, double t_stop,
                    const uint64_t *stats, uint64_t *total, int64_t n, void *stream);

/* it := 0 before time_midp[0], else find_rel(t, time_midp) + 1 (seaduck/lagrangian.py:796-802). */
int sdk_time_index(const double *t, const double *time_midp, int32_t n_midp, int32_t *it, int64_t n,
                   void *stream);

/* Crossing-log post-processing of the Lagrangian budget (seaduck/lagrangian_budget.py:245
 * find_ind_frac_tres with by_type=True) on a log that is still on the device: records [offset[p],
 * offset[p+1]) belong to particle p.  ind1 / ind2: int16 [rows][n_rec] (rows = 5 with a face dimension:
 * iw, iz, face, iy, ix; else 4), frac / tres: double [n_rec]; work: scratch of n_rec doubles.  Device
 * pointers.  Needs particles with a depth (the reference does too). */
int sdk_log_budget(const sdk_grid *grid, const sdk_log_record *records, const int64_t *offset, int64_t n_particles,
                   int64_t n_rec, int16_t *ind1, int16_t *ind2, double *frac, double *tres, double *work, void *stream);

/* calculate_budget (lagrangian_budget.py:603) on what sdk_log_budget produced: tracer concentration at every
 * record from the wall concentrations (prefetch_vector :548), its change between consecutive records of a
 * particle, the part explained by the Eulerian tendency (lhs_contribution :575) and the split of the rest
 * over the right-hand-side terms in proportion to term^2 (contr_p_relaxed :583 with p = 1).  Device pointers;
 * fields are one model time step in the dataset's layout. */
#define SDK_MAX_BUDGET_TERMS 8
typedef struct sdk_budget_fields {
    int32_t n_terms;     /* right-hand-side terms, <= SDK_MAX_BUDGET_TERMS */
    int32_t dir_of_time; /* -1: the particles ran backward in time, +1 forward */
    int32_t rows;        /* rows of ind1 / ind2: 5 (iw, iz, face, iy, ix) or 4 (no face) */
    int32_t reserved;
    int32_t cell_shape[4]; /* (nz, nface, ny, nx) of term[] and lhs; nface = 1 without faces */
    int32_t wall_shape[4]; /* same for each of the three arrays stacked in walls */
    const double *term[SDK_MAX_BUDGET_TERMS];
    const double *lhs;
    const double *walls; /* [3][wall_shape] */
} sdk_budget_fields;
/* conc: [n_rec]; contr: [n_terms + 2][n_rec - 1] = the terms in order, then "error", then the lhs
 * contribution; work: n_rec bytes of scratch.  Indices follow numpy (negative values count from the end). */
int sdk_log_calculate_budget(const sdk_log_record *records, const int64_t *offset, int64_t n_particles, int64_t n_rec,
                             const int16_t *ind1, const int16_t *ind2, const double *frac, const double *tres,
                             const sdk_budget_fields *fields, double *conc, double *contr, unsigned char *work,
                             void *stream);

/* read_wall_list (lagrangian_budget.py:406, the branch for grids with faces) and the quantities that
 * check_particle_data_compat (:470) compares: for every record of the crossing log the values of three wall fields on
 * the six walls of its cell (left, right, y-bottom, y-top, deep, shallow; the right / top wall may lie on a
 * neighbouring, rotated face), the wall velocities from (u, du, r) (_uleftright_from_udu utils.py:852), the
 * convergence of their products (crude_convergence :464) and the Eulerian convergence at the cell.  Device
 * pointers; c_list and u_list are [6][n_rec], lag_conv and eul_conv [n_rec]; u_list, lag_conv, eul_conv and
 * fields->conv may be NULL. */
typedef struct sdk_wall_fields {
    const double *x_walls, *y_walls, *z_walls; /* each [nz'][nface][ny'][nx'] with its own shape below */
    const double *conv;                        /* Eulerian convergence, shape_c */
    int32_t shape_x[4], shape_y[4], shape_z[4], shape_c[4];
    /* cos and sin of pi - DIRECTIONS[edge] + DIRECTIONS[new_edge] (topology.py:248) for edge, new_edge in
     * 0..3, index edge * 4 + new_edge, as the host's libm computes them: the reference multiplies with these
     * unrounded numbers (cos(pi/2) = 6e-17 ...), so they are data, not something to recompute on the device */
    double rot_cos[16], rot_sin[16];
    int32_t scalar; /* 1: absolute values of the coefficients (scalar fields on walls) */
    int32_t reserved;
} sdk_wall_fields;
int sdk_log_wall_values(const sdk_grid *grid, const sdk_log_record *records, int64_t n_rec, const sdk_wall_fields *fields,
                        double *c_list, double *u_list, double *lag_conv, double *eul_conv, void *stream);

/* ---- stencils of the Eulerian budget (seaduck/eulerian_budget.py); fp64, device pointers ------------- */
/* Halo gather: out[o][k] = map[k] < 0 ? 0 : src[o * src_plane + map[k]] for o < n_outer, k < n_map.  The host
 * builds `map` for buffer_x_withface :214 / buffer_y_withface :262 (LLC faces incl. rotation),
 * buffer_x_periodic :309, buffer_y_periodic :321 and buffer_z_nearest :333. */
int sdk_gather_plane(const double *src, int64_t n_outer, int64_t src_plane, const int64_t *map, int64_t n_map,
                     double *out, void *stream);
/* Tracer concentration on cell walls from a buffered field: buf [n_outer][n_out + 3][n_inner] (two halo
 * points before, one after), vel and out [n_outer][n_out][n_inner]. */
enum {
    SDK_STENCIL_DST3 = 0,      /* third_order_DST_x / _y :477-532 */
    SDK_STENCIL_LIMITER = 1,   /* second_order_flux_limiter_x / _y :363-417 (superbee) */
    SDK_STENCIL_LIMITER_Z = 2  /* second_order_flux_limiter_z :419 */
};
int sdk_wall_stencil(int scheme, const double *buf, const double *vel, double *out, int64_t n_outer, int64_t n_out,
                     int64_t n_inner, void *stream);
/* third_order_upwind_z :448: s, w, out [nz][n_inner]; sets w[0][:] = 0 like the reference does */
int sdk_upwind3_z(const double *s, double *w, double *out, int64_t nz, int64_t n_inner, void *stream);

/* ---- output records of a stop ---------------------------------------------------------------------------
 * What Particle.to_list_of_time returns per stop (seaduck/lagrangian.py:804, the deepcopy of :725) and what
 * the ranks exchange is, per particle, a position and a cell: 32 bytes. */
typedef struct sdk_snapshot_record {
    double lon, lat, dep;
    uint64_t cell; /* SDK_CELL_PACK(face, iy, ix, izl) */
} sdk_snapshot_record;
/* izl (= izl_lin) in bits 0..11, ix in 12..31, iy in 32..51, face in 52..57; bits 58..63 are zero */
#define SDK_CELL_PACK(f, iy, ix, izl) ((((uint64_t)(f) & 0x3f) << 52) | (((uint64_t)(iy) & 0xfffff) << 32) | (((uint64_t)(ix) & 0xfffff) << 12) | ((uint64_t)(izl) & 0xfff))
#define SDK_CELL_FACE(w) ((int32_t)(((w) >> 52) & 0x3f))
#define SDK_CELL_IY(w) ((int32_t)(((w) >> 32) & 0xfffff))
#define SDK_CELL_IX(w) ((int32_t)(((w) >> 12) & 0xfffff))
#define SDK_CELL_IZL(w) ((int32_t)((w) & 0xfff))

/* What a stop of Particle.to_list_of_time hands back (the deepcopy of seaduck/lagrangian.py:725,804), 32 bytes per particle:
 * out[out_index ? out_index[i] : i] = record of particle i, i < p->n (reads lon, lat, dep, face, iy, ix, izl of the
 * device state `p`; NULL members read as zero).  `out_index` restores the caller's order when the state is cell
 * sorted (sdk_cell_sort).  Device pointers. */
int sdk_pack_records(const sdk_particles *p, const int32_t *out_index, sdk_snapshot_record *out, void *stream);

/* ---- cell order ----------------------------------------------------------------------
 * perm[k] = index of the particle that comes k-th when sorted by (izl, face, iy, ix): particles of one warp then
 * read neighbouring memory.  `scratch`: device buffer of at least sdk_cell_sort_scratch_bytes(n) bytes.
 * face / izl may be NULL.  sdk_particles_gather applies a permutation to the whole device state
 * (dst[k] = src[perm[k]] for every array both structures hold; the reference's counterpart is
 * Position.subset, seaduck/eulerian.py:319). */
int64_t sdk_cell_sort_scratch_bytes(int64_t n);
int sdk_cell_sort(const sdk_grid *grid, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix,
                  const int32_t *izl, int32_t *perm, void *scratch, int64_t scratch_bytes, void *stream);
int sdk_particles_gather(const sdk_particles *src, const sdk_particles *dst, const int32_t *perm, void *stream);

/* ---- exchange of the output records between the GPUs of a node ------------------------------------------
 * Every rank owns a receive buffer [n_slots][world][n_pad] of records and SDK_SIGNAL_WORDS 64-bit signal
 * words (zero-initialised), both in memory that every rank has mapped (symmetric memory: the host side
 * allocates and exchanges the mappings, e.g. torch.distributed._symmetric_memory).  sdk_pack_ship packs this rank's
 * records (particle i at position out_index ? out_index[i] : i, as in sdk_pack_records) and stores them into block [step % n_slots][rank] of EVERY rank's buffer from inside one kernel --
 * plain stores through the peer mappings, or one multimem.st per 16 bytes through `recv_multicast` when the
 * NVSwitch can replicate (NVLS) -- then posts `step` in every rank's arrived[rank].  It first waits until every
 * rank has consumed the step that last used the slot.  sdk_exchange_wait makes `stream` wait until all
 * ranks' records of `step` have landed here (wait != 0) and / or tells every rank that this rank is done
 * reading them (release != 0).  Steps count from 1.  Waits are bounded; a timeout sets signal[rank][SDK_SIGNAL_ERROR].
 * One exchange belongs to one stream at a time (`ticket` is a device counter of this rank, zero-initialised). */
#define SDK_MAX_RANKS 16
#define SDK_SIGNAL_ARRIVED 0
#define SDK_SIGNAL_CONSUMED SDK_MAX_RANKS
#define SDK_SIGNAL_ERROR (2 * SDK_MAX_RANKS)
#define SDK_SIGNAL_WORDS (2 * SDK_MAX_RANKS + 1)
typedef struct sdk_exchange {
    int32_t world, rank, n_slots, reserved;
    int64_t n_pad;                              /* records per rank and slot */
    sdk_snapshot_record *recv[SDK_MAX_RANKS];   /* recv[r]: this rank's mapping of rank r's receive buffer */
    sdk_snapshot_record *recv_multicast;        /* multicast mapping of the receive buffers, or NULL */
    uint64_t *signal[SDK_MAX_RANKS];            /* signal[r]: this rank's mapping of rank r's signal words */
    uint32_t *ticket;                           /* device, this rank */
} sdk_exchange;
int sdk_pack_ship(const sdk_exchange *x, const sdk_particles *p, const int32_t *out_index, uint64_t step, void *stream);
int sdk_exchange_wait(const sdk_exchange *x, uint64_t step, int wait, int release, void *stream);

/* Self-test of the device math helpers used by sdk_advect (csrc/fastmath.cuh): out[i] = op(a[i], b[i]),
 * bad[i] = 1 where the helper asked for the slow path.  Device pointers; b may be NULL for unary ops.
 * SDK_SELFTEST_STEP runs one analytic in-cell step (time to the walls, first event, move: _time2wall
 * seaduck/utils.py:859, _which_early :875, _increment :777) per element instead: a = [n][10] doubles
 * (x0[3], u[3], du[3], tf), out = [n][8] (tend, t_event, newx[3], newu[3]); b and bad are not used. */
enum {
    SDK_SELFTEST_DIV = 0, SDK_SELFTEST_LOG = 1, SDK_SELFTEST_EXP = 2,
    SDK_SELFTEST_DIV_IEEE = 3, SDK_SELFTEST_LOG_CUDA = 4, SDK_SELFTEST_EXP_CUDA = 5, SDK_SELFTEST_STEP = 6
};
int sdk_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad, void *stream);
/* Self-test of the device's cell-to-cell move (the closed form behind Topology.ind_tend, seaduck/topology.py:389: LLC and
 * ASTE face tables, x-periodic wrap, closed boxes): in = [n][4] (face, iy, ix, direction 0 up / 1 down / 2 left / 3 right),
 * out = [n][5] (face, iy, ix, edge of the new face the move came in through or -1, SDK_ST_* status).  Device pointers. */
int sdk_selftest_cell_move(const sdk_grid *grid, int64_t n, const int32_t *in, int32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEADUCK_B200_H */
