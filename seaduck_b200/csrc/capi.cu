// C ABI of libseaduck_b200.so (see include/seaduck_b200.h).  Host-side glue only: argument checks,
// handle bookkeeping, launch configuration.  All arithmetic lives in the kernels.
#include <cuda_runtime.h>

#include <climits>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <cstdlib>
#include <string>
#include <vector>

#include "advect_args.cuh"
#include "interp_args.cuh"
#include "budget_args.cuh"
#include "locate_args.cuh"

namespace sdk {
cudaError_t launch_gather_plane(const double *src, int64_t n_outer, int64_t src_plane, const int64_t *map, int64_t n_map,
                                double *out, cudaStream_t s);
cudaError_t launch_wall_stencil(int scheme, const double *buf, const double *vel, double *out, int64_t n_outer, int64_t n_out,
                                int64_t n_inner, cudaStream_t s);
cudaError_t launch_upwind3_z(const double *sfield, double *w, double *out, int64_t nz, int64_t n_inner, cudaStream_t s);
cudaError_t launch_pack_records(const sdk_particles &p, int64_t n, const int32_t *out_index, sdk_snapshot_record *out,
                                cudaStream_t s);
cudaError_t launch_pack_ship(const sdk_exchange &x, const sdk_particles &p, int64_t n, const int32_t *out_index,
                             uint64_t step, int sm_count, cudaStream_t s);
cudaError_t launch_exchange_wait(const sdk_exchange &x, uint64_t step, int wait, int release, cudaStream_t s);
size_t cell_sort_scratch_bytes(int64_t n);
cudaError_t launch_cell_sort(const GridDev &g, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix,
                             const int32_t *izl, int32_t *perm, void *scratch, size_t scratch_bytes, cudaStream_t s);
cudaError_t launch_particles_gather(const sdk_particles &src, const sdk_particles &dst, const int32_t *perm, cudaStream_t s);
cudaError_t launch_point_sort(const GridDev &g, const BucketGrid &b, int64_t n, const double *lon, const double *lat, const double *dep,
                              int32_t *perm, void *scratch, size_t scratch_bytes, cudaStream_t s);
cudaError_t launch_points_gather(int64_t n, const int32_t *perm, const double *lon, const double *lat, const double *dep, const double *t,
                                 double *olon, double *olat, double *odep, double *ot, cudaStream_t s);
cudaError_t launch_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad,
                                 cudaStream_t s);
cudaError_t launch_selftest_cell_move(const GridDev &g, int64_t n, const int32_t *in, int32_t *out, cudaStream_t s);
}

namespace {
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return SDK_ECUDA;
}
}  // namespace

constexpr unsigned kQueueSlots = 64;
// measured best for 10^6 particles (tools/exp_e2e.py): 2-4 chunks within 3 %, more chunks = more kernel tails.  Two other
// ways of cutting the call up were built and measured in round 2 (profiles/README.md, profiles/r02d_fed_queue_experiment.patch):
// chunked upload + locate into one state with ONE advection launch (2.31 ms against 2.32 ms), and one advection launch fed
// through a gated queue while the upload is still running (2.43 - 2.9 ms: the locate kernels crawl beside the persistent grid).
constexpr int kTrackChunks = 3;
constexpr int kTrackChunksMax = 32;

// +1 / -1 when the n floats of a DEVICE table are strictly ascending / descending, else 0
static int table_order(const float *dev, int n) {
    if (!dev || n < 2) return 0;
    std::vector<float> h((size_t)n);
    if (cudaMemcpy(h.data(), dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    bool up = true, down = true;
    for (int k = 1; k < n; ++k) {
        up = up && h[k] > h[k - 1];
        down = down && h[k] < h[k - 1];
    }
    return up ? 1 : (down ? -1 : 0);
}

struct sdk_grid {
    sdk::GridDev dev;
    int sm_count;
    // device work-queue heads: one slot per launch, taken round robin, so that launches on
    // different streams (the chunks of sdk_track_host) never share a counter
    unsigned long long *queue;
    mutable unsigned next_queue;
    mutable cudaStream_t pipe[3];  // lazily created by sdk_track_host
    mutable cudaEvent_t pipe_ev[4];
    mutable bool has_pipe;
    sdk::BucketGrid buckets;
    bool has_buckets;
};

extern "C" {

int sdk_abi_version(void) { return SDK_ABI_VERSION; }

const char *sdk_last_error(void) { return g_err.c_str(); }

int sdk_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SDK_ECUDA;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SDK_ECUDA;
    return n;
}

int sdk_grid_create(const sdk_grid_desc *d, sdk_grid **out) {
    if (!d || !out) return fail(SDK_EINVAL, "sdk_grid_create: null argument");
    if (d->nface < 1 || d->nface > SDK_MAX_FACES) return fail(SDK_EINVAL, "sdk_grid_create: nface out of range");
    if (d->typ == SDK_TYP_LLC && (!d->nbr_face || !d->nbr_edge))
        return fail(SDK_EINVAL, "sdk_grid_create: LLC grids need the face tables");
    if (d->hmode < SDK_H_OCEANPARCEL || d->hmode > SDK_H_RECTILINEAR) return fail(SDK_EINVAL, "sdk_grid_create: unknown horizontal scheme");
    if (d->hmode != SDK_H_RECTILINEAR && (!d->XC || !d->YC)) return fail(SDK_EINVAL, "sdk_grid_create: XC / YC missing");
    if (d->hmode == SDK_H_RECTILINEAR && (!d->lon || !d->lat || !d->dlon || !d->dlat || !d->coslat32))
        return fail(SDK_EINVAL, "sdk_grid_create: the rectilinear scheme needs lon, lat, dlon, dlat and the cosine table");
    if (d->hmode == SDK_H_RECTILINEAR && d->has_face) return fail(SDK_EINVAL, "sdk_grid_create: a rectilinear grid has no faces");
    if (d->hmode == SDK_H_LOCAL_CARTESIAN && (!d->dX || !d->dY || !d->CS || !d->SN || !d->coslat32))
        return fail(SDK_EINVAL, "sdk_grid_create: the local_cartesian scheme needs dX, dY, CS, SN and the cosine table");
    if (d->hmode == SDK_H_OCEANPARCEL && (!d->XG || !d->YG))
        return fail(SDK_EINVAL, "sdk_grid_create: oceanparcel scheme needs XG / YG");
    if (d->hmode == SDK_H_OCEANPARCEL && (d->gny <= d->ny || d->gnx <= d->nx) && !d->corner_tab)
        return fail(SDK_EINVAL, "sdk_grid_create: corner table required when XG has the shape of XC");
    if (d->n_zl > 0 && (!d->Zl || !d->dZl)) return fail(SDK_EINVAL, "sdk_grid_create: Zl / dZl missing");
    if ((size_t)(d->n_zl + d->n_z) * sizeof(float) > 40 * 1024)
        return fail(SDK_EUNSUPPORTED, "sdk_grid_create: too many vertical levels for shared memory staging");
    sdk_grid *g = new (std::nothrow) sdk_grid();
    if (!g) return fail(SDK_ENOMEM, "sdk_grid_create: out of host memory");
    sdk::GridDev &v = g->dev;
    std::memset(&v, 0, sizeof v);
    v.typ = d->typ;
    v.has_face = d->has_face;
    v.nface = d->nface;
    v.ny = d->ny;
    v.nx = d->nx;
    v.gny = d->gny;
    v.gnx = d->gnx;
    v.hmode = d->hmode;
    v.n_zl = d->n_zl;
    v.n_z = d->n_z;
    v.n_zc = d->n_zc;
    for (int i = 0; i < SDK_MAX_FACES * 4; ++i) {
        v.nbr_face[i] = -1;
        v.nbr_edge[i] = -1;
    }
    if (d->nbr_face && d->nbr_edge) {
        for (int i = 0; i < d->nface * 4; ++i) {
            v.nbr_face[i] = (signed char)d->nbr_face[i];
            v.nbr_edge[i] = (signed char)d->nbr_edge[i];
        }
    }
    v.XC = d->XC; v.YC = d->YC; v.XG = d->XG; v.YG = d->YG;
    v.dX = d->dX; v.dY = d->dY; v.CS = d->CS; v.SN = d->SN;
    v.Zl = d->Zl; v.dZl = d->dZl; v.Z = d->Z; v.dZ = d->dZ;
    v.corner_tab = d->corner_tab;
    v.lon1 = d->lon; v.lat1 = d->lat; v.dlon = d->dlon; v.dlat = d->dlat;
    v.coslat32 = d->coslat32;
    v.dlon_is_f32 = d->dlon_is_f32;
    g->sm_count = sdk_sm_count();
    if (g->sm_count <= 0) {
        delete g;
        return fail(SDK_ECUDA, "sdk_grid_create: no CUDA device");
    }
    // strictly monotonic vertical tables are searched by bisection (locate.cu: nearest_index)
    v.z_order = table_order(d->Z, d->n_zc);
    v.zl_order = table_order(d->Zl, d->n_zl);
    g->next_queue = 0;
    g->has_pipe = false;
    cudaError_t e = cudaMalloc(&g->queue, kQueueSlots * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        delete g;
        return cuda_fail(e, "sdk_grid_create: cudaMalloc");
    }
    g->has_buckets = false;
    if (d->hmode == SDK_H_OCEANPARCEL) {
        // corner coordinates as (XG, YG) pairs: one 8-byte read per corner
        const int64_t ncorner = (int64_t)d->nface * d->gny * d->gnx;
        float2 *xyg = nullptr;
        if ((e = cudaMalloc(&xyg, sizeof(float2) * (size_t)ncorner)) != cudaSuccess) {
            cudaFree(g->queue);
            delete g;
            return cuda_fail(e, "sdk_grid_create: cudaMalloc (corner pairs)");
        }
        e = sdk::launch_interleave_corners(d->XG, d->YG, xyg, ncorner, nullptr);
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) {
            cudaFree(xyg);
            cudaFree(g->queue);
            delete g;
            return cuda_fail(e, "sdk_grid_create: corner pairs");
        }
        v.XYG = xyg;
    }
    *out = g;
    return SDK_OK;
}

int sdk_grid_destroy(sdk_grid *g) {
    if (!g) return SDK_OK;
    cudaFree(g->queue);
    cudaFree((void *)g->dev.XYG);
    if (g->has_pipe) {
        for (auto &st : g->pipe) cudaStreamDestroy(st);
        for (auto &ev : g->pipe_ev) cudaEventDestroy(ev);
    }
    if (g->has_buckets) {
        cudaFree(g->buckets.table);
        cudaFree((void *)g->buckets.cxyz);
        cudaFree((void *)g->buckets.rim);
        cudaFree((void *)g->buckets.wide);
    }
    delete g;
    return SDK_OK;
}

static int check_exchange(const sdk_exchange *x, const char *who);
int sdk_pack_ship(const sdk_exchange *x, const sdk_particles *p, const int32_t *out_index, uint64_t step, void *stream);

static int fill_args(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, double t_stop,
                     const sdk_advect_params *par, const sdk_log *log, int commit, sdk::AdvectArgs &a) {
    if (!grid || !vel || !p) return fail(SDK_EINVAL, "null argument");
    if (!vel->U || !vel->V) return fail(SDK_EINVAL, "velocity arrays missing");
    if (vel->transport && !vel->Vol) return fail(SDK_EINVAL, "transport mode needs Vol");
    if (vel->dtype != SDK_F64 && vel->dtype != SDK_F32) return fail(SDK_EINVAL, "bad velocity dtype");
    if (p->n < 0) return fail(SDK_EINVAL, "negative particle count");
    const int has_dep = par ? par->has_dep : 0;
    if (p->n > 0) {
        if (!p->lon || !p->lat || !p->t || !p->iy || !p->ix || !p->rx || !p->ry || !p->u || !p->v || !p->du ||
            !p->dv || !p->px || !p->py)
            return fail(SDK_EINVAL, "particle state arrays missing");
        if (grid->dev.has_face && !p->face) return fail(SDK_EINVAL, "face array missing");
        if (has_dep && (!p->dep || !p->rzl || !p->w || !p->dw || !p->bzl || !p->dzl || !p->izl))
            return fail(SDK_EINVAL, "vertical state arrays missing");
        if (has_dep && grid->dev.n_zl == 0) return fail(SDK_EINVAL, "particles have depth but the grid has no Zl");
        if (!vel->transport && grid->dev.hmode == SDK_H_OCEANPARCEL && (!p->dx || !p->dy))
            return fail(SDK_EINVAL, "dx / dy missing (velocity mode)");
        if (vel->has_time && !p->it) return fail(SDK_EINVAL, "time index array missing");
    }
    a.g = grid->dev;
    sdk::VelDev &v = a.v;
    v.U = vel->U; v.V = vel->V; v.W = vel->W; v.Vol = vel->Vol;
    v.dtype = vel->dtype; v.vol_dtype = vel->vol_dtype;
    v.has_time = vel->has_time; v.it0 = vel->it0; v.nt = vel->nt > 0 ? vel->nt : 1;
    v.has_z = vel->has_z; v.nzU = vel->nzU > 0 ? vel->nzU : 1; v.nzW = vel->nzW > 0 ? vel->nzW : 1;
    v.nyU = vel->nyU; v.nxU = vel->nxU; v.nyV = vel->nyV; v.nxV = vel->nxV;
    v.u_stag = vel->u_stag; v.v_stag = vel->v_stag; v.vol_has_z = vel->vol_has_z;
    v.transport = vel->transport;
    v.has_w = (vel->W != nullptr) && has_dep;
    v.planeU = (int64_t)grid->dev.nface * v.nyU * v.nxU;
    v.planeV = (int64_t)grid->dev.nface * v.nyV * v.nxV;
    v.planeC = (int64_t)grid->dev.nface * grid->dev.ny * grid->dev.nx;
    a.p = *p;
    if (log) a.log = *log; else std::memset(&a.log, 0, sizeof a.log);
    a.t_stop = t_stop;
    a.tol = par ? par->tol : 1e-4;
    a.trim_tol = par ? par->trim_tol : 1e-14;
    a.max_iteration = par ? par->max_iteration : 200;
    a.kick_back = par ? par->kick_back : 0;
    a.has_dep = has_dep;
    a.has_log = log != nullptr;
    a.commit = commit;
    a.keep_status = (par && par->keep_status && p->status) ? 1 : 0;
    a.ship = 0;
    a.ship_index = nullptr;
    a.ship_step = 0;
    std::memset(&a.x, 0, sizeof a.x);
    if (par && par->ship) {
        if (log || p->active) return fail(SDK_EINVAL, "the fused output exchange does not go with a crossing log or an active mask");
        if (!commit) return fail(SDK_EINVAL, "the fused output exchange needs commit != 0");
        int rc = check_exchange(par->ship, "sdk_advect");
        if (rc) return rc;
        if (p->n > par->ship->n_pad) return fail(SDK_EINVAL, "sdk_advect: particle count does not fit the exchange's n_pad");
        if (par->ship_step < 1) return fail(SDK_EINVAL, "sdk_advect: exchange steps count from 1");
        a.ship = 1;
        a.x = *par->ship;
        a.ship_index = par->ship_index;
        a.ship_step = par->ship_step;
    }
    int thr = par ? par->refill_threshold : 0;
    a.refill_threshold = thr > 0 ? (thr > 32 ? 32 : thr) : 8;
    a.queue = grid->queue + (grid->next_queue++ % kQueueSlots);
    a.stats = par ? (unsigned long long *)par->stats : nullptr;
    return SDK_OK;
}

int sdk_advect(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, double t_stop,
               const sdk_advect_params *par, const sdk_log *log, int commit, void *stream_) {
    sdk::AdvectArgs a;
    int rc = fill_args(grid, vel, p, t_stop, par, log, commit, a);
    if (rc) return rc;
    if (p->n == 0) {
        // nothing to advect, but the other ranks still wait for this rank's (empty) block
        if (par && par->ship) return sdk_pack_ship(par->ship, p, nullptr, par->ship_step, stream_);
        return SDK_OK;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_advect: memset");
    const int threads = (par && par->threads_per_block > 0) ? par->threads_per_block : sdk::advect_default_threads();
    const int per_sm = (par && par->blocks_per_sm > 0) ? par->blocks_per_sm : sdk::advect_default_blocks_per_sm(sdk::advect_profile(a, threads)) * sdk::advect_default_threads() / threads;
    int64_t need = (p->n + threads - 1) / threads;
    int sms = grid->sm_count - ((par && par->reserve_sms > 0) ? par->reserve_sms : 0);
    if (sms < 1) sms = 1;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > need) blocks = need;
    if (threads > sdk::advect_max_threads() || threads % 32)
        return fail(SDK_EINVAL, "threads_per_block must be a multiple of 32 and at most the compiled block size");
    e = sdk::launch_advect(a, (int)blocks, threads, stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_advect: launch");
    return SDK_OK;
}

int sdk_advect_profile(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, const sdk_advect_params *par,
                       const sdk_log *log) {
    sdk::AdvectArgs a;
    int rc = fill_args(grid, vel, p, 0.0, par, log, 1, a);
    if (rc) return rc < 0 ? rc : -rc;
    const int threads = (par && par->threads_per_block > 0) ? par->threads_per_block : sdk::advect_default_threads();
    return sdk::advect_profile(a, threads);
}

int sdk_get_u_du(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *p, int has_dep,
                 void *stream_) {
    sdk_advect_params par;
    std::memset(&par, 0, sizeof par);
    par.has_dep = has_dep;
    par.tol = 1e-4;
    par.trim_tol = 1e-14;
    par.max_iteration = 200;
    sdk::AdvectArgs a;
    int rc = fill_args(grid, vel, p, 0.0, &par, nullptr, 1, a);
    if (rc) return rc;
    cudaError_t e = sdk::launch_get_u_du(a, (cudaStream_t)stream_);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_get_u_du: launch");
    return SDK_OK;
}

int sdk_grid_build_locator(sdk_grid *grid, double lon0, double lat0, double span_lon, double span_lat,
                           int32_t nlon, int32_t nlat, void *stream) {
    if (!grid) return fail(SDK_EINVAL, "sdk_grid_build_locator: null grid");
    if (grid->dev.hmode == SDK_H_RECTILINEAR) return SDK_OK;  // two 1-D searches: nothing to build
    if (nlon < 1 || nlat < 1 || !(span_lon > 0) || !(span_lat > 0))
        return fail(SDK_EINVAL, "sdk_grid_build_locator: bad bucket geometry");
    if (grid->has_buckets) {
        cudaFree(grid->buckets.table);
        cudaFree((void *)grid->buckets.cxyz);
        cudaFree((void *)grid->buckets.rim);
        cudaFree((void *)grid->buckets.wide);
        grid->has_buckets = false;
    }
    sdk::BucketGrid &b = grid->buckets;
    b.lon0 = lon0;
    b.lat0 = lat0;
    b.periodic = span_lon >= 360.0;
    b.span_lon = b.periodic ? 360.0 : span_lon;
    b.span_lat = span_lat;
    b.nlon = nlon;
    b.nlat = nlat;
    b.inv_dlon = nlon / b.span_lon;
    b.inv_dlat = nlat / b.span_lat;
    b.table = nullptr;
    b.cxyz = nullptr;
    b.wide = nullptr;
    b.rim = nullptr;
    cudaError_t e = sdk::build_buckets(grid->dev, b, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_grid_build_locator");
    grid->has_buckets = true;
    return SDK_OK;
}

int sdk_locate(const sdk_grid *grid, int64_t n, const double *lon, const double *lat, const double *dep,
               const double *t, const double *ts, int32_t n_t, const sdk_located *out, void *stream) {
    if (!grid || !out) return fail(SDK_EINVAL, "sdk_locate: null argument");
    if (n < 0) return fail(SDK_EINVAL, "sdk_locate: negative count");
    if ((lon == nullptr) != (lat == nullptr)) return fail(SDK_EINVAL, "sdk_locate: lon and lat go together");
    if (lon) {
        if (grid->dev.hmode != SDK_H_RECTILINEAR && !grid->has_buckets) return fail(SDK_EINVAL, "sdk_locate: call sdk_grid_build_locator first");
        if (!out->iy || !out->ix || !out->rx || !out->ry) return fail(SDK_EINVAL, "sdk_locate: output arrays missing");
    }
    if (out->izl && (!out->rzl || !out->dzl || !out->bzl)) return fail(SDK_EINVAL, "sdk_locate: vertical output arrays missing");
    if (out->iz && (!out->rz || !out->dz || !out->bz)) return fail(SDK_EINVAL, "sdk_locate: vertical output arrays missing");
    if (out->iz_lin && (!out->rz_lin || !out->dz_lin || !out->bz_lin || !grid->dev.dZ))
        return fail(SDK_EINVAL, "sdk_locate: vertical output arrays (or dZ) missing");
    if (out->izl_near && (!out->rzl_near || !out->dzl_near || !out->bzl_near)) return fail(SDK_EINVAL, "sdk_locate: vertical output arrays missing");
    if (out->it_lin && (!out->rt_lin || !out->dt_lin || !out->bt_lin)) return fail(SDK_EINVAL, "sdk_locate: time output arrays missing");
    sdk::LocateArgs a;
    a.g = grid->dev;
    a.b = grid->buckets;
    a.n = n;
    a.lon = lon; a.lat = lat; a.dep = dep; a.t = t; a.ts = ts;
    a.n_t = n_t;
    a.max_walk = 4 * (grid->dev.ny + grid->dev.nx);
    a.out = *out;
    cudaError_t e = sdk::launch_locate(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_locate: launch");
    return SDK_OK;
}

int sdk_finish_stop(double *t, int32_t *it, const double *time_midp, int32_t n_midp, double t_stop,
                    const uint64_t *stats, uint64_t *total, int64_t n, void *stream) {
    if (!t || !stats || n < 0) return fail(SDK_EINVAL, "sdk_finish_stop: bad argument");
    if (time_midp && n_midp < 1) return fail(SDK_EINVAL, "sdk_finish_stop: empty time_midp");
    cudaError_t e = sdk::launch_finish_stop(t, it, time_midp, n_midp, t_stop, (const unsigned long long *)stats,
                                            (unsigned long long *)total, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_finish_stop: launch");
    return SDK_OK;
}

int sdk_time_index(const double *t, const double *time_midp, int32_t n_midp, int32_t *it, int64_t n,
                   void *stream) {
    if (!t || !time_midp || !it || n_midp < 1) return fail(SDK_EINVAL, "sdk_time_index: bad argument");
    cudaError_t e = sdk::launch_time_index(t, time_midp, n_midp, it, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_time_index: launch");
    return SDK_OK;
}

// ---- host-buffer variant ---------------------------------------------------------------------
namespace {
struct Field {
    size_t offset;  // of sdk_particles member
    size_t elem;    // bytes per element
    int mult;       // elements per particle
};
#define F(member, type, mult) {offsetof(sdk_particles, member), sizeof(type), mult}
const Field kFields[] = {
    F(lon, double, 1), F(lat, double, 1), F(dep, double, 1), F(t, double, 1),
    F(face, int32_t, 1), F(iy, int32_t, 1), F(ix, int32_t, 1), F(izl, int32_t, 1), F(it, int32_t, 1),
    F(rx, double, 1), F(ry, double, 1), F(rzl, double, 1),
    F(u, double, 1), F(v, double, 1), F(w, double, 1), F(du, double, 1), F(dv, double, 1), F(dw, double, 1),
    F(px, double, 4), F(py, double, 4), F(dx, float, 1), F(dy, float, 1),
    F(bzl, double, 1), F(dzl, double, 1), F(vol, double, 1), F(active, int32_t, 1),
    F(status, int32_t, 1), F(niter, int32_t, 1), F(ncross, int32_t, 1),
};
#undef F
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

int64_t sdk_particles_bytes(int64_t n) {
    size_t total = 0;
    for (const Field &f : kFields) total += align256((size_t)n * f.elem * f.mult);
    return (int64_t)total;
}

int sdk_advect_host(const sdk_grid *grid, const sdk_velocity *vel, const sdk_particles *hp, double t_stop,
                    const sdk_advect_params *par, void *scratch, void *stream_) {
    if (!hp || !scratch) return fail(SDK_EINVAL, "sdk_advect_host: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    sdk_particles dp;
    std::memset(&dp, 0, sizeof dp);
    dp.n = hp->n;
    char *cur = (char *)scratch;
    const int n_in = (int)(sizeof(kFields) / sizeof(kFields[0])) - 3;  // last three are outputs
    for (int k = 0; k < (int)(sizeof(kFields) / sizeof(kFields[0])); ++k) {
        const Field &f = kFields[k];
        void *const *hsrc = (void *const *)((const char *)hp + f.offset);
        void **ddst = (void **)((char *)&dp + f.offset);
        const size_t bytes = (size_t)hp->n * f.elem * f.mult;
        if (*hsrc) {
            *ddst = cur;
            if (k < n_in && bytes) {
                cudaError_t e = cudaMemcpyAsync(cur, *hsrc, bytes, cudaMemcpyHostToDevice, stream);
                if (e != cudaSuccess) return cuda_fail(e, "sdk_advect_host: H2D");
            }
            cur += align256(bytes);
        }
    }
    int rc = sdk_advect(grid, vel, &dp, t_stop, par, nullptr, 1, stream);
    if (rc) return rc;
    for (const Field &f : kFields) {
        void *const *hdst = (void *const *)((const char *)hp + f.offset);
        void *const *dsrc = (void *const *)((const char *)&dp + f.offset);
        const size_t bytes = (size_t)hp->n * f.elem * f.mult;
        if (*hdst && bytes && f.offset != offsetof(sdk_particles, active)) {
            cudaError_t e = cudaMemcpyAsync(*hdst, *dsrc, bytes, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return cuda_fail(e, "sdk_advect_host: D2H");
        }
    }
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_advect_host: sync");
    return SDK_OK;
}

// ---- interpolation ----------------------------------------------------------------------------------
static int check_interp(const sdk_grid *grid, const sdk_points *pts, const sdk_field *f, const sdk_kernel_desc *k) {
    if (!grid || !pts || !f || !k) return fail(SDK_EINVAL, "sdk_interp: null argument");
    if (pts->n < 0) return fail(SDK_EINVAL, "sdk_interp: negative count");
    if (pts->n > 0 && (!pts->iy || !pts->ix || !pts->rx || !pts->ry)) return fail(SDK_EINVAL, "sdk_interp: point arrays missing");
    if (grid->dev.has_face && pts->n > 0 && !pts->face) return fail(SDK_EINVAL, "sdk_interp: face array missing");
    if (!f->data) return fail(SDK_EINVAL, "sdk_interp: field data missing");
    if (k->n_nodes < 1 || k->n_nodes > SDK_MAX_NODES || k->n_levels < 1 || k->n_levels > SDK_MAX_LEVELS)
        return fail(SDK_EINVAL, "sdk_interp: kernel too large");
    if (k->reach > 0 && !k->rim) return fail(SDK_EINVAL, "sdk_interp: rim table missing");
    if (grid->dev.gny > 8191 || grid->dev.gnx > 8191 || grid->dev.nface > 32) return fail(SDK_EUNSUPPORTED, "sdk_interp: faces larger than 8191 x 8191");
    if (f->has_z && k->vmode != SDK_MODE_NEAREST && !pts->rz && k->vmode == SDK_MODE_LINEAR)
        return fail(SDK_EINVAL, "sdk_interp: rz missing");
    if (f->has_z && !pts->kz) return fail(SDK_EINVAL, "sdk_interp: vertical index missing");
    if (f->has_time && !pts->kt) return fail(SDK_EINVAL, "sdk_interp: time index missing");
    if (k->tmode == SDK_MODE_LINEAR && !pts->rt) return fail(SDK_EINVAL, "sdk_interp: rt missing");
    return SDK_OK;
}

int sdk_interp_scalar(const sdk_grid *grid, const sdk_points *pts, const sdk_field *field,
                      const sdk_kernel_desc *knw, double *out, void *stream) {
    int rc = check_interp(grid, pts, field, knw);
    if (rc) return rc;
    if (!out) return fail(SDK_EINVAL, "sdk_interp_scalar: output missing");
    sdk::InterpArgs a;
    a.g = grid->dev;
    a.pts = *pts;
    a.fu = *field;
    a.fv = *field;
    a.ku = *knw;
    a.kv = *knw;
    a.vec_transform = 0;
    a.out_u = out;
    a.out_v = nullptr;
    cudaError_t e = sdk::launch_interp(a, false, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_interp_scalar: launch");
    return SDK_OK;
}

int sdk_interp_scalar_multi(const sdk_grid *grid, const sdk_points *pts, const sdk_field *fields, int32_t n_fields,
                            const sdk_kernel_desc *knw, double *const *outs, void *stream) {
    if (!fields || !outs || n_fields < 1 || n_fields > SDK_MAX_MULTI_FIELDS)
        return fail(SDK_EINVAL, "sdk_interp_scalar_multi: between 1 and SDK_MAX_MULTI_FIELDS fields");
    sdk::InterpMultiArgs a;
    for (int k = 0; k < n_fields; ++k) {
        int rc = check_interp(grid, pts, &fields[k], knw);
        if (rc) return rc;
        if (!outs[k]) return fail(SDK_EINVAL, "sdk_interp_scalar_multi: output missing");
        const sdk_field &f = fields[k], &f0 = fields[0];
        if (f.has_time != f0.has_time || f.nt != f0.nt || f.it0 != f0.it0 || f.has_z != f0.has_z || f.nz != f0.nz || f.ny != f0.ny ||
            f.nx != f0.nx || f.xstag != f0.xstag || f.ystag != f0.ystag || f.mask != f0.mask)
            return fail(SDK_EINVAL, "sdk_interp_scalar_multi: the fields must share shape, staggering and mask");
        a.f[k] = f;
        a.out[k] = outs[k];
    }
    a.g = grid->dev;
    a.pts = *pts;
    a.k = *knw;
    a.n_fields = n_fields;
    cudaError_t e = sdk::launch_interp_multi(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_interp_scalar_multi: launch");
    return SDK_OK;
}

int sdk_interp_vector(const sdk_grid *grid, const sdk_points *pts, const sdk_field *fu, const sdk_field *fv,
                      const sdk_kernel_desc *ku, const sdk_kernel_desc *kv, int vec_transform, double *out_u,
                      double *out_v, void *stream) {
    int rc = check_interp(grid, pts, fu, ku);
    if (rc) return rc;
    rc = check_interp(grid, pts, fv, kv);
    if (rc) return rc;
    if (!out_u || !out_v) return fail(SDK_EINVAL, "sdk_interp_vector: output missing");
    if (vec_transform && pts->n > 0 && (!pts->cs || !pts->sn)) return fail(SDK_EINVAL, "sdk_interp_vector: cs / sn missing");
    sdk::InterpArgs a;
    a.g = grid->dev;
    a.pts = *pts;
    a.fu = *fu;
    a.fv = *fv;
    a.ku = *ku;
    a.kv = *kv;
    a.vec_transform = vec_transform;
    a.out_u = out_u;
    a.out_v = out_v;
    cudaError_t e = sdk::launch_interp(a, true, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_interp_vector: launch");
    return SDK_OK;
}

int sdk_build_masks(const sdk_grid *grid, const uint8_t *maskC, int32_t nz, uint8_t *maskU, uint8_t *maskV,
                    uint8_t *maskW, void *stream) {
    if (!grid || !maskC || !maskU || !maskV || !maskW || nz < 1) return fail(SDK_EINVAL, "sdk_build_masks: bad argument");
    cudaError_t e = sdk::launch_build_masks(grid->dev, maskC, nz, maskU, maskV, maskW, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_build_masks: launch");
    return SDK_OK;
}

// ---- device math self-test --------------------------------------------------------------------------
int sdk_log_budget(const sdk_grid *grid, const sdk_log_record *records, const int64_t *offset, int64_t n_particles,
                   int64_t n_rec, int16_t *ind1, int16_t *ind2, double *frac, double *tres, double *work, void *stream) {
    if (!grid || n_particles < 0 || n_rec < 0) return fail(SDK_EINVAL, "sdk_log_budget: bad argument");
    if (n_rec > 0 && (!records || !offset || !ind1 || !ind2 || !frac || !tres || !work))
        return fail(SDK_EINVAL, "sdk_log_budget: null array");
    if (grid->dev.n_zl < 2) return fail(SDK_EUNSUPPORTED, "sdk_log_budget: the grid has no vertical levels");
    sdk::BudgetArgs a;
    a.g = grid->dev;
    a.rec = records;
    a.offset = offset;
    a.n_particles = n_particles;
    a.n_rec = n_rec;
    a.ind1 = ind1;
    a.ind2 = ind2;
    a.frac = frac;
    a.tres = tres;
    a.work = work;
    a.rows = grid->dev.has_face ? 5 : 4;
    cudaError_t e = sdk::launch_budget(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_log_budget: launch");
    return SDK_OK;
}

int sdk_log_calculate_budget(const sdk_log_record *records, const int64_t *offset, int64_t n_particles, int64_t n_rec,
                             const int16_t *ind1, const int16_t *ind2, const double *frac, const double *tres,
                             const sdk_budget_fields *fields, double *conc, double *contr, unsigned char *work,
                             void *stream) {
    if (!fields || n_particles < 0 || n_rec < 0) return fail(SDK_EINVAL, "sdk_log_calculate_budget: bad argument");
    if (fields->n_terms < 1 || fields->n_terms > SDK_MAX_BUDGET_TERMS)
        return fail(SDK_EINVAL, "sdk_log_calculate_budget: between 1 and SDK_MAX_BUDGET_TERMS right-hand-side terms");
    if (fields->rows != 4 && fields->rows != 5) return fail(SDK_EINVAL, "sdk_log_calculate_budget: rows must be 4 or 5");
    if (fields->dir_of_time != 1 && fields->dir_of_time != -1)
        return fail(SDK_EINVAL, "sdk_log_calculate_budget: dir_of_time must be +1 or -1");
    for (int d = 0; d < 4; ++d)
        if (fields->cell_shape[d] < 1 || fields->wall_shape[d] < 1)
            return fail(SDK_EINVAL, "sdk_log_calculate_budget: empty field");
    if (n_rec > 0) {
        if (!records || !offset || !ind1 || !ind2 || !frac || !tres || !conc || !work || !fields->lhs || !fields->walls)
            return fail(SDK_EINVAL, "sdk_log_calculate_budget: null array");
        if (n_rec > 1 && !contr) return fail(SDK_EINVAL, "sdk_log_calculate_budget: null array");
        for (int k = 0; k < fields->n_terms; ++k)
            if (!fields->term[k]) return fail(SDK_EINVAL, "sdk_log_calculate_budget: null term");
    }
    sdk::BudgetTermArgs a;
    a.rec = records;
    a.offset = offset;
    a.n_particles = n_particles;
    a.n_rec = n_rec;
    a.ind1 = ind1;
    a.ind2 = ind2;
    a.frac = frac;
    a.tres = tres;
    a.f = *fields;
    a.conc = conc;
    a.contr = contr;
    a.is_last = work;
    cudaError_t e = sdk::launch_budget_terms(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_log_calculate_budget: launch");
    return SDK_OK;
}

int sdk_log_wall_values(const sdk_grid *grid, const sdk_log_record *records, int64_t n_rec, const sdk_wall_fields *fields,
                        double *c_list, double *u_list, double *lag_conv, double *eul_conv, void *stream) {
    if (!grid || !fields || n_rec < 0) return fail(SDK_EINVAL, "sdk_log_wall_values: bad argument");
    if (!grid->dev.has_face) return fail(SDK_EUNSUPPORTED, "sdk_log_wall_values: only grids with faces (like the reference's tested branch)");
    if (n_rec > 0 && (!records || !c_list || !fields->x_walls || !fields->y_walls || !fields->z_walls))
        return fail(SDK_EINVAL, "sdk_log_wall_values: null array");
    if (eul_conv && !fields->conv) return fail(SDK_EINVAL, "sdk_log_wall_values: eul_conv asked for without a convergence field");
    for (int d = 0; d < 4; ++d)
        if (fields->shape_x[d] < 1 || fields->shape_y[d] < 1 || fields->shape_z[d] < 1 || (fields->conv && fields->shape_c[d] < 1))
            return fail(SDK_EINVAL, "sdk_log_wall_values: empty field");
    sdk::WallValueArgs a;
    a.g = grid->dev;
    a.rec = records;
    a.n_rec = n_rec;
    a.f = *fields;
    a.c_list = c_list;
    a.u_list = u_list;
    a.lag_conv = lag_conv;
    a.eul_conv = eul_conv;
    cudaError_t e = sdk::launch_wall_values(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_log_wall_values: launch");
    return SDK_OK;
}

int sdk_gather_plane(const double *src, int64_t n_outer, int64_t src_plane, const int64_t *map, int64_t n_map,
                     double *out, void *stream) {
    if (n_outer < 0 || n_map < 0 || src_plane < 0) return fail(SDK_EINVAL, "sdk_gather_plane: negative size");
    if (n_outer * n_map > 0 && (!src || !map || !out)) return fail(SDK_EINVAL, "sdk_gather_plane: null array");
    cudaError_t e = sdk::launch_gather_plane(src, n_outer, src_plane, map, n_map, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_gather_plane: launch");
    return SDK_OK;
}

int sdk_wall_stencil(int scheme, const double *buf, const double *vel, double *out, int64_t n_outer, int64_t n_out,
                     int64_t n_inner, void *stream) {
    if (scheme < SDK_STENCIL_DST3 || scheme > SDK_STENCIL_LIMITER_Z) return fail(SDK_EINVAL, "sdk_wall_stencil: unknown scheme");
    if (n_outer < 0 || n_out < 0 || n_inner < 0) return fail(SDK_EINVAL, "sdk_wall_stencil: negative size");
    if (n_outer * n_out * n_inner > 0 && (!buf || !vel || !out)) return fail(SDK_EINVAL, "sdk_wall_stencil: null array");
    cudaError_t e = sdk::launch_wall_stencil(scheme, buf, vel, out, n_outer, n_out, n_inner, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_wall_stencil: launch");
    return SDK_OK;
}

int sdk_upwind3_z(const double *s, double *w, double *out, int64_t nz, int64_t n_inner, void *stream) {
    if (nz < 0 || n_inner < 0) return fail(SDK_EINVAL, "sdk_upwind3_z: negative size");
    if (nz * n_inner > 0 && (!s || !w || !out)) return fail(SDK_EINVAL, "sdk_upwind3_z: null array");
    cudaError_t e = sdk::launch_upwind3_z(s, w, out, nz, n_inner, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_upwind3_z: launch");
    return SDK_OK;
}

int sdk_pack_records(const sdk_particles *p, const int32_t *out_index, sdk_snapshot_record *out, void *stream) {
    if (!p || p->n < 0) return fail(SDK_EINVAL, "sdk_pack_records: bad argument");
    if (p->n > 0 && (!out || !p->lon || !p->lat || !p->iy || !p->ix)) return fail(SDK_EINVAL, "sdk_pack_records: null array");
    cudaError_t e = sdk::launch_pack_records(*p, p->n, out_index, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_pack_records: launch");
    return SDK_OK;
}

int64_t sdk_cell_sort_scratch_bytes(int64_t n) { return n < 0 ? 0 : (int64_t)sdk::cell_sort_scratch_bytes(n); }

int sdk_cell_sort(const sdk_grid *grid, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix,
                  const int32_t *izl, int32_t *perm, void *scratch, int64_t scratch_bytes, void *stream) {
    if (!grid || n < 0 || n > INT32_MAX) return fail(SDK_EINVAL, "sdk_cell_sort: bad argument");
    if (n > 0 && (!iy || !ix || !perm || !scratch)) return fail(SDK_EINVAL, "sdk_cell_sort: null array");
    if (scratch_bytes < sdk_cell_sort_scratch_bytes(n)) return fail(SDK_EINVAL, "sdk_cell_sort: scratch too small");
    cudaError_t e = sdk::launch_cell_sort(grid->dev, n, face, iy, ix, izl, perm, scratch, (size_t)scratch_bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_cell_sort");
    return SDK_OK;
}

int sdk_particles_gather(const sdk_particles *src, const sdk_particles *dst, const int32_t *perm, void *stream) {
    if (!src || !dst || dst->n < 0 || src->n < 0) return fail(SDK_EINVAL, "sdk_particles_gather: bad argument");
    if (dst->n > 0 && !perm) return fail(SDK_EINVAL, "sdk_particles_gather: permutation missing");
    cudaError_t e = sdk::launch_particles_gather(*src, *dst, perm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_particles_gather: launch");
    return SDK_OK;
}

static int check_exchange(const sdk_exchange *x, const char *who) {
    if (!x) return fail(SDK_EINVAL, std::string(who) + ": null exchange");
    if (x->world < 1 || x->world > SDK_MAX_RANKS || x->rank < 0 || x->rank >= x->world || x->n_slots < 1 || x->n_pad < 0)
        return fail(SDK_EINVAL, std::string(who) + ": bad geometry");
    for (int r = 0; r < x->world; ++r)
        if (!x->signal[r] || (!x->recv[r] && !x->recv_multicast)) return fail(SDK_EINVAL, std::string(who) + ": unmapped rank");
    if (!x->ticket) return fail(SDK_EINVAL, std::string(who) + ": ticket missing");
    return SDK_OK;
}

int sdk_pack_ship(const sdk_exchange *x, const sdk_particles *p, const int32_t *out_index, uint64_t step, void *stream) {
    int rc = check_exchange(x, "sdk_pack_ship");
    if (rc) return rc;
    if (!p || p->n < 0 || p->n > x->n_pad) return fail(SDK_EINVAL, "sdk_pack_ship: particle count does not fit n_pad");
    if (step < 1) return fail(SDK_EINVAL, "sdk_pack_ship: steps count from 1");
    if (p->n > 0 && (!p->lon || !p->lat || !p->iy || !p->ix)) return fail(SDK_EINVAL, "sdk_pack_ship: null array");
    const int sms = sdk_sm_count();
    cudaError_t e = sdk::launch_pack_ship(*x, *p, p->n, out_index, step, sms > 0 ? sms : 1, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_pack_ship: launch");
    return SDK_OK;
}

int sdk_exchange_wait(const sdk_exchange *x, uint64_t step, int wait, int release, void *stream) {
    int rc = check_exchange(x, "sdk_exchange_wait");
    if (rc) return rc;
    cudaError_t e = sdk::launch_exchange_wait(*x, step, wait, release, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_exchange_wait: launch");
    return SDK_OK;
}

int64_t sdk_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(sdk_grid_desc);
        case 1: return sizeof(sdk_velocity);
        case 2: return sizeof(sdk_particles);
        case 3: return sizeof(sdk_advect_params);
        case 4: return sizeof(sdk_log);
        case 5: return sizeof(sdk_located);
        case 6: return sizeof(sdk_track_io);
        case 7: return sizeof(sdk_kernel_desc);
        case 8: return sizeof(sdk_field);
        case 9: return sizeof(sdk_points);
        case 10: return sizeof(sdk_budget_fields);
        case 11: return sizeof(sdk_wall_fields);
        case 12: return sizeof(sdk_snapshot_record);
        case 13: return sizeof(sdk_exchange);
        case 14: return sizeof(sdk_interp_job);
        case 15: return sizeof(sdk_interp_io);
        default: return -1;
    }
}

int sdk_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad, void *stream) {
    if (op < SDK_SELFTEST_DIV || op > SDK_SELFTEST_STEP) return fail(SDK_EINVAL, "sdk_selftest_math: unknown op");
    if (n < 0 || (n > 0 && (!a || !out || (!bad && op != SDK_SELFTEST_STEP)))) return fail(SDK_EINVAL, "sdk_selftest_math: bad argument");
    if ((op == SDK_SELFTEST_DIV || op == SDK_SELFTEST_DIV_IEEE) && n > 0 && !b) return fail(SDK_EINVAL, "sdk_selftest_math: b missing");
    cudaError_t e = sdk::launch_selftest_math(op, n, a, b ? b : a, out, bad, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_selftest_math: launch");
    return SDK_OK;
}

int sdk_selftest_cell_move(const sdk_grid *grid, int64_t n, const int32_t *in, int32_t *out, void *stream) {
    if (!grid || n < 0 || (n > 0 && (!in || !out))) return fail(SDK_EINVAL, "sdk_selftest_cell_move: bad argument");
    cudaError_t e = sdk::launch_selftest_cell_move(grid->dev, n, in, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_selftest_cell_move: launch");
    return SDK_OK;
}

// the three internal streams (+ events) the *_host entry points pipeline their chunks over
static int ensure_pipe(const sdk_grid *grid) {
    if (grid->has_pipe) return SDK_OK;
    cudaError_t e;
    for (auto &st : grid->pipe)
        if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "internal stream");
    for (auto &ev : grid->pipe_ev)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "internal event");
    grid->has_pipe = true;
    return SDK_OK;
}

// ---- positions in, positions out ------------------------------------------------------------------
// per chunk: the particle state, the coordinates as they came in (sorted runs), the permutation, the sort's workspace,
// the records
static size_t track_chunk_bytes(int64_t n) {
    return (size_t)sdk_particles_bytes(n) + 4 * align256((size_t)n * 8) + align256((size_t)n * 4) +
           align256((size_t)sdk_cell_sort_scratch_bytes(n)) + align256((size_t)n * sizeof(sdk_snapshot_record));
}

// head of the scratch buffer: the call's counters, then signal words + ticket of every chunk's record output
constexpr size_t kTrackSignalBytes = 512;
constexpr size_t kTrackHeadBytes = 256 + (size_t)kTrackChunksMax * kTrackSignalBytes;

int64_t sdk_track_scratch_bytes(int64_t n) {
    // chunks are rounded up to 32 particles
    return (int64_t)(track_chunk_bytes(n + 32 * kTrackChunksMax) + (size_t)kTrackChunksMax * 80 * 256 + kTrackHeadBytes);
}

static void carve_particles(sdk_particles &dp, int64_t n, char *&cur) {
    std::memset(&dp, 0, sizeof dp);
    dp.n = n;
    for (const Field &f : kFields) {
        if (f.offset == offsetof(sdk_particles, active)) continue;
        void **ddst = (void **)((char *)&dp + f.offset);
        *ddst = cur;
        cur += align256((size_t)n * f.elem * f.mult);
    }
}

// SDK_TRACK_TIMELINE=1: CUDA events between the stages of every chunk, printed (ms since the start of the call) when
// the call has synchronised.  A measuring aid (tools/exp_timeline.py); costs nothing when the variable is not set.
namespace {
struct TrackTimeline {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    struct Mark { int chunk; const char *what; cudaEvent_t ev; };
    std::vector<Mark> marks;
    void begin(cudaStream_t s) {
        on = std::getenv("SDK_TRACK_TIMELINE") != nullptr;
        if (!on) return;
        cudaEventCreate(&t0);
        cudaEventRecord(t0, s);
    }
    void mark(int chunk, const char *what, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        marks.push_back({chunk, what, ev});
    }
    void end() {
        if (!on) return;
        for (const Mark &m : marks) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t0, m.ev);
            std::fprintf(stderr, "[track timeline] chunk %d %-10s %.3f ms\n", m.chunk, m.what, ms);
            cudaEventDestroy(m.ev);
        }
        cudaEventDestroy(t0);
        marks.clear();
        on = false;
    }
};
thread_local TrackTimeline g_timeline;
thread_local int g_timeline_chunk = 0;
}  // namespace
#define TL(what) g_timeline.mark(g_timeline_chunk, what, stream)

// One chunk of sdk_track_host: H2D, locate, (cell sort,) velocity read, advect, pack, D2H -- all asynchronous on `stream`.
// `records_dev`: device-visible address of io->out_records when that is mapped pinned host memory (the advection kernel
// then stores the records there itself, over PCIe, while it runs), else NULL (pack on the device, copy afterwards);
// `signal`: kTrackSignalBytes of zeroed device memory of this chunk.
static int track_chunk(const sdk_grid *grid, const sdk_velocity *vel, const sdk_track_io *io, const double *ts_dev,
                       int32_t n_t, double t_stop, const sdk_advect_params *par_in, uint64_t *stats_dev, void *scratch,
                       sdk_snapshot_record *records_dev, void *signal, cudaStream_t stream) {
    const int64_t n = io->n;
    sdk_advect_params par;
    if (par_in) par = *par_in;
    else {
        std::memset(&par, 0, sizeof par);
        par.tol = 1e-4, par.trim_tol = 1e-14, par.max_iteration = 200;
    }
    const int has_dep = par.has_dep;
    char *cur = (char *)scratch;
    sdk_particles dp;
    carve_particles(dp, n, cur);
    cudaError_t e;
    const bool vertical = has_dep && grid->dev.n_zl > 1;
    const bool timed = vel->has_time && ts_dev && n_t > 1;
    // with `sort` the coordinates land in a staging area first and are put into (level, bucket) order on their way
    // into the state: everything after that -- locating included -- works on coherent particles
    const bool sorted = io->sort && grid->dev.hmode != SDK_H_RECTILINEAR && grid->has_buckets;
    double *in[4] = {dp.lon, dp.lat, dp.dep, dp.t};
    if (sorted) {
        for (auto &a : in) {
            a = (double *)cur;
            cur += align256((size_t)n * 8);
        }
    }
#define H2D(dst, src) \
    if ((e = cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) \
        return cuda_fail(e, "sdk_track_host: H2D");
    TL("start");
    H2D(in[0], io->lon)
    H2D(in[1], io->lat)
    H2D(in[3], io->t)
    if (io->dep) H2D(in[2], io->dep)
#undef H2D
    TL("h2d");
    int32_t *perm = nullptr;
    if (sorted) {
        perm = (int32_t *)cur;
        cur += align256((size_t)n * 4);
        const size_t sbytes = (size_t)sdk_cell_sort_scratch_bytes(n);
        if ((e = sdk::launch_point_sort(grid->dev, grid->buckets, n, in[0], in[1], (vertical && io->dep) ? in[2] : nullptr, perm, cur, sbytes,
                                        stream)) != cudaSuccess)
            return cuda_fail(e, "sdk_track_host: sort");
        cur += align256(sbytes);
        if ((e = sdk::launch_points_gather(n, perm, in[0], in[1], io->dep ? in[2] : nullptr, in[3], dp.lon, dp.lat, dp.dep, dp.t, stream)) !=
            cudaSuccess)
            return cuda_fail(e, "sdk_track_host: gather");
        TL("sort");
    }
    // Position.from_latlon: indices + rel-coords straight into the particle state
    sdk_located loc;
    std::memset(&loc, 0, sizeof loc);
    loc.face = grid->dev.has_face ? dp.face : nullptr;
    loc.iy = dp.iy; loc.ix = dp.ix; loc.rx = dp.rx; loc.ry = dp.ry;
    loc.px = dp.px; loc.py = dp.py; loc.dx = dp.dx; loc.dy = dp.dy;
    if (vertical) {
        loc.izl = dp.izl; loc.rzl = dp.rzl; loc.dzl = dp.dzl; loc.bzl = dp.bzl;
    } else {
        dp.izl = nullptr;  // no vertical index: reads as level 0 everywhere
    }
    loc.it = dp.it;
    loc.status = dp.status;
    if (!timed && (e = cudaMemsetAsync(dp.it, 0, (size_t)n * 4, stream)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: memset");
    int rc = sdk_locate(grid, n, dp.lon, dp.lat, vertical ? dp.dep : nullptr, timed ? dp.t : nullptr, ts_dev, n_t, &loc, stream);
    if (rc) return rc;
    TL("locate");
    const sdk_particles *work = &dp;
    // Particle.__init__: get_vol + get_u_du, then to_next_stop
    rc = sdk_get_u_du(grid, vel, work, has_dep, stream);
    if (rc) return rc;
    par.keep_status = 1;  // what sdk_locate flagged stays flagged
    par.stats = stats_dev;
    sdk_exchange to_host;
    if (perm) records_dev = nullptr;  // (scattered 32-byte stores over PCIe are slow: sorted runs pack on the device and copy)
    if (records_dev) {
        // the fused output exchange with the host as the only "rank": every particle's record goes to the caller's
        // pinned buffer (in the caller's order) from the write-back of the advection kernel -- no pack pass, no copy
        std::memset(&to_host, 0, sizeof to_host);
        to_host.world = 1;
        to_host.n_slots = 1;
        to_host.n_pad = n;
        to_host.recv[0] = records_dev;
        to_host.signal[0] = (uint64_t *)signal;
        to_host.ticket = (uint32_t *)((char *)signal + SDK_SIGNAL_WORDS * sizeof(uint64_t));
        par.ship = &to_host;
        par.ship_index = nullptr;
        par.ship_step = 1;
    }
    TL("get_u_du");
    rc = sdk_advect(grid, vel, work, t_stop, &par, nullptr, 1, stream);
    if (rc) return rc;
    TL("advect");
    if (io->out_records && !records_dev) {
        sdk_snapshot_record *rec = (sdk_snapshot_record *)cur;
        rc = sdk_pack_records(work, perm, rec, stream);
        if (rc) return rc;
        if ((e = cudaMemcpyAsync(io->out_records, rec, (size_t)n * sizeof(sdk_snapshot_record), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
            return cuda_fail(e, "sdk_track_host: D2H");
    }
#define D2H(dst, src, esz) \
    if (dst && src && (e = cudaMemcpyAsync(dst, src, (size_t)n * esz, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) \
        return cuda_fail(e, "sdk_track_host: D2H");
    D2H(io->out_lon, dp.lon, 8)
    D2H(io->out_lat, dp.lat, 8)
    if (has_dep) D2H(io->out_dep, dp.dep, 8)
    if (grid->dev.has_face) D2H(io->out_face, dp.face, 4)
    D2H(io->out_iy, dp.iy, 4)
    D2H(io->out_ix, dp.ix, 4)
    if (vertical) D2H(io->out_izl, dp.izl, 4)
    D2H(io->out_ncross, dp.ncross, 4)
    D2H(io->out_status, dp.status, 4)
#undef D2H
    TL("d2h");
    return SDK_OK;
}

int sdk_track_host(const sdk_grid *grid, const sdk_velocity *vel, const sdk_track_io *io, const double *ts_dev,
                   int32_t n_t, double t_stop, const sdk_advect_params *par, void *scratch, void *stream_) {
    if (!grid || !vel || !io || !scratch) return fail(SDK_EINVAL, "sdk_track_host: null argument");
    if (!io->lon || !io->lat || !io->t) return fail(SDK_EINVAL, "sdk_track_host: lon / lat / t are required");
    if (!io->out_records && (!io->out_lon || !io->out_lat)) return fail(SDK_EINVAL, "sdk_track_host: no output requested");
    if (io->sort && (io->out_lon || io->out_lat || io->out_dep || io->out_face || io->out_iy || io->out_ix || io->out_izl ||
                     io->out_ncross || io->out_status))
        return fail(SDK_EINVAL, "sdk_track_host: with sort the results come as records (out_records) and counters (out_stats) only");
    const int64_t n = io->n;
    if (n <= 0) return n == 0 ? SDK_OK : fail(SDK_EINVAL, "sdk_track_host: negative count");
    const int has_dep = par ? par->has_dep : 0;
    if (has_dep && !io->dep) return fail(SDK_EINVAL, "sdk_track_host: depth array missing");
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e;
    // the call's counters and the chunks' signal words: head of the scratch buffer
    uint64_t *stats_dev = (uint64_t *)scratch;
    char *signals = (char *)scratch + 256;
    char *cur = (char *)scratch + kTrackHeadBytes;
    if ((e = cudaMemsetAsync(scratch, 0, kTrackHeadBytes, stream)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: memset");
    g_timeline.begin(stream);
    g_timeline_chunk = 0;
    // records straight into the caller's buffer if the device can address it (pinned, mapped host memory)
    sdk_snapshot_record *records_dev = nullptr;
    if (io->out_records && !std::getenv("SDK_TRACK_NO_ZEROCOPY")) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, io->out_records) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            records_dev = (sdk_snapshot_record *)attr.devicePointer;
        else
            cudaGetLastError();  // (pageable memory: not an error, the copy path takes it)
    }
    // Small batches: one pass.  Large ones: chunks on three internal streams, so that the upload of
    // chunk k+1 and the read-back of chunk k-1 (copy engines) overlap the kernels of chunk k.
    const int64_t min_chunk = 65536;
    int nchunks = (int)(n / min_chunk);
    int max_chunks = kTrackChunks;
    if (const char *env = std::getenv("SDK_TRACK_CHUNKS")) {  // tuning knob, 1 .. kTrackChunksMax
        const int v = std::atoi(env);
        if (v >= 1 && v <= kTrackChunksMax) max_chunks = v;
    }
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 2) {
        int rc = track_chunk(grid, vel, io, ts_dev, n_t, t_stop, par, stats_dev, cur, records_dev, signals, stream);
        if (rc) return rc;
    } else {
        if (int rc = ensure_pipe(grid)) return rc;
        // the internal streams start after whatever the caller queued on `stream`
        if ((e = cudaEventRecord(grid->pipe_ev[3], stream)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: event record");
        for (auto &st : grid->pipe)
            if ((e = cudaStreamWaitEvent(st, grid->pipe_ev[3], 0)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: wait");
        // Chunk sizes grow 1 : 2 : 3 ...: the upload of the first chunk is the only one nothing overlaps, so it is the
        // smallest (SDK_TRACK_RAMP=0: equal chunks).
        const char *ramp_env = std::getenv("SDK_TRACK_RAMP");
        const bool ramp = !(ramp_env && ramp_env[0] == '0');
        const int64_t weight_sum = ramp ? (int64_t)nchunks * (nchunks + 1) / 2 : nchunks;
        int64_t off = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int64_t upto = ramp ? (int64_t)(c + 1) * (c + 2) / 2 : c + 1;
            int64_t end = c == nchunks - 1 ? n : (n * upto / weight_sum + 31) / 32 * 32;
            if (end > n) end = n;
            const int64_t cnt = end - off;
            if (cnt <= 0) continue;
            sdk_track_io sub = *io;
            sub.n = cnt;
#define SHIFT(member) \
    if (sub.member) sub.member = sub.member + off;
            SHIFT(lon) SHIFT(lat) SHIFT(dep) SHIFT(t) SHIFT(out_records) SHIFT(out_lon) SHIFT(out_lat) SHIFT(out_dep)
            SHIFT(out_face) SHIFT(out_iy) SHIFT(out_ix) SHIFT(out_izl) SHIFT(out_ncross) SHIFT(out_status)
#undef SHIFT
            g_timeline_chunk = c;
            int rc = track_chunk(grid, vel, &sub, ts_dev, n_t, t_stop, par, stats_dev, cur, records_dev ? records_dev + off : nullptr,
                                 signals + (size_t)c * kTrackSignalBytes, grid->pipe[c % 3]);
            if (rc) return rc;
            cur += track_chunk_bytes(cnt);
            off = end;
        }
        for (int k = 0; k < 3; ++k) {
            if ((e = cudaEventRecord(grid->pipe_ev[k], grid->pipe[k])) != cudaSuccess) return cuda_fail(e, "sdk_track_host: event record");
            if ((e = cudaStreamWaitEvent(stream, grid->pipe_ev[k], 0)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: wait");
        }
    }
    if (io->out_stats && (e = cudaMemcpyAsync(io->out_stats, stats_dev, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        return cuda_fail(e, "sdk_track_host: D2H");
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, "sdk_track_host: sync");
    g_timeline.end();
    return SDK_OK;
}

// ---- coordinates in, interpolated values out ------------------------------------------------------
namespace {
constexpr int kInterpChunksDefault = 8;
constexpr int kInterpChunksMax = 64;
constexpr int64_t kInterpMinChunk = 262144;
constexpr size_t kInterpHeadBytes = 256;

// device memory of one chunk in flight (there are three: one per internal stream)
size_t interp_slot_bytes(int64_t c, int n_outputs) {
    const size_t d = align256((size_t)c * 8), w = align256((size_t)c * 4);
    return 8 * d                                                            // coordinates as they came in + in sorted order
           + w + align256((size_t)sdk_cell_sort_scratch_bytes(c))           // permutation, sort workspace
           + 3 * w + 2 * d + 2 * w                                          // face iy ix, rx ry, cs sn
           + 6 * (w + 3 * d)                                                // six index families of sdk_locate
           + w                                                              // status
           + (size_t)n_outputs * d;
}

int64_t interp_chunk_count(int64_t n, int max_chunks) {
    int64_t k = n / kInterpMinChunk;
    if (k > max_chunks) k = max_chunks;
    return k < 1 ? 1 : k;
}
}  // namespace

int64_t sdk_interp_host_scratch_bytes(int64_t n, int32_t n_outputs) {
    if (n < 0 || n_outputs < 0) return 0;
    // the largest chunk any chunk count can produce: n itself (one chunk) or n / k rounded up to 32
    int64_t c = n;
    if (n >= 2 * kInterpMinChunk) c = (n / 2 + 32);
    return (int64_t)(kInterpHeadBytes + 3 * interp_slot_bytes(c, n_outputs));
}

int sdk_status_or(const int32_t *status, int64_t n, uint32_t *flags, void *stream) {
    if (n < 0 || !flags || (n > 0 && !status)) return fail(SDK_EINVAL, "sdk_status_or: bad argument");
    cudaError_t e = sdk::launch_status_or(status, n, flags, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "sdk_status_or: launch");
    return SDK_OK;
}

// One chunk of sdk_interp_host, asynchronous on `stream`: H2D, sort, locate, one launch per job, D2H.
static int interp_chunk(const sdk_grid *grid, const sdk_interp_io *io, int64_t off, int64_t n, const double *ts_dev, int32_t n_t,
                        const sdk_interp_job *jobs, int32_t n_jobs, uint32_t *flags_dev, char *cur, cudaStream_t stream) {
    cudaError_t e;
    const bool timed = io->t && ts_dev && n_t > 1;
    const bool sorted = io->sort && grid->dev.hmode != SDK_H_RECTILINEAR && grid->has_buckets;
    auto take = [&cur](size_t bytes) {
        char *p = cur;
        cur += align256(bytes);
        return p;
    };
    double *in[4], *pt[4];
    for (auto &a : in) a = (double *)take((size_t)n * 8);
    if (sorted)
        for (auto &a : pt) a = (double *)take((size_t)n * 8);
    else
        for (int k = 0; k < 4; ++k) pt[k] = in[k];
    const double *host[4] = {io->lon, io->lat, io->dep, io->t};
    for (int k = 0; k < 4; ++k)
        if (host[k] && (e = cudaMemcpyAsync(in[k], host[k] + off, (size_t)n * 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess)
            return cuda_fail(e, "sdk_interp_host: H2D");
    // which index families the jobs ask for (the linear Zl lookup always runs when there is a depth: it is the one whose
    // "Value out of bound." Position.from_latlon raises, eulerian.py:199)
    bool want_z[5] = {false, false, false, false, false}, want_t[3] = {false, false, false};
    bool want_cs = false;
    for (int j = 0; j < n_jobs; ++j) {
        want_z[jobs[j].zsel] = true;
        want_t[jobs[j].tsel] = true;
        want_cs |= jobs[j].kind == SDK_JOB_VECTOR && jobs[j].vec_transform;
    }
    const bool vertical = io->dep != nullptr;
    if (vertical && grid->dev.n_zl > 1) want_z[SDK_SEL_ZL_LIN] = true;
    int32_t *perm = nullptr;
    if (sorted) {
        perm = (int32_t *)take((size_t)n * 4);
        const size_t sbytes = (size_t)sdk_cell_sort_scratch_bytes(n);
        char *sort_ws = take(sbytes);
        if ((e = sdk::launch_point_sort(grid->dev, grid->buckets, n, in[0], in[1], (vertical && grid->dev.n_zl > 1) ? in[2] : nullptr, perm,
                                        sort_ws, sbytes, stream)) != cudaSuccess)
            return cuda_fail(e, "sdk_interp_host: sort");
        if ((e = sdk::launch_points_gather(n, perm, in[0], in[1], io->dep ? in[2] : nullptr, io->t ? in[3] : nullptr, pt[0], pt[1], pt[2], pt[3],
                                           stream)) != cudaSuccess)
            return cuda_fail(e, "sdk_interp_host: gather");
    }
    sdk_located loc;
    std::memset(&loc, 0, sizeof loc);
    if (grid->dev.has_face) loc.face = (int32_t *)take((size_t)n * 4);
    loc.iy = (int32_t *)take((size_t)n * 4);
    loc.ix = (int32_t *)take((size_t)n * 4);
    loc.rx = (double *)take((size_t)n * 8);
    loc.ry = (double *)take((size_t)n * 8);
    if (want_cs) {
        loc.cs = (float *)take((size_t)n * 4);
        loc.sn = (float *)take((size_t)n * 4);
    }
    struct Family {
        int32_t **k;
        double **r, **d, **b;
    };
    const Family zfam[5] = {{nullptr, nullptr, nullptr, nullptr},
                            {&loc.iz, &loc.rz, &loc.dz, &loc.bz},
                            {&loc.iz_lin, &loc.rz_lin, &loc.dz_lin, &loc.bz_lin},
                            {&loc.izl_near, &loc.rzl_near, &loc.dzl_near, &loc.bzl_near},
                            {&loc.izl, &loc.rzl, &loc.dzl, &loc.bzl}};
    const Family tfam[3] = {{nullptr, nullptr, nullptr, nullptr}, {&loc.it, &loc.rt, &loc.dt, &loc.bt}, {&loc.it_lin, &loc.rt_lin, &loc.dt_lin, &loc.bt_lin}};
    auto carve = [&](const Family &f) {
        *f.k = (int32_t *)take((size_t)n * 4);
        *f.r = (double *)take((size_t)n * 8);
        *f.d = (double *)take((size_t)n * 8);
        *f.b = (double *)take((size_t)n * 8);
    };
    for (int s = 1; s < 5; ++s)
        if (want_z[s] && vertical) carve(zfam[s]);
    for (int s = 1; s < 3; ++s)
        if (want_t[s] && timed) carve(tfam[s]);
    loc.status = (int32_t *)take((size_t)n * 4);
    int rc = sdk_locate(grid, n, pt[0], pt[1], vertical ? pt[2] : nullptr, timed ? pt[3] : nullptr, ts_dev, n_t, &loc, stream);
    if (rc) return rc;
    if ((e = sdk::launch_status_or(loc.status, n, flags_dev, stream)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: status");
    for (int j = 0; j < n_jobs; ++j) {
        const sdk_interp_job &job = jobs[j];
        sdk_points p;
        std::memset(&p, 0, sizeof p);
        p.n = n;
        p.face = loc.face; p.iy = loc.iy; p.ix = loc.ix; p.rx = loc.rx; p.ry = loc.ry;
        if (job.zsel) {
            if (!vertical) return fail(SDK_EINVAL, "sdk_interp_host: the variable has a vertical dimension but the points have no depth");
            p.kz = *zfam[job.zsel].k;
            p.rz = *zfam[job.zsel].r;
        }
        if (job.tsel) {
            if (!timed) return fail(SDK_EINVAL, "sdk_interp_host: the variable has a time dimension but the points have no time");
            p.kt = *tfam[job.tsel].k;
            p.rt = *tfam[job.tsel].r;
        }
        p.cs = loc.cs; p.sn = loc.sn;
        p.out_index = perm;
        double *out[SDK_MAX_MULTI_FIELDS];
        for (int k = 0; k < job.n_fields; ++k) out[k] = (double *)take((size_t)n * 8);
        if (job.kind == SDK_JOB_VECTOR)
            rc = sdk_interp_vector(grid, &p, &job.fields[0], &job.fields[1], &job.knw[0], &job.knw[1], job.vec_transform, out[0], out[1], stream);
        else if (job.n_fields == 1)
            rc = sdk_interp_scalar(grid, &p, &job.fields[0], &job.knw[0], out[0], stream);
        else
            rc = sdk_interp_scalar_multi(grid, &p, job.fields, job.n_fields, &job.knw[0], out, stream);
        if (rc) return rc;
        for (int k = 0; k < job.n_fields; ++k)
            if ((e = cudaMemcpyAsync(job.out[k] + off, out[k], (size_t)n * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
                return cuda_fail(e, "sdk_interp_host: D2H");
    }
    return SDK_OK;
}

int sdk_interp_host(const sdk_grid *grid, const sdk_interp_io *io, const double *ts_dev, int32_t n_t, const sdk_interp_job *jobs,
                    int32_t n_jobs, void *scratch, void *stream_) {
    if (!grid || !io || !scratch || (n_jobs > 0 && !jobs) || n_jobs < 0) return fail(SDK_EINVAL, "sdk_interp_host: null argument");
    if (!io->lon || !io->lat) return fail(SDK_EINVAL, "sdk_interp_host: lon / lat are required");
    if (grid->dev.hmode != SDK_H_RECTILINEAR && !grid->has_buckets) return fail(SDK_EINVAL, "sdk_interp_host: call sdk_grid_build_locator first");
    int n_outputs = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const sdk_interp_job &job = jobs[j];
        if (job.kind != SDK_JOB_SCALARS && job.kind != SDK_JOB_VECTOR) return fail(SDK_EINVAL, "sdk_interp_host: unknown job kind");
        if (job.kind == SDK_JOB_VECTOR ? job.n_fields != 2 : (job.n_fields < 1 || job.n_fields > SDK_MAX_MULTI_FIELDS))
            return fail(SDK_EINVAL, "sdk_interp_host: bad number of fields in a job");
        if (job.zsel < 0 || job.zsel > SDK_SEL_ZL_LIN || job.tsel < 0 || job.tsel > SDK_SEL_T_LIN)
            return fail(SDK_EINVAL, "sdk_interp_host: bad index family");
        for (int k = 0; k < job.n_fields; ++k)
            if (!job.out[k]) return fail(SDK_EINVAL, "sdk_interp_host: output array missing");
        n_outputs += job.n_fields;
    }
    const int64_t n = io->n;
    if (n < 0) return fail(SDK_EINVAL, "sdk_interp_host: negative count");
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e;
    uint32_t *flags_dev = (uint32_t *)scratch;
    if ((e = cudaMemsetAsync(scratch, 0, kInterpHeadBytes, stream)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: memset");
    int max_chunks = io->max_chunks > 0 ? io->max_chunks : kInterpChunksDefault;
    if (const char *env = std::getenv("SDK_INTERP_CHUNKS")) {  // tuning knob
        const int v = std::atoi(env);
        if (v >= 1) max_chunks = v;
    }
    if (max_chunks > kInterpChunksMax) max_chunks = kInterpChunksMax;
    const int64_t nchunks = interp_chunk_count(n, max_chunks);
    char *base = (char *)scratch + kInterpHeadBytes;
    if (n > 0 && nchunks < 2) {
        int rc = interp_chunk(grid, io, 0, n, ts_dev, n_t, jobs, n_jobs, flags_dev, base, stream);
        if (rc) return rc;
    } else if (n > 0) {
        if (int rc = ensure_pipe(grid)) return rc;
        if ((e = cudaEventRecord(grid->pipe_ev[3], stream)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: event record");
        for (auto &st : grid->pipe)
            if ((e = cudaStreamWaitEvent(st, grid->pipe_ev[3], 0)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: wait");
        const int64_t per = ((n + nchunks - 1) / nchunks + 31) / 32 * 32;
        const size_t slot = interp_slot_bytes(per, n_outputs);
        int c = 0;
        for (int64_t off = 0; off < n; off += per, ++c) {
            const int64_t cnt = n - off < per ? n - off : per;
            // chunk c and chunk c + 3 share a slot and a stream: the stream's order keeps them apart
            int rc = interp_chunk(grid, io, off, cnt, ts_dev, n_t, jobs, n_jobs, flags_dev, base + (size_t)(c % 3) * slot, grid->pipe[c % 3]);
            if (rc) return rc;
        }
        for (int k = 0; k < 3; ++k) {
            if ((e = cudaEventRecord(grid->pipe_ev[k], grid->pipe[k])) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: event record");
            if ((e = cudaStreamWaitEvent(stream, grid->pipe_ev[k], 0)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: wait");
        }
    }
    if (io->out_flags && (e = cudaMemcpyAsync(io->out_flags, flags_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        return cuda_fail(e, "sdk_interp_host: D2H");
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, "sdk_interp_host: sync");
    return SDK_OK;
}

}  // extern "C"
