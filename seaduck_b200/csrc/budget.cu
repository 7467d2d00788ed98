// K5: crossing-log post-processing of the Lagrangian budget, on the device.
//
// Replaces lagrangian_budget.find_ind_frac_tres (reference seaduck/lagrangian_budget.py:245) for a
// log that is still in HBM (the 128-byte records sdk_advect writes): per record the index of the
// wall the particle is on (which_wall :85, wall_index :192 incl. the LLC face-edge cases of
// ind_tend_uv :163), for the first and last record of every particle the walls it came from / goes to
// and the fraction of the cell traverse (redo_index :223, pseudo_motion :97), and the residence time
// (residence_time :139, tres_update :144).  One thread per particle walks its records.
#include <cuda_runtime.h>

#include "budget_args.cuh"

namespace sdk {

// wall_index (lagrangian_budget.py:192): index of wall `iwall` (0 left, 1 right, 2 down, 3 up, 4 deep,
// 5 shallow) of cell (izl, f, iy, ix) as (iw, iz, [face,] iy, ix)
__device__ __forceinline__ void wall_index(const GridDev &g, int iwall, int izl, int f, int iy, int ix, int16_t *out,
                                           int64_t stride, int rows) {
    int iw = iwall / 2;
    int iz = izl + (iwall == 4) - 1;
    int ny_ = iy + (iwall == 3), nx_ = ix + (iwall == 1);
    int nf = f;
    if (g.typ == SDK_TYP_LLC && (ny_ > g.gny - 1 || nx_ > g.gnx - 1)) {
        // the right / top wall of an edge cell lives on the neighbouring face, possibly as the other
        // component (ind_tend_uv :163 -> Topology._ind_tend_U / _ind_tend_V)
        int cf = f, cy = iy, cx = ix, came;
        const int st = cell_move(g, iw == 0 ? 3 : 0, cf, cy, cx, came);
        if (st == 0) {
            nf = cf, ny_ = cy, nx_ = cx;
            if (iw == 0 && came == 1) iw = 1;       // arrived through a "down" edge: a V wall
            else if (iw == 1 && came == 2) iw = 0;  // arrived through a "left" edge: a U wall
        }
    }
    out[0] = (int16_t)iw;
    out[stride] = (int16_t)iz;
    if (rows == 5) {
        out[2 * stride] = (int16_t)nf;
        out[3 * stride] = (int16_t)ny_;
        out[4 * stride] = (int16_t)nx_;
    } else {
        out[2 * stride] = (int16_t)ny_;
        out[3 * stride] = (int16_t)nx_;
    }
}

// pseudo_motion (lagrangian_budget.py:97): first wall and time forward (sg = +1) or backward (-1)
__device__ __forceinline__ void first_wall(const sdk_log_record &r, double sg, int &tend, double &t_event) {
    double ts[7];
    wall_times(r.uu, r.du, r.rx, sg, ts[0], ts[1]);
    wall_times(r.vv, r.dv, r.ry, sg, ts[2], ts[3]);
    wall_times(r.ww, r.dw, r.rzl - 0.5, sg, ts[4], ts[5]);
    ts[6] = sg * 1e80;
    tend = 0;
    double best = INFINITY;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double d = ts[k] * sg;
        if (isnan(d) || d < 0) d = INFINITY;
        if (d < best) {
            best = d;
            tend = k;
        }
    }
    t_event = ts[0];
#pragma unroll
    for (int k = 1; k < 7; ++k)
        if (tend == k) t_event = ts[k];
}

__global__ void __launch_bounds__(128) budget_kernel(const __grid_constant__ BudgetArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_particles) return;
    const int64_t o = a.offset[p], e = a.offset[p + 1];
    for (int64_t i = o; i < e; ++i) {
        const sdk_log_record r = a.rec[i];
        const int f = a.g.has_face ? r.fc : 0;
        double frac = 1.0;
        if (i == o || i == e - 1) {
            // redo_index (:223)
            int tendf, tendb;
            double tf, tb;
            first_wall(r, 1.0, tendf, tf);
            first_wall(r, -1.0, tendb, tb);
            if (tendf == 6) tendf = 0;
            if (tendb == 6) tendb = 0;
            wall_index(a.g, tendf, r.izl, f, r.iy, r.ix, a.ind1 + i, a.n_rec, a.rows);
            wall_index(a.g, tendb, r.izl, f, r.iy, r.ix, a.ind2 + i, a.n_rec, a.rows);
            const double tim = tf - tb;
            frac = -tb / tim;
            if (frac != frac) frac = 1.0;  // np.nan_to_num(frac, nan=1)
            else frac = nan_to_num(frac);
        } else {
            // which_wall (:85): the nearest of the six walls
            const double d[6] = {fabs(0.5 + r.rx), fabs(0.5 - r.rx), fabs(0.5 + r.ry), fabs(0.5 - r.ry), fabs(r.rzl), fabs(1 - r.rzl)};
            int iwall = 0;
#pragma unroll
            for (int k = 1; k < 6; ++k)
                if (d[k] < d[iwall]) iwall = k;
            wall_index(a.g, iwall, r.izl, f, r.iy, r.ix, a.ind1 + i, a.n_rec, a.rows);
            for (int k = 0; k < a.rows; ++k) a.ind2[k * a.n_rec + i] = 1;
        }
        a.frac[i] = frac;
        // residence_time (:139): 2 / sum |u| over the six walls
        const double z0 = r.rzl - 0.5;
        const double ul = r.uu - (r.rx + 0.5) * r.du, ur = r.uu + (0.5 - r.rx) * r.du;
        const double vl = r.vv - (r.ry + 0.5) * r.dv, vr = r.vv + (0.5 - r.ry) * r.dv;
        const double wl = r.ww - (z0 + 0.5) * r.dw, wr = r.ww + (0.5 - z0) * r.dw;
        a.tres[i] = 2 / (((((fabs(ul) + fabs(ur)) + fabs(vl)) + fabs(vr)) + fabs(wl)) + fabs(wr));
        a.work[i] = 1.0;
    }
}

// tres_update (:144): fracs[first + 1] = fraction_first, then fracs[last] -= fraction_last (in this
// order, for all particles), tres = tres0 * fracs, zero where the stamp is > 6
__global__ void budget_first_kernel(const __grid_constant__ BudgetArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_particles) return;
    const int64_t o = a.offset[p];
    if (a.offset[p + 1] > o && o + 1 < a.n_rec) a.work[o + 1] = a.frac[o];
}

__global__ void budget_last_kernel(const __grid_constant__ BudgetArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_particles) return;
    const int64_t o = a.offset[p], e = a.offset[p + 1];
    if (e > o) a.work[e - 1] -= a.frac[e - 1];
}

__global__ void budget_tres_kernel(const __grid_constant__ BudgetArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rec) return;
    a.tres[i] = a.rec[i].vs > 6 ? 0.0 : a.tres[i] * a.work[i];
}

cudaError_t launch_budget(const BudgetArgs &a, cudaStream_t s) {
    if (a.n_particles == 0 || a.n_rec == 0) return cudaSuccess;
    const unsigned pb = (unsigned)((a.n_particles + 127) / 128), rb = (unsigned)((a.n_rec + 255) / 256);
    budget_kernel<<<pb, 128, 0, s>>>(a);
    budget_first_kernel<<<pb, 128, 0, s>>>(a);
    budget_last_kernel<<<pb, 128, 0, s>>>(a);
    budget_tres_kernel<<<rb, 256, 0, s>>>(a);
    return cudaGetLastError();
}

// ---- calculate_budget (lagrangian_budget.py:603) ---------------------------------------------------------

__global__ void budget_mark_last_kernel(const __grid_constant__ BudgetTermArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_particles) return;
    const int64_t end = a.offset[p + 1];
    if (end > a.offset[p]) a.is_last[end - 1] = 1;
}

// numpy fancy indexing of a [3][shape] array with one column of ind1 / ind2
__device__ __forceinline__ double wall_value(const BudgetTermArgs &a, const int16_t *ind, int64_t i) {
    const int32_t *sh = a.f.wall_shape;
    int r = 0;
    const int iw = wrap_index(ind[(r++) * a.n_rec + i], 3);
    const int iz = wrap_index(ind[(r++) * a.n_rec + i], sh[0]);
    const int fc = a.f.rows == 5 ? wrap_index(ind[(r++) * a.n_rec + i], sh[1]) : 0;
    const int iy = wrap_index(ind[(r++) * a.n_rec + i], sh[2]);
    const int ix = wrap_index(ind[r * a.n_rec + i], sh[3]);
    return __ldg(a.f.walls + ((((int64_t)iw * sh[0] + iz) * sh[1] + fc) * sh[2] + iy) * sh[3] + ix);
}

// trc_conc = s1 * frac + (1 - frac) * s2   (:652-655)
__global__ void budget_conc_kernel(const __grid_constant__ BudgetTermArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rec) return;
    const double fr = a.frac[i];
    a.conc[i] = wall_value(a, a.ind1, i) * fr + (1 - fr) * wall_value(a, a.ind2, i);
}

__global__ void budget_contr_kernel(const __grid_constant__ BudgetTermArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = a.n_rec - 1;
    if (i >= m) return;
    double delta = 0.0, span = 0.0, dt = 0.0;
    const sdk_log_record &r = a.rec[i];
    if (!a.is_last[i]) {
        delta = nan_to_num(a.conc[i + 1] - a.conc[i]);     // :656-657
        span = (double)a.f.dir_of_time * a.tres[i + 1];    // :658-659
        dt = nan_to_num(a.rec[i + 1].tt - r.tt);           // lhs_contribution :575
    }
    const int32_t *sh = a.f.cell_shape;
    const int64_t cell = (((int64_t)wrap_index(r.izl - 1, sh[0]) * sh[1] + (a.f.rows == 5 ? wrap_index(r.fc, sh[1]) : 0)) * sh[2] +
                          wrap_index(r.iy, sh[2])) * sh[3] + wrap_index(r.ix, sh[3]);
    const double lhs = dt * __ldg(a.f.lhs + cell);
    const double target = delta - lhs;  // :662
    // contr_p_relaxed :583, p = 1
    double x[SDK_MAX_BUDGET_TERMS];
    double deno = 0.0, sums = 0.0;
#pragma unroll
    for (int k = 0; k < SDK_MAX_BUDGET_TERMS; ++k)
        if (k < a.f.n_terms) {
            x[k] = __ldg(a.f.term[k] + cell);
            deno += x[k] * x[k];
            sums += x[k];
        }
    const double gap = target - sums * span;
    double total = 0.0;
#pragma unroll
    for (int k = 0; k < SDK_MAX_BUDGET_TERMS; ++k)
        if (k < a.f.n_terms) {
            const double c = x[k] * span + x[k] * x[k] / deno * gap;
            a.contr[k * m + i] = c;
            total += c;
        }
    a.contr[(int64_t)a.f.n_terms * m + i] = 0.0 + (target - total);
    a.contr[(int64_t)(a.f.n_terms + 1) * m + i] = lhs;
}

cudaError_t launch_budget_terms(const BudgetTermArgs &a, cudaStream_t s) {
    if (a.n_rec == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.is_last, 0, (size_t)a.n_rec, s);
    if (e != cudaSuccess) return e;
    const int t = 256;
    if (a.n_particles > 0) budget_mark_last_kernel<<<(unsigned)((a.n_particles + t - 1) / t), t, 0, s>>>(a);
    budget_conc_kernel<<<(unsigned)((a.n_rec + t - 1) / t), t, 0, s>>>(a);
    if (a.n_rec > 1) budget_contr_kernel<<<(unsigned)((a.n_rec - 1 + t - 1) / t), t, 0, s>>>(a);
    return cudaGetLastError();
}

// ---- read_wall_list / check_particle_data_compat (lagrangian_budget.py:406, :470) --------------------------

__device__ __forceinline__ double field_at(const double *a, const int32_t *sh, int iz, int f, int iy, int ix) {
    return __ldg(a + (((int64_t)wrap_index(iz, sh[0]) * sh[1] + wrap_index(f, sh[1])) * sh[2] + wrap_index(iy, sh[2])) * sh[3] +
                 wrap_index(ix, sh[3]));
}

__global__ void wall_values_kernel(const __grid_constant__ WallValueArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rec) return;
    const sdk_log_record &r = a.rec[i];
    const sdk_wall_fields &w = a.f;
    const int f = r.fc, iy = r.iy, ix = r.ix, lev = r.izl - 1;
    // neighbours to the right and above; where a face has none the old index stays (ind_tend_vec topology.py:554-562)
    int rf = f, ry = iy, rx = ix, in_r, uf = f, uy = iy, ux = ix, in_u;
    cell_move(a.g, 3, rf, ry, rx, in_r);
    cell_move(a.g, 0, uf, uy, ux, in_u);
    // rotation of the neighbouring face (_llc_get_uv_mask_from_face topology.py:222): identity on the same face
    double ufu = 1.0, ufv = 0.0, vfu = 0.0, vfv = 1.0;
    if (rf != f) {
        ufu = w.rot_cos[3 * 4 + in_r];
        ufv = w.rot_sin[3 * 4 + in_r];
    }
    if (uf != f) {
        vfu = -w.rot_sin[0 * 4 + in_u];
        vfv = w.rot_cos[0 * 4 + in_u];
    }
    if (w.scalar) ufu = fabs(ufu), ufv = fabs(ufv), vfu = fabs(vfu), vfv = fabs(vfv);
    const double ur = nan_to_num(field_at(w.x_walls, w.shape_x, lev, rf, ry, rx)) * ufu +
                      nan_to_num(field_at(w.y_walls, w.shape_y, lev, rf, ry, rx)) * ufv;
    const double vr = nan_to_num(field_at(w.x_walls, w.shape_x, lev, uf, uy, ux)) * vfu +
                      nan_to_num(field_at(w.y_walls, w.shape_y, lev, uf, uy, ux)) * vfv;
    double c[6];
    c[0] = nan_to_num(field_at(w.x_walls, w.shape_x, lev, f, iy, ix));
    c[1] = nan_to_num(ur);
    c[2] = nan_to_num(field_at(w.y_walls, w.shape_y, lev, f, iy, ix));
    c[3] = nan_to_num(vr);
    c[4] = nan_to_num(field_at(w.z_walls, w.shape_z, r.izl, f, iy, ix));
    c[5] = nan_to_num(field_at(w.z_walls, w.shape_z, lev, f, iy, ix));
#pragma unroll
    for (int k = 0; k < 6; ++k) a.c_list[k * a.n_rec + i] = c[k];
    // wall velocities from (u, du, r): _uleftright_from_udu utils.py:852; the vertical coordinate is rz - 1/2
    const double z = r.rzl - 0.5;
    double u[6];
    u[0] = r.uu - (r.rx + 0.5) * r.du;
    u[1] = r.uu + (0.5 - r.rx) * r.du;
    u[2] = r.vv - (r.ry + 0.5) * r.dv;
    u[3] = r.vv + (0.5 - r.ry) * r.dv;
    u[4] = r.ww - (z + 0.5) * r.dw;
    u[5] = r.ww + (0.5 - z) * r.dw;
    if (a.u_list) {
#pragma unroll
        for (int k = 0; k < 6; ++k) a.u_list[k * a.n_rec + i] = u[k];
    }
    if (a.lag_conv) {  // crude_convergence :464: the six products with alternating sign, summed in order
        double conv = (c[0] * u[0]) * 1.0;
#pragma unroll
        for (int k = 1; k < 6; ++k) conv = conv + (c[k] * u[k]) * ((k & 1) ? -1.0 : 1.0);
        a.lag_conv[i] = conv;
    }
    if (a.eul_conv && w.conv) a.eul_conv[i] = field_at(w.conv, w.shape_c, lev, f, iy, ix);
}

cudaError_t launch_wall_values(const WallValueArgs &a, cudaStream_t s) {
    if (a.n_rec == 0) return cudaSuccess;
    wall_values_kernel<<<(unsigned)((a.n_rec + 127) / 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sdk
