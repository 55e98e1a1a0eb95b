/*
 * pillar_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * Plain-C CPU restatement of the reference's dynamic pillar encoder,
 *   /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py
 * (DynamicPillarVFE :49-142, DynamicPillarVFESimple2D :146-252 and its two radar
 * subclasses :255-373, PFNLayerV2 :14-46), plus the two torch_scatter functions it
 * calls (scatter_mean / scatter_max; torch-scatter==2.1.1, third-party, not in the
 * reference tree -- semantics restated from the library's documentation, see
 * oracle/ref_loader.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (radardistill_b200/) never does.
 *
 * Parity status: the reference ships NO tests / golden vectors for this path
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the reference
 * module itself, run in the build container through oracle/ref_loader.py and committed
 * as the .npz files under tests/golden/ by oracle/gen_golden.py.
 *
 * Arithmetic modes (argument `mean_mode`):
 *   ORC_MEAN_SEQ_F32 (0): scatter_mean exactly as torch_scatter's CPU kernel does it --
 *       fp32 running sum in row order, fp32 divide by the count.  This is the
 *       "reference CPU semantics" mode used to pin against the goldens.
 *   ORC_MEAN_F64 (1): the CANONICAL order-independent definition the CUDA kernels
 *       implement: sum in fp64 (exact for any realistic pillar), one rounding to fp32.
 *       Differences to mode 0 are <= 1 ulp of the mean and are bounded in the tests.
 * The linear layer is a sequential fmaf chain over k = 0..Cin-1 (the reference's sgemm
 * has an unspecified summation order; results agree to fp32 tolerance, tests state it).
 *
 * Threads: the two per-point loops (features, linear) are split over pthreads
 * (orc_set_threads; libgomp is not in the image).  Everything else is scalar.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tiny parallel-for */
static int g_threads = 1;
void orc_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }
int orc_get_threads(void) { return g_threads; }

typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t lo, hi; } range_job;
static void *range_tramp(void *p) { range_job *j = (range_job *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void parallel_for(int64_t n, range_fn fn, void *ctx) {
    int t = g_threads;
    if (t <= 1 || n < 4096) { fn(0, n, ctx); return; }
    pthread_t th[256];
    range_job jobs[256];
    int64_t chunk = (n + t - 1) / t;
    int started = 0;
    for (int i = 0; i < t; ++i) {
        int64_t lo = i * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo >= hi) break;
        jobs[i].fn = fn; jobs[i].ctx = ctx; jobs[i].lo = lo; jobs[i].hi = hi;
        if (pthread_create(&th[i], NULL, range_tramp, &jobs[i]) != 0) { fn(lo, hi, ctx); th[i] = 0; jobs[i].fn = NULL; }
        started = i + 1;
    }
    for (int i = 0; i < started; ++i) if (jobs[i].fn) pthread_join(th[i], NULL);
}

#define ORC_MEAN_SEQ_F32 0
#define ORC_MEAN_F64 1

#define ORC_LAYOUT_SIMPLE2D 0 /* dynamic_pillar_vfe.py:219-237 */
#define ORC_LAYOUT_DYNPILLAR 1 /* dynamic_pillar_vfe.py:113-121 */

typedef struct {
    float lo[3];   /* point_cloud_range[0:3]                       (:189 / :85) */
    float vsz[3];  /* voxel_size as fp32                           (:188 / :84) */
    float off[3];  /* voxel/2 + lo, computed in double, cast once  (:180-182)   */
    int32_t nx, ny; /* grid_size[0], grid_size[1]                  (:184-187)   */
    int32_t cols;   /* 1 + C floats per row (batch idx first)                   */
    int32_t layout, use_abs, use_cluster, use_relative, with_distance;
    int32_t c_in, c_out;
} orc_cfg;

int orc_abi_version(void) { return 2; }

/* ------------------------------------------------------------------ a1-a4: index part */

/* LSD radix sort of (key, idx) pairs by key (keys are non-negative here). */
static void radix_sort_pairs(uint32_t *key, int32_t *idx, int64_t n) {
    uint32_t *k2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *i2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int pass = 0; pass < 4; ++pass) {
        int64_t hist[257];
        memset(hist, 0, sizeof(hist));
        int sh = pass * 8;
        for (int64_t i = 0; i < n; ++i) hist[((key[i] >> sh) & 255u) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (int64_t i = 0; i < n; ++i) {
            int64_t d = hist[(key[i] >> sh) & 255u]++;
            k2[d] = key[i];
            i2[d] = idx[i];
        }
        uint32_t *tk = key; key = k2; k2 = tk;
        int32_t *ti = idx; idx = i2; i2 = ti;
    }
    /* 4 passes => data is back in the caller's buffers */
    free(k2);
    free(i2);
}

/*
 * Quantise, mask, merged key, unique (sorted), inverse, counts, coords.
 *   keep[j]   : original row of the j-th kept point (stable, input order)  (:203-206)
 *   pcoord    : (n,2) int32 cx, cy of kept points                          (:201-202)
 *   inv[j]    : pillar id of kept point j                                  (:212)
 *   unq[p]    : merged key of pillar p, ascending                          (:208-212)
 *   cnt[p]    : points in pillar p
 *   coords    : (P, coord_cols) int32, [b,cy,cx] or [b,0,cy,cx]            (:243-248 / :132-138)
 * Returns 0, or -1 if a kept row has a negative merged key (batch index < 0).
 */
int orc_index(const float *pts, int64_t n0, const orc_cfg *g, int coord_cols,
              int32_t *keep, int32_t *pcoord, int32_t *inv, int32_t *unq, int32_t *cnt,
              int32_t *coords, int64_t *n_out, int64_t *p_out) {
    const int cols = g->cols;
    int64_t n = 0;
    for (int64_t i = 0; i < n0; ++i) {
        const float *r = pts + i * cols;
        /* torch: floor((x - lo) / vsz).int()   -- fp32 sub, IEEE fp32 divide, floor, cast */
        float qx = floorf((r[1] - g->lo[0]) / g->vsz[0]);
        float qy = floorf((r[2] - g->lo[1]) / g->vsz[1]);
        /* NaN / +-inf / out-of-int32 values: torch's .int() yields INT_MIN on x86 => dropped */
        int okx = (qx >= 0.0f) && (qx < (float)g->nx);
        int oky = (qy >= 0.0f) && (qy < (float)g->ny);
        if (okx && oky) {
            keep[n] = (int32_t)i;
            pcoord[2 * n] = (int32_t)qx;
            pcoord[2 * n + 1] = (int32_t)qy;
            ++n;
        }
    }
    *n_out = n;
    uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    const int32_t sxy = g->nx * g->ny, sy = g->ny;
    int bad = 0;
    for (int64_t j = 0; j < n; ++j) {
        int32_t b = (int32_t)pts[(int64_t)keep[j] * cols]; /* .int() truncates toward zero */
        int32_t k = b * sxy + pcoord[2 * j] * sy + pcoord[2 * j + 1]; /* int32, :208-210 */
        if (k < 0) bad = 1;
        key[j] = (uint32_t)k;
        idx[j] = (int32_t)j;
    }
    if (bad) { free(key); free(idx); return -1; }
    radix_sort_pairs(key, idx, n); /* stable; ascending == torch.unique(sorted=True) */
    int64_t p = -1;
    for (int64_t s = 0; s < n; ++s) {
        if (s == 0 || key[s] != key[s - 1]) {
            ++p;
            unq[p] = (int32_t)key[s];
            cnt[p] = 0;
        }
        cnt[p]++;
        inv[idx[s]] = (int32_t)p;
    }
    const int64_t P = p + 1;
    *p_out = P;
    for (int64_t q = 0; q < P; ++q) {
        int32_t u = unq[q];
        int32_t b = u / sxy, cx = (u % sxy) / sy, cy = u % sy;
        if (coord_cols == 3) {
            coords[3 * q] = b; coords[3 * q + 1] = cy; coords[3 * q + 2] = cx;
        } else {
            coords[4 * q] = b; coords[4 * q + 1] = 0; coords[4 * q + 2] = cy; coords[4 * q + 3] = cx;
        }
    }
    free(key);
    free(idx);
    return 0;
}

/* ------------------------------------------------------------------ a5: scatter_mean */
void orc_pillar_mean(const float *pts, const orc_cfg *g, const int32_t *keep, const int32_t *inv,
                     const int32_t *cnt, int64_t n, int64_t P, int mean_mode, float *mean) {
    const int cols = g->cols;
    if (mean_mode == ORC_MEAN_SEQ_F32) {
        /* torch_scatter CPU: out[index[i]] += src[i] in row order, then out / clamp(count,1) */
        for (int64_t q = 0; q < 3 * P; ++q) mean[q] = 0.0f;
        for (int64_t j = 0; j < n; ++j) {
            const float *r = pts + (int64_t)keep[j] * cols;
            float *m = mean + 3 * (int64_t)inv[j];
            m[0] = m[0] + r[1]; m[1] = m[1] + r[2]; m[2] = m[2] + r[3];
        }
        for (int64_t q = 0; q < P; ++q) {
            float c = (float)(cnt[q] < 1 ? 1 : cnt[q]);
            mean[3 * q] = mean[3 * q] / c; mean[3 * q + 1] = mean[3 * q + 1] / c; mean[3 * q + 2] = mean[3 * q + 2] / c;
        }
    } else {
        double *acc = (double *)calloc((size_t)(3 * P + 1), sizeof(double));
        for (int64_t j = 0; j < n; ++j) {
            const float *r = pts + (int64_t)keep[j] * cols;
            double *m = acc + 3 * (int64_t)inv[j];
            m[0] += (double)r[1]; m[1] += (double)r[2]; m[2] += (double)r[3];
        }
        for (int64_t q = 0; q < P; ++q) {
            double c = (double)(cnt[q] < 1 ? 1 : cnt[q]);
            mean[3 * q] = (float)(acc[3 * q] / c);
            mean[3 * q + 1] = (float)(acc[3 * q + 1] / c);
            mean[3 * q + 2] = (float)(acc[3 * q + 2] / c);
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------ a6-a8: decorated features */
/* Writes the Cin features of one kept point (same op order / roundings as the reference). */
static inline void point_features(const float *r, const int32_t *pc, const float *m, const orc_cfg *g, float *f) {
    const int C = g->cols - 1;
    const float x = r[1], y = r[2], z = r[3];
    /* :215-217  x - (cx.float()*voxel_x + x_offset): separate mul, add, sub roundings */
    float cen[3];
    cen[0] = x - ((float)pc[0] * g->vsz[0] + g->off[0]);
    cen[1] = y - ((float)pc[1] * g->vsz[1] + g->off[1]);
    cen[2] = z - g->off[2];
    int k = 0;
    if (g->layout == ORC_LAYOUT_SIMPLE2D) {
        f[k++] = cen[0]; f[k++] = cen[1]; f[k++] = cen[2];
        for (int c = g->use_abs ? 1 : 4; c <= C; ++c) f[k++] = r[c];
        if (g->use_cluster) { f[k++] = x - m[0]; f[k++] = y - m[1]; f[k++] = z - m[2]; }
        if (g->with_distance) f[k++] = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
        if (g->use_relative) { f[k++] = x - g->lo[0]; f[k++] = y - g->lo[1]; f[k++] = z - g->lo[2]; }
    } else {
        for (int c = g->use_abs ? 1 : 4; c <= C; ++c) f[k++] = r[c];
        f[k++] = x - m[0]; f[k++] = y - m[1]; f[k++] = z - m[2];
        f[k++] = cen[0]; f[k++] = cen[1]; f[k++] = cen[2];
        if (g->with_distance) f[k++] = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
    }
}

typedef struct { const float *pts; const orc_cfg *g; const int32_t *keep, *pcoord, *inv; const float *mean; float *f; } feat_ctx;
static void feat_range(int64_t lo, int64_t hi, void *p) {
    feat_ctx *c = (feat_ctx *)p;
    const int cols = c->g->cols, cin = c->g->c_in;
    for (int64_t j = lo; j < hi; ++j)
        point_features(c->pts + (int64_t)c->keep[j] * cols, c->pcoord + 2 * j, c->mean + 3 * (int64_t)c->inv[j], c->g,
                       c->f + j * cin);
}
void orc_features(const float *pts, const orc_cfg *g, const int32_t *keep, const int32_t *pcoord,
                  const int32_t *inv, const float *mean, int64_t n, float *f) {
    feat_ctx c = {pts, g, keep, pcoord, inv, mean, f};
    parallel_for(n, feat_range, &c);
}

/* ------------------------------------------------------------------ a9: PFNLayerV2 */
/* x[j][c] = sum_k W[c][k] f[j][k]  as a sequential fmaf chain (k ascending), (+ bias). */
typedef struct { const float *f, *W; float *x; int cin, cout; } lin_ctx;
static void lin_range(int64_t lo, int64_t hi, void *p) {
    lin_ctx *c = (lin_ctx *)p;
    const int cin = c->cin, cout = c->cout;
    for (int64_t j = lo; j < hi; ++j) {
        const float *fj = c->f + j * cin;
        for (int o = 0; o < cout; ++o) {
            float acc = 0.0f;
            for (int k = 0; k < cin; ++k) acc = fmaf(c->W[o * cin + k], fj[k], acc);
            c->x[j * cout + o] = acc;
        }
    }
}
void orc_linear(const float *f, int64_t n, int cin, const float *W, int cout, float *x) {
    lin_ctx c = {f, W, x, cin, cout};
    parallel_for(n, lin_range, &c);
}

/* Train-mode batch statistics over the n kept points: mean and BIASED variance, in fp64. */
void orc_bn_batch_stats(const float *x, int64_t n, int cout, double *mean, double *var) {
    double *s = (double *)calloc((size_t)(2 * cout), sizeof(double)), *s2 = s + cout;
    for (int64_t j = 0; j < n; ++j)          /* per channel: rows in ascending order */
        for (int c = 0; c < cout; ++c) {
            double v = (double)x[j * cout + c];
            s[c] += v;
            s2[c] += v * v;
        }
    for (int c = 0; c < cout; ++c) {
        double m = n > 0 ? s[c] / (double)n : 0.0;
        double vv = n > 0 ? s2[c] / (double)n - m * m : 0.0;
        mean[c] = m;
        var[c] = vv > 0.0 ? vv : 0.0;
    }
    free(s);
}

/*
 * Canonical (ORC_MEAN_F64) train-mode statistics: x = W f is linear, so the batch moments of x follow from the
 * first and second moments of the features, accumulated in fp64 (products of two fp32 values are exact in fp64):
 *   S1[k] = sum_j f[j][k],  S2[k][l] = sum_j f[j][k] f[j][l],
 *   mean_c = (w_c . S1) / n,  E[x^2]_c = (w_c^T S2 w_c) / n,  var_c = E[x^2]_c - mean_c^2   (biased, as BatchNorm1d).
 * Differs from summing the rounded fp32 x values (orc_bn_batch_stats) by the fp32 rounding of x only (~1e-7
 * relative); this is what the CUDA kernels compute (one pass over the rows without the linear layer).
 */
void orc_bn_batch_stats_moments(const float *f, int64_t n, int cin, const float *W, int cout, double *mean, double *var) {
    double *S1 = (double *)calloc((size_t)(cin + cin * cin), sizeof(double)), *S2 = S1 + cin;
    for (int64_t j = 0; j < n; ++j) {
        const float *fj = f + j * cin;
        for (int k = 0; k < cin; ++k) {
            const double a = (double)fj[k];
            S1[k] += a;
            for (int l = k; l < cin; ++l) S2[k * cin + l] += a * (double)fj[l];
        }
    }
    for (int k = 0; k < cin; ++k)
        for (int l = 0; l < k; ++l) S2[k * cin + l] = S2[l * cin + k];
    for (int c = 0; c < cout; ++c) {
        const float *w = W + c * cin;
        double m = 0.0, e2 = 0.0;
        for (int k = 0; k < cin; ++k) m += (double)w[k] * S1[k];
        for (int k = 0; k < cin; ++k) {
            double row = 0.0;
            for (int l = 0; l < cin; ++l) row += S2[k * cin + l] * (double)w[l];
            e2 += (double)w[k] * row;
        }
        m = n > 0 ? m / (double)n : 0.0;
        double vv = n > 0 ? e2 / (double)n - m * m : 0.0;
        mean[c] = m;
        var[c] = vv > 0.0 ? vv : 0.0;
    }
    free(S1);
}

/* y = x*scale + shift with scale = gamma/sqrt(var+eps), shift = beta - mean*scale (fp64, one rounding). */
void orc_bn_fold(const float *gamma, const float *beta, const double *mean, const double *var, double eps,
                 int cout, float *scale, float *shift) {
    for (int c = 0; c < cout; ++c) {
        double inv_std = 1.0 / sqrt(var[c] + eps);
        double a = (double)gamma[c] * inv_std;
        scale[c] = (float)a;
        shift[c] = (float)((double)beta[c] - mean[c] * a);
    }
}

/* z = relu(fma(x, scale, shift)); out[p][c] = max_j z ; arg = FIRST j attaining it (:39-40). */
void orc_act_max(const float *x, int64_t n, int cout, const float *scale, const float *shift,
                 const int32_t *inv, int64_t P, float *out, int32_t *arg) {
    for (int64_t q = 0; q < P * cout; ++q) { out[q] = -1.0f; arg[q] = (int32_t)n; }
    for (int64_t j = 0; j < n; ++j) {
        float *o = out + (int64_t)inv[j] * cout;
        int32_t *a = arg + (int64_t)inv[j] * cout;
        for (int c = 0; c < cout; ++c) {
            float y = fmaf(x[j * cout + c], scale[c], shift[c]);
            float z = y > 0.0f ? y : 0.0f;
            if (z > o[c]) { o[c] = z; a[c] = (int32_t)j; } /* strict '>' => first index wins */
        }
    }
}

/* ------------------------------------------------------------------ a12: backward (parameter grads)
 * Given g (P,Cout): route to argmax rows, ReLU', BatchNorm backward (train: batch stats; eval:
 * running stats => g_x = scale * g_y), dW = g_x^T f.  All reductions in fp64.
 * use_norm == 0: linear has a bias, y = x + b: dbias -> dbeta slot, dgamma = 0.
 */
typedef struct {
    const float *gy, *f, *x;
    const double *mu, *inv_std, *a, *db, *dg;
    int cin, cout, pass, dense;
    int64_t n;
    double *acc; /* per job: pass 0 -> [db(cout) | dg(cout)], pass 1 -> dW (cout*cin) */
    int64_t lo, hi;
} bwd_job;

static void bwd_range(bwd_job *J) {
    const int cin = J->cin, cout = J->cout;
    for (int64_t j = J->lo; j < J->hi; ++j) {
        const float *gyj = J->gy + j * cout, *xj = J->x + j * cout, *fj = J->f + j * cin;
        if (J->pass == 0) {
            for (int c = 0; c < cout; ++c) {
                if (gyj[c] == 0.0f) continue;
                double xh = ((double)xj[c] - J->mu[c]) * J->inv_std[c];
                J->acc[c] += (double)gyj[c];
                J->acc[cout + c] += (double)gyj[c] * xh;
            }
        } else {
            for (int c = 0; c < cout; ++c) {
                double gx = (double)gyj[c];
                if (J->dense) {
                    double xh = ((double)xj[c] - J->mu[c]) * J->inv_std[c];
                    gx = gx - J->db[c] / (double)J->n - xh * J->dg[c] / (double)J->n;
                } else if (gx == 0.0) continue;
                gx *= J->a[c];
                double *w = J->acc + c * cin;
                for (int k = 0; k < cin; ++k) w[k] += gx * (double)fj[k];
            }
        }
    }
}
static void *bwd_tramp(void *p) { bwd_range((bwd_job *)p); return NULL; }

static void bwd_pass(bwd_job *proto, int64_t n, int width, double *out) {
    int t = g_threads;
    if (n < 4096) t = 1;
    bwd_job jobs[256];
    pthread_t th[256];
    int64_t chunk = (n + t - 1) / t;
    int used = 0;
    for (int i = 0; i < t; ++i) {
        int64_t lo = i * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo >= hi) break;
        jobs[i] = *proto;
        jobs[i].lo = lo; jobs[i].hi = hi;
        jobs[i].acc = (double *)calloc((size_t)width, sizeof(double));
        used = i + 1;
    }
    for (int i = 0; i < used; ++i)
        if (used == 1 || pthread_create(&th[i], NULL, bwd_tramp, &jobs[i]) != 0) { bwd_range(&jobs[i]); th[i] = 0; jobs[i].pass |= 256; }
    for (int i = 0; i < used; ++i) if (!(jobs[i].pass & 256)) pthread_join(th[i], NULL);
    for (int q = 0; q < width; ++q) out[q] = 0.0;
    for (int i = 0; i < used; ++i) {        /* fixed combination order */
        for (int q = 0; q < width; ++q) out[q] += jobs[i].acc[q];
        free(jobs[i].acc);
    }
}

void orc_backward(const float *g, const float *f, const float *x, const float *out, const int32_t *arg,
                  int64_t n, int64_t P, int cin, int cout, const float *gamma, const double *mean,
                  const double *var, double eps, int train_bn, int use_norm,
                  float *dW, float *dgamma, float *dbeta) {
    /* g_z routed to the argmax row, times ReLU' (out > 0).  One (pillar, channel) -> one row. */
    float *gy = (float *)calloc((size_t)(n * cout + 1), sizeof(float));
    for (int64_t q = 0; q < P; ++q)
        for (int c = 0; c < cout; ++c)
            if (out[q * cout + c] > 0.0f && arg[q * cout + c] < n)
                gy[(int64_t)arg[q * cout + c] * cout + c] = g[q * cout + c];
    double *mu = (double *)calloc((size_t)(6 * cout), sizeof(double));
    double *inv_std = mu + cout, *a = mu + 2 * cout, *dbdg = mu + 3 * cout; /* db | dg */
    for (int c = 0; c < cout; ++c) {
        inv_std[c] = use_norm ? 1.0 / sqrt(var[c] + eps) : 1.0;
        mu[c] = use_norm ? mean[c] : 0.0;
        a[c] = use_norm ? (double)gamma[c] * inv_std[c] : 1.0;
    }
    bwd_job J = {gy, f, x, mu, inv_std, a, dbdg, dbdg + cout, cin, cout, 0, 0, n, NULL, 0, 0};
    bwd_pass(&J, n, 2 * cout, dbdg);
    for (int c = 0; c < cout; ++c) {
        dbeta[c] = (float)dbdg[c];
        dgamma[c] = use_norm ? (float)dbdg[cout + c] : 0.0f;
    }
    double *dw = (double *)calloc((size_t)(cout * cin), sizeof(double));
    J.pass = 1;
    J.dense = (use_norm && train_bn) ? 1 : 0;
    bwd_pass(&J, n, cout * cin, dw);
    for (int q = 0; q < cout * cin; ++q) dW[q] = (float)dw[q];
    free(dw);
    free(mu);
    free(gy);
}

/* ====================================================================================================================
 * ORC_FOLDED (mean_mode 2): the canonical arithmetic of the round-2 CUDA kernels (radardistill_b200/csrc/rdp_pfn.cuh).
 *
 * Every decorated feature of the reference (:214-237, :105-121) is affine in the row's centre offsets
 * d = xyz - centre (exactly f_center, :215-217) and in per-pillar constants:
 *     xyz = d + centre,   f_cluster = xyz - mean = d + (centre - mean),   f_rel = xyz - lo = d + (centre - lo).
 * With T the (c_in x (G+1)) matrix that writes the layout's features in the reduced basis
 *     g = [dx, dy, dz, raw features 4.., (dist) | cx, cy, cx - mx, cy - my, cz - mz | 1]        (KIN row inputs, 5 pillar constants)
 * (cx - mx etc. are evaluated as -mean(d): fp64 sum of the rows' fp32 offsets d, one division, one rounding -- see orc_pillar_consts)
 * the linear layer is  x_c = W_c . f = wg_c . g + const_c  with [wg_c | const_c] = W_c T (fp64, rounded once), evaluated as
 *     v_ic = k-ascending fmaf chain over the KIN row inputs            (first term a plain product)
 *     u_pc = fmaf chain over the 5 pillar constants, starting from const_c
 *     x_ic = v_ic + u_pc  (one rounding);   y = fma(x, scale, shift);   z = max(y, 0);   out_pc = max_i z_ic.
 * Train-mode batch statistics come from the fp64 moments of g (S1 = sum g, S2 = sum g g^T):
 *     mean_c = wg_c . S1 / n + const_c,   var_c = wg_c^T S2 wg_c / n - (wg_c . S1 / n)^2.
 * argmax rule: the row of the pillar with the largest s_c v_ic (s_c = sign(scale_c): the row that maximises the BatchNorm
 * output), lowest kept index on exact ties; a (pillar, channel) whose maximum the ReLU clamped to 0 reports the pillar's
 * lowest kept index (every row ties at 0 -- torch_scatter's CPU rule, :40) and receives no gradient.
 * Against the reference arithmetic this differs by fp32 rounding only (features <= 1e-5 norm-relative on every golden).
 * ==================================================================================================================== */
#define ORC_MAX_G 14

/* T (c_in rows, stride ORC_MAX_G + 1).  Returns c_in, or -1.  Mirrors build_T in radardistill_b200/csrc/rdp_pfn.cu. */
int orc_build_T(const orc_cfg *g, float *T, int32_t *kin_out, int32_t *g_out) {
    const int C = g->cols - 1, KIN = C + (g->with_distance ? 1 : 0), G = KIN + 5, S = ORC_MAX_G + 1;
    if (G > ORC_MAX_G) return -1;
    int j = 0;
#define ROW() (memset(T + (size_t)j * S, 0, sizeof(float) * S), T + (size_t)(j++) * S)
    float *r;
    const float off_z = g->off[2];
    for (int pass = 0; pass < 5; ++pass) {
        /* block order: Simple2D = center, points, cluster, dist, relative ; DynamicPillarVFE = points, cluster, center, dist */
        int what;
        if (g->layout == ORC_LAYOUT_SIMPLE2D) { static const int o[5] = {0, 1, 2, 3, 4}; what = o[pass]; }
        else { static const int o[5] = {1, 2, 0, 3, -1}; what = o[pass]; }
        if (what == 0) {
            for (int a = 0; a < 3; ++a) { r = ROW(); r[a] = 1.0f; }
        } else if (what == 1) {
            if (g->use_abs) {
                r = ROW(); r[0] = 1.0f; r[KIN] = 1.0f;
                r = ROW(); r[1] = 1.0f; r[KIN + 1] = 1.0f;
                r = ROW(); r[2] = 1.0f; r[G] = off_z;
            }
            for (int c = 4; c <= C; ++c) { r = ROW(); r[c - 1] = 1.0f; }
        } else if (what == 2) {
            if (g->layout != ORC_LAYOUT_SIMPLE2D || g->use_cluster)
                for (int a = 0; a < 3; ++a) { r = ROW(); r[a] = 1.0f; r[KIN + 2 + a] = 1.0f; }
        } else if (what == 3) {
            if (g->with_distance) { r = ROW(); r[C] = 1.0f; }
        } else if (what == 4) {
            if (g->use_relative) {
                r = ROW(); r[0] = 1.0f; r[KIN] = 1.0f; r[G] = -g->lo[0];
                r = ROW(); r[1] = 1.0f; r[KIN + 1] = 1.0f; r[G] = -g->lo[1];
                r = ROW(); r[2] = 1.0f; r[G] = (float)((double)off_z - (double)g->lo[2]);
            }
        }
    }
#undef ROW
    *kin_out = KIN; *g_out = G;
    return j == g->c_in ? j : -1;
}

/* [wg_c | const_c] = W_c T (+ bias): fp64 products (exact) and sums in feature order, one rounding to fp32. */
void orc_fold_weights(const float *W, const float *bias, const float *T, int cin, int cout, int G, float *wg, float *cst) {
    const int S = ORC_MAX_G + 1;
    for (int c = 0; c < cout; ++c) {
        double acc[ORC_MAX_G + 1] = {0.0};
        for (int j = 0; j < cin; ++j) {
            const double w = (double)W[c * cin + j];
            for (int m = 0; m <= G; ++m) { const double pr = w * (double)T[(size_t)j * S + m]; acc[m] = acc[m] + pr; }
        }
        if (bias) acc[G] = acc[G] + (double)bias[c];
        for (int m = 0; m < G; ++m) wg[c * G + m] = (float)acc[m];
        cst[c] = (float)acc[G];
    }
}

/* Per-pillar constants q[p] = [cx, cy, -mean(dx), -mean(dy), -mean(dz)]: the pillar centre (:215-216) and the negated
 * mean of the rows' centre offsets d = xyz - centre -- i.e. centre - mean(xyz), evaluated on the small, exactly summable
 * offsets: fp64 sum of the fp32 d values (exact, order independent), one division, one rounding to fp32.
 * f_cluster = xyz - mean (:226-227) then is d + q[2:5]. */
void orc_pillar_consts(const float *pts, const orc_cfg *g, const int32_t *keep, const int32_t *inv, const int32_t *unq,
                       const int32_t *cnt, int64_t n, int64_t P, float *q) {
    const int32_t sxy = g->nx * g->ny, sy = g->ny;
    const int cols = g->cols;
    double *acc = (double *)calloc((size_t)(3 * P + 1), sizeof(double));
    for (int64_t p = 0; p < P; ++p) {
        const int32_t u = unq[p], cx = (u % sxy) / sy, cy = u % sy;
        q[5 * p] = (float)cx * g->vsz[0] + g->off[0];       /* (:215-216): separate mul and add roundings */
        q[5 * p + 1] = (float)cy * g->vsz[1] + g->off[1];
    }
    for (int64_t j = 0; j < n; ++j) {
        const float *r = pts + (int64_t)keep[j] * cols;
        const int64_t p = inv[j];
        const float dx = r[1] - q[5 * p], dy = r[2] - q[5 * p + 1], dz = r[3] - g->off[2];
        acc[3 * p] += (double)dx; acc[3 * p + 1] += (double)dy; acc[3 * p + 2] += (double)dz;
    }
    for (int64_t p = 0; p < P; ++p) {
        const double c = (double)(cnt[p] < 1 ? 1 : cnt[p]);
        for (int k = 0; k < 3; ++k) q[5 * p + 2 + k] = (float)(-(acc[3 * p + k] / c));
    }
    free(acc);
}

typedef struct { const float *pts; const orc_cfg *g; const int32_t *keep, *inv; const float *q, *wg, *cst; int KIN, G;
                 float *rin, *v, *x; } fold_ctx;
static void fold_range(int64_t lo, int64_t hi, void *p) {
    fold_ctx *c = (fold_ctx *)p;
    const int cols = c->g->cols, KIN = c->KIN, G = c->G, cout = c->g->c_out;
    for (int64_t j = lo; j < hi; ++j) {
        const float *r = c->pts + (int64_t)c->keep[j] * cols, *qp = c->q + 5 * (int64_t)c->inv[j];
        float *rin = c->rin + j * KIN;
        const float x = r[1], y = r[2], z = r[3];
        rin[0] = x - qp[0]; rin[1] = y - qp[1]; rin[2] = z - c->g->off[2];
        for (int k = 4; k < cols; ++k) rin[k - 1] = r[k];
        if (c->g->with_distance) rin[cols - 1] = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));   /* (:230-231) */
        for (int o = 0; o < cout; ++o) {
            const float *w = c->wg + o * G;
            float v = w[0] * rin[0];
            for (int k = 1; k < KIN; ++k) v = fmaf(w[k], rin[k], v);
            float u = c->cst[o];
            for (int k = 0; k < 5; ++k) u = fmaf(w[KIN + k], qp[k], u);
            c->v[j * cout + o] = v;
            c->x[j * cout + o] = v + u;
        }
    }
}
/* rin (n x KIN), v (n x cout), x = v + u (n x cout) */
void orc_forward_folded(const float *pts, const orc_cfg *g, const int32_t *keep, const int32_t *inv, const float *q, const float *wg,
                        const float *cst, int KIN, int G, int64_t n, float *rin, float *v, float *x) {
    fold_ctx c = {pts, g, keep, inv, q, wg, cst, KIN, G, rin, v, x};
    parallel_for(n, fold_range, &c);
}

/* S1 (G), S2 (G x G) of g_i = [rin_i | q_inv[i]] in fp64 (products of two fp32 values are exact). */
void orc_moments_folded(const float *rin, const float *q, const int32_t *inv, int64_t n, int KIN, int G, double *S1, double *S2) {
    for (int k = 0; k < G; ++k) S1[k] = 0.0;
    for (int k = 0; k < G * G; ++k) S2[k] = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        double gv[ORC_MAX_G];
        for (int k = 0; k < KIN; ++k) gv[k] = (double)rin[j * KIN + k];
        for (int k = 0; k < 5; ++k) gv[KIN + k] = (double)q[5 * (int64_t)inv[j] + k];
        for (int k = 0; k < G; ++k) {
            S1[k] += gv[k];
            for (int l = k; l < G; ++l) S2[k * G + l] += gv[k] * gv[l];
        }
    }
    for (int k = 0; k < G; ++k)
        for (int l = 0; l < k; ++l) S2[k * G + l] = S2[l * G + k];
}

void orc_bn_stats_from_moments(const double *S1, const double *S2, int64_t n, int G, const float *wg, const float *cst, int cout,
                               double *mean, double *var) {
    for (int c = 0; c < cout; ++c) {
        const float *w = wg + c * G;
        double m1 = 0.0, e2 = 0.0;
        for (int k = 0; k < G; ++k) {
            m1 += (double)w[k] * S1[k];
            double row = 0.0;
            for (int l = 0; l < G; ++l) row += S2[k * G + l] * (double)w[l];
            e2 += (double)w[k] * row;
        }
        double mc = n > 0 ? m1 / (double)n : 0.0;
        double vv = n > 0 ? e2 / (double)n - mc * mc : 0.0;
        mean[c] = n > 0 ? mc + (double)cst[c] : 0.0;
        var[c] = vv > 0.0 ? vv : 0.0;
    }
}

/* out / arg under the folded argmax rule (see the header of this section). */
void orc_act_max_folded(const float *x, const float *v, int64_t n, int cout, const float *scale, const float *shift,
                        const int32_t *inv, int64_t P, float *out, int32_t *arg) {
    float *best = (float *)malloc(sizeof(float) * (size_t)(P * cout + 1));
    int32_t *first = (int32_t *)malloc(sizeof(int32_t) * (size_t)(P + 1));
    for (int64_t q = 0; q < P * cout; ++q) { out[q] = -1.0f; arg[q] = (int32_t)n; best[q] = -INFINITY; }
    for (int64_t p = 0; p < P; ++p) first[p] = -1;
    for (int64_t j = 0; j < n; ++j) {
        const int64_t p = inv[j];
        if (first[p] < 0) first[p] = (int32_t)j;
        for (int c = 0; c < cout; ++c) {
            const float y = fmaf(x[j * cout + c], scale[c], shift[c]);
            const float z = y > 0.0f ? y : 0.0f;
            if (z > out[p * cout + c]) out[p * cout + c] = z;
            const float sv = scale[c] < 0.0f ? -v[j * cout + c] : v[j * cout + c];
            if (sv > best[p * cout + c]) { best[p * cout + c] = sv; arg[p * cout + c] = (int32_t)j; } /* strict: first index */
        }
    }
    for (int64_t p = 0; p < P; ++p)
        for (int c = 0; c < cout; ++c)
            if (!(out[p * cout + c] > 0.0f)) arg[p * cout + c] = first[p];
    free(best);
    free(first);
}
