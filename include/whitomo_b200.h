/*
 * whitomo_b200.h -- C ABI of the B200-native (sm_100a) reconstruction engine.
 *
 * This is the drop-in boundary for the reference's operator layer.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repository root).  The reference
 * reaches ASTRA's single-threaded CPU algorithms through src/astra_proxy.py; a maintainer binds these
 * symbols with ctypes instead (whitomo_b200/_capi.py is that binding).
 *
 * Conventions
 *  - plain C types only; all data pointers are caller-owned DEVICE pointers (fp32, C order);
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, performs no
 *    allocation and no synchronisation (handles own a few KB of per-angle constants, created once);
 *  - the caller provides scratch (`ws`, `ws_bytes`) sized by wt_workspace_bytes();
 *  - return 0 on success, a negative WT_ERR_* otherwise; wt_last_error() gives the message
 *    (thread-local); no C++ exception crosses the boundary;
 *  - "images" are independent 2D problems sharing one geometry: slices of a stack and/or the K
 *    element maps of the white-beam model.  Volumes are (nimages, rows, cols), sinograms
 *    (nimages, nangles, ndet), both dense.
 */
#ifndef WHITOMO_B200_H
#define WHITOMO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_OK 0
#define WT_ERR_INVALID -1      /* bad argument (null pointer, non-positive size, misaligned pointer) */
#define WT_ERR_UNSUPPORTED -2  /* geometry the kernels do not implement (det_width < 1, non-parallel) */
#define WT_ERR_WORKSPACE -3    /* ws_bytes smaller than wt_workspace_bytes() */
#define WT_ERR_CUDA -4         /* a CUDA runtime call failed; message in wt_last_error() */

typedef struct wt_geom wt_geom;         /* parallel-beam geometry + per-angle device constants */
typedef struct wt_spectrum wt_spectrum; /* polychromatic source / absorption tables on the device */

const char *wt_last_error(void);
int wt_version(void);
/* Diagnostic (no counterpart in the reference): shared-memory gather indices of K1/K2 found outside their staged
 * window since the library was loaded.  Always 0 unless the library was built with -DWT_BOUNDS_CHECK=1 (the checks
 * cost ~10 % and are compiled out of the product build); the GPU test suite is run once against such a build in place
 * of compute-sanitizer, which is not available on the measurement pool.  Synchronises the device. */
long long wt_debug_oob_count(void);

/* Diagnostic, host only (no CUDA call): what wt_geom_create plans for a geometry -- the forward projector's angle
 * tiles (consecutive angles of one orientation that share a staging window; first angle, angle count, orientation) and
 * the order the tiles are launched in (phase by phase; phase_start holds 5 prefix offsets into launch_order).  Arrays
 * may be null; at most max_tiles entries are written, *n_tiles receives the tile count. */
int wt_geom_plan(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, int max_tiles,
                 int *tile_first, int *tile_count, int *tile_vertical, int *launch_order, int *n_tiles, int *n_phases,
                 int *phase_start);

/* Diagnostic, host only: how small problems are spread over the SMs.  A geometry with fewer (angle tile, detector
 * block) pairs than the SMs hold CTAs has each pair's lines split over a thread-block cluster of *fp_split CTAs (K1); a
 * volume with fewer pixel tiles than SMs has each tile's angles split over *bp_split CTAs (K2); rank 0 of the cluster adds
 * the partial sums in rank order over distributed shared memory, so results stay deterministic.  1 = no split.  Both
 * depend on the geometry alone.  n_tiles: the tile count wt_geom_plan reports. */
int wt_split_plan(int rows, int cols, int ndet, int nangles, int n_tiles, int *fp_split, int *bp_split, int *bp_tile_rows);

/* Replaces astra.create_proj_geom('parallel', det_width, ndet, angles) + astra.create_vol_geom(rows, cols)
 * + astra.create_projector('linear', pg, vg)  (src/white_projection.py:41-63, src/white_rec.py:66-74,
 * teeth_recon/experiments.py:228-235).  `angles` is a HOST array (radians, cast to float32 like ASTRA). */
int wt_geom_create(int rows, int cols, int ndet, float det_width, const float *angles, int nangles,
                   wt_geom **out);
/* Position arithmetic of both projectors.
 *   WT_POS_ASTRA (default): the reference's.  ASTRA's 'linear' projector carries a ray's position as a float32 running
 *     sum, c += delta once per line (the projector behind src/astra_proxy.py:35-66), which drifts away from
 *     c_0 + l * delta by up to N * ulp(N) / 2 positions (0.1 pixel at 2048^2, 2.5 at 8192^2).  K1 repeats that sum
 *     literally; K2 inverts its closed form.  This is the mode that matches the reference to 1e-4 at every size.
 *   WT_POS_EXACT: c_0 + l * delta evaluated in closed form (closer to the continuous operator, 1e-4 .. 1e-2 away from
 *     the reference at N >= 1024); enables the mirror-paired single-image kernels (1.7x faster for 1-2 images).
 * wt_geom_create uses WT_POS_ASTRA unless the environment variable WT_POSITIONS is "exact". */
#define WT_POS_ASTRA 0
#define WT_POS_EXACT 1
int wt_geom_create_ex(int rows, int cols, int ndet, float det_width, const float *angles, int nangles, int positions,
                      wt_geom **out);
int wt_geom_positions(const wt_geom *g);
void wt_geom_destroy(wt_geom *g);
int wt_geom_dims(const wt_geom *g, int *rows, int *cols, int *ndet, int *nangles);

/* bytes of scratch any call below needs for `nimages` images */
size_t wt_workspace_bytes(const wt_geom *g, int nimages);

/* p = W x.  Replaces gpu_fp (src/astra_proxy.py:35-49: ASTRA 'FP' with the 'linear' projector).
 * vol (nimages, rows, cols) -> sino (nimages, nangles, ndet).  Only sinogram rows [angle_begin, angle_end) are
 * computed and written (the angle-range split of one oversized slice); pass 0, nangles for the full operator. */
int wt_fp(const wt_geom *g, const float *vol, float *sino, int nimages, int angle_begin, int angle_end,
          void *ws, size_t ws_bytes, void *stream);

/* g = scale * W^T q (exact transpose).  Replaces gpu_bp (src/astra_proxy.py:52-66: ASTRA 'BP'); scale = 1 is
 * the plain operator, -1 gives the white-beam gradient sign (src/white_projection.py:112), pi/(2A) the FBP
 * normalisation (src/astra_proxy.py:87-102).  sino (nimages, nangles, ndet) -> vol (nimages, rows, cols).
 * Only angles [angle_begin, angle_end) contribute (the angle-range split of one oversized slice); pass
 * 0, nangles for the full operator. */
int wt_bp(const wt_geom *g, const float *sino, float *vol, int nimages, int angle_begin, int angle_end,
          float scale, void *ws, size_t ws_bytes, void *stream);

/* The same operator, row block by row block: only rows [row_begin, row_end) of the rows the backprojector assigns a
 * thread to are computed and written (row_end < 0: through the last one), so that a caller can hand finished blocks to
 * a collective while later blocks are still being computed (the angle-range split of one oversized slice sums the
 * ranks' partial backprojections: whitomo_b200/distributed.py).  wt_bp_row_info tells how many rows are processed
 * (all `rows`; under WT_POS_EXACT with one or two large images only the upper half, rows [0, ceil(rows/2)), and a
 * block then also writes its point reflection, rows [rows - row_end, rows - row_begin)), the alignment row_begin must
 * have, and whether blocks are mirrored.  reuse_packed != 0: `ws` still holds the packed sinogram of a previous call
 * with the same sino / nimages / angle range (skips the re-pack). */
int wt_bp_rows(const wt_geom *g, const float *sino, float *vol, int nimages, int angle_begin, int angle_end,
               int row_begin, int row_end, int reuse_packed, float scale, void *ws, size_t ws_bytes, void *stream);
int wt_bp_row_info(const wt_geom *g, int nimages, int *rows_processed, int *row_align, int *mirrored);

/* x = SIRT(p, n_iters), x0 = 0.  Replaces cpu_sirt (src/astra_proxy.py:69-84: ASTRA 'SIRT').
 * sino (nimages, nangles, ndet) -> vol (nimages, rows, cols). */
int wt_sirt(const wt_geom *g, const float *sino, float *vol, int nimages, int n_iters, void *ws,
            size_t ws_bytes, void *stream);
size_t wt_sirt_workspace_bytes(const wt_geom *g, int nimages);

/* Polychromatic tables.  Replaces the constructor state of WhiteProjection
 * (src/white_projection.py:23-73): source[E] (already normalised by its integral), bin widths w[E]
 * (= energy_grid[:,1]), element_absorptions[K][E], pixel_size.  HOST arrays, float64. */
int wt_spectrum_create(int n_elements, int n_bins, const double *source, const double *widths,
                       const double *absorptions, double pixel_size, wt_spectrum **out);
void wt_spectrum_destroy(wt_spectrum *s);

/* White-beam forward model + residual, fused into the projector epilogue.
 * Replaces FP_white + calc_mu + the residual of func/grad
 * (src/white_projection.py:86-98,116-127,130-139; src/white_rec.py:111-148):
 *   P_k = W c_k;  I[r] = sum_e s_e w_e exp(-pix sum_k ea[k,e] P_k[r]);  q = meas - I;
 *   mu_k[r] = -pix sum_e s_e w_e ea[k,e] exp(...);  qmu_k = q * mu_k;  loss = 1/2 sum_r q^2.
 * conc (nslices, K, rows, cols); meas (nslices, nangles, ndet) or NULL (then q = -I);
 * outputs, each optional (NULL to skip): intensity (nslices, nangles, ndet), qmu (nslices, K, nangles, ndet),
 * loss (nslices) as float64.  Only rays of angles [angle_begin, angle_end) are computed, written and summed into
 * the loss (angle-range split: the model is local to a ray, the partial losses of the ranks add up). */
int wt_white_fp(const wt_geom *g, const wt_spectrum *s, const float *conc, const float *meas,
                float *intensity, float *qmu, double *loss, int nslices, int angle_begin, int angle_end,
                void *ws, size_t ws_bytes, void *stream);

/* Mono metal-artifact-reduction residual, fused into the projector epilogue.  Replaces, per inner iteration of
 * barrier_least_squares (teeth_recon/experiments.py:651-749), the FP inside cost/grad (:712-722) and the FP inside
 * HoughBarrier.eval (teeth_recon/barrier.py:50-69).  With P = W x, per image i and ray r:
 *   good ray  (mask == 0):  u = 2 (P - b) * inv_ngood[i]                (the reference's grad = 2 BP(d / n_good), B8)
 *   metal ray (mask != 0):  u = -bt / (P - b_hb)                        (beta*t * grad phi of  b_hb - P <= 0)
 * so that ONE backprojection W^T u is  grad cost + beta*t*grad phi_hough.  stats (nimages, 3) float64, optional:
 *   [0] cost = sum_good ((P - b) inv_ngood)^2,  [1] sum_metal -log(P - b_hb) (+inf if infeasible),
 *   [2] number of infeasible metal rays (P < b_hb).
 * vol (nimages, rows, cols); b (nimages, nangles, ndet); mask (nimages, nangles, ndet) bytes; inv_ngood (nimages);
 * u (nimages, nangles, ndet), optional; proj (nimages, nangles, ndet), optional plain projections. */
int wt_mar_fp(const wt_geom *g, const float *vol, const float *b, const unsigned char *mask,
              const float *inv_ngood, float bt, float b_hb, float *u, float *proj, double *stats, int nimages,
              void *ws, size_t ws_bytes, void *stream);

/* K3: fused gradient-combine + regularisers + log-barrier + step + projection, one streaming pass.
 * Replaces, per inner iteration,
 *   - src/barrier_method.py:164-206 with the (f, g, proj) box tuples of src/barrier_method_gogo.py:79-94 and
 *     the |c0 c1| regulariser (calc_uneq_reg_grad, src/white_rec.py:180-186);
 *   - the 'morezero' constraint of the mono MAR run (teeth_recon/experiments.py:704-706) with lower_only;
 *   - the plain gradient step of Iteration (src/white_rec.py:223-242) with t = 0, reg_w = triv_w = 1.
 * For every slice s, element k, pixel i (x, grad: (nslices, K, npix), x updated in place):
 *   G  = grad + beta_reg*(reg_w * d|c0 c1|/dc_k + triv_w * c(c(4c-3)+1)) + beta_reg*t*(-1/x [+ 1/(1-x)])
 *   x' = x - alpha*G          if bit k of elem_mask is set and the slice is active
 *   x' = clip(x', 0, 1 | +inf) if mode != WT_UPD_STEP_ONLY
 * t == 0 drops the barrier terms altogether (no 0*inf).  mode WT_UPD_CLIP_ONLY only projects.
 * active (nslices bytes, optional): slices with active[s] == 0 are left untouched (per-slice early stop).
 * penalty (nslices float64, optional): sum_k,i beta_reg*reg_w*|c0 c1| + beta_reg*t*phi(x') at the NEW x -- the
 * terms barrier_method adds to f(x) for its early-stop test (src/barrier_method.py:208-213); +inf on the boundary. */
#define WT_UPD_STEP_CLIP 0
#define WT_UPD_STEP_ONLY 1
#define WT_UPD_CLIP_ONLY 2
typedef struct {
    float alpha;
    float beta_reg;
    float t;
    float reg_w;
    float triv_w;
    int lower_only;
    int mode;
    unsigned elem_mask;
} wt_update_params;

/* One whole inner iteration of the white-beam loops in three launches: K2 (backprojection of the packed q*mu of the
 * previous forward model, its epilogue taking the wt_barrier_update step on x and refreshing the forward projector's
 * packed copies of x), K1 (white-beam forward model at the new x, its epilogue writing q*mu straight into K2's packed
 * layout and the partial sums of the data term), one reduction.  Replaces, per inner iteration, the same reference
 * code as wt_white_fp + wt_bp(scale = -1) + wt_barrier_update together (src/barrier_method.py:163-229 with
 * src/white_projection.py:116-127; src/white_rec.py:199-244) -- without the gradient array, the q*mu array and the two
 * re-packing passes those separate calls exchange.  n_elements K <= 2; not for mirror-paired calls.
 *   flags & WT_STEP_INIT: (re)start -- pack x, run the forward model only; loss = f(x), no step, penalty untouched.
 *   otherwise:            x <- step(x, -W^T(q mu)) with *prm (see wt_barrier_update), then the forward model at the new
 *                         x;  loss (nslices, float64, optional) = f(x_new),  penalty (optional) = the wt_barrier_update
 *                         penalty at x_new.
 * State (the packed x and q*mu) lives in `ws` between calls: the same `ws` (>= wt_step_workspace_bytes) must be passed,
 * untouched, from a WT_STEP_INIT call on; any other use of it, a change of nslices, or writing x from outside
 * requires a new WT_STEP_INIT. */
#define WT_STEP_INIT 1
#define WT_STEP_TIMED 2   /* diagnostic: record CUDA events around K2 and K1 of this step (see wt_step_times) */
#define WT_STEP_RING 64
size_t wt_step_workspace_bytes(const wt_geom *g, int nslices, int n_elements);
int wt_white_barrier_step(const wt_geom *g, const wt_spectrum *s, float *x, const float *meas,
                          const wt_update_params *prm, const unsigned char *active, double *loss, double *penalty,
                          int nslices, int flags, void *ws, size_t ws_bytes, void *stream);

/* Diagnostic (no counterpart in the reference): device durations of K2 (with the update epilogue) and K1 (white-beam
 * epilogue) of the most recent WT_STEP_TIMED steps on this geometry, oldest first, at most max_steps (<= WT_STEP_RING);
 * waits for those steps to finish.  Returns the number of steps written, or a negative error. */
int wt_step_times(const wt_geom *g, int max_steps, float *bp_ms, float *fp_ms);

int wt_barrier_update(float *x, const float *grad, int nslices, int n_elements, size_t npix,
                      const wt_update_params *prm, const unsigned char *active, double *penalty, void *ws,
                      size_t ws_bytes, void *stream);

/* sum_i (c0_i c1_i)^2 of every slice, as WT_PAIR_NORM_BLOCKS float64 partial sums per slice in a fixed summation order
 * (the caller adds them, e.g. once after its loop): the dissimilarity term ||beta c0 c1|| of the loss the reference's
 * Iteration reports every step (src/white_rec.py:236-237: np.linalg.norm(beta_reg * c[0] * c[1])) -- one launch, no
 * temporaries, no host sync.  conc: (nslices, 2, npix); out: (nslices, WT_PAIR_NORM_BLOCKS) float64. */
#define WT_PAIR_NORM_BLOCKS 32
int wt_pair_norm2(const float *conc, int nslices, size_t npix, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
