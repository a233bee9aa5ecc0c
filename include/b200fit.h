/* b200fit.h -- C ABI of the B200-native per-pixel spectral fitter (libb200fit.so).
 *
 * The reference (mcyc/mufasa) is pure Python and has no FFI for this path; its plugin boundary is
 *     pcube.specfit.Registry.add_fitter(fittype, fitter, npars)        mufasa/multi_v_fit.py:119-120
 *     pcube.fiteach(fittype=, guesses=, limitedmin/max=, minpars/maxpars=, maskmap=, ...)
 *                                                                       mufasa/multi_v_fit.py:374-393, :192-206
 *     pcube.get_modelcube()                                             mufasa/UltraCube.py:324, :978
 *     UltraCube.get_rss / get_AICc / get_all_lnk_maps / get_best_2c_parcube
 *                                                                       mufasa/UltraCube.py:432-498, :1218-1379
 * and, either side of the fit, the voxel work of
 *     signals.get_moments / get_rms_robust / get_snr                    mufasa/signals.py:24-121, :329-396
 *     convolve_tools.convolve_sky_byfactor                              mufasa/convolve_tools.py:37-138
 * Each entry point below names the reference call it stands in for.  A reference-side binding is a
 * ctypes stub.  Plain pointers and sizes only; no torch / CUDA types in signatures
 * (``stream`` is a cudaStream_t passed as void*; NULL = default stream).
 *
 * Conventions
 *   - Every function returns 0 on success, <0 on error (B200FIT_E_*).  Errors are per call, never per pixel:
 *     per-pixel outcomes are reported in ``status[]`` (MINPACK-style, see b200fit_fit).
 *   - Pointers named d_* are DEVICE pointers owned by the caller; h_* are HOST pointers.
 *   - Pixel-major device layout: spectra[npix][nchan_stride] fp32, velocity ascending,
 *     nchan_stride = b200fit_nchan_stride(nchan) (multiple of 32; pad channels hold NaN).
 *   - Parameter order per pixel: [v1, sigma1, Tex1, tau1, (v2, sigma2, Tex2, tau2)]   BaseModel.py:71-73;
 *     component 1 is the rear slab (HyperfineModel.py:46-56).
 */
#ifndef B200FIT_H
#define B200FIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FIT_ABI_VERSION 5

enum {
    B200FIT_OK = 0,
    B200FIT_E_ARG = -1,      /* bad argument (NULL pointer, ncomp not in {1,2}, nchan too large, ...) */
    B200FIT_E_CUDA = -2,     /* a CUDA runtime call failed; see b200fit_last_error()                  */
    B200FIT_E_NOGPU = -3,    /* no usable CUDA device: there is NO CPU fallback                        */
    B200FIT_E_MODEL = -4     /* unknown model_id (FitTypeError on the Python side, meta_model.py:135) */
};

enum { B200FIT_MODEL_NH3_11 = 0, B200FIT_MODEL_N2HP_10 = 1 };
#define B200FIT_MAX_NCHAN 2048
#define B200FIT_MAX_PAR 8

typedef struct b200fit_ctx b200fit_ctx;      /* one per device */

/* Context: replaces the implicit per-process state of pyspeckit's fitter Registry (multi_v_fit.py:119-120). */
int b200fit_create(b200fit_ctx** out, int device);
void b200fit_destroy(b200fit_ctx* ctx);
const char* b200fit_last_error(const b200fit_ctx* ctx);
int b200fit_abi_version(void);
int b200fit_num_sms(const b200fit_ctx* ctx);

/* Measurement hooks (no reference analogue; bench.py's roofline and launch accounting).
 *   b200fit_launch_count: kernels launched by this context since creation.
 *   b200fit_host_launch_count: launches the HOST issued for them (a replayed CUDA graph of a round group counts once).
 *   b200fit_measure_xu_peak: ex2.approx operations per second of this GPU (micro-benchmark: 8 independent chains per
 *     thread, 8 blocks of 256 threads per SM) -- the measured peak of the pipe that bounds the fit's evaluation kernel.
 *   b200fit_set_profiling(on): record CUDA events around the kernels of b200fit_fit on the launching stream.
 *   b200fit_get_profile: waits for the last profiled fit and returns out[4] = {prologue ms, sum of evaluation-kernel
 *   ms, sum of solver-kernel ms, rounds launched}. */
uint64_t b200fit_launch_count(const b200fit_ctx* ctx);
uint64_t b200fit_host_launch_count(const b200fit_ctx* ctx);
int b200fit_measure_xu_peak(b200fit_ctx* ctx, double* ex2_per_second, void* stream);
int b200fit_set_profiling(b200fit_ctx* ctx, int32_t on);
int b200fit_get_profile(b200fit_ctx* ctx, double* out);

/* Description of one fit call == the arguments MUFASA hands to pcube.fiteach (multi_v_fit.py:374-383) plus the
 * spectral axis pyspeckit takes from pcube.xarr. */
typedef struct {
    int32_t model_id;        /* B200FIT_MODEL_*  ('nh3_multi_v' / 'n2hp_multi_v', meta_model.py:107,118)   */
    int32_t ncomp;           /* 1 or 2 -> P = 4*ncomp                                                        */
    int32_t nchan;           /* channels per spectrum (<= B200FIT_MAX_NCHAN)                                 */
    int32_t max_iter;        /* accepted-step cap, mpfit maxiter = 200                                       */
    int64_t npix;            /* spectra in this call (already masked / compacted)                            */
    double v0, dv;           /* V_i = v0 + i*dv, km/s, radio convention, dv > 0 (ascending; the transpose
                                kernel flips descending cubes)                                               */
    double nu_ref_ghz;       /* rest frequency of the cube's velocity axis (0 => the model's nu0,
                                multi_v_fit.py:109-113)                                                      */
    double lo[B200FIT_MAX_PAR], hi[B200FIT_MAX_PAR];   /* minpars / maxpars, all limits active (:376-379)    */
    double ftol;             /* stop when predicted chi2 reduction <= ftol*chi2 (mpfit ftol = 1e-10)         */
    double signal_cut;       /* pyspeckit fiteach signal_cut: skip a pixel whose max(y/err) < cut; 0 = off   */
    int32_t weight_mode;     /* 0 = err := std(finite channels) per pixel (what fiteach does for MUFASA);
                                1 = use d_err[pix]                                                           */
    int32_t reserved;
} b200fit_problem;

int64_t b200fit_nchan_stride(int32_t nchan);
size_t b200fit_workspace_bytes(const b200fit_problem* prob);

/* Cube layout pass: FITS order [nchan][ny][nx] (channel-major) -> compacted pixel-major rows.
 * Replaces pcube.get_spectrum(x, y) per pixel inside Cube.fiteach.  pix_index[k] = flat index y*nx+x of the
 * k-th selected pixel (host builds it from maskmap, multi_v_fit.py:347-350).  flip != 0 reverses the channel
 * order (descending-velocity cubes). */
int b200fit_gather_spectra(b200fit_ctx* ctx, const float* d_cube, int32_t nchan, int64_t nplane,
                           const int64_t* d_pix_index, int64_t npix, int32_t flip,
                           float* d_spectra, int64_t nchan_stride, void* stream);

/* The fit: replaces pcube.fiteach(...) -> Specfit -> SpectralModel.fitter -> mpfit for every pixel.
 *   d_row_index [npix] int64 or NULL: pixel k reads spectrum row d_row_index[k] (NULL = row k).  Lets masked refits
 *              (cubefit_simp on a maskmap, multi_v_fit.py:126-208) run on the resident rows without a re-gather.
 *   d_guesses [npix][P] fp64   must lie inside [lo,hi] (else status 0 like mpfit); NaN => status 0
 *   d_err     [npix] fp64 or NULL (weight_mode 0)
 *   d_params  [npix][P] fp64   zeros on failure              (parcube)
 *   d_perr    [npix][P] fp64   0 for pegged parameters       (errcube = mpfit perror)
 *   d_chi2    [npix] fp64      sum(((y-M)/err)^2) at the returned parameters
 *   d_err_used[npix] fp64      the weight actually used
 *   d_status  [npix] int32     1,2,3,4 converged (MINPACK meaning), 6 = converged to the precision of the fp32
 *                              model evaluation (MINPACK info 6, "ftol too small"), 5 = max_iter, 7 = xtol too small,
 *                              0 = not fitted / bad guess, -16 = non-finite; has_fit = status > 0
 *   d_counts  [npix][4] uint32 {evaluations, rejected steps, executed gaussian-lane evals / 32, active chunks}
 *                              (roofline accounting); may be NULL
 *   d_workspace: b200fit_workspace_bytes(prob) bytes of device scratch (two active lists + ~0.6 / 1.1 KB of LM state
 *              per pixel for ncomp 1 / 2)
 * The fit is a loop of rounds {evaluation kernel over the active pixels; solver kernel}; the call polls the
 * active count through a pinned counter and returns once every pixel finished (kernels may still be draining on
 * ``stream``: results are ordered after them on that stream). */
int b200fit_fit(b200fit_ctx* ctx, const b200fit_problem* prob,
                const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                const double* d_guesses, const double* d_err,
                double* d_params, double* d_perr, double* d_chi2, double* d_err_used,
                int32_t* d_status, uint32_t* d_counts,
                void* d_workspace, void* stream);

/* Several fits in flight at once (typically the 1- and the 2-component fit of the same spectra, which
 * UltraCube.fit_cube loops over, UltraCube.py:312-320): the rounds of the jobs are interleaved on internal streams, so
 * the latency-bound late rounds of one fit overlap the other's work.  Results are identical to separate b200fit_fit
 * calls.  Each job needs its own workspace and output buffers; ``stream`` orders the whole batch (fork / join). */
#define B200FIT_MAX_BATCH 8
typedef struct {
    const b200fit_problem* prob;
    const float* d_spectra;
    int64_t nchan_stride;
    const int64_t* d_row_index;
    const double* d_guesses;
    const double* d_err;
    double* d_params;
    double* d_perr;
    double* d_chi2;
    double* d_err_used;
    int32_t* d_status;
    uint32_t* d_counts;
    void* d_workspace;
    void* wait_event;   /* cudaEvent_t or NULL: this job's kernels are ordered after the event (inputs of the job that
                         * are still being uploaded on another stream, see b200fit_upload_slab) */
} b200fit_fit_args;
int b200fit_fit_batch(b200fit_ctx* ctx, const b200fit_fit_args* jobs, int32_t njobs, void* stream);

/* Host -> device copy of a row slab of a FITS-order cube [nchan][ny][nx] (pinned host memory): the ``count``
 * elements starting ``first`` elements into every plane, i.e. h_cube[c * plane + first + k] -> d_dst[c * count + k]
 * (one cudaMemcpy2DAsync on ``stream``).  Lets the fit of the first rows start while the rest of the cube is still
 * crossing PCIe; replaces the single read of the cube file in UltraCube.load_pcube (mufasa/UltraCube.py:185-200). */
int b200fit_upload_slab(b200fit_ctx* ctx, const float* h_cube, int32_t nchan, int64_t plane, int64_t first, int64_t count,
                        float* d_dst, void* stream);

/* Model evaluation: replaces pcube.get_modelcube() (one multi_v_spectrum call per pixel,
 * HyperfineModel.py:22-109), in fp64 with the reference's frequency-space operation order (except that the gaussian
 * argument of a hyperfine term multiplies by a per-line reciprocal and uses a lean exp: last-bit differences).
 *   d_params [npix][P]; all-zero or non-finite rows produce NaN rows (pyspeckit convention)
 *   d_model  [npix][nchan_stride] fp64 */
int b200fit_model(b200fit_ctx* ctx, const b200fit_problem* prob, const double* d_params,
                  double* d_model, int64_t nchan_stride, void* stream);

/* Model selection: replaces UltraCube.get_all_lnk_maps + get_best_2c_parcube's selection
 * (UltraCube.py:1218-1379 => get_rss :1384-1475, expand_mask :1658-1698, aic.py:136-201).
 *   d_params1 [npix][4], d_params2 [npix][8]  (all-zero row = no fit); d_row_index as in b200fit_fit
 *   model_f32: != 0 => model cube takes the data cube's dtype (fp32), as pyspeckit's full_like does
 *   d_rss [npix][3] (ncomp 0,1,2), d_nsamp [npix], d_lnk [npix][3] (lnk10, lnk20, lnk21),
 *   d_ncomp [npix] int8 (0/1/2 with thresholds thr[3] = {lnk10, lnk20, lnk21}; NULL = 5,5,5)
 * The include_nosamp rule (UltraCube.py:1423-1438) is GLOBAL: pixels whose own model mask is empty borrow the
 * spectral mask of the first (row-major) pixel with the largest mask.  It is therefore split in two passes:
 *   pass 1  only_empty = 0, d_fill_bits = NULL: all pixels; d_mask_counts[npix] (int32, out) = own mask sizes;
 *           empty-mask pixels get NaN statistics for now
 *   ......  b200fit_mask_argmax over d_mask_counts (and over all ranks on the host side), b200fit_mask_bits
 *   pass 2  only_empty = 1, d_fill_bits = the 64-word bit set: only pixels with d_mask_counts == 0 are redone.
 * Passing d_fill_bits in pass 1 applies a caller-chosen fill mask directly. */
int b200fit_aicc(b200fit_ctx* ctx, const b200fit_problem* prob,
                 const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                 const double* d_params1, const double* d_params2,
                 const uint32_t* d_fill_bits, int32_t only_empty, int32_t expand, int32_t model_f32, const double* thr,
                 double* d_rss, double* d_nsamp, double* d_lnk, signed char* d_ncomp, int32_t* d_mask_counts,
                 void* stream);

/* The same comparison for any pair of fits A (ncomp_a components, d_params1 [npix][4 ncomp_a]) and B (ncomp_b,
 * d_params2 [npix][4 ncomp_b]) of the same spectra -- e.g. a NEW fit against the OLD one with the same number of
 * components, expand = 0, which is what replace_bad_pix decides on (master_fitter.py:1449 ->
 * calc_AICc_likelihood(ucube_new, nc, nc, ucube_B=old), UltraCube.py:1120-1216).  Outputs: d_rss [npix][3] = rss of
 * the zero model, A, B over the union mask; d_lnk [npix][3] = lnk(A, 0), lnk(B, 0), lnk(B, A); d_ncomp = ncomp_b where
 * lnk(B, A) > thr[2] and lnk(B, 0) > thr[1], else ncomp_a, and 0 where lnk(A, 0) <= thr[0]. */
int b200fit_aicc_pair(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                      const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                      const double* d_params1, const double* d_params2, const uint32_t* d_fill_bits,
                      int32_t only_empty, int32_t expand, int32_t model_f32, const double* thr, double* d_rss,
                      double* d_nsamp, double* d_lnk, signed char* d_ncomp, int32_t* d_mask_counts, void* stream);
/* The same with the caller's spectral mask: d_mask_bits [npix][B200FIT_MAX_NCHAN / 32] (bit b of word w = channel
 * 32 w + b; NULL = b200fit_aicc_pair) replaces (model A > 0) | (model B > 0) as the un-dilated mask of every pixel --
 * UltraCube.get_rss(ncomp, mask=...) / get_AICc(..., mask=...), mufasa/UltraCube.py:432-451, :474-486.  The
 * include_nosamp rule and the dilation apply to it as they do to the model mask (:1423-1451). */
int b200fit_aicc_pair_masked(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                             const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                             const double* d_params1, const double* d_params2, const uint32_t* d_mask_bits,
                             const uint32_t* d_fill_bits, int32_t only_empty, int32_t expand, int32_t model_f32,
                             const double* thr, double* d_rss, double* d_nsamp, double* d_lnk, signed char* d_ncomp,
                             int32_t* d_mask_counts, void* stream);
int b200fit_mask_bits_pair(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                           const double* d_params1, const double* d_params2, int32_t model_f32, const int64_t* d_best,
                           int64_t pix, uint32_t* d_bits, void* stream);

/* d_best[2] (int64, device) = {largest mask count, index_offset + lowest pixel index holding it} -- the argmax the
 * reference takes with np.where(nsamp_map == nanmax(nsamp_map)) (UltraCube.py:1428-1429).  index_offset lets
 * each rank of a partitioned cube report global pixel indices. */
int b200fit_mask_argmax(b200fit_ctx* ctx, const int32_t* d_mask_counts, int64_t npix, int64_t index_offset,
                        int64_t* d_best, void* d_workspace, void* stream);

/* Un-dilated union mask (model1 > 0) | (model2 > 0) of one pixel as a bit set d_bits[64] (bit b of word w =
 * channel 32 w + b).  The pixel is d_best[1] when d_best != NULL (empty set if d_best[0] == 0), else ``pix``. */
int b200fit_mask_bits(b200fit_ctx* ctx, const b200fit_problem* prob,
                      const double* d_params1, const double* d_params2, int32_t model_f32,
                      const int64_t* d_best, int64_t pix, uint32_t* d_bits, void* stream);

/* The guess conditioning of cubefit_simp (multi_v_fit.py:156-190) on the device: a velocity guess of exactly 0 becomes
 * NaN ("no guess", :158-159), every value outside [lo, hi] is moved just inside (lo + eps / hi - eps, :178-190).
 *   d_raw [npar][npix] field-major (the caller's [npar, ny, nx] maps), d_guesses [npix][npar] (input of b200fit_fit);
 *   lo / hi: HOST arrays of npar doubles (minpars / maxpars); npar = 4 or 8.  Pixels left with a non-finite guess are
 *   reported "not fitted" by b200fit_fit (status 0, zeros), which is what the reference's maskmap does to them. */
int b200fit_prepare_guesses(b200fit_ctx* ctx, int64_t npix, int32_t npar, const double* d_raw, const double* lo,
                            const double* hi, double eps, double* d_guesses, void* stream);

/* UltraCube.get_best_2c_parcube (UltraCube.py:1283-1379; include_1c=True, return_lnks=True) assembled on the device
 * from the two fits and the output of b200fit_aicc, field-major d_out[B200FIT_PACK_FIELDS][out_stride]:
 *   [0,4) parcube1  [4,8) errcube1  [8,16) parcube2  [16,24) errcube2  [24,32) best parcube  [32,40) best errcube
 *   (the 2-component fit where ncomp == 2; the 1-component fit in planes 0-3, NaN in 4-7 where ncomp == 1; NaN where
 *   ncomp == 0 -- :1350-1367)  [40,43) lnk10, lnk20, lnk21  [43] ncomp.   These rows are what a row-slab partitioned
 *   run gathers over NCCL (mufasa_b200/dist.py); d_lnk [npix][3] and d_ncomp [npix] come from b200fit_aicc. */
#define B200FIT_PACK_FIELDS 44
int b200fit_pack_selection(b200fit_ctx* ctx, int64_t npix, const double* d_params1, const double* d_perr1,
                           const double* d_params2, const double* d_perr2, const double* d_lnk,
                           const signed char* d_ncomp, double* d_out, int64_t out_stride, void* stream);

/* Residual statistics of one fitted model: replaces UltraCube.get_rms / get_chisq(reduced=True)
 * (UltraCube.py:1701-1740, :1478-1565; prob->ncomp selects the model, d_params [npix][4*ncomp]).
 *   d_rms [npix] = 1.4826*nanmedian(|r - roll(r,2)|)/sqrt(2);  d_chisq_red [npix]; NaN where there is no fit */
int b200fit_resid_stats(b200fit_ctx* ctx, const b200fit_problem* prob,
                        const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                        const double* d_params, int32_t expand, int32_t model_f32,
                        double* d_rms, double* d_chisq_red, void* stream);

/* Robust noise and peak per pixel -- the step right before the fit in cubefit_gen / cubefit_simp:
 * replaces signals.get_rms_robust (mufasa/signals.py:48-121: mad_std, then refine_rms with the sigma_cut signal mask
 * dilated by ``expand`` channels), the per-pixel half of signals.get_snr (:24-45) and multi_v_fit.snr_estimate's
 * errmap = mad_std(cube, axis=0) (:677-702).  fp64, bit-identical to numpy on a float64 copy of the spectrum.
 *   d_planemask [npix] uint8 or NULL: pixels excluded by trim_cube_edge (signals.py:199-226) -> NaN
 *   d_rms0 [npix] = 1.4826 nanmedian|y - nanmedian y|  (first pass);  d_rms [npix] = the refined value;
 *   d_peak [npix] = max over the included channels (peak S/N = d_peak / d_rms) */
int b200fit_rms_snr(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                    const int64_t* d_row_index, const unsigned char* d_planemask, double sigma_cut, int32_t expand,
                    double* d_rms0, double* d_rms, double* d_peak, void* stream);

/* Moment / signal-mask pre-pass of cubefit_gen (signals.get_moments, mufasa/signals.py:329-396) as primitives on
 * BIT CUBES: bits[ny*nx][nchan_stride/32] uint32, bit b of word w <-> channel 32 w + b, pixels in row-major plane
 * order (NOT compacted: the morphology needs the neighbours).  Composition: mufasa_b200/prepass.py.
 *   b200fit_signal_bits     bits = (y > snr_cut * rms[pix])                                  (signals.py:264, :316)
 *   b200fit_bits_morph      op 0: binary_erosion with ball(2) (v_estimate :266); 1 / 2: erosion / dilation with the
 *                           6-connected cross (binary_opening, get_v_mask :322); 3: dilation with disk(3) in the plane
 *                           (:386).  Erosion sees True outside the cube, dilation False.  d_in != d_out.
 *   b200fit_bits_window     out = in & (vc - hwidth < V_i < vc + hwidth); vc per pixel (d_vcenter) or vcenter_all
 *   b200fit_v_at_peak       velocity of the first maximum over the channels set in d_bits_a (& d_bits_b, optional),
 *                           finite and in the planemask; NaN without one                    (get_v_at_peak :229-248)
 *   b200fit_masked_moments  moment0/1/2 of spectral_cube over the set, finite, planemask channels       (:388-391) */
int b200fit_signal_bits(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                        const double* d_rms, double snr_cut, uint32_t* d_bits, void* stream);
int b200fit_bits_morph(b200fit_ctx* ctx, int32_t op, int32_t nchan, int32_t ny, int32_t nx, int64_t nchan_stride,
                       const uint32_t* d_in, uint32_t* d_out, void* stream);
int b200fit_bits_window(b200fit_ctx* ctx, int32_t nchan, int64_t npix, int64_t nchan_stride, double v0, double dv,
                        const uint32_t* d_in, const double* d_vcenter, double vcenter_all, double hwidth,
                        uint32_t* d_out, void* stream);
int b200fit_v_at_peak(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                      const uint32_t* d_bits_a, const uint32_t* d_bits_b, const unsigned char* d_planemask, double v0,
                      double dv, double* d_vpeak, void* stream);
int b200fit_masked_moments(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                           const uint32_t* d_bits, const unsigned char* d_planemask, double v0, double dv, double* d_m0,
                           double* d_m1, double* d_m2, void* stream);

/* Convolved-cube stage of iter_2comp_fit: the spatial smoothing of convolve_tools.convolve_sky
 * (mufasa/convolve_tools.py:90-138, spectral_cube ``convolve_to`` = astropy ``convolve`` per channel plane) and the
 * regridding of convolve_sky_byfactor (:72-79, ``reproject(..., order='bilinear')``).  Cubes are channel-major planes
 * [nchan][ny][nx] float32 (the layout of the FITS cube), NOT the compacted fit rows.
 *   b200fit_convolve_planes   out = sum K v / sum K f per voxel: kernel normalised over the finite, included voxels
 *                             (nan_treatment='interpolate'), zeros beyond the image (boundary='fill'); NaN where the
 *                             voxel itself is NaN or its pixel excluded (d_include [ny*nx] uint8, NULL = all
 *                             included: the output keeps the input's mask) or nothing finite lies under the kernel.  ksize odd, <= 65.  Kernel: separable (d_kx, d_ky [ksize], d_k2 NULL:
 *                             K(dy,dx) = ky[dy] kx[dx], a circular beam) or general (d_k2 [ksize*ksize] row-major,
 *                             d_kx = d_ky = NULL).  The kernel need not be normalised; kernel_sum = the sum of its
 *                             K*K weights (sum(kx) * sum(ky)), the normalisation of a fully finite neighbourhood.
 *                             d_out != d_cube.
 *   b200fit_resample_bilinear out(c, oy, ox) = bilinear interpolation of plane c at (y, x) = (ay oy + by, ax ox + bx)
 *                             (0-based pixel centres); edge replicated within half a pixel of the image, NaN beyond
 *                             and wherever one of the four neighbours is NaN. */
int b200fit_convolve_planes(b200fit_ctx* ctx, const float* d_cube, int32_t nchan, int32_t ny, int32_t nx,
                            const unsigned char* d_include, int32_t ksize, const float* d_kx, const float* d_ky,
                            const float* d_k2, double kernel_sum, float* d_out, void* stream);
int b200fit_resample_bilinear(b200fit_ctx* ctx, const float* d_in, int32_t nchan, int32_t ny, int32_t nx, float* d_out,
                              int32_t out_ny, int32_t out_nx, double ay, double by, double ax, double bx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FIT_H */
