/*
 * spim_fusion.h -- device-side fusion pre-step for the multi-view deconvolution session
 * (the callers on the input side of the hot path).  C ABI, additive:
 * a host that already builds its transformed views, weights and PSFs itself keeps using mvd_set_view.
 *
 * What it replaces in the reference (paths under /root/reference/src/main/java/):
 *   spim/process/fusion/deconvolution/ProcessForDeconvolution.java:180-312  (per-view transform loop)
 *   spim/process/fusion/deconvolution/TransformInput.java:70-116            (affine + tri-linear + minValue clamp)
 *   spim/process/fusion/deconvolution/TransformInputAndWeights.java:76-135  (the same + blending weight)
 *   spim/process/fusion/deconvolution/TransformWeights.java:73-111          (weights only)
 *   spim/process/fusion/weights/BlendingRealRandomAccess.java:44-54, 91-121 (cosine blending, 1001-entry table)
 *   spim/process/fusion/deconvolution/WeightNormalizer.java:73-253          (sum-of-weights normalisation, overlap statistics)
 *   spim/process/fusion/weights/NormalizingRandomAccess.java:57-66          (virtual weights: min(1, w / sum * osem))
 *   spim/process/fusion/deconvolution/ProcessForDeconvolution.java:353-405  (OSEM factor from the overlap statistics, clamp)
 *   spim/fiji/spimdata/imgloaders/AbstractImgLoader.java:164-184            (min-max normalisation of a loaded stack)
 *   spim/process/fusion/deconvolution/ExtractPSF.java:277-457               (extractPSFLocal, normalize, transformPSF, transform)
 *
 * Conventions: volumes are C-ordered [z][y][x] fp32 and every dims triple is (z, y, x), as everywhere in this
 * library.  Coordinates, offsets, blending borders / ranges and affine matrices keep the reference's (x, y, z)
 * order; an affine is the 12 row-packed doubles of AffineTransform3D.getRowPackedCopy().
 * All functions return 0 on success; mvd_last_error() gives the message.  There is no CPU fallback.
 */
#ifndef SPIM_FUSION_H
#define SPIM_FUSION_H

#include "spim_mvdecon.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvd_transform {
    int struct_size;         /* sizeof(mvd_transform) */
    double inverse[12];      /* transform.inverse().getRowPackedCopy(): bounding-box coordinate -> raw-stack coordinate */
    long long offset[3];     /* bb.min (x, y, z), ProcessForDeconvolution.java:218 */
    int want_image;          /* 1: write the transformed image of the view (TransformInput) */
    int want_weight;         /* 1: write the blending weight of the view (TransformInputAndWeights / TransformWeights) */
    float border[3];         /* blendingBorder (x, y, z), may be negative */
    float range[3];          /* blendingRange  (x, y, z) */
    int reserved[8];
} mvd_transform;

/* weight normalisation modes of mvd_normalize_weights */
enum { MVD_WEIGHTS_PRECOMPUTED = 0,   /* WeightNormalizer(weights): w <- (float)(w / sum) everywhere (0/0 = NaN as in the reference) */
       MVD_WEIGHTS_VIRTUAL = 1 };     /* WeightNormalizer(weights, factory): sum image = sum > 1 ? sum : 1; the division and the
                                         OSEM clamp happen together, in double, when mvd_init fixes the OSEM factor */

/* Upload one raw stack (host, (z,y,x) dims) into the session's stack buffer; normalize != 0 applies the loader's
 * (v - min) / (max - min).  stack == NULL releases the buffer. */
int mvd_load_stack(mvd_session* s, const float* stack, const int dims[3], int normalize);

/* Resample the loaded stack into view `view` of the session (image and / or blending weight), on the device. */
int mvd_transform_view(mvd_session* s, int view, const mvd_transform* t);

/* Set only the PSF of a view (the image / weight came from mvd_transform_view). */
int mvd_set_psf(mvd_session* s, int view, const float* psf, const int psf_dims[3]);

/* WeightNormalizer.process() over the weights of all views.  num_portions = Threads.numThreads() * 2 of the
 * reference (the average is the mean of the per-portion means); min_views / avg_views receive
 * getMinOverlappingViews() / getAvgOverlappingViews().  With mvd_params.generation == 2 and osem_index 1 / 2,
 * mvd_init takes the OSEM factor from these statistics (max(1, .)). */
int mvd_normalize_weights(mvd_session* s, int mode, int num_portions, int* min_views, double* avg_views);

/* Copy a view's device-resident image (which = 0) or weight (which = 1) back to the host. */
int mvd_get_view(mvd_session* s, int view, int which, float* out);

/* ExtractPSF.extractPSFLocal (+ normalize when normalize != 0) on the loaded stack: locations are n_beads x (x, y, z)
 * doubles in stack coordinates, size is (z, y, x); out receives prod(size) floats. */
int mvd_extract_psf(mvd_session* s, int n_beads, const double* locations_xyz, const int size[3], int normalize, float* out);

/* ExtractPSF.transformPSF: geometry (host logic) and resampling (device).  model / inverse are the view's
 * AffineTransform3D and its inverse; out_dims is (z, y, x), offset (x, y, z). */
int mvd_transform_psf_size(const int dims[3], const double model[12], int out_dims[3], double offset_xyz[3]);
int mvd_transform_psf(const float* psf, const int dims[3], const double model[12], const double inverse[12],
                      float* out, const int out_dims[3], int device);

/* the 1001-entry cosine table of BlendingRealRandomAccess (for binding-side checks) */
int mvd_blending_lookup(double out[1001]);

#ifdef __cplusplus
}
#endif
#endif
