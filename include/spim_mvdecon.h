/*
 * spim_mvdecon.h -- additive, device-resident session API for the multi-view deconvolution
 * iteration (the fused path the thin Java host code calls instead of one JNA round trip per block
 * convolution).  C ABI, plain pointers and sizes.
 *
 * What it replaces in the reference (paths under /root/reference/src/main/java/):
 *   mpicbg/spim/postprocessing/deconvolution2/BayesMVDeconvolution.java:79-178   (gen-1 constructor loop)
 *   mpicbg/spim/postprocessing/deconvolution2/BayesMVDeconvolution.java:261-382  (runIteration)
 *   mpicbg/spim/postprocessing/deconvolution2/LRFFT.java:214-359, 423-626        (init, convolve1/2)
 *   spim/process/fusion/deconvolution/MVDeconvolution.java:94-211, 354-465       (gen-2 constructor, runIteration)
 *   spim/process/fusion/deconvolution/MVDeconFFT.java:183-324, 384-560           (init, convolve1/2)
 *
 * Array convention: every volume is C-ordered [z][y][x] fp32 (x fastest) and every dims triple is
 * (z, y, x) -- the same order the legacy entry convolution3DfftCUDAInPlace uses.
 * All functions return 0 on success, non-zero on failure; mvd_last_error() gives the message.
 */
#ifndef SPIM_MVDECON_H
#define SPIM_MVDECON_H

#ifdef __cplusplus
extern "C" {
#endif

/* PSFTYPE ordinals, LRFFT.java:54 / MVDeconFFT.java:49 (dialog index order,
 * fiji/plugin/Multi_View_Deconvolution.java:675-682) */
enum { MVD_OPTIMIZATION_II = 0, MVD_OPTIMIZATION_I = 1, MVD_EFFICIENT_BAYESIAN = 2, MVD_INDEPENDENT = 3 };

/* out-of-bounds rules */
enum { MVD_EXT_ZERO = 0, MVD_EXT_CONSTANT = 1, MVD_EXT_MIRROR_SINGLE = 2, MVD_EXT_MIRROR_DOUBLE = 3, MVD_EXT_PERIODIC = 4 };

typedef struct mvd_session mvd_session;

typedef struct mvd_params {
    int struct_size;       /* sizeof(mvd_params), for forward compatibility */
    int dims[3];           /* (z, y, x) of every view image / of psi */
    int num_views;
    int iteration_type;    /* MVD_* PSFTYPE ordinal */
    int generation;        /* 1 = BayesMVDeconvolution/LRFFT semantics, 2 = MVDeconvolution/MVDeconFFT */
    double lambda;         /* Tikhonov parameter, 0 = off (default 0.006) */
    float min_value;       /* LRInput.minValue = 0.0001f */
    double osem_speedup;   /* weights become min(1, w * osem) */
    int osem_index;        /* gen-1: 0 = use osem_speedup, 1 = min #overlap, 2 = avg #overlap, 3 = manual */
    int conv1_ext;         /* -1 = default (mirror-single) */
    int conv2_ext;         /* -1 = default (gen-2: constant 1.0, gen-1: mirror-single) */
    int device;            /* CUDA device ordinal */
    int haloed;            /* 1 = brick mode: psi / ratio carry a kernel-sized halo that the caller
                              fills (neighbour exchange + mvd_fill_halo) before every convolution */
    int exact_tikhonov;    /* 0 (default): Tikhonov step in the algebraically identical, cancellation-free fp32 form
                              2v/(1+sqrt(1+2*lambda*v)) (<= 2 ulp from the reference's fp64 expression);
                              1: evaluate (sqrt(1+2*lambda*v)-1)/lambda in fp64 exactly like the Java code */
    int fast_epilogue;     /* 1 (set by mvd_params_default): division / square root of the fused ratio and update epilogues as a
                              branch-free FMA refinement of the hardware approximations (csrc/fast_math.h) -- the correctly
                              rounded IEEE values for operands in the normal range; the update step as a whole is identical to
                              the IEEE evaluation for EVERY input below 2^126 (tests/cpp/fast_epilogue_composite.cpp, exhaustive);
                              a blurred value that is zero, denormal or infinite gives a NaN quotient where IEEE gives +-inf / 0
                              (both poison the next FFT convolution, in the reference too);
                              0: IEEE fp32 division / square root intrinsics.  SPIM_FAST_EPI=0/1 overrides either way */
    int reserved[6];
} mvd_params;

typedef struct mvd_info {
    double avg;            /* value psi was initialised with */
    double osem;           /* OSEM factor actually applied */
    int min_overlap;       /* gen-1 statistics (0 if not computed) */
    double avg_overlap;
    int fft_dims[3];       /* padded circular FFT size (z, y, x) */
    int pitch;             /* complex row pitch of the half spectrum */
    long long n_voxels;    /* N  = prod(dims) */
    long long np_voxels;   /* Np = prod(dims + psf - 1): the volume the roofline accounting uses */
    long long device_bytes;/* device memory held by the session */
    int halo_lo[3];        /* brick mode: halo before / after the brick on each axis */
    int halo_hi[3];
} mvd_info;

void mvd_params_default(mvd_params* p);
int mvd_session_create(const mvd_params* p, mvd_session** out);
void mvd_session_destroy(mvd_session* s);

/* LRInput.add(new LRFFT(image, weight, kernel, ...)) / MVDeconInput.add(new MVDeconFFT(...)):
 * host pointers (pageable or pinned), copied to the device; pointers into device memory are accepted as well (unified
 * addressing), for views that were produced on a GPU.  weight == NULL means constant 1 (LRFFT.java:201-204). */
int mvd_set_view(mvd_session* s, int view, const float* img, const float* weight,
                 const float* psf, const int psf_dims[3]);

/* The same for volumes handed over in cells (imglib2 CellImg) or larger than Java's 2^31-element arrays:
 * upload the box [lo, lo + ext) (z, y, x) of the view's image (which = 0) or weight (which = 1) from a tightly
 * packed host buffer.  The first call for a buffer creates it zero-filled; the PSF is set with mvd_set_psf
 * (spim_fusion.h) or a previous mvd_set_view. */
int mvd_upload_region(mvd_session* s, int view, int which, const float* data, const int lo[3], const int ext[3]);

/* views.init(iterationType) + psi initialisation + OSEM clamp
 * (BayesMVDeconvolution.java:91-117, MVDeconvolution.java:114-148). */
int mvd_init(mvd_session* s);

/* n_iterations x runIteration().  sum_change / max_change: optional arrays of
 * n_iterations*num_views doubles receiving the per view-step statistics the reference logs
 * (MVDeconvolution.java:441-457). */
int mvd_run(mvd_session* s, int n_iterations, double* sum_change, double* max_change);

/* gen-2 only: psi <- 0 where no view has data (MVDeconvolution.java:201-208); no-op for gen-1. */
int mvd_finish(mvd_session* s);

int mvd_get_psi(mvd_session* s, float* out);          /* Deconvolver.getPsi() */
int mvd_set_psi(mvd_session* s, const float* in);     /* 'initialImage' hook */
int mvd_get_kernel(mvd_session* s, int view, int which /*1 or 2*/, float* out);  /* getKernel1/2 */
int mvd_get_info(mvd_session* s, mvd_info* out);
int mvd_sync(mvd_session* s);
/* the CUDA stream (cudaStream_t) every kernel of this session is launched on, for event timing */
int mvd_get_stream(mvd_session* s, void** stream);

/* per-kernel-class CUDA-event timing (bench.py's roofline leg): ids 0 x-fwd, 1 y-fwd,
 * 2 z-fwd*K*z-inv, 3 y-inv, 4 x-inv+epilogue, 5 z-fwd (kernel spectra), 6 misc */
int mvd_set_timing(mvd_session* s, int on);
int mvd_get_timing(mvd_session* s, double ms[8], long long launches[8]);

/* ---- brick mode (one session per GPU) -------------------------------- */
/* device pointer + geometry of a session buffer: which 0 = psi, 1 = ratio/tmp */
int mvd_get_device_buffer(mvd_session* s, int which, void** dptr, int dims[3], int origin[3]);
/* declare which sides of the brick have a neighbour (bit d = axis d of (z,y,x)): the convolution loader
 * reads the halo there (the caller refreshes it before every convolution) and applies the convolution's
 * own out-of-bounds rule on all other sides (volume faces).  Default: no neighbours. */
int mvd_set_halo_mask(mvd_session* s, int lo_mask, int hi_mask);
/* gather (pack) / scatter (unpack) up to 26 box-shaped pieces of buffer `which` to / from one flat DEVICE staging
 * buffer in a single kernel launch.  regions: npieces x 6 ints (z0, y0, x0, nz, ny, nx) in array coordinates of
 * the haloed buffer; pieces are laid out back to back in `flat` in the order given. */
int mvd_halo_pack(mvd_session* s, int which, int npieces, const int* regions, void* flat);
int mvd_halo_unpack(mvd_session* s, int which, int npieces, const int* regions, void* flat);
/* fill the halo faces flagged in lo_mask / hi_mask (bit d = axis d of (z,y,x)) of buffer `which`
 * from the brick's own interior using the convolution's out-of-bounds rule (volume faces) */
int mvd_fill_halo(mvd_session* s, int which, int lo_mask, int hi_mask);
/* Direct halo push (the exchange as ONE fused copy + signal kernel over NVLink peer memory -- no staging buffer, no
 * NCCL call on the data path; replaces the 3-phase ncclSend/ncclRecv plan of SURVEY.md section 8e):
 *   mvd_p2p_export   after mvd_init: a 288-byte record (process id, device, pointers and CUDA IPC handles of the psi
 *                    buffer, the ratio buffer and a flag page) for the other ranks of the node; the caller all-gathers it
 *   mvd_p2p_connect  per neighbour piece i: the neighbour's record; boxes[i] = 9 ints (z0, y0, x0 of the box in MY
 *                    buffer, nz, ny, nx, z0, y0, x0 of the same box in the NEIGHBOUR's halo -- all bricks share one
 *                    geometry); slots[i] = 2 ints in [0, 27): the flag I raise at the neighbour, the flag it raises here
 *                    (by convention (dz+1)*9 + (dy+1)*3 + (dx+1) of the sender as seen from the receiver).  Same-process
 *                    peers are reached through their raw pointers (peer access is enabled), others through the IPC handles
 *   mvd_p2p_push     asynchronous on the session stream: store every piece of buffer `which` into the neighbours' halos,
 *                    then raise this buffer's next epoch flag at each neighbour (system-scope release)
 *   mvd_p2p_wait     asynchronous: a one-block kernel that spins (system-scope acquire) until every neighbour has raised
 *                    the same epoch here; after SPIM_P2P_TIMEOUT_S (30) seconds it gives up and sets an error word
 *   mvd_p2p_status   synchronous: *timed_out = 1 if any wait gave up (results are then invalid)
 * Pushes of the two buffers must alternate between two pushes of the same buffer (they do in an iteration: psi, ratio,
 * psi, ...) or be separated by a barrier across ranks -- that is what makes overwriting a neighbour's halo safe. */
#define MVD_P2P_RECORD_BYTES 288
int mvd_p2p_export(mvd_session* s, unsigned char record[MVD_P2P_RECORD_BYTES]);
int mvd_p2p_connect(mvd_session* s, int npieces, const unsigned char* records, const int* boxes, const int* slots);
int mvd_p2p_push(mvd_session* s, int which);
int mvd_p2p_wait(mvd_session* s, int which);
int mvd_p2p_status(mvd_session* s, int* timed_out);
int mvd_p2p_disconnect(mvd_session* s);
/* one half of a view-step: phase 0 = conv1 + quotient, phase 1 = conv2 + update.
 * stats: optional 2 doubles (sum, max) accumulated for phase 1. */
int mvd_view_phase(mvd_session* s, int view, int phase, double* stats);
/* brick-local partial sums for the psi initialisation; the caller all-reduces and calls mvd_set_avg */
int mvd_init_partials(mvd_session* s, double partial[6]);
int mvd_set_avg(mvd_session* s, double avg, double osem);

/* ---- stand-alone helpers --------------------------------------------------------------------- */
/* out = ext(img) (*) kernel (kernel origin at dim/2), host buffers, on `device`. */
int mvd_convolve(const float* img, const int im_dims[3], const float* kernel, const int kernel_dims[3],
                 int ext, float ext_value, float* out, int device);
/* smallest supported FFT length >= min_n */
int mvd_fft_size(int min_n, int need_even);
/* host-side launch counters for tests: which 0 = column passes launched with narrow (8-column) tiles */
long long mvd_debug_counter(int which);
const char* mvd_last_error(void);
const char* mvd_version(void);

#ifdef __cplusplus
}
#endif
#endif
