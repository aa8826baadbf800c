/*
 * spim_fftconv.h -- the legacy native boundary of fiji/SPIM_Registration's multi-view deconvolution.
 *
 * These eight symbols are exactly what the reference binds through JNA
 * (interface spim.process.cuda.CUDAFourierConvolution extends CUDAStandardFunctions):
 *   /root/reference/src/main/java/spim/process/cuda/CUDAStandardFunctions.java:36-44
 *   /root/reference/src/main/java/spim/process/cuda/CUDAFourierConvolution.java:30-31
 * loaded as "Convolution3D_fftCUDAlib" (fiji/plugin/Multi_View_Deconvolution.java:858) or from a
 * user-picked file whose name contains "fftCUDA" / "FourierConvolutionCUDA"
 * (spim/process/fusion/deconvolution/EfficientBayesianBased.java:1127-1131).
 *
 * C calling convention, unmangled names, plain pointers and sizes.  All buffers are owned by the
 * caller and are not retained past return.  See INTEGRATION.md for the JNA side.
 */
#ifndef SPIM_FFTCONV_H
#define SPIM_FFTCONV_H

#ifdef __cplusplus
extern "C" {
#endif

/* CUDAStandardFunctions.java:36-37 -- compute capability of device devCUDA (-1 on error). */
int getCUDAcomputeCapabilityMinorVersion(int devCUDA);
int getCUDAcomputeCapabilityMajorVersion(int devCUDA);

/* CUDAStandardFunctions.java:38-41 -- number of CUDA devices; -1 if the driver failed, 0 if none. */
int getNumDevicesCUDA(void);

/* CUDAStandardFunctions.java:42 -- writes the NUL-padded device name into the caller's 256-byte
 * buffer (Java passes byte[256], spim/process/cuda/CUDATools.java:76,84-90). */
void getNameDeviceCUDA(int devCUDA, char* name);

/* CUDAStandardFunctions.java:43-44 -- total / free device memory in bytes (Java long = 64 bit). */
long long getMemDeviceCUDA(int devCUDA);
long long getFreeMemDeviceCUDA(int devCUDA);

/* CUDAFourierConvolution.java:31 -- THE hot call (call sites
 * mpicbg/spim/postprocessing/deconvolution2/LRFFTThreads.java:70-71,85-86 and
 * spim/process/fusion/deconvolution/MVDeconFFTThreads.java:86-89,109-112).
 *
 * im        : imDim[0]*imDim[1]*imDim[2] floats, C order [z][y][x], overwritten with the result
 * imDim     : 3 ints (z, y, x) -- slowest first (Java reverses its (x,y,z), MVDeconFFTThreads.java:157-165)
 * kernel    : kernelDim[0]*kernelDim[1]*kernelDim[2] floats, same order; never modified
 * result    : circular convolution over exactly imDim of im with the kernel zero-padded to imDim and
 *             shifted so that element kernelDim/2 sits at the origin, normalised.
 * errors    : void by contract.  On failure a diagnostic goes to stderr, im is left untouched and
 *             spim_fftconv_last_error() returns the message.
 */
void convolution3DfftCUDAInPlace(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA);

/* CUDAFourierConvolution.java:30 -- declared by the reference but never called (JNA cannot map a
 * primitive-array return).  Exported for ABI completeness: returns a malloc()ed result the caller
 * frees with free(); NULL on failure. */
float* convolution3DfftCUDA(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA);

/* additive: last error message of the calling thread ("" if none). */
const char* spim_fftconv_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
