package spim.process.cuda;

import com.sun.jna.Library;
import com.sun.jna.Pointer;
import com.sun.jna.Structure;
import com.sun.jna.ptr.IntByReference;
import com.sun.jna.ptr.PointerByReference;

import java.util.Arrays;
import java.util.List;

/**
 * JNA binding of the additive session API (include/spim_mvdecon.h) exported by the same shared
 * library that implements {@link CUDAFourierConvolution}.  NOT compiled in this repository (no JVM in
 * the build image); shipped as the stub a maintainer adds next to
 * spim/process/cuda/CUDAFourierConvolution.java.
 *
 * Load with:  Native.load( "Convolution3D_fftCUDAlib", MVDeconSession.class )
 */
public interface MVDeconSession extends Library
{
	public static class Params extends Structure
	{
		public int struct_size;
		public int[] dims = new int[ 3 ];      // z, y, x
		public int num_views;
		public int iteration_type;             // PSFTYPE.ordinal()
		public int generation;                 // 1 = BayesMVDeconvolution/LRFFT, 2 = MVDeconvolution/MVDeconFFT
		public double lambda;
		public float min_value;
		public double osem_speedup;
		public int osem_index;
		public int conv1_ext;
		public int conv2_ext;
		public int device;
		public int haloed;
		public int exact_tikhonov;
		public int fast_epilogue;
		public int[] reserved = new int[ 6 ];

		@Override
		protected List< String > getFieldOrder()
		{
			return Arrays.asList( "struct_size", "dims", "num_views", "iteration_type", "generation", "lambda", "min_value",
					"osem_speedup", "osem_index", "conv1_ext", "conv2_ext", "device", "haloed", "exact_tikhonov", "fast_epilogue", "reserved" );
		}
	}

	void mvd_params_default( Params p );
	int mvd_session_create( Params p, PointerByReference session );
	void mvd_session_destroy( Pointer session );
	int mvd_set_view( Pointer session, int view, float[] img, float[] weight, float[] psf, int[] psfDimsZYX );
	int mvd_init( Pointer session );
	int mvd_run( Pointer session, int nIterations, double[] sumChange, double[] maxChange );
	int mvd_finish( Pointer session );
	int mvd_get_psi( Pointer session, float[] out );
	int mvd_set_psi( Pointer session, float[] in );
	int mvd_get_kernel( Pointer session, int view, int which, float[] out );
	/** views held in cells (CellImg) or beyond Java's 2^31-element arrays: one box of the image (which = 0) / weight (1) */
	int mvd_upload_region( Pointer session, int view, int which, float[] data, int[] loZYX, int[] extZYX );

	// ---- multi-device mode without host round trips (one Java thread per device, MVDeconFFT.java:447-469) ----------------
	// one brick-mode session (Params.haloed = 1) per device; halos travel over NVLink peer memory, see spim_mvdecon.h
	int mvd_set_halo_mask( Pointer session, int loMask, int hiMask );
	int mvd_init_partials( Pointer session, double[] partial6 );
	int mvd_set_avg( Pointer session, double avg, double osem );
	int mvd_view_phase( Pointer session, int view, int phase, double[] stats2 );
	int mvd_p2p_export( Pointer session, byte[] record288 );
	int mvd_p2p_connect( Pointer session, int nPieces, byte[] records, int[] boxes9, int[] slots2 );
	int mvd_p2p_push( Pointer session, int which );
	int mvd_p2p_wait( Pointer session, int which );
	int mvd_p2p_status( Pointer session, IntByReference timedOut );
	int mvd_p2p_disconnect( Pointer session );

	String mvd_last_error();
}
