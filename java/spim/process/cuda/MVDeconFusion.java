package spim.process.cuda;

import com.sun.jna.Pointer;
import com.sun.jna.Structure;
import com.sun.jna.ptr.DoubleByReference;
import com.sun.jna.ptr.IntByReference;

import java.util.Arrays;
import java.util.List;

/**
 * JNA binding of the device-side fusion pre-step (include/spim_fusion.h) exported by the same shared library:
 * the per-view loop of ProcessForDeconvolution.fuseStacksAndGetPSFs
 * (spim/process/fusion/deconvolution/ProcessForDeconvolution.java:180-312: TransformInput / TransformInputAndWeights /
 * TransformWeights), WeightNormalizer.process() (:320-343), the OSEM factor (:346-358) and
 * ExtractPSF.extractNextImg / transformPSF run on the GPU and leave their results inside the session.
 * NOT compiled in this repository (no JVM in the build image).
 *
 * Per view:   mvd_load_stack( s, stack, dimsZYX, 1 );
 *             t.inverse = transform.inverse().getRowPackedCopy();  t.offset = { bb.min(0), bb.min(1), bb.min(2) };
 *             mvd_transform_view( s, v, t );   [ mvd_extract_psf( ... ) ]   mvd_set_psf( s, v, psf, psfDimsZYX );
 * then:       mvd_normalize_weights( s, mode, Threads.numThreads() * 2, min, avg );  mvd_init( s );  mvd_run( ... )
 */
public interface MVDeconFusion extends MVDeconSession
{
	public static class Transform extends Structure
	{
		public int struct_size;
		public double[] inverse = new double[ 12 ];   // transform.inverse().getRowPackedCopy()
		public long[] offset = new long[ 3 ];          // bb.min (x, y, z)
		public int want_image;
		public int want_weight;
		public float[] border = new float[ 3 ];        // blendingBorder (x, y, z)
		public float[] range = new float[ 3 ];         // blendingRange (x, y, z)
		public int[] reserved = new int[ 8 ];

		@Override
		protected List< String > getFieldOrder()
		{
			return Arrays.asList( "struct_size", "inverse", "offset", "want_image", "want_weight", "border", "range", "reserved" );
		}
	}

	public static final int MVD_WEIGHTS_PRECOMPUTED = 0, MVD_WEIGHTS_VIRTUAL = 1;

	int mvd_load_stack( Pointer session, float[] stack, int[] dimsZYX, int normalize );
	int mvd_transform_view( Pointer session, int view, Transform t );
	int mvd_set_psf( Pointer session, int view, float[] psf, int[] psfDimsZYX );
	int mvd_upload_region( Pointer session, int view, int which, float[] data, int[] loZYX, int[] extZYX );
	int mvd_normalize_weights( Pointer session, int mode, int numPortions, IntByReference minViews, DoubleByReference avgViews );
	int mvd_get_view( Pointer session, int view, int which, float[] out );
	int mvd_extract_psf( Pointer session, int nBeads, double[] locationsXYZ, int[] sizeZYX, int normalize, float[] out );
	int mvd_transform_psf_size( int[] dimsZYX, double[] model, int[] outDimsZYX, double[] offsetXYZ );
	int mvd_transform_psf( float[] psf, int[] dimsZYX, double[] model, double[] inverse, float[] out, int[] outDimsZYX, int device );
	int mvd_blending_lookup( double[] out1001 );
}
