package spim.process.fusion.deconvolution;

import com.sun.jna.Native;
import com.sun.jna.Pointer;
import com.sun.jna.ptr.PointerByReference;

import net.imglib2.img.Img;
import net.imglib2.img.array.ArrayImg;
import net.imglib2.img.array.ArrayImgs;
import net.imglib2.img.basictypeaccess.array.FloatArray;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.view.Views;
import spim.process.cuda.MVDeconSession;
import spim.process.fusion.deconvolution.MVDeconFFT.PSFTYPE;

/**
 * Thin caller that replaces the body of MVDeconvolution's constructor
 * (spim/process/fusion/deconvolution/MVDeconvolution.java:94-211) by one device-resident session:
 * views are uploaded once, every iteration runs on the GPU, psi is downloaded at the end.
 * NOT compiled in this repository (no JVM in the build image).
 */
public class MVDeconvolutionGPU
{
	final Img< FloatType > psi;

	public MVDeconvolutionGPU( final MVDeconInput views, final PSFTYPE iterationType, final int numIterations,
			final double lambda, final int device )
	{
		final MVDeconSession lib = Native.load( "Convolution3D_fftCUDAlib", MVDeconSession.class );
		final MVDeconFFT first = views.getViews().get( 0 );
		final long[] d = new long[ 3 ];
		first.getImage().dimensions( d );

		final MVDeconSession.Params p = new MVDeconSession.Params();
		lib.mvd_params_default( p );
		p.dims[ 0 ] = (int)d[ 2 ]; p.dims[ 1 ] = (int)d[ 1 ]; p.dims[ 2 ] = (int)d[ 0 ];   // (z, y, x), like getCUDACoordinates
		p.num_views = views.getNumViews();
		p.iteration_type = iterationType.ordinal();
		p.generation = 2;
		p.lambda = lambda;
		p.device = device;

		final PointerByReference ref = new PointerByReference();
		check( lib, lib.mvd_session_create( p, ref ) );
		final Pointer s = ref.getValue();
		try
		{
			int v = 0;
			for ( final MVDeconFFT view : views.getViews() )
			{
				final long[] k = new long[ 3 ];
				view.getKernel1().dimensions( k );
				check( lib, lib.mvd_set_view( s, v++, toArray( view.getImage() ), toArray( view.getWeight() ),
						( (FloatArray)view.getKernel1().update( null ) ).getCurrentStorageArray(),
						new int[]{ (int)k[ 2 ], (int)k[ 1 ], (int)k[ 0 ] } ) );
			}
			check( lib, lib.mvd_init( s ) );                       // views.init( iterationType ) + psi = avg
			check( lib, lib.mvd_run( s, numIterations, null, null ) );
			check( lib, lib.mvd_finish( s ) );                     // "Masking never updated pixels."
			final float[] out = new float[ (int)( d[ 0 ] * d[ 1 ] * d[ 2 ] ) ];
			check( lib, lib.mvd_get_psi( s, out ) );
			this.psi = ArrayImgs.floats( out, d );
		}
		finally
		{
			lib.mvd_session_destroy( s );
		}
	}

	public Img< FloatType > getPsi() { return psi; }

	private static void check( final MVDeconSession lib, final int rc )
	{
		if ( rc != 0 )
			throw new RuntimeException( "mvdecon: " + lib.mvd_last_error() );
	}

	/** materialise a (possibly virtual) RandomAccessibleInterval into the float[] hand-over format */
	private static float[] toArray( final net.imglib2.RandomAccessibleInterval< FloatType > img )
	{
		if ( img instanceof ArrayImg )
			return ( (FloatArray)( (ArrayImg< FloatType, ? >)img ).update( null ) ).getCurrentStorageArray();
		final float[] a = new float[ (int)Views.iterable( img ).size() ];
		int i = 0;
		for ( final FloatType t : Views.flatIterable( img ) )
			a[ i++ ] = t.get();
		return a;
	}
}
