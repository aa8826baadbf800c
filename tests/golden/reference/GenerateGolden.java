/*
 * Golden-vector generator for the parity tests of spim_registration_b200 -- to be run by anyone with a Fiji / SPIM_Registration
 * class path (this repository's build environment has no JVM, so its oracle is otherwise "parity unpinned", DESIGN.md section 2).
 *
 *   python tests/golden/reference/export_inputs.py
 *   javac -cp "$FIJI/jars/*:$FIJI/plugins/*" -d /tmp/gg tests/golden/reference/GenerateGolden.java
 *   java  -cp "/tmp/gg:$FIJI/jars/*:$FIJI/plugins/*" GenerateGolden tests/golden/reference/inputs tests/golden/reference/out
 *   python -m pytest tests/test_reference_golden.py -q
 *
 * It runs the REFERENCE's own CPU code on the exported inputs (3 views 20x18x16, 5^3 PSFs, lambda 0.006, 2 iterations):
 *   gen-2  spim.process.fusion.deconvolution.MVDeconvolution      (MVDeconvolution.java:94-211), all four PSFTYPEs
 *   gen-1  mpicbg.spim.postprocessing.deconvolution2.BayesMVDeconvolution (BayesMVDeconvolution.java:79-251), all four
 *   one FFT convolution per out-of-bounds rule the two generations use:
 *     ImgLib2 FFTConvolution with its default image extension            (MVDeconFFT.java:412-426, conv1 of gen-2)
 *     ImgLib2 FFTConvolution on Views.extendValue(img, 1)                (MVDeconFFT.java:514-517, conv2 of gen-2)
 *     ImgLib1 FourierConvolution with its default strategy               (LRFFT.java:461-466, both convolutions of gen-1)
 * and writes raw little-endian float32 files (x fastest) plus the kernel2 of every view / type.
 */
import java.io.*;
import java.nio.*;
import java.nio.file.*;
import java.util.*;

import mpicbg.imglib.algorithm.fft.FourierConvolution;
import mpicbg.imglib.image.Image;
import mpicbg.spim.postprocessing.deconvolution2.BayesMVDeconvolution;
import mpicbg.spim.postprocessing.deconvolution2.LRFFT;
import mpicbg.spim.postprocessing.deconvolution2.LRInput;
import net.imglib2.Cursor;
import net.imglib2.algorithm.fft2.FFTConvolution;
import net.imglib2.img.Img;
import net.imglib2.img.array.ArrayImg;
import net.imglib2.img.array.ArrayImgFactory;
import net.imglib2.img.array.ArrayImgs;
import net.imglib2.img.basictypeaccess.array.FloatArray;
import net.imglib2.type.numeric.complex.ComplexFloatType;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.view.Views;
import spim.process.fusion.deconvolution.MVDeconFFT;
import spim.process.fusion.deconvolution.MVDeconInput;
import spim.process.fusion.deconvolution.MVDeconvolution;

public class GenerateGolden
{
	static float[] read( final File f ) throws IOException
	{
		final ByteBuffer b = ByteBuffer.wrap( Files.readAllBytes( f.toPath() ) ).order( ByteOrder.LITTLE_ENDIAN );
		final float[] a = new float[ b.remaining() / 4 ];
		b.asFloatBuffer().get( a );
		return a;
	}

	static void write( final File f, final float[] a ) throws IOException
	{
		final ByteBuffer b = ByteBuffer.allocate( a.length * 4 ).order( ByteOrder.LITTLE_ENDIAN );
		b.asFloatBuffer().put( a );
		Files.write( f.toPath(), b.array() );
	}

	static ArrayImg< FloatType, FloatArray > img( final float[] a, final long... dims ) { return ArrayImgs.floats( a.clone(), dims ); }

	static float[] flat( final Iterable< FloatType > it, final int n )
	{
		final float[] a = new float[ n ];
		int i = 0;
		for ( final FloatType t : it )
			a[ i++ ] = t.get();   // array containers iterate in linear order, x fastest
		return a;
	}

	static float[] flat1( final Image< mpicbg.imglib.type.numeric.real.FloatType > im )
	{
		final float[] a = new float[ im.getNumPixels() ];
		final mpicbg.imglib.cursor.Cursor< mpicbg.imglib.type.numeric.real.FloatType > c = im.createCursor();
		int i = 0;
		while ( c.hasNext() ) { c.fwd(); a[ i++ ] = c.getType().get(); }
		c.close();
		return a;
	}

	public static void main( final String[] args ) throws Exception
	{
		final File in = new File( args.length > 0 ? args[ 0 ] : "tests/golden/reference/inputs" );
		final File out = new File( args.length > 1 ? args[ 1 ] : "tests/golden/reference/out" );
		out.mkdirs();
		final Scanner sc = new Scanner( new File( in, "meta.txt" ) );
		final long nx = sc.nextLong(), ny = sc.nextLong(), nz = sc.nextLong();
		final int V = sc.nextInt();
		final long kx = sc.nextLong(), ky = sc.nextLong(), kz = sc.nextLong();
		final long cx = sc.nextLong(), cy = sc.nextLong(), cz = sc.nextLong(), qx = sc.nextLong(), qy = sc.nextLong(), qz = sc.nextLong();
		sc.close();
		final int N = (int)( nx * ny * nz );
		final float[][] im = new float[ V ][], w = new float[ V ][], psf = new float[ V ][];
		for ( int v = 0; v < V; ++v )
		{
			im[ v ] = read( new File( in, "img" + v + ".raw" ) );
			w[ v ] = read( new File( in, "w" + v + ".raw" ) );
			psf[ v ] = read( new File( in, "psf" + v + ".raw" ) );
		}
		final int iterations = 2;
		final double lambda = 0.006;
		final int[] cpu = new int[]{ -1 };   // deviceList { -1 } = the multithreaded Java CPU path

		MVDeconvolution.debug = false;
		BayesMVDeconvolution.debug = false;

		for ( final MVDeconFFT.PSFTYPE type : MVDeconFFT.PSFTYPE.values() )
		{
			// ---- gen-2 ----
			final MVDeconInput input = new MVDeconInput( new ArrayImgFactory< FloatType >() );
			for ( int v = 0; v < V; ++v )
				input.add( new MVDeconFFT( img( im[ v ], nx, ny, nz ), img( w[ v ], nx, ny, nz ), img( psf[ v ], kx, ky, kz ),
						new ArrayImgFactory< FloatType >(), cpu, false, null, false ) );
			final MVDeconvolution d2 = new MVDeconvolution( input, type, iterations, lambda, 1.0, 0, "golden" );
			write( new File( out, "psi_g2_t" + type.ordinal() + ".raw" ), flat( d2.getPsi(), N ) );
			for ( int v = 0; v < V; ++v )
				write( new File( out, "k2_g2_t" + type.ordinal() + "_v" + v + ".raw" ), flat( input.getViews().get( v ).getKernel2(), (int)( kx * ky * kz ) ) );

			// ---- gen-1 ----
			final LRInput lr = new LRInput();
			for ( int v = 0; v < V; ++v )
				lr.add( new LRFFT( (Img< FloatType >)img( im[ v ], nx, ny, nz ), (Img< FloatType >)img( w[ v ], nx, ny, nz ),
						(Img< FloatType >)img( psf[ v ], kx, ky, kz ), cpu, false, null ) );
			final BayesMVDeconvolution d1 = new BayesMVDeconvolution( lr, LRFFT.PSFTYPE.valueOf( type.name() ), iterations, lambda, 1.0, 0, "golden" );
			write( new File( out, "psi_g1_t" + type.ordinal() + ".raw" ), flat1( d1.getPsi() ) );
		}

		// ---- single convolutions, one per out-of-bounds rule on the path ----
		final float[] ci = read( new File( in, "conv_img.raw" ) ), ck = read( new File( in, "conv_kernel.raw" ) );
		final int CN = (int)( cx * cy * cz );
		{
			final Img< FloatType > a = img( ci, cx, cy, cz ), r = img( new float[ CN ], cx, cy, cz );
			final FFTConvolution< FloatType > f = new FFTConvolution< FloatType >( a, img( ck, qx, qy, qz ), new ArrayImgFactory< ComplexFloatType >() );
			f.setComputeComplexConjugate( false );
			f.setOutput( r );
			f.convolve();
			write( new File( out, "conv_imglib2_default.raw" ), flat( r, CN ) );
		}
		{
			final Img< FloatType > a = img( ci, cx, cy, cz ), r = img( new float[ CN ], cx, cy, cz );
			final FFTConvolution< FloatType > f = new FFTConvolution< FloatType >( a, img( ck, qx, qy, qz ), new ArrayImgFactory< ComplexFloatType >() );
			f.setComputeComplexConjugate( false );
			f.setImg( Views.extendValue( a, new FloatType( 1.0f ) ), a );
			f.setOutput( r );
			f.convolve();
			write( new File( out, "conv_imglib2_value1.raw" ), flat( r, CN ) );
		}
		{
			final Image< mpicbg.imglib.type.numeric.real.FloatType > a = LRFFT.wrap( (Img< FloatType >)img( ci, cx, cy, cz ) );
			final Image< mpicbg.imglib.type.numeric.real.FloatType > k = LRFFT.wrap( (Img< FloatType >)img( ck, qx, qy, qz ) );
			final FourierConvolution< mpicbg.imglib.type.numeric.real.FloatType, mpicbg.imglib.type.numeric.real.FloatType > f =
					new FourierConvolution< mpicbg.imglib.type.numeric.real.FloatType, mpicbg.imglib.type.numeric.real.FloatType >( a, k );
			f.setNumThreads();
			f.setKeepImgFFT( false );
			f.process();
			write( new File( out, "conv_imglib1_default.raw" ), flat1( f.getResult() ) );
		}
		System.out.println( "golden vectors written to " + out.getAbsolutePath() );
		System.exit( 0 );
	}
}
