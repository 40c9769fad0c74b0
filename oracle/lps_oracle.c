/*
 * lps_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C restatement of the reference LPS front end Wav2LPS_be at 16 kHz:
 *   framing / frame loop        Feature_prepare/SourceCode_Wav2LogSpec_be/Wav2LogSpec_be.c:395-404, 413-563
 *   ReadWave (int16 -> float)   fileio.c:268-282
 *   InitializeHamming / Window  FEfunc.c:80-118
 *   rfft (split-radix, Sorensen et al. 1987)  FEfunc.c:146-293
 *   power + floored ln          Wav2LogSpec_be.c:469-479  (floor exp(-50) -> -50, :54, :303)
 *
 * Pinned bit-exactly against the reference's two golden wav/lps pairs
 * (tests/golden/TEST_DR8_MPAM0_SX{289,379}.{wav,lps}; tests/test_oracle_cpu.py) and, in the
 * build container, against the reference binary itself (oracle/_ref/Wav2LPS_be_ref).
 *
 * Arithmetic notes mirrored from the reference: window in double then stored float; twiddles
 * (float)cos((double)a) with a = j*e in float; the 1/sqrt(2) butterflies divide in double;
 * everything else is float with no FMA contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define LPS_FRAME 512
#define LPS_SHIFT 256
#define LPS_BINS  257
#define ORACLE_PIx2   6.28318530717958647692   /* FEfunc.h:28 */
#define ORACLE_PI     3.14159265358979323846   /* FEfunc.h:24 */
#define ORACLE_SQRT2  1.41421356237309504880   /* FEfunc.h:25 */

long lps_oracle_nframes(long n_samples)
{
    /* first 256 samples prime the buffer (:401); every further complete hop of 256 makes a frame (:413) */
    if (n_samples < LPS_FRAME) return 0;
    return (n_samples - (LPS_FRAME - LPS_SHIFT)) / LPS_SHIFT;
}

/* FEfunc.c:80-87 (half window; Window() mirrors it, :106-118) */
void lps_oracle_hamming(float *win /* [256] */)
{
    for (int i = 0; i < LPS_FRAME / 2; i++)
        win[i] = (float)(0.54 - 0.46 * cos(ORACLE_PIx2 * i / (LPS_FRAME - 1)));
}

/* FEfunc.c:146-293, restated.  Output order Re(0..n/2), Im(n/2-1..1). */
void lps_oracle_rfft(float *x, int n, int m)
{
    /* bit reversal (:157-181) */
    for (int i = 0, j = 0; i < n - 1; i++) {
        if (i < j) { float t = x[j]; x[j] = x[i]; x[i] = t; }
        int k = n >> 1;
        while (k <= j) { j -= k; k >>= 1; }
        j += k;
    }
    /* length-2 butterflies (:184-199) */
    for (int is = 0, id = 4; is < n - 1; is = 2 * id - 2, id *= 4)
        for (int i0 = is; i0 < n; i0 += id) {
            float a0 = x[i0];
            x[i0] = a0 + x[i0 + 1];
            x[i0 + 1] = a0 - x[i0 + 1];
        }
    /* L-shaped butterflies (:202-292) */
    int n2 = 2;
    for (int k = 1; k < m; k++) {
        n2 <<= 1;
        const int n4 = n2 >> 2, n8 = n2 >> 3;
        const float e = (float)((ORACLE_PI * 2) / n2);
        for (int is = 0, id = n2 << 1; is < n; is = 2 * id - n2, id *= 4)
            for (int i = is; i <= n - 1; i += id) {
                int i1 = i, i2 = i1 + n4, i3 = i2 + n4, i4 = i3 + n4;
                float t1 = x[i4] + x[i3];
                x[i4] = x[i4] - x[i3];
                x[i3] = x[i1] - t1;
                x[i1] = x[i1] + t1;
                if (n4 != 1) {
                    i1 += n8; i2 += n8; i3 += n8; i4 += n8;
                    t1 = (float)((x[i3] + x[i4]) / ORACLE_SQRT2);
                    float t2 = (float)((x[i3] - x[i4]) / ORACLE_SQRT2);
                    x[i4] = x[i2] - t1;
                    x[i3] = -x[i2] - t1;
                    x[i2] = x[i1] - t2;
                    x[i1] = x[i1] + t2;
                }
            }
        for (int j = 1; j < n8; j++) {
            const float a = j * e, a3 = 3 * a;
            const float cc1 = (float)cos(a), ss1 = (float)sin(a), cc3 = (float)cos(a3), ss3 = (float)sin(a3);
            for (int is = 0, id = n2 << 1; is < n; is = 2 * id - n2, id *= 4)
                for (int i = is; i <= n - 1; i += id) {
                    const int i1 = i + j, i2 = i1 + n4, i3 = i2 + n4, i4 = i3 + n4;
                    const int i5 = i + n4 - j, i6 = i5 + n4, i7 = i6 + n4, i8 = i7 + n4;
                    float t1 = x[i3] * cc1 + x[i7] * ss1;
                    float t2 = x[i7] * cc1 - x[i3] * ss1;
                    float t3 = x[i4] * cc3 + x[i8] * ss3;
                    float t4 = x[i8] * cc3 - x[i4] * ss3;
                    float t5 = t1 + t3, t6 = t2 + t4;
                    t3 = t1 - t3; t4 = t2 - t4;
                    t2 = x[i6] + t6; x[i3] = t6 - x[i6]; x[i8] = t2;
                    t2 = x[i2] - t3; x[i7] = -x[i2] - t3; x[i4] = t2;
                    t1 = x[i1] + t5; x[i6] = x[i1] - t5; x[i1] = t1;
                    t1 = x[i5] + t4; x[i5] = x[i5] - t4; x[i2] = t1;
                }
        }
    }
}

/* One frame: window, FFT, power, floored natural log (Wav2LogSpec_be.c:448-479). */
void lps_oracle_frame(const int16_t *pcm /* 512 samples */, const float *win /* [256] */, float *lps /* [257] */)
{
    float buf[LPS_FRAME + 1];
    const float floor_fb = (float)exp((double)-50.0);
    for (int i = 0; i < LPS_FRAME; i++) buf[i] = (float)pcm[i];
    for (int i = 0; i < LPS_FRAME / 2; i++) buf[i] *= win[i];
    for (int i = LPS_FRAME / 2; i < LPS_FRAME; i++) buf[i] *= win[LPS_FRAME - 1 - i];
    lps_oracle_rfft(buf, LPS_FRAME, 9);
    buf[0] = buf[0] * buf[0];
    for (int i = 1; i < LPS_FRAME / 2; i++) buf[i] = buf[i] * buf[i] + buf[LPS_FRAME - i] * buf[LPS_FRAME - i];
    buf[LPS_FRAME / 2] = buf[LPS_FRAME / 2] * buf[LPS_FRAME / 2];
    for (int i = 0; i <= LPS_FRAME / 2; i++)
        lps[i] = (buf[i] < floor_fb) ? -50.0f : (float)log((double)buf[i]);
}

/* Whole utterance: frame n covers samples [256n, 256n+512). Returns the number of frames written. */
long lps_oracle_extract(const int16_t *pcm, long n_samples, float *out /* [nframes][257] */)
{
    float win[LPS_FRAME / 2];
    lps_oracle_hamming(win);
    const long nf = lps_oracle_nframes(n_samples);
    for (long f = 0; f < nf; f++) lps_oracle_frame(pcm + f * LPS_SHIFT, win, out + f * LPS_BINS);
    return nf;
}
