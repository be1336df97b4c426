/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <liquid/liquid.h>.
 *
 * liquid-dsp (git a4d7c80d3, pinned by /root/reference/HardwareSetup/Install_liquid-dsp.sh:36)
 * is not vendored in the reference tree and is not installed in this image.  This header declares
 * only the handful of liquid names that the reference's *unmodified* headers and
 * cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp mention, so that file can be compiled
 * from where it lies under /root/reference into oracle/_ref/ (see oracle/Makefile).
 * The two functions the engine actually calls (fft_create_plan, fft_execute:
 * CE_Predictive_Node.cpp:42-45,150) are defined in oracle/liquid_fft_restated.c.
 */
#ifndef CRN_ORACLE_COMPAT_LIQUID_H
#define CRN_ORACLE_COMPAT_LIQUID_H

#include <string.h>
#include <complex.h>

#ifdef __cplusplus
#include <complex>
typedef std::complex<float> liquid_float_complex;
extern "C" {
/* g++13's <complex.h> in C++11 mode does not pull in the C99 prototypes */
float cabsf(float _Complex);
#else
typedef float _Complex liquid_float_complex;
#endif

#define LIQUID_FFT_FORWARD (+1)
#define LIQUID_FFT_BACKWARD (-1)

typedef struct fftplan_s *fftplan;
fftplan fft_create_plan(unsigned int n, liquid_float_complex *x, liquid_float_complex *y,
                        int dir, int flags);
void fft_destroy_plan(fftplan p);
void fft_execute(fftplan p);

/* opaque framing types the ECR header stores by value / by handle */
typedef struct {
  float evm, rssi, cfo;
  liquid_float_complex *framesyms;
  unsigned int num_framesyms, mod_scheme, mod_bps, check, fec0, fec1;
} framesyncstats_s;
typedef struct {
  unsigned int check, fec0, fec1, mod_scheme;
} ofdmflexframegenprops_s;
typedef struct ofdmflexframesync_s *ofdmflexframesync;
typedef struct ofdmflexframegen_s *ofdmflexframegen;

#ifdef __cplusplus
}
#endif
#endif
