/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into libcrnsense or any product path.
 *
 * The oracle's OWN statement of the configurations bench.py and the tests run, so that the reference arm
 * (`bench.py --impl reference`) and the CPU checker never have to load the product library to learn what
 * "configs[1]" means.  Restated from the reference, independently of csrc/crn_config.cpp; the two are compared
 * byte for byte in tests/test_oracle.py::test_oracle_configs_equal_the_product_fillers.
 *
 *   engine constants        cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.hpp:31-32   (N = 512, K = 10)
 *   weight literals         cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp:78-120
 *   bin ranges              cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp:173-191
 *   feature order           cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp:200      (NF, CH1, CH2, CH3)
 *   threshold 0.8           cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp:245,250,255
 *   PU waveform defaults    src/crts.cpp:501-514, src/extensible_cognitive_radio.cpp:59, scenarios/predictive_model.cfg:39
 */
#include <string.h>

#include "../include/crnsense.h"

/* .cpp:78-120 in file order: 'I' = WeightIH[a][b], 'O' = WeightHO[a][b] */
static const struct { char layer; int a, b; double v; } kLiterals[] = {
    {'I', 0, 1, -0.188208}, {'I', 1, 1, -0.106634}, {'I', 2, 1, 0.005650},  {'I', 3, 1, -0.057578}, {'I', 4, 1, 0.092680},
    {'I', 0, 2, -0.170684}, {'I', 1, 2, -0.415470}, {'I', 2, 2, 0.741944},  {'I', 3, 2, 0.621154},  {'I', 4, 2, 0.809336},
    {'I', 0, 3, -0.024726}, {'I', 1, 3, 0.309261},  {'I', 2, 3, 0.006133},  {'I', 3, 3, -0.048268}, {'I', 4, 3, -0.010821},
    {'I', 0, 4, 0.001448},  {'I', 1, 4, 0.159974},  {'I', 2, 4, -0.620100}, {'I', 3, 4, -0.249186}, {'I', 4, 4, -0.546496},
    {'I', 0, 5, 0.015983},  {'I', 1, 5, 0.212781},  {'I', 2, 5, 0.669892},  {'I', 3, 5, 0.734475},  {'I', 4, 5, 0.609384},
    {'O', 0, 1, -7.033320}, {'O', 1, 1, 10.857465}, {'O', 2, 1, -6.848443}, {'O', 3, 1, 17.053079}, {'O', 4, 1, 0.087664},
    {'O', 5, 1, -6.552455}, {'O', 0, 2, 2.726400},  {'O', 1, 2, -18.452471}, {'O', 2, 2, 2.053071}, {'O', 3, 2, -13.375309},
    {'O', 4, 2, -0.269499}, {'O', 5, 2, 2.655529},  {'O', 0, 3, -2.590206}, {'O', 1, 3, 15.609466}, {'O', 2, 3, -2.929559},
    {'O', 3, 3, -15.703407}, {'O', 4, 3, 0.407028}, {'O', 5, 3, -2.552555},
};

/* Loop bounds of .cpp:173-190 as written there (inclusive), with the feature slot each sum feeds (.cpp:200). */
static const struct { int slot, first, last; } kLoops[] = {
    {0, 300, 309}, /* NF:  for (i = 300; i < 310; i++)                     */
    {1, 0, 15},    /* M1:  for (i = 0; i < 16; i++)                        */
    {1, 496, 510}, /* M1:  for (i = 496; i < 511; i++)   (511 left out)    */
    {2, 55, 84},   /* M2:  for (i = 55; i < 85; i++)                       */
    {3, 189, 221}, /* M3:  for (i = 189; i < 222; i++)                     */
};

/* The engine as the reference ships it: 512 points, rectangular, |X| averaged over 10 frames, (sum)^2, ANN. */
void crn_oracle_config_reference(crn_config *c) {
  memset(c, 0, sizeof(*c));
  c->nfft = c->frame_len = 512;
  c->navg = 10;
  c->window = CRN_WINDOW_RECT;
  c->detector = CRN_DET_MAG;
  c->postop = CRN_POST_SQUARE_OF_SUM;
  c->decide = CRN_DECIDE_ANN;
  c->nbands = 4;
  c->nsegs = (int32_t)(sizeof(kLoops) / sizeof(kLoops[0]));
  for (int s = 0; s < c->nsegs; s++) {
    c->segs[s].band = kLoops[s].slot;
    c->segs[s].lo = kLoops[s].first;
    c->segs[s].hi = kLoops[s].last + 1;
  }
  for (size_t i = 0; i < sizeof(kLiterals) / sizeof(kLiterals[0]); i++) {
    if (kLiterals[i].layer == 'I') c->ann_wih[kLiterals[i].a][kLiterals[i].b] = kLiterals[i].v;
    else c->ann_who[kLiterals[i].a][kLiterals[i].b] = kLiterals[i].v;
  }
  c->ann_threshold = 0.8;
  c->energy_factor = 4.0;
  c->ring_slots = 4;
  c->iq_format = CRN_IQ_CF32;
}

/* BASELINE configs[1]/[3]/[4]: the same bands in Hz (bin indices scale with N / 512), Hann, |X|^2, plain band sums. */
int crn_oracle_config_welch(crn_config *c, int32_t nfft, int32_t navg) {
  if (nfft < 256 || nfft > 8192 || (nfft & (nfft - 1)) || navg < 1) return -1;
  crn_oracle_config_reference(c);
  c->nfft = c->frame_len = nfft;
  c->navg = navg;
  c->window = CRN_WINDOW_HANN;
  c->detector = CRN_DET_MAGSQ;
  c->postop = CRN_POST_SUM;
  for (int s = 0; s < c->nsegs; s++) {
    c->segs[s].lo = (int32_t)((int64_t)c->segs[s].lo * nfft / 512);
    c->segs[s].hi = (int32_t)((int64_t)c->segs[s].hi * nfft / 512);
  }
  return 0;
}

/* BASELINE configs[2]: nch equal sub-channels, energy detection against the quietest sub-channel. */
int crn_oracle_config_wideband(crn_config *c, int32_t nfft, int32_t navg, int32_t nch) {
  if (nfft < 256 || nfft > 8192 || (nfft & (nfft - 1)) || navg < 1 || nch < 1 || nch > CRN_MAX_BANDS || nfft % nch) return -1;
  crn_oracle_config_welch(c, nfft, navg);
  c->decide = CRN_DECIDE_ENERGY;
  c->nbands = c->nsegs = nch;
  memset(c->segs, 0, sizeof(c->segs));
  for (int b = 0; b < nch; b++) {
    c->segs[b].band = b;
    c->segs[b].lo = b * (nfft / nch);
    c->segs[b].hi = c->segs[b].lo + nfft / nch;
  }
  return 0;
}

/* Synthetic capture defaults: 13 MS/s receiver (.cpp:70), 1.4 MS/s PU (predictive_model.cfg:39) on 833 / 835 / 838 MHz
 * (.hpp:55-57), soft gain -12 dB (ecr.cpp:59), seed 12 (crts_cognitive_radio.cpp:754), no interferer. */
void crn_oracle_synth_config_default(crn_synth_config *sc, int32_t group_samples) {
  memset(sc, 0, sizeof(*sc));
  sc->seed = 12;
  sc->fs = 13e6;
  sc->pu_rate = 1.4e6;
  sc->offsets_hz[1] = 835e6 - 833e6;
  sc->offsets_hz[2] = 838e6 - 833e6;
  sc->snr_db = 10.0;
  sc->pu_gain_db = -12.0;
  sc->dwell_groups = 64;
  sc->group_samples = group_samples;
  sc->intf_type = CRN_INTF_NONE;
  sc->intf_rate = 1e6;
  sc->intf_gain_db = -3.0;
  sc->intf_duty = 1.0;
}
