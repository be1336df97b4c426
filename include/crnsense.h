/*
 * libcrnsense — C-ABI of the B200-native spectrum-sensing engine.
 *
 * This is the drop-in boundary for ONE hot path of 0xastro/Cognitive-Radio-Network: the
 * CE_Predictive_Node sensing loop (IQ frame -> [window] -> FFT -> |X| or |X|^2 -> K-frame average ->
 * per-band sums -> power features -> 4-5-3 logistic MLP -> occupancy decision).  Every entry point
 * below cites the reference code (paths under the reference tree) it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; pointers + sizes only.
 *   - every function returns an int status (CRN_OK == 0, errors < 0); nothing here calls exit()
 *     (the reference printf()s and exit()s on fatal conditions, e.g. src/crts.cpp:306-310).
 *   - IQ is interleaved complex-float32 (re,im), 8 bytes per sample: the layout of
 *     ExtensibleCognitiveRadio::ce_usrp_rx_buffer (include/extensible_cognitive_radio.hpp:547).
 *   - one handle == one sensing stream on one GPU.  A handle is not thread-safe; different handles are
 *     independent (same rule as the reference: execute() only ever runs on the single ECR_ce_worker
 *     thread, src/extensible_cognitive_radio.cpp:1761-1808).
 *   - there is NO CPU fallback.  Without a CUDA device crn_create() fails with CRN_ERR_NO_DEVICE.
 */
#ifndef CRNSENSE_H
#define CRNSENSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRN_VERSION_MAJOR 0
#define CRN_VERSION_MINOR 1

#define CRN_MAX_BANDS 64   /* features per decision (reference: 4 = NF, CH1, CH2, CH3) */
#define CRN_MAX_SEGS 128   /* contiguous bin ranges; a band is the union of >=1 segments */
#define CRN_ANN_INPUTS 4   /* INPUTS          CE_Predictive_Node.hpp:20 */
#define CRN_ANN_HIDDEN 5   /* HIDDEN_NEURONS  CE_Predictive_Node.hpp:21 */
#define CRN_ANN_OUTPUTS 3  /* OUTPUT_NEURONS  CE_Predictive_Node.hpp:22 */

/* status codes */
enum {
  CRN_OK = 0,
  CRN_ERR_INVALID = -1,     /* bad argument / inconsistent configuration */
  CRN_ERR_NO_DEVICE = -2,   /* no CUDA device, or cfg.device out of range */
  CRN_ERR_CUDA = -3,        /* a CUDA runtime call failed (see crn_last_error) */
  CRN_ERR_NOMEM = -4,
  CRN_ERR_OVERRUN = -5,     /* ring full: producer outran the GPU (cf. UHD_OVERFLOW event,
                               src/extensible_cognitive_radio.cpp:1327-1336) */
  CRN_ERR_NOT_READY = -6,   /* crn_poll: no finished decision yet */
  CRN_ERR_UNSUPPORTED = -7  /* e.g. nfft not one of the compiled sizes */
};

enum crn_window { CRN_WINDOW_RECT = 0, CRN_WINDOW_HANN = 1 };
enum crn_detector { CRN_DET_MAG = 0, CRN_DET_MAGSQ = 1 };
/* SUM_DB: the reported feature is 10 log10(sum) in dB (Welch band power in dB); the MLP / energy detector
   still see the linear sum. */
enum crn_postop { CRN_POST_SQUARE_OF_SUM = 0, CRN_POST_SUM = 1, CRN_POST_SUM_DB = 2 };
enum crn_decide { CRN_DECIDE_NONE = 0, CRN_DECIDE_ANN = 1, CRN_DECIDE_ENERGY = 2 };
/* Sample format of every IQ buffer handed to the library.  CF32: interleaved float32 (re,im), 8 B/sample -
   what UHD's fc32 host format and ce_usrp_rx_buffer hold.  SC16: interleaved int16 (I,Q), 4 B/sample - the
   USRP's over-the-wire format (UHD otw_format "sc16"); sample value = int16 / 32768.  Converting on the GPU
   halves the bytes that cross PCIe on the live-ingest path. */
enum crn_iq_format { CRN_IQ_CF32 = 0, CRN_IQ_SC16 = 1 };

/* Outcome of the reference's first-match chain, CE_Predictive_Node.cpp:245-261. */
enum crn_decision {
  CRN_ALL_BUSY = 0,      /* no output >= threshold: "ALL BUSY, SENSE AND OBSERVE AGAIN" (:260-261) */
  CRN_CH1_OCCUPIED = 1,  /* Output[1] >= 0.8 -> set_tx_freq(CHANNEL2 = 835e6)  (:245-248) */
  CRN_CH2_OCCUPIED = 2,  /* Output[2] >= 0.8 -> set_tx_freq(CHANNEL1 = 833e6)  (:250-253) */
  CRN_CH3_OCCUPIED = 3   /* Output[3] >= 0.8 -> set_tx_freq(CHANNEL2 = 835e6)  (:255-258) */
};

/* bins [lo, hi) of the N-point spectrum contribute to feature `band` (summed in the order listed). */
typedef struct crn_seg {
  int32_t band, lo, hi;
} crn_seg;

/*
 * Sensing configuration.  Replaces the reference's compile-time constants
 * (CE_Predictive_Node.hpp:30-33,42-43,55-57), its hard-coded bin table (CE_Predictive_Node.cpp:173-191)
 * and its weight literals (CE_Predictive_Node.cpp:78-120).
 */
typedef struct crn_config {
  int32_t nfft;          /* N, power of two in [256, 8192]            (reference: fft_length = 512)   */
  int32_t frame_len;     /* L <= N valid samples per frame; tail [L,N) is zero (the reference's memcpy
                            of ce_usrp_rx_buffer_length samples into a zeroed buffer[512], .cpp:37,149) */
  int32_t frame_stride;  /* samples between consecutive frame starts in a batch buffer (>= L; 0 -> L) */
  int32_t navg;          /* K frames averaged per decision             (reference: fft_averaging = 10) */
  int32_t window;        /* enum crn_window                            (reference: none == RECT)       */
  int32_t detector;      /* enum crn_detector: per-bin |X| (reference, .cpp:153) or |X|^2             */
  int32_t postop;        /* enum crn_postop: feature = (sum)^2 (reference, .cpp:194-197) or sum        */
  int32_t decide;        /* enum crn_decide                                                            */
  int32_t nbands;        /* number of features, 1..CRN_MAX_BANDS                                       */
  int32_t nsegs;         /* number of entries used in segs[], 1..CRN_MAX_SEGS                          */
  crn_seg segs[CRN_MAX_SEGS];
  /* MLP weights with the reference's own 1-based indexing: ann_wih[i][j] == WeightIH[i][j],
     i = 0 (bias) .. 4, j = 1..5;  ann_who[j][k] == WeightHO[j][k], j = 0 (bias) .. 5, k = 1..3.
     Row/column 0 of the "other" index is unused, as in CE_Predictive_Node.hpp:66-72.
     ANN inputs are features 0..3 in order, i.e. Features_Buffer[1..4] = {NF^2, CH1, CH2, CH3}
     (.cpp:200) -> band 0 must be the noise-floor band. */
  double ann_wih[CRN_ANN_INPUTS + 1][CRN_ANN_HIDDEN + 1];
  double ann_who[CRN_ANN_HIDDEN + 1][CRN_ANN_OUTPUTS + 1];
  double ann_threshold;  /* 0.8, .cpp:245,250,255 */
  double energy_factor;  /* CRN_DECIDE_ENERGY: band c is occupied iff feat[c] > energy_factor * min_c feat */
  int32_t device;        /* CUDA device ordinal */
  int32_t ring_slots;    /* streaming API: number of K-frame decision slots in the pinned ring (>= 2) */
  int32_t iq_format;     /* enum crn_iq_format of all IQ buffers (ring slots, batch host/device)            */
  int32_t reserved_;     /* keep the struct a multiple of 8 bytes */
} crn_config;

/* One decision (K frames).  Replaces the locals/members the reference only printf()s
   (CE_Predictive_Node.cpp:163-261; Output[] is CE_Predictive_Node.hpp:72). */
typedef struct crn_result {
  uint64_t first_frame;       /* index of the first of the K frames (streaming API: running count) */
  int32_t decision;           /* enum crn_decision (CRN_DECIDE_ANN), else 0 */
  int32_t nfeat;              /* == cfg.nbands */
  uint64_t occupancy_mask;    /* CRN_DECIDE_ENERGY: bit c set <=> band c occupied; ANN: 1<<(ch-1) */
  double ann_out[CRN_ANN_OUTPUTS]; /* Output[1..3] */
  float feat[CRN_MAX_BANDS];  /* feat[0..nfeat): reference order NF^2, CH1, CH2, CH3 */
} crn_result;

typedef struct crn_handle crn_handle;

/* ---- configuration helpers ------------------------------------------------------------------ */

/* Fill *cfg with the reference-exact mode: N=512, L=512, K=10, rect, |X|, square-of-sum, the four
   bands of CE_Predictive_Node.cpp:173-191 (band 0 = NF [300,310), 1 = CH1 [0,16)+[496,511),
   2 = CH2 [55,85), 3 = CH3 [189,222)), the 43 weights of .cpp:78-120, threshold 0.8. */
int crn_config_reference(crn_config *cfg);

/* Same band plan scaled to an N-point FFT (bin indices x N/512), Hann window, |X|^2, K frames,
   feature = sum (Welch band power); ANN on.  This is BASELINE config 2 at nfft=1024, navg=64. */
int crn_config_welch(crn_config *cfg, int32_t nfft, int32_t navg);

/* nbands equal, contiguous sub-channels covering all N bins (BASELINE config 3 at nfft=8192,
   nbands=64); |X|^2, Hann, feature = sum, CRN_DECIDE_ENERGY. */
int crn_config_wideband(crn_config *cfg, int32_t nfft, int32_t navg, int32_t nbands);

/* Validate a configuration without touching the GPU. Returns CRN_OK or CRN_ERR_INVALID/UNSUPPORTED. */
int crn_config_validate(const crn_config *cfg);

/* ---- lifetime --------------------------------------------------------------------------------- */

/* Replaces the CE_Predictive_Node constructor's buffer zeroing + fft_create_plan
   (CE_Predictive_Node.cpp:36-45) and the one-shot weight load (.cpp:78-120): uploads window /
   twiddle / band / weight tables, allocates the pinned host ring and its device mirror. */
int crn_create(const crn_config *cfg, crn_handle **out);
int crn_destroy(crn_handle *h);

/* ---- streaming path: what a CognitiveEngine::execute() calls per USRP_RX_SAMPS event ---------- */

/* Pointer to the pinned host slot for the NEXT frame (frame_len samples; 8*L bytes CF32, 4*L bytes SC16).  This is the
   memcpy target that replaces ECR->ce_usrp_rx_buffer in the rx-worker handoff
   (src/extensible_cognitive_radio.cpp:1316-1317) / the engine's own memcpy (.cpp:149). */
int crn_ring_acquire(crn_handle *h, void **slot);

/* Commit `nframes` (normally 1) frames written through crn_ring_acquire.  Non-blocking and stream
   ordered.  When the K-th frame of a decision has been committed the K frames are copied to the GPU
   and the fused sensing kernel is enqueued (replaces .cpp:148-154 and, on the K-th frame, :157-261).
   Returns CRN_ERR_OVERRUN if every ring slot still holds an unread decision. */
int crn_submit(crn_handle *h, int32_t nframes);

/* ---- many co-located radios: one launch per decision round ------------------------------------------------------
   The reference runs one engine per process, one process per radio (src/crts_cognitive_radio.cpp:754-812); a host
   that senses for R radios (a CORNET rack, BASELINE configs[3]) would otherwise pay R launches per decision round.
   crn_create_many makes R streaming handles of ONE configuration on one GPU that share their tables, their stream
   and one pinned ring (slot layout [radio][K][L]); each is used exactly like a crn_create handle
   (crn_ring_acquire / crn_submit / crn_poll / crn_wait / crn_reset / crn_destroy; the batch calls are refused).
   crn_submit_many commits `nframes` frames on each of the n handles; when that completes their decisions and the
   handles are ALL the members of one pool, in step (same slot, same frame count), the n decisions are sensed by one
   host->device copy and ONE launch of the fused kernel over n decision groups.  Handles that do not qualify are
   submitted one by one - same results, n launches.  Results are fetched per handle with crn_poll / crn_wait. */
int crn_create_many(const crn_config *cfg, int32_t n, crn_handle **out /* [n] */);
int crn_submit_many(crn_handle *const *handles, int32_t n, int32_t nframes);

/* Non-blocking / blocking fetch of the oldest finished decision. */
int crn_poll(crn_handle *h, crn_result *out);
int crn_wait(crn_handle *h, crn_result *out);

/* Drop frames of a partially filled decision (the reference zeroes fft_avg / fft_counter, .cpp:287-288). */
int crn_reset(crn_handle *h);

/* ---- batch path ------------------------------------------------------------------------------- */

/* ngroups decisions from HOST memory: iq holds ngroups*K frames in cfg.iq_format (frame f starts at sample
   f*frame_stride).  Stages through pinned buffers, host->device copy, kernel, device->host read of the
   results; returns when results[0..ngroups) are filled. */
int crn_sense_batch_host(crn_handle *h, const void *iq, int64_t ngroups, crn_result *results);

/* ngroups decisions from DEVICE memory, asynchronous on `cuda_stream` (a cudaStream_t; NULL = the
   legacy default stream).  Output arrays are device pointers; any of d_ann / d_decision / d_mask may
   be NULL.  d_feat: float[ngroups][nbands]; d_ann: double[ngroups][3]; d_decision: int32[ngroups];
   d_mask: uint64[ngroups].  IQ is read from HBM exactly once; only these features are written. */
/* Reproducibility: a launch is deterministic - the same handle configuration and the same ngroups give the same
   bits, every time.  Launches of different shapes (one 300-group batch vs chunks of 128, the one-decision
   launches of the streaming path) may deal a group's frames to a different number of CTAs, which re-associates
   the fp32 band sums: such results agree to fp32 rounding (~1e-7 relative), far inside the 1e-4 parity bar.
   The call never allocates and never synchronises (scratch for split launches is sized at crn_create), so it may be
   captured into a CUDA graph.  Launches of one handle on DIFFERENT streams are safe: the few-group launches that
   share the handle's scratch are chained with an event (the later one waits for the earlier); large batches run
   concurrently.  A handle is still not thread-safe: call it from one host thread at a time. */
int crn_sense_batch_device(crn_handle *h, const void *d_iq, int64_t ngroups, float *d_feat,
                           double *d_ann, int32_t *d_decision, uint64_t *d_mask, void *cuda_stream);

/* ---- cooperative sensing -------------------------------------------------------------- */

enum crn_fusion { CRN_FUSE_OR = 0, CRN_FUSE_MAJORITY = 1, CRN_FUSE_AND = 2 };

/* Hard-decision fusion across cooperating radios (the CSS scheme of the reference's project documentation):
   d_masks is uint64[nradios][nslots] (the occupancy_mask of radio r in time slot s, e.g. gathered from
   several GPUs); d_fused[s] gets bit c set when band c is reported occupied by any / more than half / all
   of the radios.  Asynchronous on cuda_stream. */
int crn_fuse_masks_device(const uint64_t *d_masks, int64_t nradios, int64_t nslots, int32_t nbands,
                          int32_t mode, uint64_t *d_fused, int32_t device, void *cuda_stream);

/* ---- the occupancy predictor on its own (north_star: batched FMA kernel; SURVEY 8f-4: retraining) ---- */

/* The 43 weights of the 4-5-3 logistic MLP with the reference's 1-based indexing (same layout as
   crn_config.ann_wih / ann_who; CE_Predictive_Node.hpp:66-72, literals .cpp:78-120). */
typedef struct crn_ann_weights {
  double wih[CRN_ANN_INPUTS + 1][CRN_ANN_HIDDEN + 1];
  double who[CRN_ANN_HIDDEN + 1][CRN_ANN_OUTPUTS + 1];
} crn_ann_weights;

/* Batched forward pass + first-match chain (CE_Predictive_Node.cpp:214-261) over n feature vectors that are
   already in device memory: d_feat is float[n][feat_stride], inputs are columns 0..3 = NF^2, CH1, CH2, CH3
   (.cpp:200).  d_out: double[n][3] = Output[1..3] (may be NULL); d_decision: int32[n] enum crn_decision (may
   be NULL).  One thread per decision, fp64 FMA + exp, the reference's summation order.  Asynchronous. */
int crn_ann_forward_device(const crn_ann_weights *w, double threshold, const float *d_feat, int64_t n,
                           int32_t feat_stride, double *d_out, int32_t *d_decision, int32_t device,
                           void *cuda_stream);

/* On-device (re)training of the predictor from labelled feature vectors ("Array of features + label",
   Data Generation/TODO.md:1-7).  The reference ships only the outcome of its offline training
   ("Error = 0.000100 after 63.145737 Milion Epoch", CE_Predictive_Node.cpp:74); the trainer here is plain
   batch back-propagation for the same network: logistic units, sum-of-squares error
   E = 1/2 sum_p sum_k (t_pk - Output_pk)^2, update dW = eta * (-dE/dW) / n + alpha * dW_previous. */
typedef struct crn_ann_train_config {
  int32_t max_epochs;   /* passes over the n examples */
  int32_t check_every;  /* epochs per CUDA-graph replay; E is read back (and target_error tested) after each */
  double eta, alpha;    /* learning rate, momentum */
  double target_error;  /* stop once E <= target_error (<= 0: run max_epochs) */
  double input_scale[CRN_ANN_INPUTS]; /* the network trains on feat[i] * input_scale[i] (raw powers are
                           1e4..1e9 and saturate every unit); the scale is folded back into wih[i][*] on
                           return so the weights apply to raw features, as the engine feeds them.  0 -> 1 */
  double init_range;    /* > 0: start from weights uniform in (-init_range, init_range) drawn from seed;
                           0: start from *w as passed in (given for raw features, unfolded internally) */
  uint64_t seed;
} crn_ann_train_config;

int crn_ann_train_config_default(crn_ann_train_config *tc);

/* d_feat: float[n][feat_stride] (device), d_labels: int32[n] (device) enum crn_decision - target is 1 for the
   labelled channel's output and 0 elsewhere (CRN_ALL_BUSY: all 0).  *w is updated in place; *final_error = E of
   the last epoch run; *epochs_run = how many ran.  Synchronous (returns when training has finished). */
int crn_ann_train_device(const crn_ann_train_config *tc, const float *d_feat, int32_t feat_stride,
                         const int32_t *d_labels, int64_t n, crn_ann_weights *w, double *final_error,
                         int32_t *epochs_run, int32_t device, void *cuda_stream);

/* ---- synthetic primary-user IQ (stands in for the USRP) ----------------------- */

typedef struct crn_synth_config {
  uint64_t seed;
  double fs;            /* capture rate, 13e6 (scenarios/predictive_model.cfg:76) */
  double pu_rate;       /* PU sample rate, 1.4e6 (scenarios/predictive_model.cfg:39) */
  double offsets_hz[3]; /* PU centre offsets from fc for CH1..CH3: 0, +2e6, +5e6 (CE_Predictive_Node.hpp:55-57) */
  double snr_db;        /* in-band SNR of the PU against the AWGN */
  double pu_gain_db;    /* soft gain, -12 dB (src/extensible_cognitive_radio.cpp:59) */
  int32_t hop_mode;     /* 0: Markov as documented (README.md:70-74), 1: Markov as coded
                           (CE_PU_MARKOV_Chain_Tx.cpp:97-128), 2: uniform (CE_Random_Behaviour_PU.cpp:47-49) */
  int32_t dwell_groups; /* decisions per PU dwell */
  int32_t group_samples;/* samples per decision group (K * frame_stride) */
  /* Optional interferer node (src/interferer.cpp), added on top of the PU and the noise: */
  int32_t intf_type;    /* enum crn_interferer: 0 none */
  int32_t intf_period_groups; /* duty-cycle period in decision groups (its wall-clock `period`, interferer.cpp:28,
                           395-409, scaled like the PU dwell); <= 0: always on */
  int32_t pu_framed;    /* 0: the PU sends payload OFDM symbols back to back; 1: flex-frame structure of transmit_frame
                           (src/extensible_cognitive_radio.cpp:883-949): frames of 32 symbols = S0, S0, S1, 7 header
                           symbols (BPSK), 22 payload symbols (QPSK) */
  double intf_offset_hz;/* interferer tx_freq - fc */
  double intf_rate;     /* its sample rate (tx_rate); every generated sample is held for fs/intf_rate receiver samples */
  double intf_gain_db;  /* soft gain (default -3 dB, interferer.cpp:32) plus whatever path loss is wanted */
  double intf_duty;     /* on for duty*period, then off for (1-duty)*period (interferer.cpp:395-409) */
} crn_synth_config;

/* Interference waveforms of src/interferer.cpp that need no modem: CW = the constant 0.5+0.5j of
   BuildCWTransmission (:128-134), NOISE = uniform in [-0.25, 0.25) per component (BuildNOISETransmission
   :136-142), AWGN = Gaussian with MEAN 5 and sigma 5 per component, as coded (dist(5.0, 5.0) :24, :144-154). */
/* ... and the three that need one (BuildGMSKTransmission :156-221, BuildRRCTransmission :223-253,
   BuildOFDMTransmission :255-288).  liquid-dsp's frame generators are absent, so frames are restated at the level the
   sensing path can see - pulse shape, symbol rate, frame cadence, constant-envelope / PAPR character - with random
   symbols in place of liquid's coded header/payload bits:
     GMSK  BT = 0.5, modulation index 1/2, 2 samples/symbol interpolated by 2 (:189-205) = 4 samples/symbol; frames
           of 1024 symbols followed by the 12 padding samples of :212-219;
     RRC   QPSK symbols +-0.25 +- 0.25j (:237-240) zero-stuffed to 2 samples/symbol through the 129-tap root raised
           cosine (beta 0.35, semi-length 32, :61-65), filter reset every 200-sample frame (:231); as coded the
           "stop a filter length before the end" test (:236) never fires (unsigned arithmetic), so all 100 symbols
           of a frame are sent;
     OFDM  64 subcarriers, cyclic prefix 16, taper 6 (:23-24,66-71), liquid's default allocation; frames of 22
           symbols = S0, S0, S1, 7 header symbols (BPSK), 12 payload symbols (QPSK: 128 bytes + CRC32 over 44 data
           subcarriers), back to back.
   Like the other waveforms they are generated at the interferer's rate and held to the receiver's. */
enum crn_interferer { CRN_INTF_NONE = 0, CRN_INTF_CW = 1, CRN_INTF_NOISE = 2, CRN_INTF_AWGN = 3,
                      CRN_INTF_GMSK = 4, CRN_INTF_RRC = 5, CRN_INTF_OFDM = 6 };

int crn_synth_config_default(crn_synth_config *sc, int32_t group_samples);

/* Fill d_iq (device, nsamples complex-float) with the synthetic capture, asynchronously on
   cuda_stream; if d_state != NULL also write the PU channel (0..2) active in each group
   (int32[ceil(nsamples/group_samples)]).  first_sample offsets the stream (sharding across GPUs). */
int crn_synth_generate_device(const crn_synth_config *sc, int32_t device, void *d_iq,
                              int64_t first_sample, int64_t nsamples, int32_t *d_state,
                              void *cuda_stream);

/* Multi-radio variant (BASELINE configs[3]: independent sensing streams, simulated CORNET nodes): stream
   first_stream + i has its own seed, hop chain and noise, starts at its sample 0 and is written to
   d_iq + i * samples_per_stream.  d_state (optional): int32[nstreams][samples_per_stream / group_samples]. */
int crn_synth_generate_streams_device(const crn_synth_config *sc, int32_t device, void *d_iq,
                                      int64_t first_stream, int64_t nstreams, int64_t samples_per_stream,
                                      int32_t *d_state, void *cuda_stream);

/* ---- diagnostics ------------------------------------------------------------------------------ */

const char *crn_strerror(int status);
/* Text of the most recent failure on this thread (CUDA error string etc.). */
const char *crn_last_error(void);
int crn_version(int32_t *major, int32_t *minor);
/* Number of CUDA devices visible, or a negative status. */
int crn_device_count(void);
/* Kernels launched by this handle so far (sensing + synth), for launch accounting. */
int64_t crn_launch_count(const crn_handle *h);
/* Static facts about the kernel variant this handle dispatches to. */
typedef struct crn_kernel_info {
  int32_t nfft, threads_per_frame, elems_per_thread, teams_per_cta, threads_per_cta;
  int32_t ctas_per_sm, grid, smem_bytes, regs_per_thread, num_sms;
  char name[64];
} crn_kernel_info;
int crn_get_kernel_info(const crn_handle *h, crn_kernel_info *info);

#ifdef __cplusplus
}
#endif
#endif /* CRNSENSE_H */
