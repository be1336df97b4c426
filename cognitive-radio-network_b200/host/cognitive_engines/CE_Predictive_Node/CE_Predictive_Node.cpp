// See CE_Predictive_Node.hpp.  Line references are to the reference's CE_Predictive_Node.cpp.
#include "CE_Predictive_Node.hpp"

#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// ce_args (scenario key `ce_args`, parsed getopt-style like CE_Template.cpp:16-24; the reference engine
// ignores argc/argv):
//   -n <nfft>      FFT length            (default 512,  .hpp:31)
//   -k <frames>    frames per decision   (default 10,   .hpp:32)
//   -w <0|1>       0 rectangular (default, = upstream), 1 Hann
//   -p <0|1>       0 |X| and (sum)^2 (default, = upstream), 1 |X|^2 Welch band power
//   -d <ms>        re-arm period         (default 100,  .hpp:30)
//   -g <device>    CUDA device ordinal   (default 0)
//   -o <path>      append every crn_result (binary) to this file
//   -m <path>      MLP weights to use instead of the built-in literals, written exactly like the reference's
//                  assignment block (.cpp:78-120): lines `WeightIH[i][j] = v;` / `WeightHO[j][k] = v;`
//                  (e.g. the outcome of crn_ann_train_device); weights not listed keep the reference's value
//   -q             do not print the per-decision banner
CE_Predictive_Node::CE_Predictive_Node(int argc, char **argv, ExtensibleCognitiveRadio *_ECR) {
  ECR = _ECR;
  sensing_delay_ms = 1e2;
  fft_length = 512;
  fft_averaging = 10;
  Desired_fc = 833e6;
  Desired_BW = 13e6;
  CHANNEL1 = 833e6;
  CHANNEL2 = 835e6;
  CHANNEL3 = 838e6;
  fft_counter = 0;
  config = 0;
  quiet = 0;
  sense = NULL;
  result_log = NULL;
  int window = CRN_WINDOW_RECT, power = 0, device = 0;
  const char *log_path = NULL, *weights_path = NULL;

  int o;
  optind = 0;  // str2argcargv leaves it at 0 as well (src/crts.cpp:80)
  while (argc > 0 && argv && (o = getopt(argc, argv, "n:k:w:p:d:g:o:m:q")) != -1) {
    switch (o) {
      case 'n': fft_length = atoi(optarg); break;
      case 'k': fft_averaging = atoi(optarg); break;
      case 'w': window = atoi(optarg) ? CRN_WINDOW_HANN : CRN_WINDOW_RECT; break;
      case 'p': power = atoi(optarg); break;
      case 'd': sensing_delay_ms = (float)atof(optarg); break;
      case 'g': device = atoi(optarg); break;
      case 'o': log_path = optarg; break;
      case 'm': weights_path = optarg; break;
      case 'q': quiet = 1; break;
      default: break;
    }
  }

  struct timeval tv;
  gettimeofday(&tv, NULL);
  sense_time_s = tv.tv_sec;
  sense_time_us = tv.tv_usec;

  // replaces the buffer zeroing + fft_create_plan (.cpp:36-45) and the weight literals (.cpp:78-120)
  if (fft_length == 512 && !power) {
    crn_config_reference(&cfg);
  } else if (crn_config_welch(&cfg, fft_length, fft_averaging) != CRN_OK) {
    printf("CE_Predictive_Node: %s\n", crn_last_error());
    exit(EXIT_FAILURE);
  }
  cfg.navg = fft_averaging;
  cfg.window = window;
  if (!power) {
    cfg.detector = CRN_DET_MAG;
    cfg.postop = CRN_POST_SQUARE_OF_SUM;
  }
  cfg.device = device;
  custom_weights = weights_path != NULL;
  if (weights_path && load_weights(weights_path) < 0) exit(EXIT_FAILURE);  // the reference's error convention
  if (log_path) result_log = fopen(log_path, "wb");
}

// Weight file in the reference's own syntax (.cpp:78-120).  Returns the number of weights read, -1 on error.
int CE_Predictive_Node::load_weights(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) {
    printf("CE_Predictive_Node: cannot open weight file %s\n", path);
    return -1;
  }
  char line[256];
  int n = 0, lineno = 0;
  while (fgets(line, sizeof(line), f)) {
    lineno++;
    int a, b;
    double v;
    const char *p = line;
    while (*p == ' ' || *p == '\t') p++;
    if (*p == '\0' || *p == '\n' || *p == '#' || (p[0] == '/' && p[1] == '/')) continue;
    if (sscanf(p, "WeightIH[%d][%d] = %lf", &a, &b, &v) == 3) {
      if (a < 0 || a > CRN_ANN_INPUTS || b < 1 || b > CRN_ANN_HIDDEN) goto bad;
      cfg.ann_wih[a][b] = v;
    } else if (sscanf(p, "WeightHO[%d][%d] = %lf", &a, &b, &v) == 3) {
      if (a < 0 || a > CRN_ANN_HIDDEN || b < 1 || b > CRN_ANN_OUTPUTS) goto bad;
      cfg.ann_who[a][b] = v;
    } else {
      goto bad;
    }
    n++;
  }
  fclose(f);
  return n;
bad:
  printf("CE_Predictive_Node: %s:%d: not a WeightIH[i][j] / WeightHO[j][k] assignment: %s", path, lineno, line);
  fclose(f);
  return -1;
}

// Receiver-side hook (ExtensibleCognitiveRadio::set_rx_slot_provider): where the next packet should be received.
// Runs on the rx thread with CE_mutex held, i.e. never concurrently with execute().
std::complex<float> *CE_Predictive_Node::rx_slot(void *self, size_t nsamples, int *overflow) {
  CE_Predictive_Node *ce = (CE_Predictive_Node *)self;
  void *slot = NULL;
  if (!ce->sense || (int)nsamples != ce->cfg.frame_len) return NULL;
  const int st = crn_ring_acquire(ce->sense, &slot);
  if (st == CRN_ERR_OVERRUN && overflow) *overflow = 1;
  return st == CRN_OK ? (std::complex<float> *)slot : NULL;
}

CE_Predictive_Node::~CE_Predictive_Node() {
  if (sense) crn_destroy(sense);
  if (result_log) fclose(result_log);
}

void CE_Predictive_Node::execute() {
  // one-shot configuration (.cpp:66-123)
  if (config == 0) {
    ECR->stop_tx();
    ECR->set_rx_freq(Desired_fc);
    ECR->set_rx_rate(Desired_BW);
    // The packet length is only known once the receiver runs (upstream reads
    // ce_usrp_rx_buffer_length on every frame, .cpp:149).  Upstream overruns buffer[512] when a packet
    // is longer than the FFT; here that is a configuration error.
    cfg.frame_len = ECR->ce_usrp_rx_buffer_length;
    if (cfg.frame_len < 1) return;  // receiver not started yet: try again on the next event
    int st = crn_create(&cfg, &sense);
    if (st != CRN_OK) {
      printf("CE_Predictive_Node: crn_create failed: %s: %s\n", crn_strerror(st), crn_last_error());
      exit(EXIT_FAILURE);
    }
    config = 1;
#ifdef CRN_ECR_HAS_RX_SLOT_PROVIDER
    // from now on the receiver may recv() straight into the pinned ring (no memcpy here or under CE_mutex)
    if (!getenv("CRN_ENGINE_COPY"))  // CRN_ENGINE_COPY=1: keep upstream's two copies per packet (A/B of the handoff)
      ECR->set_rx_slot_provider(&CE_Predictive_Node::rx_slot, this);
#endif
    if (!quiet && (cfg.nfft != 512 || cfg.detector != CRN_DET_MAG || cfg.postop != CRN_POST_SQUARE_OF_SUM ||
                   cfg.window != CRN_WINDOW_RECT) && !custom_weights)
      printf("CE_Predictive_Node: note: the built-in MLP literals (.cpp:78-120) were trained on (sum|X|)^2 features of "
             "the 512-point rectangular mode; with -n/-w/-p changed, supply retrained weights with -m\n");
  }

  // sensing gate (.cpp:127-141), including upstream's habit of not carrying microseconds into seconds
  struct timeval tv;
  gettimeofday(&tv, NULL);
  if ((tv.tv_sec > sense_time_s) || ((tv.tv_sec == sense_time_s) && (tv.tv_usec >= sense_time_us))) {
    ECR->stop_tx();
    ECR->set_ce_sensing(1);
    sense_time_s = tv.tv_sec + (long int)floorf(sensing_delay_ms / 1e3);
    sense_time_us = tv.tv_usec + (long int)floorf(sensing_delay_ms * 1e3);
  }

  // handle samples (.cpp:146-154): stage the packet in the pinned ring and commit it
  if (ECR->CE_metrics.CE_event == ExtensibleCognitiveRadio::USRP_RX_SAMPS) {
    fft_counter++;
    void *slot = NULL;
    int st = crn_ring_acquire(sense, &slot);
    if (st == CRN_OK) {
      // the receiver may already have put the packet where it belongs (rx_slot below); otherwise copy it (.cpp:149)
      if ((void *)ECR->ce_usrp_rx_buffer != slot)
        memmove(slot, ECR->ce_usrp_rx_buffer, (size_t)ECR->ce_usrp_rx_buffer_length * sizeof(float) * 2);
      st = crn_submit(sense, 1);
    }
    if (st != CRN_OK) {
      printf("CE_Predictive_Node: %s: %s\n", crn_strerror(st), crn_last_error());
      if (st == CRN_ERR_OVERRUN) ECR->CE_metrics.CE_event = ExtensibleCognitiveRadio::UHD_OVERFLOW;
      crn_reset(sense);
      fft_counter = 0;
      return;
    }

    if (fft_counter == fft_averaging) {
      ECR->set_ce_sensing(0);  // .cpp:159
      crn_result r;
      st = crn_wait(sense, &r);  // features, MLP outputs and decision computed on the GPU (.cpp:163-261)
      if (st != CRN_OK) {
        printf("CE_Predictive_Node: crn_wait: %s: %s\n", crn_strerror(st), crn_last_error());
        exit(EXIT_FAILURE);
      }
      decisions.push_back(r);
      if (result_log) {
        fwrite(&r, sizeof(r), 1, result_log);
        fflush(result_log);
      }
      if (!quiet) {
        printf("--------------------------------------------------------------\n");
        printf("-            \t\tFEATURES BUFFER \t               -\n");
        printf("--------------------------------------------------------------\n");
        printf("NOISE FLOOR   %.2e\nCH1           %.2e\nCH2           %.2e\nCH3           %.2e\n ",
               r.feat[0], r.feat[1], r.feat[2], r.feat[3]);
        printf("\n \n \n --------------------------------------------------------------\n");
        printf("-            \t\t REAL TIME PREDICTION                  -\n");
        printf("--------------------------------------------------------------\n");
      }
      // first-match chain and retune (.cpp:245-261)
      switch (r.decision) {
        case CRN_CH1_OCCUPIED:
          if (!quiet) printf("Channel_State[1]: OCCUPIED \nChannel_State[2]: FREE \nChannel_State[3]: FREE \n \n \n");
          ECR->set_tx_freq(CHANNEL2);
          break;
        case CRN_CH2_OCCUPIED:
          if (!quiet) printf("Channel_State[1]: FREE \nChannel_State[2]: OCCUPIED \nChannel_State[3]: FREE \n \n \n");
          ECR->set_tx_freq(CHANNEL1);
          break;
        case CRN_CH3_OCCUPIED:
          if (!quiet) printf("Channel_State[1]: FREE \nChannel_State[2]: FREE \nChannel_State[3]: OCCUPIED \n \n \n");
          ECR->set_tx_freq(CHANNEL2);
          break;
        default:
          if (!quiet) printf("ALL BUSY, SENSE AND OBSERVE AGAIN \n");
          break;
      }
      fft_counter = 0;  // .cpp:287-288 (the averaging buffer lives on the GPU and is per decision)
    }
  }
}
