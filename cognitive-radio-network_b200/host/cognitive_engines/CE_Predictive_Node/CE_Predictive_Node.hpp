// GPU-backed CE_Predictive_Node: same directory name, class name, constructor convention and cfg string
// (cognitive_engine = "CE_Predictive_Node", scenarios/predictive_model.cfg:60) as the reference engine
// (cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.{hpp,cpp}), so the shipped scenario selects
// it unmodified.  The control flow of execute() is the reference's; the arithmetic
// (.cpp:149-154 per frame, .cpp:163-261 per decision) runs in libcrnsense on the B200.
#ifndef _CE_PREDICTIVE_NODE_
#define _CE_PREDICTIVE_NODE_

#include <sys/time.h>

#include <complex>
#include <vector>

#include "cognitive_engine.hpp"
#include "crnsense.h"
#include "extensible_cognitive_radio.hpp"

class CE_Predictive_Node : public CognitiveEngine {
private:
  // sensing parameters: defaults are the reference's constants (.hpp:30-33,42-43,55-57);
  // unlike upstream they can be overridden per scenario through ce_args (see the constructor)
  float sensing_delay_ms;
  int fft_length;
  int fft_averaging;
  float Desired_fc;
  float Desired_BW;
  float CHANNEL1, CHANNEL2, CHANNEL3;

  long int sense_time_s;
  long int sense_time_us;
  int config;
  int fft_counter;
  int quiet;
  bool custom_weights;

  crn_config cfg;
  crn_handle *sense;
  FILE *result_log;

  int load_weights(const char *path);  // -m <file>: the reference's `WeightIH[i][j] = v;` syntax
  // receiver hook: the pinned ring slot the next packet should be received into (see extensible_cognitive_radio.hpp)
  static std::complex<float> *rx_slot(void *self, size_t nsamples, int *overflow);

public:
  CE_Predictive_Node(int argc, char **argv, ExtensibleCognitiveRadio *_ECR);
  ~CE_Predictive_Node();
  virtual void execute();

  // what upstream only printf()s: every decision taken so far, in order
  std::vector<crn_result> decisions;
};

#endif
