// crn_replay: run one cognitive-radio node of a CRTS scenario against a recorded/synthetic IQ capture,
// without a USRP, a controller or a network.  It stands in for src/crts_cognitive_radio.cpp on the
// sensing path: read the node's parameters from the scenario .cfg, build the radio, apply the setters in
// the order Initialize_CR does (src/crts_cognitive_radio.cpp:404-430), set_ce(), start_rx(), start_ce()
// (:810-812), and let the rx worker feed packets to the engine until the capture ends.
//
//   crn_replay --scenario scenarios/predictive_model.cfg --node 2 --iq capture.c64 --packet-len 363
//              [--ce-args "-d 0 -o decisions.bin -q"] [--free-run]
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "extensible_cognitive_radio.hpp"
#include "scenario_cfg.hpp"

int main(int argc, char **argv) {
  std::string scenario, iq, ce_args_override;
  int node = 2, packet_len = 512;
  bool lockstep = true, have_override = false, loop = false;
  double patience_ms = -1.0;
  long max_packets = 0;
  static struct option opts[] = {{"scenario", 1, 0, 's'}, {"node", 1, 0, 'n'},       {"iq", 1, 0, 'i'},
                                 {"packet-len", 1, 0, 'l'}, {"ce-args", 1, 0, 'a'}, {"free-run", 0, 0, 'f'},
                                 {"lockstep-patience-ms", 1, 0, 'p'}, {"repeat-packets", 1, 0, 'r'},
                                 {0, 0, 0, 0}};
  int o;
  while ((o = getopt_long(argc, argv, "s:n:i:l:a:fp:r:", opts, NULL)) != -1) {
    switch (o) {
      case 's': scenario = optarg; break;
      case 'n': node = atoi(optarg); break;
      case 'i': iq = optarg; break;
      case 'l': packet_len = atoi(optarg); break;
      case 'a': ce_args_override = optarg; have_override = true; break;
      case 'f': lockstep = false; break;
      case 'p': patience_ms = atof(optarg); break;
      case 'r': max_packets = atol(optarg); loop = true; break;  // replay the capture in a loop for this many packets
      default:
        fprintf(stderr, "usage: %s --scenario file.cfg --node N --iq capture.c64 [--packet-len L] [--ce-args \"...\"] [--free-run] "
                        "[--lockstep-patience-ms MS] [--repeat-packets N]\n", argv[0]);
        return 2;
    }
  }
  if (scenario.empty() || iq.empty()) {
    fprintf(stderr, "crn_replay: --scenario and --iq are required\n");
    return 2;
  }
  CfgGroup root;
  std::string err;
  if (!cfg_parse_file(scenario, &root, &err)) {
    fprintf(stderr, "crn_replay: %s: %s\n", scenario.c_str(), err.c_str());
    return 1;
  }
  NodeParams np;
  if (!cfg_node_params(root, node, &np, &err)) {
    fprintf(stderr, "crn_replay: %s\n", err.c_str());
    return 1;
  }
  if (have_override) np.ce_args = ce_args_override;

  FileIqSource src(iq, loop, max_packets);
  if (!src.ok()) {
    fprintf(stderr, "crn_replay: cannot open IQ capture %s\n", iq.c_str());
    return 1;
  }

  ExtensibleCognitiveRadio *ECR = new ExtensibleCognitiveRadio();
  // Initialize_CR order (src/crts_cognitive_radio.cpp:404-430), restricted to what exists without a PHY
  ECR->set_ce_timeout_ms(np.ce_timeout_ms);
  ECR->set_rx_freq(np.rx_freq);
  ECR->set_rx_rate(np.rx_rate);
  ECR->set_rx_gain_uhd(np.rx_gain);
  ECR->set_tx_freq(np.tx_freq);
  ECR->set_tx_rate(np.tx_rate);
  ECR->set_tx_gain_soft(np.tx_gain_soft);
  ECR->set_tx_gain_uhd(np.tx_gain);
  int ce_argc = 0;
  char **ce_argv = NULL;
  cfg_str2argcargv(np.ce_args, "CE", &ce_argc, &ce_argv);
  ECR->set_ce((char *)np.cognitive_engine.c_str(), ce_argc, ce_argv);
  ECR->set_iq_source(&src, packet_len);
  ECR->set_lockstep(lockstep);
  if (patience_ms >= 0.0) ECR->set_lockstep_patience_ms(patience_ms);

  struct timeval t0, t1;
  gettimeofday(&t0, NULL);
  ECR->start_rx();
  ECR->start_ce();
  ECR->wait_for_end_of_capture();
  gettimeofday(&t1, NULL);
  ECR->stop_ce();
  const double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_usec - t0.tv_usec);
  printf("crn_replay: node%d engine=%s rx=%.0f Hz @ %.0f S/s: %lu packets received, %lu forwarded to the CE, "
         "%lu CE executions, final tx_freq=%.0f Hz\n",
         node, np.cognitive_engine.c_str(), ECR->get_rx_freq(), ECR->get_rx_rate(), ECR->packets_received(),
         ECR->packets_forwarded(), ECR->ce_executions(), ECR->get_tx_freq());
  printf("crn_replay: %lu packets received straight into engine slots (no copy), %.3f s, %.2f us per forwarded packet\n",
         ECR->packets_direct(), secs, ECR->packets_forwarded() ? 1e6 * secs / (double)ECR->packets_forwarded() : 0.0);
  delete ECR;
  return 0;
}
