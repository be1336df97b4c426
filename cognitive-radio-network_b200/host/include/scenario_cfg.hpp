// Minimal reader for the scenario files of the reference (libconfig syntax subset; libconfig itself is
// not available here).  It understands what scenarios/*.cfg use: `name = value;` settings with string,
// integer or floating-point values, `name : { ... };` groups, and //, # and /* */ comments.
// The keys consumed for a node are the ones src/crts.cpp:229-689 reads for the sensing path
// (SURVEY 8b "Scenario keys"); unknown keys are kept and ignored, as upstream does with
// `generate_octave_log_file`.
#ifndef CRN_HOST_SCENARIO_CFG_HPP
#define CRN_HOST_SCENARIO_CFG_HPP

#include <map>
#include <string>

struct CfgGroup {
  std::map<std::string, std::string> values;  // raw text of scalar settings (strings unquoted)
  std::map<std::string, CfgGroup> groups;
  bool has(const std::string &k) const { return values.count(k) != 0; }
  std::string str(const std::string &k, const std::string &dflt = "") const;
  double num(const std::string &k, double dflt = 0.0) const;
};

// Returns false and fills *err on a syntax error.
bool cfg_parse_file(const std::string &path, CfgGroup *root, std::string *err);
bool cfg_parse_text(const std::string &text, CfgGroup *root, std::string *err);

// The node parameters of the sensing path, with the defaults of src/crts.cpp (e.g. ce_timeout_ms).
struct NodeParams {
  std::string cognitive_engine, ce_args, node_type, cognitive_radio_type;
  double ce_timeout_ms, tx_freq, tx_rate, tx_gain, tx_gain_soft, rx_freq, rx_rate, rx_gain;
};
bool cfg_node_params(const CfgGroup &root, int node, NodeParams *np, std::string *err);

// `ce_args` string -> argc/argv with argv[0] = prog, as str2argcargv does (src/crts.cpp:43-81).
void cfg_str2argcargv(const std::string &args, const std::string &prog, int *argc, char ***argv);

#endif
