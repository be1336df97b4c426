#include "scenario_cfg.hpp"

#include <ctype.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <vector>

std::string CfgGroup::str(const std::string &k, const std::string &dflt) const {
  std::map<std::string, std::string>::const_iterator it = values.find(k);
  return it == values.end() ? dflt : it->second;
}
double CfgGroup::num(const std::string &k, double dflt) const {
  std::map<std::string, std::string>::const_iterator it = values.find(k);
  if (it == values.end()) return dflt;
  char *end = NULL;
  double v = strtod(it->second.c_str(), &end);
  return end == it->second.c_str() ? dflt : v;
}

namespace {
struct Parser {
  const std::string &s;
  size_t i;
  std::string err;
  explicit Parser(const std::string &t) : s(t), i(0) {}

  void skip() {
    for (;;) {
      while (i < s.size() && isspace((unsigned char)s[i])) i++;
      if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
        while (i < s.size() && s[i] != '\n') i++;
      } else if (i < s.size() && s[i] == '#') {
        while (i < s.size() && s[i] != '\n') i++;
      } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
        size_t e = s.find("*/", i + 2);
        i = (e == std::string::npos) ? s.size() : e + 2;
      } else {
        return;
      }
    }
  }
  bool fail(const std::string &m) {
    std::ostringstream o;
    size_t line = 1;
    for (size_t k = 0; k < i && k < s.size(); k++) line += s[k] == '\n';
    o << "line " << line << ": " << m;
    err = o.str();
    return false;
  }
  bool group(CfgGroup *g, bool top) {
    for (;;) {
      skip();
      if (i >= s.size()) return top ? true : fail("unterminated group");
      if (s[i] == '}') {
        if (top) return fail("unexpected '}'");
        i++;
        return true;
      }
      size_t b = i;
      while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == '*')) i++;
      if (i == b) return fail("expected a setting name");
      std::string name = s.substr(b, i - b);
      skip();
      if (i >= s.size() || (s[i] != '=' && s[i] != ':')) return fail("expected '=' or ':' after " + name);
      i++;
      skip();
      if (i < s.size() && s[i] == '{') {
        i++;
        if (!group(&g->groups[name], false)) return false;
      } else if (i < s.size() && s[i] == '"') {
        std::string v;
        i++;
        while (i < s.size() && s[i] != '"') {
          if (s[i] == '\\' && i + 1 < s.size()) i++;
          v += s[i++];
        }
        if (i >= s.size()) return fail("unterminated string");
        i++;
        g->values[name] = v;
      } else {
        size_t vb = i;
        while (i < s.size() && s[i] != ';' && s[i] != ',' && s[i] != '\n' && s[i] != '}') i++;
        std::string v = s.substr(vb, i - vb);
        while (!v.empty() && isspace((unsigned char)v[v.size() - 1])) v.erase(v.size() - 1);
        if (v.empty()) return fail("missing value for " + name);
        g->values[name] = v;
      }
      skip();
      if (i < s.size() && (s[i] == ';' || s[i] == ',')) i++;
    }
  }
};
}  // namespace

bool cfg_parse_text(const std::string &text, CfgGroup *root, std::string *err) {
  Parser p(text);
  bool ok = p.group(root, true);
  if (!ok && err) *err = p.err;
  return ok;
}

bool cfg_parse_file(const std::string &path, CfgGroup *root, std::string *err) {
  std::ifstream f(path.c_str());
  if (!f) {
    if (err) *err = "cannot open " + path;
    return false;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return cfg_parse_text(ss.str(), root, err);
}

bool cfg_node_params(const CfgGroup &root, int node, NodeParams *np, std::string *err) {
  char name[32];
  snprintf(name, sizeof(name), "node%d", node);
  std::map<std::string, CfgGroup>::const_iterator it = root.groups.find(name);
  if (it == root.groups.end()) {
    if (err) *err = std::string("scenario has no group ") + name;
    return false;
  }
  const CfgGroup &g = it->second;
  np->node_type = g.str("node_type", "cognitive radio");
  np->cognitive_radio_type = g.str("cognitive_radio_type", "ecr");
  np->cognitive_engine = g.str("cognitive_engine");
  if (np->cognitive_engine.empty()) {
    // upstream: "A cognitive engine must be specified" + exit (src/crts.cpp:306-310); here an error
    if (err) *err = std::string(name) + ": a cognitive engine (cognitive_engine) must be specified";
    return false;
  }
  np->ce_args = g.str("ce_args", "");
  np->ce_timeout_ms = g.num("ce_timeout_ms", 1000.0);
  np->tx_freq = g.num("tx_freq", 460e6);
  np->tx_rate = g.num("tx_rate", 500e3);
  np->tx_gain = g.num("tx_gain", 0.0);
  np->tx_gain_soft = g.num("tx_gain_soft", -12.0);
  np->rx_freq = g.num("rx_freq", 460e6);
  np->rx_rate = g.num("rx_rate", 500e3);
  np->rx_gain = g.num("rx_gain", 0.0);
  return true;
}

void cfg_str2argcargv(const std::string &args, const std::string &prog, int *argc, char ***argv) {
  std::vector<std::string> tok;
  tok.push_back(prog);
  std::istringstream is(args);
  std::string t;
  while (is >> t) tok.push_back(t);
  *argc = (int)tok.size();
  *argv = (char **)malloc(sizeof(char *) * (tok.size() + 1));
  for (size_t k = 0; k < tok.size(); k++) (*argv)[k] = strdup(tok[k].c_str());
  (*argv)[tok.size()] = NULL;
  optind = 0;  // as upstream (src/crts.cpp:80): engines call getopt() on this argv from scratch
}
