// Plugin plumbing of the host runtime: the CognitiveEngine base (its three members are declared out of
// line upstream, include/cognitive_engine.hpp:23-25,44, so engines link against them) and the registry
// that maps the scenario's `cognitive_engine = "CE_<Name>"` string to a factory.  Upstream generates a
// strcmp chain inside set_ce() for this (src/extensible_cognitive_radio.cpp:356-367); here each engine is
// announced by one CRN_REGISTER_CE line in the generated lib/ce_registry_generated.cpp.
#include <map>
#include <string>

#include "extensible_cognitive_radio.hpp"

namespace {
typedef std::map<std::string, crn_ce_factory> Registry;
Registry &registry() {
  static Registry r;  // function-local: safe to use from other translation units' static initialisers
  return r;
}
}  // namespace

bool crn_register_ce(const char *name, crn_ce_factory make) {
  registry()[name] = make;
  return true;
}

// nullptr when no engine of that name was registered
CognitiveEngine *crn_create_ce(const char *name, int argc, char **argv, ExtensibleCognitiveRadio *ecr) {
  Registry::const_iterator it = registry().find(name ? name : "");
  return it == registry().end() ? nullptr : it->second(argc, argv, ecr);
}

// An engine that overrides nothing does nothing per event.
void CognitiveEngine::execute() {}
CognitiveEngine::~CognitiveEngine() {}
CognitiveEngine::CognitiveEngine() : ECR(nullptr) {}
