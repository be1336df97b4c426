#include "cognitive_engine.hpp"

CognitiveEngine::CognitiveEngine() : ECR(nullptr) {}
CognitiveEngine::~CognitiveEngine() {}
void CognitiveEngine::execute() {}
