// Plugin boundary, kept as the reference defines it (include/cognitive_engine.hpp:21-44,
// src/cognitive_engine.cpp:4-6): a cognitive engine is a class with a public ECR pointer and a
// virtual, argument-less execute() that the radio's CE thread calls once per event.
// Engines written against the reference compile against this header unchanged.
#ifndef CRN_HOST_COGNITIVE_ENGINE_HPP
#define CRN_HOST_COGNITIVE_ENGINE_HPP

class ExtensibleCognitiveRadio;

class CognitiveEngine {
public:
  CognitiveEngine();
  ~CognitiveEngine();  // non-virtual upstream too: engines are created once and never deleted
  ExtensibleCognitiveRadio *ECR;
  virtual void execute();
};

#endif
