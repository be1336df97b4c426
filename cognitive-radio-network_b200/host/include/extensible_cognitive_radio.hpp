// UHD-free ExtensibleCognitiveRadio for the sensing path.
//
// The reference class (include/extensible_cognitive_radio.hpp, src/extensible_cognitive_radio.cpp) owns
// a USRP, an OFDM modem, a TUN device and three pthreads.  This one keeps, with the same names and
// meaning, exactly the part cognitive engines on the sensing path touch:
//   - the CE event enum and CE_metrics.CE_event                      (hpp:65-91,538)
//   - ce_usrp_rx_buffer / ce_usrp_rx_buffer_length / set_ce_sensing   (hpp:543-550, cpp:389-391)
//   - the frequency/rate/gain setters+getters and start/stop calls the shipped engines use
//   - set_ce(name, argc, argv), start_ce/stop_ce, set_ce_timeout_ms   (cpp:354-369,371-387)
//   - the rx-worker -> CE-worker handoff: copy the packet under CE_mutex, raise USRP_RX_SAMPS, signal;
//     the CE thread waits with a timeout, marks TIMEOUT if it expires, and runs CE->execute() with
//     CE_mutex held                                                    (cpp:1310-1324,1761-1808)
// and replaces the USRP by an IqSource (file replay or any callback), one packet of
// rx_buffer_len samples per recv().  PHY/TX/TUN/logging are out of scope (SURVEY 2 rows 9,16,17).
//
// Deterministic frame selection: upstream, which packets reach the engine is a race (a timed-out wait
// overwrites CE_event).  With set_lockstep(true) the rx worker hands a packet over only after
// the engine has consumed the previous one, so a replayed capture is sensed frame by frame.
#ifndef CRN_HOST_ECR_HPP
#define CRN_HOST_ECR_HPP

// Upstream's header drags these in (through crts.hpp, liquid and UHD); engines written against it rely on
// that, e.g. CE_Template.cpp uses getopt()/atoi() without including anything itself.
#include <getopt.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <complex>
#include <fstream>
#include <iostream>
#include <cstddef>
#include <string>

#include "cognitive_engine.hpp"
#include "timer.h"

// transmitter / receiver states, as upstream (include/extensible_cognitive_radio.hpp:31-44)
enum tx_states { TX_STOPPED = 0, TX_CONTINUOUS, TX_BURST };
enum rx_states { RX_STOPPED = 0, RX_CONTINUOUS };

// A stand-in for uhd::device::recv(..., RECV_MODE_ONE_PACKET) (cpp:1304-1306).
class IqSource {
public:
  virtual ~IqSource() {}
  // Fill up to max_samps samples; return how many were produced, 0 at end of capture.
  virtual size_t recv(std::complex<float> *buf, size_t max_samps) = 0;
};

// Flat little-endian complex64 file (what the synthetic generator and any SDR recorder write).
class FileIqSource : public IqSource {
public:
  // loop: start over at the end of the file; max_packets > 0 ends a looped replay after that many packets
  explicit FileIqSource(const std::string &path, bool loop = false, long max_packets = 0);
  ~FileIqSource();
  bool ok() const { return fp_ != nullptr; }
  size_t recv(std::complex<float> *buf, size_t max_samps) override;

private:
  void *fp_;
  bool loop_;
  long left_;
};

class ExtensibleCognitiveRadio {
public:
  ExtensibleCognitiveRadio();
  ~ExtensibleCognitiveRadio();

  enum CE_Event {
    TIMEOUT = 0,         // no event for ce_timeout_ms
    PHY_FRAME_RECEIVED,  // (never raised here: no PHY)
    TX_COMPLETE,
    UHD_OVERFLOW,        // raised when a streaming consumer reports a ring overrun
    UHD_UNDERRUN,
    USRP_RX_SAMPS        // one packet of samples is in ce_usrp_rx_buffer
  };
  struct metric_s {
    int CE_event;
  };

  // --- cognitive engine plumbing (same names as upstream) ---
  void set_ce(char *ce, int argc, char **argv);
  void start_ce();
  void stop_ce();
  void set_ce_timeout_ms(double new_timeout_ms);
  double get_ce_timeout_ms();
  struct metric_s CE_metrics;
  void set_ce_sensing(int ce_sensing);
  std::complex<float> *ce_usrp_rx_buffer;
  int ce_usrp_rx_buffer_length;

  // --- direct-to-slot receive (the "RX buffer handoff" change to cpp:1310-1324) ---
  // An engine that stages packets in memory of its own (libcrnsense's pinned ring) can ask the receiver to recv()
  // STRAIGHT into it: while sensing is armed and no handoff is pending, the rx worker calls the provider under
  // CE_mutex (so it never runs concurrently with execute()) for the address the next packet of `nsamples` samples
  // should land in, receives there, and hands it over by pointing ce_usrp_rx_buffer at it - no memcpy under the
  // mutex, and none in the engine.  The provider returns NULL to decline (the packet then takes the usual
  // rx_buffer -> ce_usrp_rx_buffer copy); if it also sets *overflow the receiver raises UHD_OVERFLOW on the CE
  // (cpp:1327-1336: what upstream does when the USRP reports an overflow).  Engines that never register one -
  // every unmodified reference engine - see ce_usrp_rx_buffer exactly as before.
#define CRN_ECR_HAS_RX_SLOT_PROVIDER 1
  typedef std::complex<float> *(*rx_slot_provider)(void *ctx, size_t nsamples, int *overflow);
  void set_rx_slot_provider(rx_slot_provider fn, void *ctx);

  // --- radio parameters: stored, reported back, counted; no hardware behind them ---
  void set_tx_freq(double f);
  void set_tx_rate(double r);
  void set_tx_gain_soft(double g);
  void set_tx_gain_uhd(double g);
  void set_rx_freq(double f);
  void set_rx_rate(double r);
  void set_rx_gain_uhd(double g);
  double get_tx_freq();
  double get_tx_rate();
  double get_rx_freq();
  double get_rx_rate();
  void start_tx();
  void stop_tx();
  int get_tx_state();
  int get_rx_state();
  void start_rx();
  void stop_rx();

  // --- replay specific ---
  void set_iq_source(IqSource *src, int packet_len);  // packet_len = get_max_recv_samps_per_packet()
  void set_lockstep(bool on);
  // Lock-step holds a packet until the engine has consumed the previous one AND armed sensing.  An engine that never
  // senses (CE_Template, the PU engines) would hold the receiver forever, so until sensing has been armed at least
  // once the wait is bounded by this patience (default 1000 ms of wall clock - ten of the reference engine's 100 ms
  // re-arm periods); after it expires packets are dropped while sensing is off, as upstream does (cpp:1310).
  void set_lockstep_patience_ms(double ms);
  unsigned long packets_direct() const { return direct_; }   // packets received straight into an engine slot
  void wait_for_end_of_capture();  // returns when the source is exhausted and the CE drained it
  unsigned long packets_received() const { return packets_; }
  unsigned long packets_forwarded() const { return forwarded_; }
  unsigned long tx_retunes() const { return tx_retunes_; }
  unsigned long ce_executions() const { return executions_; }

private:
  CognitiveEngine *CE;
  double ce_timeout_ms;
  // flags the workers read outside their mutexes (upstream uses plain ints / bools for these, e.g. the unlocked
  // read of ce_sensing_flag at src/extensible_cognitive_radio.cpp:1310): atomics here, so the replay runtime is
  // free of data races by construction
  std::atomic<int> ce_sensing_flag;
  pthread_t CE_process, rx_process;
  pthread_mutex_t CE_mutex, rx_params_mutex, tx_params_mutex;
  pthread_cond_t CE_cond, CE_execute_sig, rx_cond, consumed_sig, done_sig;
  std::atomic<bool> ce_thread_running, ce_running, rx_thread_running, rx_running;
  bool capture_done;
  bool lockstep_, handoff_pending_;
  bool ce_ever_started_;  // lock-step replay: the rx worker holds the first packet until start_ce() has been called
  std::atomic<bool> ever_sensed_;   // set_ce_sensing(1) has been called at least once
  bool patience_spent_;             // lock-step: the bounded wait for a never-sensing engine has expired once
  double lockstep_patience_ms_;
  rx_slot_provider slot_fn_;
  void *slot_ctx_;
  std::complex<float> *ce_buffer_own_;  // the buffer ce_usrp_rx_buffer points at when a packet is copied (cpp:1269)
  unsigned long direct_;
  IqSource *src_;
  std::complex<float> *rx_buffer;
  size_t rx_buffer_len;
  double tx_freq_, tx_rate_, tx_gain_soft_, tx_gain_uhd_, rx_freq_, rx_rate_, rx_gain_uhd_;
  bool tx_on_;
  unsigned long packets_, forwarded_, tx_retunes_, executions_;
  friend void *ECR_rx_worker(void *);
  friend void ECR_lockstep_gate(ExtensibleCognitiveRadio *);
  friend void *ECR_ce_worker(void *);
};

// Engine registry.  Upstream regenerates a strcmp chain in set_ce() from the directory listing
// (src/config_cognitive_engines.cpp:202-249 rewriting extensible_cognitive_radio.cpp:356-367); here an
// engine translation unit registers its own factory, and set_ce() looks the name up.
typedef CognitiveEngine *(*crn_ce_factory)(int argc, char **argv, ExtensibleCognitiveRadio *ecr);
bool crn_register_ce(const char *name, crn_ce_factory f);
#define CRN_REGISTER_CE(NAME)                                                                     \
  static CognitiveEngine *crn_make_##NAME(int argc, char **argv, ExtensibleCognitiveRadio *ecr) { \
    return new NAME(argc, argv, ecr);                                                             \
  }                                                                                               \
  static bool crn_registered_##NAME = crn_register_ce(#NAME, crn_make_##NAME);

#endif
