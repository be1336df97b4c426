// UHD-free ExtensibleCognitiveRadio: see include/extensible_cognitive_radio.hpp for scope and the
// reference lines each piece mirrors.
#include "extensible_cognitive_radio.hpp"

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>


void *ECR_rx_worker(void *arg);
void *ECR_ce_worker(void *arg);

// engine registry: src/plugin.cpp
CognitiveEngine *crn_create_ce(const char *name, int argc, char **argv, ExtensibleCognitiveRadio *ecr);

// ---- file source ------------------------------------------------------------------------------------
FileIqSource::FileIqSource(const std::string &path, bool loop, long max_packets)
    : fp_(fopen(path.c_str(), "rb")), loop_(loop), left_(max_packets) {}
FileIqSource::~FileIqSource() {
  if (fp_) fclose((FILE *)fp_);
}
size_t FileIqSource::recv(std::complex<float> *buf, size_t max_samps) {
  if (!fp_) return 0;
  if (left_ < 0) return 0;  // a looped replay delivered its max_packets
  size_t n = fread(buf, sizeof(std::complex<float>), max_samps, (FILE *)fp_);
  if (n < max_samps && loop_) {
    rewind((FILE *)fp_);
    n += fread(buf + n, sizeof(std::complex<float>), max_samps - n, (FILE *)fp_);
  }
  // a trailing partial packet is dropped: upstream always hands over full rx_buffer_len packets
  if (n != max_samps) return 0;
  if (left_ > 0 && --left_ == 0) left_ = -1;
  return n;
}

// ---- radio ------------------------------------------------------------------------------------------
ExtensibleCognitiveRadio::ExtensibleCognitiveRadio()
    : ce_usrp_rx_buffer(nullptr), ce_usrp_rx_buffer_length(0), CE(nullptr), ce_timeout_ms(1000.0),
      ce_sensing_flag(0), ce_thread_running(true), ce_running(false), rx_thread_running(true),
      rx_running(false), capture_done(false), lockstep_(false), handoff_pending_(false), ce_ever_started_(false), ever_sensed_(false),
      patience_spent_(false), lockstep_patience_ms_(1000.0), slot_fn_(nullptr), slot_ctx_(nullptr), ce_buffer_own_(nullptr),
      direct_(0), src_(nullptr),
      rx_buffer(nullptr), rx_buffer_len(0), tx_freq_(460e6), tx_rate_(1e6), tx_gain_soft_(-12.0),
      tx_gain_uhd_(0.0), rx_freq_(460e6), rx_rate_(1e6), rx_gain_uhd_(0.0), tx_on_(false), packets_(0),
      forwarded_(0), tx_retunes_(0), executions_(0) {
  CE_metrics.CE_event = TIMEOUT;
  pthread_mutex_init(&CE_mutex, NULL);
  pthread_mutex_init(&rx_params_mutex, NULL);
  pthread_mutex_init(&tx_params_mutex, NULL);
  pthread_cond_init(&CE_cond, NULL);
  pthread_cond_init(&CE_execute_sig, NULL);
  pthread_cond_init(&rx_cond, NULL);
  pthread_cond_init(&consumed_sig, NULL);
  pthread_cond_init(&done_sig, NULL);
  pthread_create(&CE_process, NULL, ECR_ce_worker, (void *)this);
  pthread_create(&rx_process, NULL, ECR_rx_worker, (void *)this);
}

ExtensibleCognitiveRadio::~ExtensibleCognitiveRadio() {
  pthread_mutex_lock(&CE_mutex);
  ce_running = false;
  ce_thread_running = false;
  pthread_cond_broadcast(&CE_cond);
  pthread_cond_broadcast(&CE_execute_sig);
  pthread_cond_broadcast(&consumed_sig);
  pthread_mutex_unlock(&CE_mutex);
  pthread_mutex_lock(&rx_params_mutex);
  rx_running = false;
  rx_thread_running = false;
  pthread_cond_broadcast(&rx_cond);
  pthread_mutex_unlock(&rx_params_mutex);
  pthread_join(rx_process, NULL);
  pthread_join(CE_process, NULL);
  free(rx_buffer);
  free(ce_buffer_own_);
}

void ExtensibleCognitiveRadio::set_ce(char *ce, int argc, char **argv) {
  CE = crn_create_ce(ce, argc, argv, this);
  if (!CE) {
    // same convention as upstream (src/crts.cpp:306-310): an unknown engine is fatal
    printf("The cognitive engine %s is not registered with this radio.\n", ce);
    exit(EXIT_FAILURE);
  }
}

void ExtensibleCognitiveRadio::start_ce() {
  pthread_mutex_lock(&CE_mutex);
  ce_running = true;
  ce_ever_started_ = true;
  pthread_cond_signal(&CE_cond);
  pthread_mutex_unlock(&CE_mutex);
}
void ExtensibleCognitiveRadio::stop_ce() {
  pthread_mutex_lock(&CE_mutex);
  ce_running = false;
  pthread_mutex_unlock(&CE_mutex);
}
void ExtensibleCognitiveRadio::set_ce_timeout_ms(double t) { ce_timeout_ms = t; }
double ExtensibleCognitiveRadio::get_ce_timeout_ms() { return ce_timeout_ms; }
void ExtensibleCognitiveRadio::set_ce_sensing(int on) {  // no lock, as upstream (cpp:389-391)
  if (on) ever_sensed_ = true;
  ce_sensing_flag = on;
}
// called from the engine (CE thread, CE_mutex held) or before start_ce()
void ExtensibleCognitiveRadio::set_rx_slot_provider(rx_slot_provider fn, void *ctx) {
  slot_fn_ = fn;
  slot_ctx_ = ctx;
}
void ExtensibleCognitiveRadio::set_lockstep_patience_ms(double ms) { lockstep_patience_ms_ = ms; }

#define LOCKED(m, stmt)        \
  do {                         \
    pthread_mutex_lock(&m);    \
    stmt;                      \
    pthread_mutex_unlock(&m);  \
  } while (0)

void ExtensibleCognitiveRadio::set_tx_freq(double f) { LOCKED(tx_params_mutex, tx_freq_ = f; tx_retunes_++); }
void ExtensibleCognitiveRadio::set_tx_rate(double r) { LOCKED(tx_params_mutex, tx_rate_ = r); }
void ExtensibleCognitiveRadio::set_tx_gain_soft(double g) { LOCKED(tx_params_mutex, tx_gain_soft_ = g); }
void ExtensibleCognitiveRadio::set_tx_gain_uhd(double g) { LOCKED(tx_params_mutex, tx_gain_uhd_ = g); }
void ExtensibleCognitiveRadio::set_rx_freq(double f) { LOCKED(rx_params_mutex, rx_freq_ = f); }
void ExtensibleCognitiveRadio::set_rx_rate(double r) { LOCKED(rx_params_mutex, rx_rate_ = r); }
void ExtensibleCognitiveRadio::set_rx_gain_uhd(double g) { LOCKED(rx_params_mutex, rx_gain_uhd_ = g); }
double ExtensibleCognitiveRadio::get_tx_freq() { double v; LOCKED(tx_params_mutex, v = tx_freq_); return v; }
double ExtensibleCognitiveRadio::get_tx_rate() { double v; LOCKED(tx_params_mutex, v = tx_rate_); return v; }
double ExtensibleCognitiveRadio::get_rx_freq() { double v; LOCKED(rx_params_mutex, v = rx_freq_); return v; }
double ExtensibleCognitiveRadio::get_rx_rate() { double v; LOCKED(rx_params_mutex, v = rx_rate_); return v; }
void ExtensibleCognitiveRadio::start_tx() { LOCKED(tx_params_mutex, tx_on_ = true); }
void ExtensibleCognitiveRadio::stop_tx() { LOCKED(tx_params_mutex, tx_on_ = false); }

int ExtensibleCognitiveRadio::get_tx_state() { int v; LOCKED(tx_params_mutex, v = tx_on_ ? TX_CONTINUOUS : TX_STOPPED); return v; }
int ExtensibleCognitiveRadio::get_rx_state() { int v; LOCKED(rx_params_mutex, v = rx_running ? RX_CONTINUOUS : RX_STOPPED); return v; }

void ExtensibleCognitiveRadio::start_rx() {
  pthread_mutex_lock(&rx_params_mutex);
  rx_running = true;
  pthread_cond_signal(&rx_cond);
  pthread_mutex_unlock(&rx_params_mutex);
}
void ExtensibleCognitiveRadio::stop_rx() { LOCKED(rx_params_mutex, rx_running = false); }

void ExtensibleCognitiveRadio::set_iq_source(IqSource *src, int packet_len) {
  pthread_mutex_lock(&rx_params_mutex);
  src_ = src;
  // upstream sizes both buffers from get_max_recv_samps_per_packet() (cpp:1263-1269)
  rx_buffer_len = (size_t)packet_len;
  ce_usrp_rx_buffer_length = packet_len;
  free(rx_buffer);
  free(ce_buffer_own_);
  rx_buffer = (std::complex<float> *)malloc(rx_buffer_len * sizeof(std::complex<float>));
  ce_buffer_own_ = (std::complex<float> *)malloc(rx_buffer_len * sizeof(std::complex<float>));
  ce_usrp_rx_buffer = ce_buffer_own_;
  capture_done = false;
  pthread_mutex_unlock(&rx_params_mutex);
}
void ExtensibleCognitiveRadio::set_lockstep(bool on) { lockstep_ = on; }

void ExtensibleCognitiveRadio::wait_for_end_of_capture() {
  pthread_mutex_lock(&CE_mutex);
  while (!capture_done) pthread_cond_wait(&done_sig, &CE_mutex);
  pthread_mutex_unlock(&CE_mutex);
}

// Lock-step gate (CE_mutex held): wait until the engine took the previous packet and (re-)armed sensing.  start_rx()
// precedes start_ce() (src/crts_cognitive_radio.cpp:810-812), so "CE not started YET" must wait as well - or a short
// capture is drained before the engine ever runs; a CE that was stopped again lets packets drop.  An engine that has
// never armed sensing is waited for at most lockstep_patience_ms_, once; after that only the consumption of the
// previous event is awaited.
void ECR_lockstep_gate(ExtensibleCognitiveRadio *ECR) {
  struct timespec until;
  clock_gettime(CLOCK_REALTIME, &until);
  const double ns = (double)until.tv_nsec + ECR->lockstep_patience_ms_ * 1e6;
  until.tv_sec += (time_t)(ns / 1e9);
  until.tv_nsec = (long)fmod(ns, 1e9);
  while (ECR->ce_thread_running && (ECR->ce_running || !ECR->ce_ever_started_)) {
    const bool wait_for_arming = !ECR->ce_sensing_flag && (ECR->ever_sensed_ || !ECR->patience_spent_);
    if (!ECR->handoff_pending_ && !wait_for_arming) break;
    if (ECR->handoff_pending_ || ECR->ever_sensed_) {
      pthread_cond_wait(&ECR->consumed_sig, &ECR->CE_mutex);
    } else if (pthread_cond_timedwait(&ECR->consumed_sig, &ECR->CE_mutex, &until) == ETIMEDOUT) {
      ECR->patience_spent_ = true;  // this engine does not sense: stop holding packets for it
    }
  }
}

// receiver worker: recv one packet, hand it to the CE when sensing is on (cpp:1258-1324)
void *ECR_rx_worker(void *arg) {
  ExtensibleCognitiveRadio *ECR = (ExtensibleCognitiveRadio *)arg;
  while (ECR->rx_thread_running) {
    pthread_mutex_lock(&ECR->rx_params_mutex);
    while (ECR->rx_thread_running && !(ECR->rx_running && ECR->src_)) pthread_cond_wait(&ECR->rx_cond, &ECR->rx_params_mutex);
    pthread_mutex_unlock(&ECR->rx_params_mutex);
    if (!ECR->rx_thread_running) break;

    // where this packet lands: an engine slot (direct) when the engine offers one and is ready for it, else rx_buffer
    std::complex<float> *dst = ECR->rx_buffer;
    bool direct = false, gated = false;
    if (ECR->slot_fn_ && (ECR->ce_sensing_flag || ECR->lockstep_)) {
      pthread_mutex_lock(&ECR->CE_mutex);
      if (ECR->lockstep_) {
        ECR_lockstep_gate(ECR);
        gated = true;
      }
      if (ECR->slot_fn_ && ECR->ce_sensing_flag && !ECR->handoff_pending_) {
        int overflow = 0;
        std::complex<float> *p = ECR->slot_fn_(ECR->slot_ctx_, ECR->rx_buffer_len, &overflow);
        if (p) {
          dst = p;
          direct = true;
        } else if (overflow) {  // the consumer's ring is full: tell the engine as upstream reports a USRP overflow
          ECR->CE_metrics.CE_event = ExtensibleCognitiveRadio::UHD_OVERFLOW;
          ECR->handoff_pending_ = true;
          pthread_cond_signal(&ECR->CE_execute_sig);
        }
      }
      pthread_mutex_unlock(&ECR->CE_mutex);
    }

    size_t n = ECR->src_->recv(dst, ECR->rx_buffer_len);
    if (n == 0) {  // end of capture: let the engine finish, then report
      pthread_mutex_lock(&ECR->CE_mutex);
      while (ECR->lockstep_ && ECR->handoff_pending_ && ECR->ce_thread_running && ECR->ce_running)
        pthread_cond_wait(&ECR->consumed_sig, &ECR->CE_mutex);
      ECR->capture_done = true;
      pthread_cond_broadcast(&ECR->done_sig);
      pthread_mutex_unlock(&ECR->CE_mutex);
      pthread_mutex_lock(&ECR->rx_params_mutex);
      ECR->rx_running = false;
      pthread_mutex_unlock(&ECR->rx_params_mutex);
      continue;
    }
    ECR->packets_++;

    if (ECR->ce_sensing_flag || ECR->lockstep_) {
      pthread_mutex_lock(&ECR->CE_mutex);
      if (ECR->lockstep_ && !gated) ECR_lockstep_gate(ECR);
      if (ECR->ce_sensing_flag) {  // re-check under the mutex, as upstream does (cpp:1313)
        if (direct) {
          ECR->ce_usrp_rx_buffer = dst;  // already in the engine's slot
          ECR->direct_++;
        } else {
          ECR->ce_usrp_rx_buffer = ECR->ce_buffer_own_;
          memcpy(ECR->ce_buffer_own_, ECR->rx_buffer, ECR->rx_buffer_len * sizeof(std::complex<float>));
        }
        ECR->CE_metrics.CE_event = ExtensibleCognitiveRadio::USRP_RX_SAMPS;
        ECR->handoff_pending_ = true;
        ECR->forwarded_++;
        pthread_cond_signal(&ECR->CE_execute_sig);
      }
      pthread_mutex_unlock(&ECR->CE_mutex);
    }
  }
  return NULL;
}

// CE worker (cpp:1761-1808): wait for an event or the timeout, run execute() with CE_mutex held
void *ECR_ce_worker(void *arg) {
  ExtensibleCognitiveRadio *ECR = (ExtensibleCognitiveRadio *)arg;
  while (ECR->ce_thread_running) {
    pthread_mutex_lock(&ECR->CE_mutex);
    while (ECR->ce_thread_running && !ECR->ce_running) pthread_cond_wait(&ECR->CE_cond, &ECR->CE_mutex);
    pthread_mutex_unlock(&ECR->CE_mutex);

    while (ECR->ce_running && ECR->ce_thread_running) {
      struct timeval now;
      gettimeofday(&now, NULL);
      double t_ns = (double)now.tv_usec * 1e3 + (double)now.tv_sec * 1e9 + ECR->ce_timeout_ms * 1e6;
      double s_part;
      double ns_part = modf(t_ns / 1e9, &s_part);
      struct timespec until;
      until.tv_sec = (long)s_part;
      until.tv_nsec = (long)(ns_part * 1e9);

      pthread_mutex_lock(&ECR->CE_mutex);
      if (!ECR->handoff_pending_) {
        if (ETIMEDOUT == pthread_cond_timedwait(&ECR->CE_execute_sig, &ECR->CE_mutex, &until) &&
            !ECR->handoff_pending_)
          ECR->CE_metrics.CE_event = ExtensibleCognitiveRadio::TIMEOUT;
      }
      if (ECR->CE && ECR->ce_running) {
        ECR->CE->execute();
        ECR->executions_++;
      }
      ECR->handoff_pending_ = false;
      ECR->CE_metrics.CE_event = ExtensibleCognitiveRadio::TIMEOUT;  // an event is delivered once
      pthread_cond_broadcast(&ECR->consumed_sig);
      pthread_mutex_unlock(&ECR->CE_mutex);
    }
  }
  return NULL;
}
