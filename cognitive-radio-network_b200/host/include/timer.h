// tic/toc stopwatch with the interface the reference engines include as "timer.h"
// (include/timer.h:31-43 there, itself borrowed from liquid-dsp): an opaque handle, seconds as float.
#ifndef CRN_HOST_TIMER_H
#define CRN_HOST_TIMER_H

typedef struct timer_s *timer;

timer timer_create();
void timer_destroy(timer q);
void timer_tic(timer q);   // start / restart
float timer_toc(timer q);  // seconds since the last tic

#endif
