#include "timer.h"

#include <stdlib.h>
#include <time.h>

struct timer_s {
  struct timespec t0;
};

timer timer_create() {
  timer q = (timer)malloc(sizeof(struct timer_s));
  clock_gettime(CLOCK_MONOTONIC, &q->t0);
  return q;
}
void timer_destroy(timer q) { free(q); }
void timer_tic(timer q) { clock_gettime(CLOCK_MONOTONIC, &q->t0); }
float timer_toc(timer q) {
  struct timespec t1;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (float)((double)(t1.tv_sec - q->t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - q->t0.tv_nsec));
}
