// libcrnsense internals shared between the host-only and CUDA translation units.
#pragma once
#include <cstdarg>
#include <cstdint>

#include "crnsense.h"

namespace crn {
// Record a message for crn_last_error() and return `status` (never exits, unlike the reference's
// printf + exit(EXIT_FAILURE) convention, e.g. src/crts.cpp:306-310).
int fail(int status, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
const char *last_error();
}  // namespace crn
