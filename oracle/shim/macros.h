/* Stand-in for uc_tools macros.h (linux/test_cproc.c:2).  LOG is routed to a
 * capture hook so the harness can read what cproc_output() reported
 * (test_cproc.c:5-7 logs "output %d %d"). */
#ifndef MACROS_H
#define MACROS_H
#include <stddef.h>
void ref_log_capture(const char *fmt, ...);
#define LOG(...) ref_log_capture(__VA_ARGS__)
#endif
