/* util.c -- status / warn / die conventions and number formatting of the host driver.
 * Reference behaviour mirrored: src/global/ctx_output.c:27-88 (die = message + exit(1),
 * status lines are time-stamped and carry a per-run 3-letter code, -q silences status),
 * src/global/util.c:206-222 (memory suffixes), :251-264 (thousands separators). */
#include "mcx_host.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <unistd.h>

FILE *mcx_msg_out = NULL;
static char run_code[4] = "";

static void run_code_init(void)
{
  static const char cons[] = "bcdfghjklmnpqrstvwxyzBCDFGHJKLMNPQRSTVWXYZ", vow[] = "aeiouAEIOU";
  unsigned s = (unsigned)time(NULL) * 2654435761u + (unsigned)getpid();
  run_code[0] = cons[s % (sizeof(cons) - 1)]; s /= 41;
  run_code[1] = vow[s % (sizeof(vow) - 1)]; s /= 11;
  run_code[2] = cons[s % (sizeof(cons) - 1)]; run_code[3] = 0;
}
/* status lines come from several threads (one per file being loaded): one line at a time */
static pthread_mutex_t msg_lock = PTHREAD_MUTEX_INITIALIZER;
static void timestamp(FILE *fh)
{
  static pthread_once_t once = PTHREAD_ONCE_INIT;
  pthread_once(&once, run_code_init);
  time_t t; char ts[100]; struct tm tmv;
  time(&t);
  strftime(ts, sizeof(ts), "[%d %b %Y %H:%M:%S", localtime_r(&t, &tmv));
  fprintf(fh, "%s-%s]", ts, run_code);
}

void mcx_status(const char *fmt, ...)
{
  if(!mcx_msg_out) return;
  va_list ap;
  pthread_mutex_lock(&msg_lock);
  timestamp(mcx_msg_out);
  if(fmt[0] != ' ' && fmt[0] != '[') fputc(' ', mcx_msg_out);
  va_start(ap, fmt); vfprintf(mcx_msg_out, fmt, ap); va_end(ap);
  if(fmt[strlen(fmt) - 1] != '\n') fputc('\n', mcx_msg_out);
  fflush(mcx_msg_out);
  pthread_mutex_unlock(&msg_lock);
}

static void err_msg(const char *type, const char *fmt, va_list ap)
{
  fflush(stdout);
  pthread_mutex_lock(&msg_lock);
  timestamp(stderr);
  fprintf(stderr, "[mccortex-b200] %s: ", type);
  vfprintf(stderr, fmt, ap);
  if(fmt[strlen(fmt) - 1] != '\n') fputc('\n', stderr);
  pthread_mutex_unlock(&msg_lock);
}

void mcx_warn(const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt); err_msg("Warn", fmt, ap); va_end(ap);
}

void mcx_die(const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt); err_msg("Fatal Error", fmt, ap); va_end(ap);
  exit(EXIT_FAILURE);
}

void mcx_print_usage(const char *usage, const char *errfmt, ...)
{
  if(errfmt) {
    va_list ap;
    fprintf(stderr, "Error: ");
    va_start(ap, errfmt); vfprintf(stderr, errfmt, ap); va_end(ap);
    if(errfmt[strlen(errfmt) - 1] != '\n') fputc('\n', stderr);
  }
  fputs(usage, stderr);
  exit(EXIT_FAILURE);
}

/* MCX_TIMING=1: wall clock of the driver's phases on stderr (not part of the reference's log) */
void mcx_phase(const char *what)
{
  static int on = -1; static struct timespec t0, last;
  struct timespec now;
  pthread_mutex_lock(&msg_lock);
  if(on < 0) { on = getenv("MCX_TIMING") != NULL; clock_gettime(CLOCK_MONOTONIC, &t0); last = t0; }
  if(!on) { pthread_mutex_unlock(&msg_lock); return; }
  clock_gettime(CLOCK_MONOTONIC, &now);
  fprintf(stderr, "[phase] %-28s +%.3f s  (at %.3f s)\n", what, (double)(now.tv_sec - last.tv_sec) + 1e-9 * (double)(now.tv_nsec - last.tv_nsec),
          (double)(now.tv_sec - t0.tv_sec) + 1e-9 * (double)(now.tv_nsec - t0.tv_nsec));
  last = now;
  pthread_mutex_unlock(&msg_lock);
}

void mcx_ulong_to_str(uint64_t num, char *out)
{
  char tmp[32]; int n = snprintf(tmp, sizeof(tmp), "%llu", (unsigned long long)num), i, j = 0;
  for(i = 0; i < n; i++) { out[j++] = tmp[i]; if((n - 1 - i) % 3 == 0 && i + 1 < n) out[j++] = ','; }
  out[j] = 0;
}

void mcx_bytes_to_str(uint64_t bytes, char *out)
{
  static const char *units[7] = {"B", "KB", "MB", "GB", "TB", "PB", "EB"};
  int u = 0; double v = (double)bytes;
  while(v >= 1024.0 && u < 6) { v /= 1024.0; u++; }
  if(u == 0) sprintf(out, "%lluB", (unsigned long long)bytes);
  else {
    sprintf(out, "%.1f", v);
    size_t n = strlen(out);
    if(n > 2 && out[n - 1] == '0' && out[n - 2] == '.') out[n - 2] = 0; /* the reference trims ".0" */
    strcat(out, units[u]);
  }
}

bool mcx_mem_to_integer(const char *arg, size_t *bytes)
{
  char *end;
  unsigned long num = strtoul(arg, &end, 10);
  if(end == arg) return false;
  if(!strcasecmp(end, "T") || !strcasecmp(end, "TB")) { *bytes = num << 40; return true; }
  if(!strcasecmp(end, "G") || !strcasecmp(end, "GB")) { *bytes = num << 30; return true; }
  if(!strcasecmp(end, "M") || !strcasecmp(end, "MB")) { *bytes = num << 20; return true; }
  if(!strcasecmp(end, "K") || !strcasecmp(end, "KB")) { *bytes = num << 10; return true; }
  if(*end != '\0') return false;
  *bytes = num;
  return true;
}
