/*
 * oracle/ref_main.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Dispatch shim (written for this repo) that links the UNMODIFIED reference
 * sources under /root/reference into a small binary exposing only the
 * sub-commands the parity tests need:  build, sort, join, view, check, hashtest.
 * It replaces src/main/mccortex.c:279-332 (whose command table pulls in all
 * 30 sub-commands and, through them, htslib/VCF and seq-align), and follows
 * the same start-up sequence: cortex_init -> cmd_init -> cmd_set_usage ->
 * strip -q flags -> cmd_print_status_header -> cmd->func(argc-1, argv+1).
 */
#include "global.h"
#include "commands.h"
#include "cmd.h"
#include "util.h"
#include <strings.h>

typedef struct { const char *cmd; int (*func)(int, char **); const char *usage; } RefCmd;

static const RefCmd cmds[] = {
  {"build",    ctx_build,        build_usage},
  {"sort",     ctx_sort,         sort_usage},
  {"join",     ctx_join,         join_usage},
  {"view",     ctx_view,         view_usage},
  {"check",    ctx_health_check, health_usage},
  {"hashtest", ctx_exp_hashtest, exp_hashtest_usage},
};

static int strip_quiet(int *argcp, char **argv)
{
  int found = 0, i, j, argc = *argcp;
  for(i = j = 1; i < argc; i++) {
    if(!strcmp(argv[i], "--quiet") || !strcmp(argv[i], "-q")) { found = 1; continue; }
    if(argv[i][0] == '-' && argv[i][1] != '-') {
      char *p, *q;
      for(p = q = argv[i] + 1; *p; p++) { if(*p == 'q') found = 1; else *q++ = *p; }
      *q = '\0';
    }
    argv[j++] = argv[i];
  }
  *argcp = j;
  return found;
}

int main(int argc, char **argv)
{
  size_t i, n = sizeof(cmds) / sizeof(cmds[0]);
  const RefCmd *cmd = NULL;
  ctx_msg_out = stderr;
  cortex_init();
  cmd_init(argc, argv);
  if(argc < 2) { fprintf(stderr, "usage: %s <build|sort|join|view|check|hashtest> ...\n", argv[0]); return 1; }
  for(i = 0; i < n; i++) if(!strcasecmp(cmds[i].cmd, argv[1])) cmd = &cmds[i];
  if(cmd == NULL) { fprintf(stderr, "unknown command: %s\n", argv[1]); return 1; }
  cmd_set_usage(cmd->usage);
  if(argc == 2) cmd_print_usage(NULL);
  if(strip_quiet(&argc, argv)) ctx_msg_out = NULL;
  cmd_print_status_header();
  char *tmp = argv[1]; argv[1] = argv[0]; argv[0] = tmp;
  int ret = cmd->func(argc - 1, argv + 1);
  cmd_destroy();
  status(ret == 0 ? "Done." : "Fail.");
  cortex_destroy();
  return ret;
}
