/* oracle/shim/htslib/version.h -- stands in for the version.h htslib's own
 * Makefile generates (only used in the status banner, src/global/global.h:49-50). */
#define HTS_VERSION "stub"
