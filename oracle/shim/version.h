/* oracle/shim/version.h -- stands in for the file the reference Makefile
 * generates at build time (Makefile:268-269: echo '#define CTX_VERSION ...'). */
#ifndef CTX_VERSION
#define CTX_VERSION "ref"
#endif
