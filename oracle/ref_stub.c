/*
 * oracle/ref_stub.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference's LOG() macro (include/audiosync/audiosync.h:88-94) reads the
 * global `global_debug`, whose real definition lives in src/audiosync.c:37,
 * a translation unit the oracle build does not include.  This provides it.
 */
volatile int global_debug = 0;

int oracle_ref_get_debug(void) { return global_debug; }
void oracle_ref_set_debug(int v) { global_debug = v; }
