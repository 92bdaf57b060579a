// The scripted engine of ref_stubs/random needs a source of words; the reference's own unit
// tests (oracle/Makefile, target ref_gtests) never draw a random number, so theirs is empty.
// TEST INFRASTRUCTURE.
#include <cstdint>
extern "C" uint64_t qmcl_ref_next_word(void) { return 0; }
extern "C" double qmcl_ref_next_normal(void) { return 0; }
