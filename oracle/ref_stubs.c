/* ORACLE -- TEST INFRASTRUCTURE ONLY.  The vendored oomph-lib references METIS (mesh partitioning) and SuperLU (direct solver) from
 * third-party libraries that are not part of /root/reference; nothing the pins call reaches them.  Defined here so that
 * oracle/_ref/liboomph_ref.so links; reaching one is a bug and aborts. */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void) { fprintf(stderr, "oracle/_ref: unexpected call of " #name "\n"); abort(); }
STUB(METIS_PartGraphKway)
STUB(METIS_PartGraphVKway)
STUB(superlu)
STUB(superlu_complex)
