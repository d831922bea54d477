/* Minimal stand-in for R's C API: just enough declarations for `gcc -fsyntax-only r/rcall.c`
 * (tests/test_abi.py).  R itself is not installed in this image; nothing links against this. */
#ifndef BNMF_STUB_R_H
#define BNMF_STUB_R_H
#include <stddef.h>
#include <string.h>
#endif
