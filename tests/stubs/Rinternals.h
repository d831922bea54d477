#ifndef BNMF_STUB_RINTERNALS_H
#define BNMF_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
enum { REALSXP = 14, VECSXP = 19, RAWSXP = 24 };
extern SEXP R_NilValue;
void Rf_error(const char*, ...);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), Rboolean);
int Rf_asInteger(SEXP); int Rf_asLogical(SEXP); double Rf_asReal(SEXP);
int Rf_nrows(SEXP); int Rf_ncols(SEXP); int Rf_isMatrix(SEXP); int Rf_isNull(SEXP);
double* REAL(SEXP); Rbyte* RAW(SEXP); R_xlen_t XLENGTH(SEXP);
SEXP Rf_coerceVector(SEXP, int); SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_allocMatrix(int, int, int);
SEXP Rf_alloc3DArray(int, int, int, int); SEXP Rf_ScalarInteger(int);
SEXP STRING_ELT(SEXP, R_xlen_t); const char* CHAR(SEXP); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#endif
