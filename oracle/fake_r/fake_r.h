/*
 * fake_r.h -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
 *
 * A tiny stand-in for the slice of the R C API that the reference's native layer touches,
 * so that the UNMODIFIED reference sources (/root/reference/src/randomForestSRC.c and
 * splitCustom.c) can be compiled into oracle/_ref/librfslam_ref.so without an R toolchain.
 * SURVEY.md section 8(c) lists the 27 API names; the reference maps them in
 * src/randomForestSRC.h:2-8 (RF_nativeNaN = NA_REAL, RF_nativeIsNaN = ISNA, ...).
 *
 * Nothing here is copied from R or from the reference: it is a from-scratch box type with the
 * same spelling as the R macros.
 */
#ifndef RFSLAM_FAKE_R_H
#define RFSLAM_FAKE_R_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NILSXP = 0, CHARSXP = 9, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };

typedef struct fr_box {
  int            type;
  long           len;
  void          *data;     /* int[] / double[] / SEXP[] / char[] */
  struct fr_box *names;    /* names attribute (STRSXP) or NULL */
} fr_box;

typedef fr_box *SEXP;

extern SEXP   R_NilValue;
extern SEXP   R_NamesSymbol;
extern double R_NaReal;                 /* NaN with low word 1954, as in R */

SEXP   allocVector(int type, long n);
SEXP   mkChar(const char *s);
SEXP   setAttrib(SEXP x, SEXP which, SEXP val);
int    R_IsNA(double x);
void   Rprintf(const char *fmt, ...);
void   error(const char *fmt, ...);     /* longjmps back to fr_guard() */

#define NA_REAL            R_NaReal
#define ISNA(x)            R_IsNA(x)
#define INTEGER(x)         ((int *)    ((x)->data))
#define REAL(x)            ((double *) ((x)->data))
#define VECTOR_ELT(x, i)   (((SEXP *) ((x)->data))[i])
#define STRING_ELT(x, i)   (((SEXP *) ((x)->data))[i])
#define SET_VECTOR_ELT(x, i, v) (((SEXP *) ((x)->data))[i] = (v))
#define SET_STRING_ELT(x, i, v) (((SEXP *) ((x)->data))[i] = (v))
#define CHAR(x)            ((const char *) ((x)->data))
#define AS_CHARACTER(x)    (x)
#define NEW_NUMERIC(n)     allocVector(REALSXP, (long) (n))
#define NEW_INTEGER(n)     allocVector(INTSXP,  (long) (n))
#define NEW_CHARACTER(n)   allocVector(STRSXP,  (long) (n))
#define NUMERIC_POINTER(x)   REAL(x)
#define INTEGER_POINTER(x)   INTEGER(x)
#define CHARACTER_POINTER(x) ((SEXP *) ((x)->data))
#define length(x)          ((int) ((x)->len))
#define LENGTH(x)          ((int) ((x)->len))
#define PROTECT(x)         (x)
#define UNPROTECT(n)       ((void) (n))

/* Rdynload.h slice, only so that R_init_randomForestSRC.c would also compile if wanted. */
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct fr_dllinfo DllInfo;

/* ---- driver-side helpers (used by oracle/refdriver.py through ctypes) ---- */
SEXP  fr_int(const int *v, long n);
SEXP  fr_real(const double *v, long n);
SEXP  fr_real_nocopy(double *v, long n);  /* aliases caller memory (big x matrices) */
SEXP  fr_int_nocopy(int *v, long n);
SEXP  fr_str(const char *const *v, long n);
SEXP  fr_list(long n);
void  fr_list_set(SEXP l, long i, SEXP v);
SEXP  fr_nil(void);
int   fr_type(SEXP x);
long  fr_len(SEXP x);
void *fr_data(SEXP x);
const char *fr_name(SEXP l, long i);
SEXP  fr_elt(SEXP l, long i);
void  fr_release_all(void);               /* free every box allocated since the last release */
/* call fn(args[0..nargs)) under a setjmp guard; returns NULL and fills msg on error() */
SEXP  fr_guard_call(void *fn, int nargs, SEXP *args, char *msg, int msglen);

#ifdef __cplusplus
}
#endif
#endif
