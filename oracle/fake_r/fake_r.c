/*
 * fake_r.c -- TEST INFRASTRUCTURE ONLY. Implementation of the R-API stand-in declared in
 * fake_r.h; linked with the unmodified reference sources into oracle/_ref/librfslam_ref.so.
 */
#include "fake_r.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static fr_box fr_nil_box   = { NILSXP, 0, NULL, NULL };
static fr_box fr_names_box = { NILSXP, 0, NULL, NULL };
SEXP   R_NilValue    = &fr_nil_box;
SEXP   R_NamesSymbol = &fr_names_box;
double R_NaReal;

/* arena of everything handed out, so a driver can drop a whole call's worth at once */
typedef struct fr_track { fr_box *box; int owns; struct fr_track *next; } fr_track;
static fr_track *fr_head = NULL;

static jmp_buf  fr_jmp;
static int      fr_jmp_armed = 0;
static char     fr_errmsg[2048];

__attribute__((constructor)) static void fr_init(void) {
  /* R's NA_real_: quiet NaN whose low 32 bits are 1954 */
  union { double d; uint64_t u; } v;
  v.u = 0x7FF00000000007A2ULL;
  R_NaReal = v.d;
}

int R_IsNA(double x) {
  union { double d; uint64_t u; } v;
  v.d = x;
  if (x == x) return 0;
  return (uint32_t) (v.u & 0xFFFFFFFFu) == 1954u;
}

static fr_box *fr_new(int type, long n, void *alias) {
  fr_box *b = (fr_box *) calloc(1, sizeof(fr_box));
  fr_track *t = (fr_track *) malloc(sizeof(fr_track));
  size_t esz = 0;
  switch (type) {
    case INTSXP:  esz = sizeof(int);    break;
    case REALSXP: esz = sizeof(double); break;
    case STRSXP:
    case VECSXP:  esz = sizeof(SEXP);   break;
    case CHARSXP: esz = 1;              break;
    default:      esz = 1;              break;
  }
  b->type = type;
  b->len = n;
  b->names = NULL;
  if (alias != NULL) {
    b->data = alias;
    t->owns = 0;
  } else {
    b->data = calloc((size_t) (n > 0 ? n : 1) + (type == CHARSXP ? 1 : 0), esz);
    if (b->data == NULL) {
      fprintf(stderr, "fake_r: out of memory allocating %ld x %zu bytes\n", n, esz);
      abort();
    }
    t->owns = 1;
    if (type == STRSXP || type == VECSXP) {
      for (long i = 0; i < n; i++) ((SEXP *) b->data)[i] = R_NilValue;
    }
  }
  t->box = b;
  t->next = fr_head;
  fr_head = t;
  return b;
}

SEXP allocVector(int type, long n) { return fr_new(type, n, NULL); }

SEXP mkChar(const char *s) {
  long n = (long) strlen(s);
  fr_box *b = fr_new(CHARSXP, n, NULL);
  memcpy(b->data, s, (size_t) n + 1);
  return b;
}

SEXP setAttrib(SEXP x, SEXP which, SEXP val) {
  if (which == R_NamesSymbol) x->names = val;
  return val;
}

void Rprintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
}

void error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(fr_errmsg, sizeof fr_errmsg, fmt, ap);
  va_end(ap);
  if (fr_jmp_armed) longjmp(fr_jmp, 1);
  fprintf(stderr, "fake_r error(): %s\n", fr_errmsg);
  abort();
}

SEXP fr_int(const int *v, long n)      { fr_box *b = fr_new(INTSXP, n, NULL);  if (n > 0) memcpy(b->data, v, (size_t) n * sizeof(int));    return b; }
SEXP fr_real(const double *v, long n)  { fr_box *b = fr_new(REALSXP, n, NULL); if (n > 0) memcpy(b->data, v, (size_t) n * sizeof(double)); return b; }
SEXP fr_real_nocopy(double *v, long n) { return fr_new(REALSXP, n, v); }
SEXP fr_int_nocopy(int *v, long n)     { return fr_new(INTSXP, n, v); }
SEXP fr_str(const char *const *v, long n) {
  fr_box *b = fr_new(STRSXP, n, NULL);
  for (long i = 0; i < n; i++) ((SEXP *) b->data)[i] = mkChar(v[i]);
  return b;
}
SEXP fr_list(long n)                       { return fr_new(VECSXP, n, NULL); }
void fr_list_set(SEXP l, long i, SEXP v)   { ((SEXP *) l->data)[i] = v; }
SEXP fr_nil(void)                          { return R_NilValue; }
int  fr_type(SEXP x)                       { return x->type; }
long fr_len(SEXP x)                        { return x->len; }
void *fr_data(SEXP x)                      { return x->data; }
SEXP fr_elt(SEXP l, long i)                { return ((SEXP *) l->data)[i]; }
const char *fr_name(SEXP l, long i) {
  if (l->names == NULL || i >= l->names->len) return "";
  SEXP c = ((SEXP *) l->names->data)[i];
  if (c == NULL || c == R_NilValue) return "";
  return (const char *) c->data;
}

void fr_release_all(void) {
  while (fr_head != NULL) {
    fr_track *t = fr_head;
    fr_head = t->next;
    if (t->owns) free(t->box->data);
    free(t->box);
    free(t);
  }
}

typedef SEXP (*fr_fn44)(SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,
                        SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP);
typedef SEXP (*fr_fn58)(SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,
                        SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,
                        SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP,SEXP);

SEXP fr_guard_call(void *fn, int nargs, SEXP *a, char *msg, int msglen) {
  SEXP out = NULL;
  if (msg != NULL && msglen > 0) msg[0] = 0;
  fr_jmp_armed = 1;
  if (setjmp(fr_jmp) == 0) {
    if (nargs == 44) {
      out = ((fr_fn44) fn)(a[0],a[1],a[2],a[3],a[4],a[5],a[6],a[7],a[8],a[9],a[10],a[11],a[12],a[13],a[14],a[15],a[16],a[17],a[18],a[19],a[20],a[21],
                           a[22],a[23],a[24],a[25],a[26],a[27],a[28],a[29],a[30],a[31],a[32],a[33],a[34],a[35],a[36],a[37],a[38],a[39],a[40],a[41],a[42],a[43]);
    } else if (nargs == 58) {
      out = ((fr_fn58) fn)(a[0],a[1],a[2],a[3],a[4],a[5],a[6],a[7],a[8],a[9],a[10],a[11],a[12],a[13],a[14],a[15],a[16],a[17],a[18],a[19],a[20],a[21],
                           a[22],a[23],a[24],a[25],a[26],a[27],a[28],a[29],a[30],a[31],a[32],a[33],a[34],a[35],a[36],a[37],a[38],a[39],a[40],a[41],a[42],a[43],
                           a[44],a[45],a[46],a[47],a[48],a[49],a[50],a[51],a[52],a[53],a[54],a[55],a[56],a[57]);
    } else {
      snprintf(fr_errmsg, sizeof fr_errmsg, "fr_guard_call: unsupported arity %d", nargs);
      out = NULL;
      if (msg != NULL) snprintf(msg, (size_t) msglen, "%s", fr_errmsg);
    }
  } else {
    if (msg != NULL) snprintf(msg, (size_t) msglen, "%s", fr_errmsg);
    out = NULL;
  }
  fr_jmp_armed = 0;
  return out;
}
