// cpiu_ref_shim.cpp -- TEST INFRASTRUCTURE ONLY: C entry points over the UNMODIFIED reference helpers of
// /root/reference/utilities/cpiu.cpp (assemble_indicator / _accumulator / _time_elapsed / _continuous), compiled where the
// file lies against oracle/fake_rcpp/Rcpp.h.  The reference source is included by path, never copied.
#include CPIU_REF_SOURCE

extern "C" {
void cpiu_ref_indicator(const double* ev, int ne, const double* ts, const double* te, int ni, int* out) {
  IntegerVector r = assemble_indicator(NumericVector(ev, ne), NumericVector(ts, ni), NumericVector(te, ni));
  for (int i = 0; i < ni; i++) out[i] = r[i];
}
void cpiu_ref_accumulator(const double* ev, int ne, const double* ts, const double* te, int ni, int flag, int* out) {
  IntegerVector r = assemble_accumulator(NumericVector(ev, ne), NumericVector(ts, ni), NumericVector(te, ni), flag);
  for (int i = 0; i < ni; i++) out[i] = r[i];
}
void cpiu_ref_time_elapsed(const double* ev, int ne, const double* ts, const double* te, int ni, double* out) {
  NumericVector r = assemble_time_elapsed(NumericVector(ev, ne), NumericVector(ts, ni), NumericVector(te, ni));
  for (int i = 0; i < ni; i++) out[i] = r[i];
}
void cpiu_ref_continuous(const double* mt, const double* mv, int nm, const double* ts, const double* te, int ni, int k_to_persist, double* out) {
  List m; m.set("times", NumericVector(mt, nm)); m.set("values", NumericVector(mv, nm));
  NumericVector r = assemble_continuous(m, NumericVector(ts, ni), NumericVector(te, ni), k_to_persist);
  for (int i = 0; i < ni; i++) out[i] = r[i];
}
}
