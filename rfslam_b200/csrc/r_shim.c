/*
 * r_shim.c -- the reference's two hot-path .Call entry points, re-pointed at librfslam_b200.so.
 *
 *   SEXP rfsrcGrow(44 x SEXP)     replaces src/randomForestSRC.c:22110-22217
 *   SEXP rfsrcPredict(58 x SEXP)  replaces src/randomForestSRC.c:22219-22402
 *
 * Same symbol names, same argument order, same named output list (names src/randomForestSRC.c:9-79, list
 * order as the reference builds it), so src/R_init_randomForestSRC.c:19-38 registers them unchanged and
 * R/rfsrc.R / R/generic.predict.rfsrc.R need no edit:
 *   - forest arrays (treeID nodeID parmID contPT mwcpSZ) hold sum(2*leafCount-1) records; R reads
 *     [1:nativeArraySize] of them (R/rfsrc.R:1766-1794), so the reference's theoretical-maximum padding is
 *     simply absent; trees appear in ascending treeID order (the reference appends in thread-completion order);
 *   - tnRCNT / tnACNT are emitted at the reference's stride treeTheoreticalMaximum = maxBootstrapSize -
 *     2*(nodeSize-1) per tree, leafCount[b] valid entries each (R/rfsrc.R:1846-1861), zero elsewhere.
 * This file only uses the R API (R.h / Rinternals.h / Rdefines.h): in an R installation it is compiled by
 * R CMD SHLIB next to the package sources and linked with -lrfslam_b200; the test suite
 * compiles it against the same stand-in headers the reference oracle is built with (oracle/Makefile, target
 * rshim) and drives both libraries with one argument packer (tests/test_rshim.py).
 */
#include <R.h>
#include <Rinternals.h>
#include <Rdefines.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rfslam_b200.h"

#define RF_OPT_PERF       0x00000004u     /* optLow  (rf.h:110) */
#define RF_OPT_TREE       0x00000020u     /* optLow  (rf.h:113): return the forest */
#define RF_OPTH_MEMB_USER 0x00000040u     /* optHigh: nodeMembership / bootMembership */
#define RF_OPTH_MEMB_OUTG 0x00010000u     /* optHigh: rmbr / ambr / tnRCNT / tnACNT outgoing */
#define RF_OPTH_MEMB_INCG 0x00020000u     /* optHigh: the same, incoming (predict) */
#define RF_OPTH_TERM_OUTG 0x00040000u     /* optHigh: terminal-node quantities (tnCLAS) outgoing */
#define RF_OPTH_TERM_INCG 0x00080000u     /* optHigh: terminal-node quantities (tnCLAS) incoming */

typedef struct { SEXP list, names; int k, cap; } outlist;

static void out_begin(outlist *o, int cap) {
  o->cap = cap; o->k = 0;
  o->list = PROTECT(allocVector(VECSXP, cap));
  o->names = PROTECT(allocVector(STRSXP, cap));
}
static int *out_int(outlist *o, const char *nm, const int32_t *v, size_t n) {
  SEXP x = PROTECT(allocVector(INTSXP, (long) n));
  if (n && v) memcpy(INTEGER(x), v, n * sizeof(int));
  SET_VECTOR_ELT(o->list, o->k, x); SET_STRING_ELT(o->names, o->k, mkChar(nm)); o->k++;
  UNPROTECT(1);
  return INTEGER(x);
}
static double *out_real(outlist *o, const char *nm, const double *v, size_t n) {
  SEXP x = PROTECT(allocVector(REALSXP, (long) n));
  if (n && v) memcpy(REAL(x), v, n * sizeof(double));
  SET_VECTOR_ELT(o->list, o->k, x); SET_STRING_ELT(o->names, o->k, mkChar(nm)); o->k++;
  UNPROTECT(1);
  return REAL(x);
}
/* the list was allocated with room to spare: hand back one of exactly k elements */
static SEXP out_end(outlist *o) {
  SEXP l = PROTECT(allocVector(VECSXP, o->k)), nm = PROTECT(allocVector(STRSXP, o->k));
  for (int i = 0; i < o->k; i++) { SET_VECTOR_ELT(l, i, VECTOR_ELT(o->list, i)); SET_STRING_ELT(nm, i, STRING_ELT(o->names, i)); }
  setAttrib(l, R_NamesSymbol, nm);
  UNPROTECT(4);
  return l;
}

static char *type_chars(SEXP v, int n) {          /* first character of each type string (rf.c:22430) */
  char *s = (char *) malloc((size_t) n + 1);
  for (int i = 0; i < n; i++) s[i] = CHAR(STRING_ELT(AS_CHARACTER(v), i))[0];
  s[n] = 0;
  return s;
}

/* per-leaf arrays at the reference's per-tree stride (R/rfsrc.R:1846-1861) */
static void put_strided(outlist *o, const char *nm, const int32_t *v, const int32_t *leafCount, int ntree, long stride, int width) {
  const size_t row = (size_t) stride * (size_t) width;
  int *dst = out_int(o, nm, NULL, (size_t) ntree * row);
  memset(dst, 0, (size_t) ntree * row * sizeof(int));
  size_t at = 0;
  for (int b = 0; b < ntree; b++) {
    const size_t cnt = (size_t) leafCount[b] * (size_t) width;
    memcpy(dst + (size_t) b * row, v + at, cnt * sizeof(int)); at += cnt;
  }
}

SEXP rfsrcGrow(SEXP traceFlag, SEXP seedPtr, SEXP optLow, SEXP optHigh, SEXP splitRule, SEXP nsplit,
               SEXP mtry, SEXP htry, SEXP ytry, SEXP nodeSize, SEXP nodeDepth, SEXP crWeightSize,
               SEXP crWeight, SEXP ntree, SEXP observationSize, SEXP ySize, SEXP rType, SEXP rLevels,
               SEXP rData, SEXP xSize, SEXP xType, SEXP xLevels, SEXP riskTime, SEXP stratifier,
               SEXP maxStrat, SEXP varCategories, SEXP categoryWeights, SEXP nCategories, SEXP toFix,
               SEXP nToFix, SEXP k_for_alpha, SEXP maxBootstrapSize, SEXP bootstrapSize, SEXP bootstrap,
               SEXP caseWeight, SEXP xSplitStatWeight, SEXP yWeight, SEXP xWeight, SEXP xData,
               SEXP timeInterestSize, SEXP timeInterest, SEXP nImpute, SEXP perfBlock, SEXP numThreads) {
  rfslam_grow_args a;
  memset(&a, 0, sizeof a);
  const int p = INTEGER(xSize)[0], ny = INTEGER(ySize)[0];
  char *rt = type_chars(rType, ny), *xt = type_chars(xType, p);
  a.trace_flag = INTEGER(traceFlag)[0];  a.seed = INTEGER(seedPtr)[0];
  a.opt_low = (uint32_t) INTEGER(optLow)[0];  a.opt_high = (uint32_t) INTEGER(optHigh)[0];
  a.split_rule = INTEGER(splitRule)[0];  a.nsplit = INTEGER(nsplit)[0];  a.mtry = INTEGER(mtry)[0];
  a.htry = INTEGER(htry)[0];  a.ytry = INTEGER(ytry)[0];
  a.node_size = INTEGER(nodeSize)[0];  a.node_depth = INTEGER(nodeDepth)[0];
  a.cr_weight_size = INTEGER(crWeightSize)[0];  a.cr_weight = REAL(crWeight);
  a.ntree = INTEGER(ntree)[0];  a.observation_size = INTEGER(observationSize)[0];  a.y_size = ny;
  a.r_type = rt;  a.r_levels = INTEGER(rLevels);  a.r_data = REAL(rData);
  a.x_size = p;  a.x_type = xt;  a.x_levels = INTEGER(xLevels);
  a.risk_time = REAL(riskTime);  a.stratifier = INTEGER(stratifier);  a.max_strat = INTEGER(maxStrat)[0];
  a.var_categories = INTEGER(varCategories);  a.category_weights = REAL(categoryWeights);
  a.n_categories = INTEGER(nCategories)[0];  a.to_fix = INTEGER(toFix);  a.n_to_fix = INTEGER(nToFix)[0];
  a.k_for_alpha = REAL(k_for_alpha)[0];  a.max_bootstrap_size = INTEGER(maxBootstrapSize)[0];
  a.bootstrap_size = INTEGER(bootstrapSize);  a.bootstrap = INTEGER(bootstrap);
  a.case_weight = REAL(caseWeight);  a.x_split_stat_weight = REAL(xSplitStatWeight);
  a.y_weight = REAL(yWeight);  a.x_weight = REAL(xWeight);  a.x_data = REAL(xData);
  a.time_interest_size = INTEGER(timeInterestSize)[0];  a.time_interest = REAL(timeInterest);
  a.n_impute = INTEGER(nImpute)[0];  a.perf_block = INTEGER(perfBlock)[0];
  a.num_threads = INTEGER(numThreads)[0];
  a.device = -1;  a.finalize_ensembles = 1;         /* tree_first / tree_step = 0: this process grows every tree */
  a.want_member_lists = (a.opt_high & RF_OPTH_MEMB_OUTG) ? 1 : 0;

  rfslam_forest_out f;
  const int rc = rfslam_b200_grow(&a, &f);
  free(rt); free(xt);
  if (rc != RFSLAM_OK) error("\nRF-SRC (b200):  %s\nRF-SRC:  The application will now exit.\n", rfslam_b200_last_error());

  const size_t n = (size_t) a.observation_size, T = (size_t) f.ntree, NN = (size_t) f.total_nodes;
  outlist o;
  out_begin(&o, 24);
  out_real(&o, "oobEnsbCLS", f.oobEnsbCLS, 2 * n);
  out_real(&o, "allEnsbCLS", f.allEnsbCLS, 2 * n);
  if (f.perfClas) out_real(&o, "perfClas", f.perfClas, 3 * T);                       /* rf.c:17454 */
  if (f.vimpClas) out_real(&o, "vimpClas", f.vimpClas, 3 * (long) f.vimp_count);    /* rf.c:17537: [xSize][1 + levels] */
  out_int(&o, "leafCount", f.leafCount, T);
  if (a.opt_low & RF_OPT_TREE) out_int(&o, "seed", f.seed, T);
  if (f.nodeMembership) {
    out_int(&o, "nodeMembership", f.nodeMembership, T * n);
    out_int(&o, "bootMembership", f.bootMembership, T * n);
  }
  if (a.opt_low & RF_OPT_TREE) {
    out_int(&o, "treeID", f.treeID, NN);
    out_int(&o, "nodeID", f.nodeID, NN);
    out_int(&o, "parmID", f.parmID, NN);
    out_real(&o, "contPT", f.contPT, NN);
    out_int(&o, "mwcpSZ", f.mwcpSZ, NN);
    out_int(&o, "mwcpPT", f.mwcpPT, (size_t) f.mwcp_total);
    out_int(&o, "mwcpCT", f.mwcpCT, T);
  }
  if (a.opt_high & RF_OPTH_MEMB_OUTG) {
    long stride = (long) a.max_bootstrap_size - 2L * ((long) a.node_size - 1L);       /* R/rfsrc.R:1846-1853 */
    if (stride < 1) stride = 1;
    for (size_t b = 0; b < T; b++) if ((long) f.leafCount[b] > stride) stride = (long) f.leafCount[b];
    out_int(&o, "rmbrMembership", f.rmbrMembership, T * (size_t) f.rmbr_stride);
    out_int(&o, "ambrMembership", f.ambrMembership, T * n);
    put_strided(&o, "tnRCNT", f.tnRCNT, f.leafCount, (int) T, stride, 1);
    put_strided(&o, "tnACNT", f.tnACNT, f.leafCount, (int) T, stride, 1);
    if (a.opt_high & RF_OPTH_TERM_OUTG)                                                 /* [leaf][class], R/rfsrc.R:1903-1909 */
      put_strided(&o, "tnCLAS", f.tnCLAS, f.leafCount, (int) T, stride, 2);
  }
  rfslam_b200_free_forest(&f);
  return out_end(&o);
}

SEXP rfsrcPredict(SEXP traceFlag, SEXP seedPtr, SEXP optLow, SEXP optHigh, SEXP ntree, SEXP observationSize,
                  SEXP ySize, SEXP rType, SEXP rLevels, SEXP rData, SEXP xSize, SEXP xType, SEXP xLevels,
                  SEXP xData, SEXP k_for_alpha, SEXP maxBootstrapSize, SEXP bootstrapSize, SEXP bootstrap,
                  SEXP caseWeight, SEXP timeInterestSize, SEXP timeInterest, SEXP totalNodeCount, SEXP seed,
                  SEXP htry, SEXP treeID, SEXP nodeID, SEXP hc_zero, SEXP hc_one, SEXP hc_parmID,
                  SEXP hc_contPT, SEXP hc_contPTR, SEXP hc_mwcpSZ, SEXP hc_mwcpPT, SEXP tnRMBR, SEXP tnAMBR,
                  SEXP tnRCNT, SEXP tnACNT, SEXP tnSURV, SEXP tnMORT, SEXP tnNLSN, SEXP tnCSHZ, SEXP tnCIFN,
                  SEXP tnREGR, SEXP tnCLAS, SEXP rTarget, SEXP rTargetCount, SEXP ptnCount,
                  SEXP intrPredictorSize, SEXP intrPredictor, SEXP partial, SEXP sobservationSize,
                  SEXP sobservationIndv, SEXP fobservationSize, SEXP frSize, SEXP frData, SEXP fxData,
                  SEXP perfBlock, SEXP numThreads) {
  rfslam_predict_args a;
  memset(&a, 0, sizeof a);
  const int p = INTEGER(xSize)[0], ny = INTEGER(ySize)[0];
  char *rt = type_chars(rType, ny), *xt = type_chars(xType, p);
  a.trace_flag = INTEGER(traceFlag)[0];  a.seed = INTEGER(seedPtr)[0];
  a.opt_low = (uint32_t) INTEGER(optLow)[0];  a.opt_high = (uint32_t) INTEGER(optHigh)[0];
  a.ntree = INTEGER(ntree)[0];  a.observation_size = INTEGER(observationSize)[0];  a.y_size = ny;
  a.r_type = rt;  a.r_levels = INTEGER(rLevels);  a.r_data = REAL(rData);
  a.x_size = p;  a.x_type = xt;  a.x_levels = INTEGER(xLevels);  a.x_data = REAL(xData);
  a.k_for_alpha = REAL(k_for_alpha)[0];  a.max_bootstrap_size = INTEGER(maxBootstrapSize)[0];
  a.bootstrap_size = INTEGER(bootstrapSize);  a.bootstrap = INTEGER(bootstrap);  a.case_weight = REAL(caseWeight);
  a.time_interest_size = INTEGER(timeInterestSize)[0];  a.time_interest = REAL(timeInterest);
  a.total_node_count = INTEGER(totalNodeCount)[0];  a.forest_seed = INTEGER(seed);  a.htry = INTEGER(htry)[0];
  a.treeID = INTEGER(treeID);  a.nodeID = INTEGER(nodeID);
  a.parmID = INTEGER(VECTOR_ELT(hc_zero, 0));  a.contPT = REAL(VECTOR_ELT(hc_zero, 1));      /* rf.c:22361-22364 */
  a.mwcpSZ = INTEGER(VECTOR_ELT(hc_zero, 2));
  if (VECTOR_ELT(hc_zero, 3) != R_NilValue) { a.mwcpPT = INTEGER(VECTOR_ELT(hc_zero, 3)); a.mwcp_total = (int64_t) length(VECTOR_ELT(hc_zero, 3)); }
  (void) hc_one; (void) hc_parmID; (void) hc_contPT; (void) hc_contPTR; (void) hc_mwcpSZ; (void) hc_mwcpPT;   /* htry > 0 only; refused below */
  a.tnRMBR = INTEGER(tnRMBR);  a.tnAMBR = INTEGER(tnAMBR);  a.tnRCNT = INTEGER(tnRCNT);  a.tnACNT = INTEGER(tnACNT);
  a.tnSURV = REAL(tnSURV);  a.tnMORT = REAL(tnMORT);  a.tnNLSN = REAL(tnNLSN);  a.tnCSHZ = REAL(tnCSHZ);
  a.tnCIFN = REAL(tnCIFN);  a.tnREGR = REAL(tnREGR);
  a.tnCLAS = INTEGER(tnCLAS);  a.tnCLAS_len = (int64_t) length(tnCLAS);
  a.r_target = INTEGER(rTarget);  a.r_target_count = INTEGER(rTargetCount)[0];  a.ptn_count = INTEGER(ptnCount)[0];
  a.intr_predictor_size = INTEGER(intrPredictorSize)[0];  a.intr_predictor = INTEGER(intrPredictor);
  a.partial_type = INTEGER(VECTOR_ELT(partial, 0))[0];  a.partial_xvar = INTEGER(VECTOR_ELT(partial, 1))[0];
  a.partial_length = INTEGER(VECTOR_ELT(partial, 2))[0];
  a.sobservation_size = INTEGER(sobservationSize)[0];  a.sobservation_indv = INTEGER(sobservationIndv);
  a.fobservation_size = INTEGER(fobservationSize)[0];  a.fr_size = INTEGER(frSize)[0];
  a.fr_data = REAL(frData);  a.fx_data = REAL(fxData);
  a.perf_block = INTEGER(perfBlock)[0];  a.num_threads = INTEGER(numThreads)[0];
  a.device = -1;  a.finalize_ensembles = 1;

  rfslam_predict_out r;
  const int rc = rfslam_b200_predict(&a, &r);
  free(rt); free(xt);
  if (rc != RFSLAM_OK) error("\nRF-SRC (b200):  %s\nRF-SRC:  The application will now exit.\n", rfslam_b200_last_error());

  const size_t n = (size_t) a.observation_size, T = (size_t) a.ntree, m = (size_t) r.m;
  outlist o;
  out_begin(&o, 12);
  if (r.oobEnsbCLS) out_real(&o, "oobEnsbCLS", r.oobEnsbCLS, 2 * m);                 /* restore mode only (rf.c:22361) */
  out_real(&o, "allEnsbCLS", r.allEnsbCLS, 2 * m);
  if (r.perfClas) out_real(&o, "perfClas", r.perfClas, 3 * T);
  if (r.vimpClas) out_real(&o, "vimpClas", r.vimpClas, 3 * (long) r.vimp_count);
  out_int(&o, "leafCount", r.leafCount, T);
  if (r.nodeMembership) {
    out_int(&o, "nodeMembership", r.nodeMembership, T * m);
    out_int(&o, "bootMembership", r.bootMembership, T * n);
  }
  rfslam_b200_free_predict(&r);
  return out_end(&o);
}
