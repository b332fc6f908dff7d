/*
 * rfslam_oracle.c -- CPU restatement of the rfSLAM Poisson-split tree-growing path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg may build, load or call this file; the product path (rfslam_b200/) never does.
 *
 * This is a from-scratch restatement (plain C99, scalar, one thread) of what the reference
 * computes on the hot path, written against the behaviour of the files below (paths relative to
 * /root/reference/src; rf.c = randomForestSRC.c, sc.c = splitCustom.c).  It is pinned against
 * the reference itself: tests/test_oracle.py runs the UNMODIFIED reference
 * (oracle/_ref/librfslam_ref.so, built by oracle/Makefile) and this file on the same seeded
 * inputs and requires identical forests; tests/golden/ holds outputs of the reference for the
 * GPU box, where /root/reference does not exist.
 *
 *   RNG + seeding .......... rf.c:5976-6117 (ran1_generic, lcgenerator), rf.c:6664-6692
 *   bootstrap .............. rf.c:492-627 (by.root SWR rf.c:529-533, by.user rf.c:520-526)
 *   node gate / purity ..... rf.c:15342-15457, rf.c:6370-6426, rf.c:13235-13270
 *   covariate draws ........ rf.c:14793-15150, rf.c:8362-8386, rf.c:8305-8308, rf.c:8206-8214
 *   cut enumeration ........ rf.c:14433-14568, rf.c:15165-15216 (LEFT iff cut - x >= 0)
 *   factor (MWCP) cuts ..... rf.c:14452-14518 (deterministic / random), rf.c:1416-1559 (makeFactor, bookPair: the
 *                            complementary pairs by left-side cardinality, lexicographic inside a cardinality, the
 *                            middle cardinality of an even level count halved), rf.c:15232-15330 (getRandomPair,
 *                            createRandomBinaryPair, convertRelToAbsBinaryPair), rf.c:1607-1616 (splitOnFactor)
 *   split statistics ....... sc.c:639-1687 (poissonSplit1..6), called from rf.c:13404-13529
 *   best-split tracker ..... rf.c:15471-15551 (EPSILON = 1e-9 absolute, rf.h:169)
 *   fork / node ids ........ rf.c:10568-10734, rf.c:21620-21691
 *   leaf statistics ........ rf.c:755-858, rf.c:7585-7657
 *   ensembles .............. rf.c:859-973, rf.c:8113-8125
 *   forest record order .... rf.c:10735-10806 (pre-order)
 *
 * Deliberate differences from the reference (none changes a result beyond floating-point
 * summation order, which the reference itself does not fix: its quicksort is unstable,
 * rf.c:5575-5640):
 *   - per-cut per-side sufficient statistics (E, N, T, W per stratum) come from one forward and
 *     one backward pass over the covariate-sorted rows instead of a fresh O(m) pass per cut;
 *   - trees are stored in tree order 1..ntree (the reference appends in thread-completion order);
 *   - outputs are sized by the actual node counts.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rfslam_oracle.h"

/* ------------------------------------------------------------------ RNG (rf.c:5976-6117) */

#define RAN_IA 16807
#define RAN_IM 2147483647
#define RAN_IQ 127773
#define RAN_IR 2836
#define RAN_NTAB 32
#define RAN_NDIV (1 + (RAN_IM - 1) / RAN_NTAB)

typedef struct { int iy; int iv[RAN_NTAB]; int idum; } ran1_t;

static void ran1_seed(ran1_t *s, int seed) { s->iy = 0; s->idum = seed; memset(s->iv, 0, sizeof s->iv); }

static int ran1_step(int v) {           /* Schrage: v * 16807 mod (2^31 - 1) */
  int k = v / RAN_IQ;
  v = RAN_IA * (v - k * RAN_IQ) - RAN_IR * k;
  if (v < 0) v += RAN_IM;
  return v;
}

float orc_ran1(ran1_t *s) {
  const double am = 1.0 / RAN_IM;
  const double rnmx = 1.0 - 1.2e-7;
  if (s->idum <= 0 || s->iy == 0) {
    s->idum = (-(s->idum) < 1) ? 1 : -(s->idum);
    for (int j = RAN_NTAB + 7; j >= 0; j--) {
      s->idum = ran1_step(s->idum);
      if (j < RAN_NTAB) s->iv[j] = s->idum;
    }
    s->iy = s->iv[1];                 /* sic: the reference's generic variant reads slot 1 (rf.c:6095) */
  }
  s->idum = ran1_step(s->idum);
  int j = s->iy / RAN_NDIV;
  s->iy = s->iv[j];
  s->iv[j] = s->idum;
  float t = (float) (am * s->iy);
  if ((double) t > rnmx) return (float) rnmx;
  return t;
}

static unsigned lcg_next(unsigned s) { return (1366u * s + 150889u) % 714025u; }

/* chain seeds for trees 1..ntree, three chains A (bootstrap), B (covariates), C (unused here) */
void orc_chain_seeds(int seed, int ntree, int *seedA, int *seedB) {
  unsigned s = (unsigned) abs(seed);
  if (s >= 714025u) s %= 714025u;
  for (int pass = 0; pass < 2; pass++) {
    for (int b = 0; b < ntree; b++) {
      s = lcg_next(s);
      s = lcg_next(s);
      while (s == 0) s = lcg_next(s);
      if (pass == 0) seedA[b] = -(int) s; else seedB[b] = -(int) s;
    }
  }
}

/* in predict / restore mode the LCG starts from 0 and chain A comes from forest$seed
   (rf.c:6453, rf.c:6665, rf.c:7046-7050) */
void orc_chain_seeds_restore(int ntree, int *seedB) {
  unsigned s = 0;
  for (int b = 0; b < ntree; b++) {
    s = lcg_next(s);
    s = lcg_next(s);
    while (s == 0) s = lcg_next(s);
    seedB[b] = -(int) s;
  }
}

/* ------------------------------------------------------------ split statistic (sc.c:639-1687) */

typedef struct { double E, N, T, W; } suff_t;     /* per (side, stratum) sufficient statistics */

static int rule_stratified(int rule) { return rule == 1 || rule == 3 || rule == 4 || rule == 6; }
static int rule_unit_time(int rule)  { return rule == 3 || rule == 6; }
static int rule_bayes(int rule)      { return rule <= 3; }

/* one side's statistic: sum over strata 1..K in ascending order (sc.c:751-764, sc.c:1262-1290) */
static double side_stat(int rule, int K, const suff_t *s, double alpha, double beta) {
  double stat = 0.0;
  if (rule_bayes(rule)) {
    for (int j = 1; j <= K; j++) {
      double lam = (s[j].E + alpha) / (beta + s[j].T);
      stat += s[j].E * log(lam);
    }
  } else {
    for (int j = 1; j <= K; j++) {
      double mu = (s[j].T == 0) ? 0.0 : s[j].W / s[j].T;
      if (mu != 0) stat += (s[j].E * log(mu)) - (s[j].N * mu);
    }
  }
  return stat;
}

/* Restatement of one plugin call (sc.c: poissonSplitN) for direct comparison with the reference
   symbol; arrays are 0-based here, membership 1 = LEFT. */
double orc_poisson_stat(int rule, int n, double k_for_alpha, const char *membership,
                        const double *response, const double *strat, const double *rt) {
  int K = 1;
  if (rule_stratified(rule)) { K = 0; for (int i = 0; i < n; i++) if (strat[i] > K) K = (int) strat[i]; }
  double ysum = 0, rtsum = 0;
  for (int i = 0; i < n; i++) { ysum += response[i]; rtsum += rt[i]; }
  double alpha = 1 / (k_for_alpha * k_for_alpha);
  double beta = alpha / (ysum / rtsum);
  suff_t *P = calloc(3 * (size_t) (K + 1), sizeof(suff_t)), *L = P + (K + 1), *R = L + (K + 1);
  for (int i = 0; i < n; i++) {
    int j = rule_stratified(rule) ? (int) strat[i] : 1;
    double t = rule_unit_time(rule) ? 1.0 : rt[i];
    double e = response[i] - 1;
    suff_t *S = (membership[i] == 1) ? L : R;
    if (response[i] == 2.0) { P[j].E++; S[j].E++; }
    P[j].N++; S[j].N++;
    P[j].T += t; S[j].T += t;
    P[j].W += e * t; S[j].W += e * t;
  }
  double d = side_stat(rule, K, L, alpha, beta) + side_stat(rule, K, R, alpha, beta) - side_stat(rule, K, P, alpha, beta);
  free(P);
  return d;
}

/* -------------------------------------------------------------------------- tree growth */

typedef struct {
  const orc_args *a;
  int b;                       /* 0-based tree */
  ran1_t chainB;
  int leafCount;
  int Kmax;
  /* growing per-tree outputs */
  int *nodeID, *parmID; double *contPT, *stat; unsigned *mwcp; long nnodes, cap;   /* mwcp[k] != 0: factor split, the LEFT level mask */
  int *nodeMembership;         /* [n] leaf id of every row */
  int *tnRCNT, *tnACNT, *tnE;  /* indexed by leaf id (1-based), capacity n+1 */
  int *rmbr, *ambr; long rmbrIter, ambrIter;
  long candidates;
  /* scratch */
  double *xs; int *ord; suff_t *L, *R, *P; double *statL, *statR; double *gval; int *gend;
} tree_ctx;

static void push_node(tree_ctx *c, int nodeID, int parm, double cut, double stat, unsigned mwcp) {
  if (c->nnodes == c->cap) {
    c->cap = c->cap ? 2 * c->cap : 1024;
    c->nodeID = realloc(c->nodeID, c->cap * sizeof(int));
    c->parmID = realloc(c->parmID, c->cap * sizeof(int));
    c->contPT = realloc(c->contPT, c->cap * sizeof(double));
    c->stat   = realloc(c->stat,   c->cap * sizeof(double));
    c->mwcp   = realloc(c->mwcp,   c->cap * sizeof(unsigned));
  }
  c->mwcp[c->nnodes] = mwcp;
  c->nodeID[c->nnodes] = nodeID; c->parmID[c->nnodes] = parm; c->contPT[c->nnodes] = cut; c->stat[c->nnodes] = stat;
  c->nnodes++;
}

static const double *g_sort_key;
static int cmp_ord(const void *pa, const void *pb) {
  int a = *(const int *) pa, b = *(const int *) pb;
  double xa = g_sort_key[a], xb = g_sort_key[b];
  if (xa < xb) return -1;
  if (xa > xb) return 1;
  return (a > b) - (a < b);
}

/* Score every cut of covariate `cov` in a node whose in-bag rows are rep[0..m); feed each delta
   (already multiplied by the split weight) to the sequential tracker of rf.c:15471-15551.
   Returns the number of distinct values. */
static int score_factor(tree_ctx *c, int cov, const int *rep, int m, int D,
                        int *have_best, double *best_delta, int *best_cov, double *best_cut, unsigned *best_mask);

static int score_covariate(tree_ctx *c, int cov, const int *rep, int m,
                           int *have_best, double *best_delta, int *best_cov, double *best_cut, unsigned *best_mask) {
  const orc_args *a = c->a;
  const double *col = a->x + (size_t) cov * a->n;
  const int rule = a->rule;
  double *xs = c->xs; int *ord = c->ord;
  for (int i = 0; i < m; i++) { xs[i] = col[rep[i]]; ord[i] = i; }
  g_sort_key = xs;
  qsort(ord, m, sizeof(int), cmp_ord);
  /* distinct-value groups (rf.c:14989-14996): gend[g] = one past the last sorted position of group g */
  int D = 0;
  for (int i = 0; i < m; i++) {
    if (i == 0 || xs[ord[i]] > c->gval[D - 1]) { c->gval[D] = xs[ord[i]]; D++; }
    c->gend[D - 1] = i + 1;
  }
  if (D < 2) return D;
  if (a->xtype && a->xtype[cov] == 'C') return score_factor(c, cov, rep, m, D, have_best, best_delta, best_cov, best_cut, best_mask);
  /* nsplit > 0 (stackAndConstructSplitVector, randomForestSRC.c:14533-14564): with more than nsplit distinct values,
     nsplit of the D - 1 cut positions are drawn from chain B without replacement (swap-remove over 1..D-1,
     slot = ceil(ran1B * size)) and evaluated in ascending order; otherwise every position is a cut */
  char *sel = NULL;
  if (a->nsplit > 0 && D > a->nsplit) {
    sel = calloc(D, 1);
    int *swor = malloc((size_t) D * sizeof(int)); int sz = D - 1;
    for (int j = 1; j <= sz; j++) swor[j] = j;
    for (int j = 1; j <= a->nsplit; j++) {
      int idx = (int) ceil((double) orc_ran1(&c->chainB) * (sz * 1.0));
      sel[swor[idx] - 1] = 1;
      swor[idx] = swor[sz]; sz--;
    }
    free(swor);
    c->candidates += a->nsplit;
  } else {
    c->candidates += D - 1;
  }

  int K = 1;
  if (rule_stratified(rule)) { K = 0; for (int i = 0; i < m; i++) if (a->strat[rep[i]] > K) K = a->strat[rep[i]]; }
  double ysum = 0, rtsum = 0;
  for (int i = 0; i < m; i++) { int r = rep[ord[i]]; ysum += a->y[r]; rtsum += a->rt[r]; }
  double alpha = 1 / (a->k_for_alpha * a->k_for_alpha);
  double beta = alpha / (ysum / rtsum);

  suff_t *L = c->L, *R = c->R, *P = c->P;
  memset(L, 0, (K + 1) * sizeof(suff_t)); memset(R, 0, (K + 1) * sizeof(suff_t)); memset(P, 0, (K + 1) * sizeof(suff_t));
#define ADD_ROW(S, r) do { \
    int j_ = rule_stratified(rule) ? a->strat[r] : 1; \
    double t_ = rule_unit_time(rule) ? 1.0 : a->rt[r]; \
    double e_ = a->y[r] - 1; \
    if (a->y[r] == 2.0) (S)[j_].E++; \
    (S)[j_].N++; (S)[j_].T += t_; (S)[j_].W += e_ * t_; } while (0)
  /* forward pass: left-side statistic after each group except the last */
  int pos = 0;
  for (int g = 0; g < D; g++) {
    for (; pos < c->gend[g]; pos++) { int r = rep[ord[pos]]; ADD_ROW(L, r); ADD_ROW(P, r); }
    if (g < D - 1) c->statL[g] = side_stat(rule, K, L, alpha, beta);
  }
  double statP = side_stat(rule, K, P, alpha, beta);
  /* backward pass: right-side statistic for the cut after group g = rows of groups g+1..D-1 */
  pos = m;
  for (int g = D - 1; g >= 1; g--) {
    for (; pos > c->gend[g - 1]; pos--) { int r = rep[ord[pos - 1]]; ADD_ROW(R, r); }
    c->statR[g - 1] = side_stat(rule, K, R, alpha, beta);
  }
#undef ADD_ROW
  double w = a->split_weight ? a->split_weight[cov] : 1.0;
  for (int g = 0; g < D - 1; g++) {
    if (sel && !sel[g]) continue;
    double delta = (c->statL[g] + c->statR[g] - statP) * w;
    int flag;
    if (!*have_best) flag = 1;
    else flag = (delta - *best_delta) > 1.0e-9;
    if (flag) { *have_best = 1; *best_delta = delta; *best_cov = cov; *best_cut = c->gval[g]; *best_mask = 0u; }
  }
  free(sel);
  return D;
}

/* n choose k exactly (rf.c:1531-1560 computes the same integers with fraction reduction) */
static double choose_d(int n, int k) {
  double t = 1.0;
  if (k > n - k) k = n - k;
  for (int i = 1; i <= k; i++) t = t * (double) (n - k + i) / (double) i;
  return floor(t + 0.5);
}

/* the next combination of g of the r relative levels 1..r in lexicographic order; lv[0..g) ascending */
static int next_combination(int *lv, int g, int r) {
  int i = g - 1;
  while (i >= 0 && lv[i] == r - (g - 1 - i)) i--;
  if (i < 0) return 0;
  lv[i]++;
  for (int j = i + 1; j < g; j++) lv[j] = lv[j - 1] + 1;
  return 1;
}

/* Unordered factor: the r levels present in the node (c->gval[0..D), ascending) are split into a LEFT set and its
   complement.  Deterministic mode (rf.c:14463-14478, rf.c:14502-14517): all 2^(r-1) - 1 complementary pairs when that
   is fewer than the node's in-bag size, grouped by the cardinality of the LEFT set (1 .. floor(r/2)), lexicographic
   inside a group, only the first half of the middle group when r is even (bookPair, rf.c:1498-1530).  Otherwise as
   many RANDOM pairs as the node has in-bag members (rf.c:14479-14481, nsplit > 0: rf.c:14486-14498): a cardinality
   drawn in proportion to the group sizes, then that many levels without replacement, all from chain B
   (getRandomPair / createRandomBinaryPair, rf.c:15232-15302).  A cut's LEFT mask holds ABSOLUTE levels: bit
   (level - 1) (convertRelToAbsBinaryPair rf.c:15303-15330; one 32-bit word: levels <= 32). */
static int score_factor(tree_ctx *c, int cov, const int *rep, int m, int D,
                        int *have_best, double *best_delta, int *best_cov, double *best_cut, unsigned *best_mask) {
  const orc_args *a = c->a;
  const double *col = a->x + (size_t) cov * a->n;
  const int rule = a->rule, r = D;
  const double pairs = ldexp(1.0, r - 1) - 1.0;
  int det, ncuts;
  if (a->nsplit == 0) {
    det = !(pairs >= (double) m);
    ncuts = det ? (int) pairs : m;
  } else {
    int lim = a->nsplit <= m ? a->nsplit : m;
    det = pairs <= (double) lim;
    ncuts = det ? (int) pairs : lim;
  }
  unsigned *masks = malloc((size_t) ncuts * sizeof(unsigned));
  int G = r / 2;
  if (det) {
    int j = 0, lv[32];
    for (int g = 1; g <= G; g++) {
      double gs = choose_d(r, g);
      if (!(r & 1) && g == G) gs = gs / 2;
      for (int i = 0; i < g; i++) lv[i] = i + 1;
      for (long k = 0; k < (long) gs; k++) {
        unsigned mk = 0;
        for (int i = 0; i < g; i++) mk |= 1u << ((unsigned) c->gval[lv[i] - 1] - 1u);
        masks[j++] = mk;
        next_combination(lv, g, r);
      }
    }
  } else {
    double cdf[17];
    for (int k = 1; k <= G; k++) { cdf[k] = choose_d(r, k); if (!(r & 1) && k == G) cdf[k] = floor(cdf[k] / 2); }
    for (int k = 2; k <= G; k++) cdf[k] += cdf[k - 1];
    for (int j = 0; j < ncuts; j++) {
      double rv = ceil(orc_ran1(&c->chainB) * cdf[G]);
      int g = 1; while (rv > cdf[g]) g++;
      int lvv[33], size = r; unsigned mk = 0;
      for (int k = 1; k <= r; k++) lvv[k] = k;
      for (int k = 1; k <= g; k++) {
        int slot = (int) ceil((double) orc_ran1(&c->chainB) * (size * 1.0));
        int rel = lvv[slot];
        lvv[slot] = lvv[size]; size--;
        mk |= 1u << ((unsigned) c->gval[rel - 1] - 1u);
      }
      masks[j] = mk;
    }
  }
  c->candidates += ncuts;

  int K = 1;
  if (rule_stratified(rule)) { K = 0; for (int i = 0; i < m; i++) if (a->strat[rep[i]] > K) K = a->strat[rep[i]]; }
  double ysum = 0, rtsum = 0;
  for (int i = 0; i < m; i++) { int rr = rep[c->ord[i]]; ysum += a->y[rr]; rtsum += a->rt[rr]; }
  double alpha = 1 / (a->k_for_alpha * a->k_for_alpha);
  double beta = alpha / (ysum / rtsum);
  suff_t *L = c->L, *R = c->R, *P = c->P;
  double w = a->split_weight ? a->split_weight[cov] : 1.0;
  for (int j = 0; j < ncuts; j++) {
    memset(L, 0, (K + 1) * sizeof(suff_t)); memset(R, 0, (K + 1) * sizeof(suff_t)); memset(P, 0, (K + 1) * sizeof(suff_t));
    for (int i = 0; i < m; i++) {                       /* rows in covariate-sorted order, as the plugin sees them */
      int rr = rep[c->ord[i]];
      int left = (masks[j] >> ((unsigned) col[rr] - 1u)) & 1u;
      suff_t *S = left ? L : R;
      int j_ = rule_stratified(rule) ? a->strat[rr] : 1;
      double t_ = rule_unit_time(rule) ? 1.0 : a->rt[rr];
      double e_ = a->y[rr] - 1;
      if (a->y[rr] == 2.0) { S[j_].E++; P[j_].E++; }
      S[j_].N++; S[j_].T += t_; S[j_].W += e_ * t_;
      P[j_].N++; P[j_].T += t_; P[j_].W += e_ * t_;
    }
    double delta = (side_stat(rule, K, L, alpha, beta) + side_stat(rule, K, R, alpha, beta) - side_stat(rule, K, P, alpha, beta)) * w;
    int flag;
    if (!*have_best) flag = 1;
    else flag = (delta - *best_delta) > 1.0e-9;
    if (flag) { *have_best = 1; *best_delta = delta; *best_cov = cov; *best_cut = NAN; *best_mask = masks[j]; }
  }
  free(masks);
  return D;
}

static void grow_node(tree_ctx *c, int *rep, int m, int *all, int na, int depth, int nodeID, char *perm) {
  const orc_args *a = c->a;
  int have_best = 0, best_cov = -1; double best_delta = 0, best_cut = 0; unsigned best_mask = 0u;
  int attempt = (m >= 2 * a->nodesize) && (a->nodedepth < 0 || depth < a->nodedepth);
  if (attempt) {
    /* purity: variance of the response codes over in-bag rows (rf.c:6391-6416) */
    double mean = 0; for (int i = 0; i < m; i++) mean += a->y[rep[i]];
    mean /= (double) m;
    double var = 0; for (int i = 0; i < m; i++) var += pow(mean - a->y[rep[i]], 2.0);
    var /= (double) m;
    if (var <= 1.0e-9) attempt = 0;
  }
  if (attempt) {
    int *pool = malloc((a->p + 1) * sizeof(int)); int size = 0;
    for (int k = 0; k < a->p; k++) if (perm[k]) pool[++size] = k;     /* 1-based slots, rf.c:8209-8213 */
    int count = 0;
    while (count < a->mtry && size > 0) {
      count++;
      int slot = (int) ceil((double) orc_ran1(&c->chainB) * (size * 1.0));
      int cov = pool[slot];
      pool[slot] = pool[size]; size--;
      int D = score_covariate(c, cov, rep, m, &have_best, &best_delta, &best_cov, &best_cut, &best_mask);
      if (D < 2) perm[cov] = 0;                                        /* rf.c:15004-15007 */
    }
    free(pool);
  }
  if (have_best) {
    push_node(c, nodeID, best_cov + 1, best_cut, best_delta, best_mask);
    c->leafCount++;
    int leftID = nodeID, rightID = c->leafCount;
    const double *col = a->x + (size_t) best_cov * a->n;
    int *repL = malloc((size_t) m * sizeof(int)), *repR = malloc((size_t) m * sizeof(int)); int mL = 0, mR = 0;
#define GOES_LEFT(row) (best_mask ? (int) ((best_mask >> ((unsigned) col[row] - 1u)) & 1u) : (best_cut - col[row] >= 0.0))   /* rf.c:1607-1616 / rf.c:15194 */
    for (int i = 0; i < m; i++) { if (GOES_LEFT(rep[i])) repL[mL++] = rep[i]; else repR[mR++] = rep[i]; }
    int *allL = malloc((size_t) na * sizeof(int)), *allR = malloc((size_t) na * sizeof(int)); int aL = 0, aR = 0;
    for (int i = 0; i < na; i++) { if (GOES_LEFT(all[i])) allL[aL++] = all[i]; else allR[aR++] = all[i]; }
#undef GOES_LEFT
    char *permL = malloc(a->p), *permR = malloc(a->p);
    memcpy(permL, perm, a->p); memcpy(permR, perm, a->p);
    grow_node(c, repL, mL, allL, aL, depth + 1, leftID, permL);
    grow_node(c, repR, mR, allR, aR, depth + 1, rightID, permR);
    free(repL); free(repR); free(allL); free(allR); free(permL); free(permR);
  } else {
    push_node(c, nodeID, 0, NAN, NAN, 0u);
    int E = 0; for (int i = 0; i < m; i++) if (a->y[rep[i]] == 2.0) E++;
    c->tnRCNT[nodeID] = m; c->tnACNT[nodeID] = na; c->tnE[nodeID] = E;
    for (int i = 0; i < na; i++) { c->nodeMembership[all[i]] = nodeID; c->ambr[c->ambrIter++] = all[i] + 1; }
    for (int i = 0; i < m; i++) c->rmbr[c->rmbrIter++] = rep[i] + 1;
  }
}

orc_out *orc_grow(const orc_args *a) {
  const int n = a->n, ntree = a->ntree;
  orc_out *o = calloc(1, sizeof *o);
  o->ntree = ntree; o->n = n;
  o->leafCount = calloc(ntree, sizeof(int));
  o->mwcpCT = calloc(ntree, sizeof(int));
  o->seed = calloc(ntree, sizeof(int));
  o->bootMembership = calloc((size_t) ntree * n, sizeof(int));
  o->nodeMembership = calloc((size_t) ntree * n, sizeof(int));
  o->rmbr_stride = (a->max_boot > 0 ? a->max_boot : n);           /* RF-SRC sizes rmbr by maxBootstrapSize (rf.c:18296) */
  o->rmbr = calloc((size_t) ntree * o->rmbr_stride, sizeof(int));
  o->ambr = calloc((size_t) ntree * n, sizeof(int));
  o->oobNum = calloc(2 * (size_t) n, sizeof(double)); o->allNum = calloc(2 * (size_t) n, sizeof(double));
  o->oobDen = calloc(n, sizeof(int)); o->allDen = calloc(n, sizeof(int));
  o->oobEns = calloc(2 * (size_t) n, sizeof(double)); o->allEns = calloc(2 * (size_t) n, sizeof(double));
  int *seedA = malloc(ntree * sizeof(int)), *seedB = malloc(ntree * sizeof(int));
  if (a->tree_seeds) { memcpy(seedA, a->tree_seeds, ntree * sizeof(int)); orc_chain_seeds_restore(ntree, seedB); }
  else orc_chain_seeds(a->seed, ntree, seedA, seedB);
  int Kmax = 1; for (int i = 0; i < n; i++) if (a->strat[i] > Kmax) Kmax = a->strat[i];

  long leaf_cap = 0, node_cap = 0;
  for (int b = 0; b < ntree; b++) {
    tree_ctx c; memset(&c, 0, sizeof c);
    c.a = a; c.b = b; c.Kmax = Kmax;
    o->seed[b] = seedA[b];
    ran1_t chainA; ran1_seed(&chainA, seedA[b]); ran1_seed(&c.chainB, seedB[b]);
    /* bootstrap (rf.c:520-533) */
    int bs; int *rep;
    int *boot = o->bootMembership + (size_t) b * n;
    if (a->samp) {
      const int *cnt = a->samp + (size_t) b * n; bs = 0;
      for (int k = 0; k < n; k++) bs += cnt[k];
      rep = malloc((size_t) (bs > 0 ? bs : 1) * sizeof(int)); int i = 0;
      for (int k = 0; k < n; k++) for (int j = 0; j < cnt[k]; j++) rep[i++] = k;
    } else {
      bs = a->boot_size ? a->boot_size[b] : n;
      rep = malloc((size_t) bs * sizeof(int));
      for (int i = 0; i < bs; i++) {
        int k = (int) ceil((double) orc_ran1(&chainA) * (n * 1.0));
        rep[i] = k - 1;
      }
    }
    for (int i = 0; i < bs; i++) boot[rep[i]]++;
    int *all = malloc((size_t) n * sizeof(int)); for (int i = 0; i < n; i++) all[i] = i;
    char *perm = malloc(a->p); memset(perm, 1, a->p);
    c.nodeMembership = o->nodeMembership + (size_t) b * n;
    c.tnRCNT = calloc((size_t) n + 2, sizeof(int)); c.tnACNT = calloc((size_t) n + 2, sizeof(int)); c.tnE = calloc((size_t) n + 2, sizeof(int));
    c.rmbr = o->rmbr + (size_t) b * o->rmbr_stride; c.ambr = o->ambr + (size_t) b * n;
    c.xs = malloc((size_t) bs * sizeof(double)); c.ord = malloc((size_t) bs * sizeof(int));
    c.L = malloc((Kmax + 1) * sizeof(suff_t)); c.R = malloc((Kmax + 1) * sizeof(suff_t)); c.P = malloc((Kmax + 1) * sizeof(suff_t));
    c.statL = malloc((size_t) bs * sizeof(double)); c.statR = malloc((size_t) bs * sizeof(double));
    c.gval = malloc((size_t) bs * sizeof(double)); c.gend = malloc((size_t) bs * sizeof(int));
    c.leafCount = 1;
    grow_node(&c, rep, bs, all, n, 0, 1, perm);
    o->leafCount[b] = c.leafCount;
    o->candidates += c.candidates;
    /* append forest records */
    if (o->total_nodes + c.nnodes > node_cap) {
      node_cap = 2 * (o->total_nodes + c.nnodes);
      o->treeID = realloc(o->treeID, node_cap * sizeof(int)); o->nodeID = realloc(o->nodeID, node_cap * sizeof(int));
      o->parmID = realloc(o->parmID, node_cap * sizeof(int)); o->contPT = realloc(o->contPT, node_cap * sizeof(double));
      o->splitStat = realloc(o->splitStat, node_cap * sizeof(double));
      o->mwcpSZ = realloc(o->mwcpSZ, node_cap * sizeof(int)); o->mwcpPT = realloc(o->mwcpPT, node_cap * sizeof(unsigned));
    }
    for (long k = 0; k < c.nnodes; k++) {
      long q = o->total_nodes + k;
      o->treeID[q] = b + 1; o->nodeID[q] = c.nodeID[k]; o->parmID[q] = c.parmID[k]; o->contPT[q] = c.contPT[k]; o->splitStat[q] = c.stat[k];
      o->mwcpSZ[q] = c.mwcp[k] ? 1 : 0;                               /* one word: levels <= 32 (rf.c:10773-10779) */
      if (c.mwcp[k]) { o->mwcpPT[o->mwcp_total++] = c.mwcp[k]; o->mwcpCT[b]++; }
    }
    o->total_nodes += c.nnodes;
    if (o->total_leaves + c.leafCount > leaf_cap) {
      leaf_cap = 2 * (o->total_leaves + c.leafCount);
      o->tnRCNT = realloc(o->tnRCNT, leaf_cap * sizeof(int)); o->tnACNT = realloc(o->tnACNT, leaf_cap * sizeof(int));
      o->tnCLAS = realloc(o->tnCLAS, 2 * leaf_cap * sizeof(int));
    }
    for (int l = 1; l <= c.leafCount; l++) {
      long q = o->total_leaves + l - 1;
      o->tnRCNT[q] = c.tnRCNT[l]; o->tnACNT[q] = c.tnACNT[l];
      o->tnCLAS[2 * q] = c.tnRCNT[l] - c.tnE[l]; o->tnCLAS[2 * q + 1] = c.tnE[l];
    }
    o->total_leaves += c.leafCount;
    /* ensembles (rf.c:928-962): class proportion of the leaf's in-bag members */
    for (int i = 0; i < n; i++) {
      int l = c.nodeMembership[i];
      double p1 = (double) (c.tnRCNT[l] - c.tnE[l]) / (double) c.tnRCNT[l];
      double p2 = (double) c.tnE[l] / (double) c.tnRCNT[l];
      o->allDen[i]++; o->allNum[i] += p1; o->allNum[n + i] += p2;
      if (boot[i] == 0) { o->oobDen[i]++; o->oobNum[i] += p1; o->oobNum[n + i] += p2; }
    }
    free(rep); free(all); free(perm); free(c.tnRCNT); free(c.tnACNT); free(c.tnE);
    free(c.xs); free(c.ord); free(c.L); free(c.R); free(c.P); free(c.statL); free(c.statR); free(c.gval); free(c.gend);
    free(c.nodeID); free(c.parmID); free(c.contPT); free(c.stat); free(c.mwcp);
  }
  for (int i = 0; i < n; i++) {                     /* rf.c:8113-8125 */
    for (int k = 0; k < 2; k++) {
      o->allEns[(size_t) k * n + i] = o->allDen[i] ? o->allNum[(size_t) k * n + i] / o->allDen[i] : 0.0;
      o->oobEns[(size_t) k * n + i] = o->oobDen[i] ? o->oobNum[(size_t) k * n + i] / o->oobDen[i] : 0.0;  /* never-OOB rows keep the 0 the output was allocated with (rf.c:8115) */
    }
  }
  free(seedA); free(seedB);
  return o;
}

/* Predict: route rows of fx ([p][m] columns) down every tree of a pre-order forest and average the
   leaf class proportions (rf.c:3687-3747 routing, rf.c:928-962 ensemble).  leaf class counts come
   from tnCLAS-style arrays (per tree, indexed by leaf id). */
void orc_predict(int ntree, int p, long total_nodes, const int *treeID, const int *nodeID, const int *parmID,
                 const double *contPT, const int *leafCount, const int *tnRCNT, const int *tnE,
                 const double *fx, int m, double *ens /* [2][m] */, int *membership /* [ntree][m] */,
                 const int *mwcpSZ /* [total_nodes] or NULL */, const unsigned *mwcpPT /* the words, in record order */) {
  (void) p;
  long *mwoff = NULL;                                  /* word index of a factor split */
  if (mwcpSZ) { mwoff = malloc(total_nodes * sizeof(long)); long at = 0; for (long q = 0; q < total_nodes; q++) { mwoff[q] = at; at += mwcpSZ[q]; } }
  long *child = malloc(total_nodes * sizeof(long));   /* right-child record index for internal nodes */
  /* pre-order: left child of record q is q+1; right child follows the whole left subtree */
  long *stack = malloc((total_nodes + 1) * sizeof(long)); long sp = 0;
  for (long q = 0; q < total_nodes; q++) child[q] = -1;
  for (long q = 0; q < total_nodes; q++) {
    /* a record closes the most recent internal node still waiting for its right child when the
       previous record ended a left subtree */
    if (q > 0 && treeID[q] == treeID[q - 1] && parmID[q - 1] == 0) {
      long par = stack[--sp];
      child[par] = q;
    }
    if (q > 0 && treeID[q] != treeID[q - 1]) sp = 0;
    if (parmID[q] != 0) stack[sp++] = q;
  }
  memset(ens, 0, 2 * (size_t) m * sizeof(double));
  long root = 0, leafoff = 0;
  for (int b = 0; b < ntree; b++) {
    for (int i = 0; i < m; i++) {
      long q = root;
      while (parmID[q] != 0) {
        double xv = fx[(size_t) (parmID[q] - 1) * m + i];
        if (mwcpSZ && mwcpSZ[q] > 0) q = ((mwcpPT[mwoff[q]] >> ((unsigned) xv - 1u)) & 1u) ? q + 1 : child[q];   /* rf.c:1607-1616 */
        else q = (contPT[q] - xv >= 0.0) ? q + 1 : child[q];
      }
      int l = nodeID[q];
      if (membership) membership[(size_t) b * m + i] = l;
      double cnt = (double) tnRCNT[leafoff + l - 1], e = (double) tnE[leafoff + l - 1];
      ens[i] += (cnt - e) / cnt; ens[(size_t) m + i] += e / cnt;
    }
    root += 2 * (long) leafCount[b] - 1;
    leafoff += leafCount[b];
  }
  for (size_t i = 0; i < 2 * (size_t) m; i++) ens[i] /= (double) ntree;
  free(child); free(stack); free(mwoff);
}

void orc_free(orc_out *o) {
  if (!o) return;
  free(o->treeID); free(o->nodeID); free(o->parmID); free(o->contPT); free(o->splitStat);
  free(o->mwcpSZ); free(o->mwcpPT); free(o->mwcpCT);
  free(o->leafCount); free(o->seed); free(o->bootMembership); free(o->nodeMembership);
  free(o->tnRCNT); free(o->tnACNT); free(o->tnCLAS); free(o->rmbr); free(o->ambr);
  free(o->oobNum); free(o->allNum); free(o->oobDen); free(o->allDen); free(o->oobEns); free(o->allEns);
  free(o);
}

/* perfClas of the finalized OOB ensemble (default performance of the class family):
   out[0] = getClassificationIndex (randomForestSRC.c:1067-1095) over getMaxVote (randomForestSRC.c:1175-1218:
   the last class whose probability is >= the running maximum wins, i.e. class 2 on a tie);
   out[1..2] = getConditionalClassificationIndex (randomForestSRC.c:1025-1066): correct OOB votes of the class
   over ALL rows of the class.  na = the caller's NA_REAL.  y in {1.0, 2.0}; ens = [2][n] class-major. */
void orc_perf_clas(int n, const double *y, const double *ens, const int *den, double na, double *out) {
  double correct = 0.0, cpv[3] = {0.0, 0.0, 0.0};
  unsigned cum = 0, count[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    count[(unsigned) y[i]]++;
    if (den[i] != 0) {
      double maxValue = 0.0, maxClass = 0.0;
      for (int k = 1; k <= 2; k++)
        if (maxValue <= ens[(size_t)(k - 1) * n + i]) { maxValue = ens[(size_t)(k - 1) * n + i]; maxClass = (double) k; }
      cum++;
      if (y[i] == maxClass) { correct += 1.0; cpv[(unsigned) y[i]] += 1.0; }
    }
  }
  if (cum == 0) { out[0] = out[1] = out[2] = na; return; }
  out[0] = 1.0 - correct / (double) cum;
  for (int k = 1; k <= 2; k++) out[k] = count[k] ? 1.0 - cpv[k] / (double) count[k] : na;
}

/* The other performance types of the class family and the RFQ vote (getPerformance rf.c:7954-8010):
   mode 0 = misclassification (as orc_perf_clas), 1 = Brier (getBrierScore rf.c:974-1024), 2 = g-mean (getGMeanIndex
   rf.c:1096-1140; per-class slots stay NA); rfq != 0: vote = minority class iff its probability >= its prevalence
   (getMaxVote rf.c:1181-1197, minority / majority / threshold rf.c:16859-16895). */
void orc_perf_clas2(int n, const double *y, const double *ens, const int *den, double na, int mode, int rfq, double *out) {
  unsigned count[3] = {0, 0, 0}, cum = 0;
  for (int i = 0; i < n; i++) count[(unsigned) y[i]]++;
  unsigned minority = 1, majority = 1;
  for (unsigned k = 1; k <= 2; k++) if (count[k] < count[minority]) minority = k;
  for (unsigned k = 1; k <= 2; k++) if (count[k] >= count[majority]) majority = k;
  double threshold = (double) count[minority] / (double) n;
  double *vote = malloc((size_t) n * sizeof(double));
  for (int i = 0; i < n; i++) {
    vote[i] = NAN;
    if (den[i] == 0) continue;
    cum++;
    if (rfq) vote[i] = (ens[(size_t)(minority - 1) * n + i] >= threshold) ? (double) minority : (double) majority;
    else {
      double maxValue = 0.0, maxClass = 0.0;
      for (int k = 1; k <= 2; k++)
        if (maxValue <= ens[(size_t)(k - 1) * n + i]) { maxValue = ens[(size_t)(k - 1) * n + i]; maxClass = (double) k; }
      vote[i] = maxClass;
    }
  }
  out[0] = out[1] = out[2] = na;
  if (cum == 0) { free(vote); return; }
  if (mode == 1) {
    double result = 0.0;
    for (int against = 1; against <= 2; against++) {
      double cpv = 0.0;
      for (int i = 0; i < n; i++)
        if (den[i] != 0) cpv += pow(((double) ((unsigned) y[i] == (unsigned) against) - ens[(size_t)(against - 1) * n + i]), 2.0);
      cpv /= (double) cum;
      out[against] = cpv; result += cpv;
    }
    out[0] = result * 2 / (2 - 1);
  } else if (mode == 2) {
    double tr[3] = {0, 0, 0}, fr[3] = {0, 0, 0}, result = 1.0;
    for (int i = 0; i < n; i++)
      if (den[i] > 0) { if (y[i] == vote[i]) tr[(unsigned) y[i]] += 1.0; else fr[(unsigned) y[i]] += 1.0; }
    for (int k = 1; k <= 2; k++) {
      double denom = tr[k] + fr[k];
      if (denom > 0) result = result * tr[k] / denom; else result = result * (1 + tr[k]) / (1 + denom);
    }
    out[0] = 1.0 - sqrt(result);
  } else {
    double correct = 0.0, cpv[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; i++)
      if (den[i] != 0 && y[i] == vote[i]) { correct += 1.0; cpv[(unsigned) y[i]] += 1.0; }
    out[0] = 1.0 - correct / (double) cum;
    for (int k = 1; k <= 2; k++) out[k] = count[k] ? 1.0 - cpv[k] / (double) count[k] : na;
  }
  free(vote);
}

/* ------------------------------------------------------------ permutation importance (rf.c:2106-2218, 2258-2697)
   Chain C (rf.c:6685-6692 grow; rf.c:6665 + 6685-6692 restore, where the seed LCG starts from 0 and chain A is
   skipped): per tree, in the order of `intr`, the OOB rows' values of the target covariate are permuted among
   themselves (permute(), rf.c:1974-2007: for i = m..1 the k-th still-empty slot receives i, k = ceil(ran1C * i)),
   the OOB rows are dropped down the tree again (identifyPerturbedMembership rf.c:1713-1783) and either
     leob = 1 ("permute"):          the tree's own OOB performance with and without the permutation is compared and the
                                    differences averaged over trees (x e for the per-class slots, rf.c:2607-2633), or
     leob = 0 ("permute.ensemble"): a perturbed OOB ensemble per covariate is accumulated over the trees and its
                                    performance compared with the forest's (rf.c:2685-2694). */
void orc_chain_seed_c(int seed, int ntree, int restore, int *seedC) {
  unsigned s = restore ? 0u : (unsigned) abs(seed);
  if (s >= 714025u) s %= 714025u;
  for (int pass = restore ? 1 : 0; pass < 3; pass++)
    for (int b = 0; b < ntree; b++) {
      s = lcg_next(s);
      s = lcg_next(s);
      while (s == 0) s = lcg_next(s);
      if (pass == 2) seedC[b] = -(int) s;
    }
}

static long vimp_walk(long root, const int *parmID, const double *contPT, const long *child, const int *mwcpSZ, const long *mwoff,
                      const unsigned *mwcpPT, const double *x, int n, int i, int target /* 1-based or 0 */, double shadow) {
  long q = root;
  while (parmID[q] != 0) {
    double xv = (parmID[q] == target) ? shadow : x[(size_t) (parmID[q] - 1) * n + i];
    if (mwcpSZ && mwcpSZ[q] > 0) q = ((mwcpPT[mwoff[q]] >> ((unsigned) xv - 1u)) & 1u) ? q + 1 : child[q];
    else q = (contPT[q] - xv >= 0.0) ? q + 1 : child[q];
  }
  return q;
}

void orc_vimp_permute(int ntree, int p, int n, long total_nodes, const int *treeID, const int *nodeID, const int *parmID,
                      const double *contPT, const int *leafCount, const int *tnCLAS /* [leaf][2] */, const int *mwcpSZ,
                      const unsigned *mwcpPT, const double *x /* [p][n] */, const double *y, const int *boot /* [ntree][n] */,
                      const int *seedC, int n_intr, const int *intr /* 1-based */, int leob, int perf_mode, int rfq, double na,
                      double *vimp /* [n_intr][3] */) {
  (void) p;
  long *mwoff = NULL;
  if (mwcpSZ) { mwoff = malloc(total_nodes * sizeof(long)); long at = 0; for (long q = 0; q < total_nodes; q++) { mwoff[q] = at; at += mwcpSZ[q]; } }
  long *child = malloc(total_nodes * sizeof(long)), *stack = malloc((total_nodes + 1) * sizeof(long)); long sp = 0;
  for (long q = 0; q < total_nodes; q++) {
    child[q] = -1;
    if (q > 0 && treeID[q] == treeID[q - 1] && parmID[q - 1] == 0) child[stack[--sp]] = q;
    if (q > 0 && treeID[q] != treeID[q - 1]) sp = 0;
    if (parmID[q] != 0) stack[sp++] = q;
  }
  double *ens = calloc((size_t) 2 * n, sizeof(double)); int *den = calloc(n, sizeof(int));
  double *vens = leob ? NULL : calloc((size_t) n_intr * 2 * n, sizeof(double)); int *vden = leob ? NULL : calloc((size_t) n_intr * n, sizeof(int));
  double *fens = calloc((size_t) 2 * n, sizeof(double)); int *fden = calloc(n, sizeof(int));      /* the forest's OOB ensemble */
  double *sum = calloc((size_t) n_intr * 3, sizeof(double)); int *cnt = calloc((size_t) n_intr * 3, sizeof(int));
  int *index = malloc((size_t) (n + 1) * sizeof(int)), *perm = malloc((size_t) (n + 2) * sizeof(int));
  double *shadow = malloc((size_t) n * sizeof(double));
  long root = 0, leafoff = 0;
  for (int b = 0; b < ntree; b++) {
    const int *bag = boot + (size_t) b * n;
    int m = 0;
    for (int i = 0; i < n; i++) if (bag[i] == 0) index[++m] = i;
    ran1_t chainC; ran1_seed(&chainC, seedC[b]);
    double perf0[3] = {na, na, na};
    if (m > 0) {
      /* the tree's own OOB estimate (summarizeTreePerformance rf.c:2834-2881) and its share of the forest ensemble */
      memset(ens, 0, (size_t) 2 * n * sizeof(double)); memset(den, 0, (size_t) n * sizeof(int));
      for (int k = 1; k <= m; k++) {
        int i = index[k];
        long q = vimp_walk(root, parmID, contPT, child, mwcpSZ, mwoff, mwcpPT, x, n, i, 0, 0.0);
        const int *c = tnCLAS + 2 * (leafoff + nodeID[q] - 1);
        double rc = (double) c[0] + (double) c[1];
        ens[i] = (double) c[0] / rc; ens[(size_t) n + i] = (double) c[1] / rc; den[i] = 1;
        fens[i] += ens[i]; fens[(size_t) n + i] += ens[(size_t) n + i]; fden[i]++;
      }
      if (leob) orc_perf_clas2(n, y, ens, den, na, perf_mode, rfq, perf0);
      for (int v = 0; v < n_intr; v++) {
        const int target = intr[v];
        for (int k = 1; k <= m; k++) perm[k] = 0;
        for (int i = m; i > 0; i--) {                                   /* permute(3, b, m, perm) */
          unsigned k = (unsigned) ceil(orc_ran1(&chainC) * (i * 1.0));
          int j;
          for (j = 1; k > 0; j++) if (perm[j] == 0) k--;
          perm[j - 1] = i;
        }
        for (int k = 1; k <= m; k++) shadow[index[k]] = x[(size_t) (target - 1) * n + index[perm[k]]];
        if (leob) { memset(ens, 0, (size_t) 2 * n * sizeof(double)); }
        for (int k = 1; k <= m; k++) {
          int i = index[k];
          long q = vimp_walk(root, parmID, contPT, child, mwcpSZ, mwoff, mwcpPT, x, n, i, target, shadow[i]);
          const int *c = tnCLAS + 2 * (leafoff + nodeID[q] - 1);
          double rc = (double) c[0] + (double) c[1];
          if (leob) { ens[i] = (double) c[0] / rc; ens[(size_t) n + i] = (double) c[1] / rc; }
          else { vens[((size_t) v * 2) * n + i] += (double) c[0] / rc; vens[((size_t) v * 2 + 1) * n + i] += (double) c[1] / rc; vden[(size_t) v * n + i]++; }
        }
        if (leob) {
          double pv[3]; orc_perf_clas2(n, y, ens, den, na, perf_mode, rfq, pv);
          for (int k = 0; k < 3; k++)
            if (!isnan(pv[k]) && !isnan(perf0[k])) { sum[v * 3 + k] += pv[k] - perf0[k]; cnt[v * 3 + k]++; }
        }
      }
    }
    root += 2 * (long) leafCount[b] - 1; leafoff += leafCount[b];
  }
  if (leob) {
    for (int v = 0; v < n_intr; v++) for (int k = 0; k < 3; k++)
      vimp[v * 3 + k] = cnt[v * 3 + k] ? (k > 0 ? M_E * sum[v * 3 + k] / (double) cnt[v * 3 + k] : sum[v * 3 + k] / (double) cnt[v * 3 + k]) : na;
  } else {
    for (int i = 0; i < n; i++) if (fden[i]) { fens[i] /= fden[i]; fens[(size_t) n + i] /= fden[i]; }
    double perf0[3]; orc_perf_clas2(n, y, fens, fden, na, perf_mode, rfq, perf0);
    for (int v = 0; v < n_intr; v++) {
      double *e = vens + (size_t) v * 2 * n; int *d = vden + (size_t) v * n;
      for (int i = 0; i < n; i++) if (d[i]) { e[i] /= d[i]; e[(size_t) n + i] /= d[i]; } else { e[i] = e[(size_t) n + i] = na; }
      double pv[3]; orc_perf_clas2(n, y, e, d, na, perf_mode, rfq, pv);
      for (int k = 0; k < 3; k++) vimp[v * 3 + k] = (!isnan(pv[k]) && !isnan(perf0[k])) ? pv[k] - perf0[k] : pv[k];
    }
  }
  free(mwoff); free(child); free(stack); free(ens); free(den); free(vens); free(vden); free(fens); free(fden);
  free(sum); free(cnt); free(index); free(perm); free(shadow);
}

/* bootstrap-only entry (chain A replay) for the in-bag parity tests */
void orc_bootstrap(int seed, int ntree, int n, int tree /* 0-based */, int *draws /* [n], 1-based row ids */) {
  int *seedA = malloc(ntree * sizeof(int)), *seedB = malloc(ntree * sizeof(int));
  orc_chain_seeds(seed, ntree, seedA, seedB);
  ran1_t s; ran1_seed(&s, seedA[tree]);
  for (int i = 0; i < n; i++) draws[i] = (int) ceil((double) orc_ran1(&s) * (n * 1.0));
  free(seedA); free(seedB);
}

/* first `count` floats of a ran1 chain started from `seed` (unit test of the generator) */
void orc_ran1_stream(int seed, int count, float *out) {
  ran1_t s; ran1_seed(&s, seed);
  for (int i = 0; i < count; i++) out[i] = orc_ran1(&s);
}
