/* rfslam_oracle.h -- TEST INFRASTRUCTURE ONLY: interface of the CPU restatement (see rfslam_oracle.c). */
#ifndef RFSLAM_ORACLE_H
#define RFSLAM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int n, p, ntree, mtry, nodesize, nodedepth, rule, seed, max_boot;
  double k_for_alpha;
  const double *x;            /* [p][n] */
  const double *y;            /* [n] in {1,2} */
  const double *rt;           /* [n] */
  const int    *strat;        /* [n] >= 1 */
  const int    *samp;         /* [ntree][n] in-bag counts (bootstrap = by.user) or NULL */
  const int    *boot_size;    /* [ntree] or NULL (= n) */
  const double *split_weight; /* [p] or NULL */
  const int    *tree_seeds;   /* [ntree] chain-A seeds (restore) or NULL (grow) */
  int           nsplit;       /* 0 = every distinct value but the last is a cut; > 0: at most nsplit random cuts per covariate */
  const char   *xtype;        /* [p] 'R' ordered / 'C' unordered factor (levels 1..xlevels[c], at most 32), or NULL (all 'R') */
  const int    *xlevels;      /* [p] factor sizes (0 for 'R') or NULL */
} orc_args;

typedef struct {
  int ntree, n;
  long total_nodes, total_leaves, candidates, rmbr_stride;
  int *treeID, *nodeID, *parmID; double *contPT, *splitStat;      /* pre-order, tree order */
  int *mwcpSZ; unsigned *mwcpPT; int *mwcpCT; long mwcp_total;     /* factor splits: words per node, the words, words per tree */
  int *leafCount, *seed;
  int *bootMembership, *nodeMembership;                            /* [ntree][n] */
  int *tnRCNT, *tnACNT, *tnCLAS;                                   /* per tree, leafCount[b] entries */
  int *rmbr, *ambr;                                                /* [ntree][stride], [ntree][n] */
  double *oobNum, *allNum; int *oobDen, *allDen;
  double *oobEns, *allEns;                                         /* [2][n] class-major */
} orc_out;

orc_out *orc_grow(const orc_args *a);
void     orc_free(orc_out *o);
void     orc_predict(int ntree, int p, long total_nodes, const int *treeID, const int *nodeID, const int *parmID,
                     const double *contPT, const int *leafCount, const int *tnRCNT, const int *tnE,
                     const double *fx, int m, double *ens, int *membership, const int *mwcpSZ, const unsigned *mwcpPT);
double   orc_poisson_stat(int rule, int n, double k_for_alpha, const char *membership,
                          const double *response, const double *strat, const double *rt);
void     orc_chain_seeds(int seed, int ntree, int *seedA, int *seedB);
void     orc_chain_seeds_restore(int ntree, int *seedB);
void     orc_bootstrap(int seed, int ntree, int n, int tree, int *draws);
void     orc_ran1_stream(int seed, int count, float *out);
void     orc_perf_clas(int n, const double *y, const double *ens, const int *den, double na, double *out);
void     orc_chain_seed_c(int seed, int ntree, int restore, int *seedC);
void     orc_vimp_permute(int ntree, int p, int n, long total_nodes, const int *treeID, const int *nodeID, const int *parmID,
                          const double *contPT, const int *leafCount, const int *tnCLAS, const int *mwcpSZ, const unsigned *mwcpPT,
                          const double *x, const double *y, const int *boot, const int *seedC, int n_intr, const int *intr,
                          int leob, int perf_mode, int rfq, double na, double *vimp);
void     orc_perf_clas2(int n, const double *y, const double *ens, const int *den, double na, int mode, int rfq, double *out);

#ifdef __cplusplus
}
#endif
#endif
