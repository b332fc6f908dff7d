/*
 * rfslam_b200.h -- C-ABI of the B200-native rfSLAM hot path (librfslam_b200.so).
 *
 * This is the drop-in boundary: plain C, raw pointers and sizes, no torch
 * or CUDA types.  Every entry point names the reference interface it replaces; paths are
 * relative to /root/reference/src (rf.c = randomForestSRC.c, sc.c = splitCustom.c).
 *
 *   rfslam_b200_grow      <->  SEXP rfsrcGrow(44 x SEXP)      rf.c:22110-22217
 *   rfslam_b200_predict   <->  SEXP rfsrcPredict(58 x SEXP)   rf.c:22219-22402
 *   rfslam_b200_rule_of_slot / rfslam_b200_register_this
 *                         <->  registerCustomFunctions / registerThis  sc.c:23-67, rf.c:8856-8866
 *   rfslam_b200_score_node<->  one covariate's worth of customFunction calls
 *                              (rf.c:13404-13529 -> sc.c poissonSplit1..6)
 *
 * All host arrays are caller-owned and read-only for the duration of the call (as R vectors are,
 * rf.c:22169-22204); outputs are allocated by the callee and released with rfslam_b200_free_*().
 * Every function returns 0 on success, non-zero on error; rfslam_b200_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every compute entry fails with RFSLAM_ENODEV.
 */
#ifndef RFSLAM_B200_H
#define RFSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFSLAM_OK        0
#define RFSLAM_EINVAL    1   /* bad argument (the reference would call RF_nativeExit) */
#define RFSLAM_ENODEV    2   /* no CUDA device / CUDA failure */
#define RFSLAM_ENOSUP    3   /* valid for the reference but outside the accelerated path */
#define RFSLAM_ENOMEM    4

/* option bits the path understands; same values as the reference (rf.h:108-158) */
#define RFSLAM_OPT_PERF        0x00000004u   /* optLow: performance of the OOB ensemble (rf.h:110) */
#define RFSLAM_OPT_PERF_CALB   0x00000008u   /* optLow: perf.type "brier" (rf.h:111, R/utilities.R:267-269) */
#define RFSLAM_OPT_PERF_GMN2   0x00004000u   /* optLow: perf.type "g.mean" (rf.h:122) */
#define RFSLAM_OPT_VIMP_TYP1   0x00000100u   /* optLow: importance by permutation (rf.h:116); TYP2 = random daughters, neither = anti */
#define RFSLAM_OPT_VIMP_TYP2   0x00000200u   /* optLow (rf.h:117): not on the accelerated path */
#define RFSLAM_OPT_VIMP_JOIN   0x00000400u   /* optLow: joint importance (rf.h:118): not on the accelerated path */
#define RFSLAM_OPT_VIMP_LEOB   0x01000000u   /* optLow: per-tree (Breiman-Cutler) importance (rf.h:131); clear = ensemble importance */
#define RFSLAM_OPT_VIMP        0x02000000u   /* optLow: importance requested (rf.h:132); needs RFSLAM_OPT_PERF (rf.c:1282-1286) */
#define RFSLAM_OPT_CLAS_RFQ    0x00008000u   /* optLow: rfq = TRUE, quantile-classifier vote (rf.h:123, rf.c:1181-1197) */
#define RFSLAM_OPT_BOOT_TYP1   0x00080000u   /* optLow  (by.user = TYP1|TYP2, rf.c:520) */
#define RFSLAM_OPT_BOOT_TYP2   0x00100000u
#define RFSLAM_OPTH_MEMB_USER  0x00000040u   /* optHigh: return nodeMembership / bootMembership */
#define RFSLAM_OPTH_SPLT_CUST  0x00000F00u   /* optHigh: custom split slot - 1 (rf.h:147) */
#define RFSLAM_OPTH_MEMB_OUTG  0x00010000u   /* optHigh: return rmbr/ambr/tnRCNT/tnACNT */

/* ---------------------------------------------------------------- grow ------------------------ */

/* Mirrors the 44 .Call arguments of rfsrcGrow in their order (rf.c:22110-22155); scalars by value,
   vectors as pointers.  Fields the Poisson class path never reads are kept so that a binding can
   forward every argument; they are validated, not ignored (see rfslam_b200_grow). */
struct rfslam_comm;
typedef struct rfslam_grow_args {
  int32_t        trace_flag;
  int32_t        seed;                /* must be < 0 (rf.c:6454) */
  uint32_t       opt_low;
  uint32_t       opt_high;
  int32_t        split_rule;          /* 16 = custom (rf.h:189) */
  int32_t        nsplit;              /* 0 = every distinct value but the last is a cut; > 0 = at most nsplit random cuts per covariate
                                         (rf.c:14533-14564; accelerated for covariates of <= 256 distinct values) */
  int32_t        mtry;
  int32_t        htry;                /* must be 0 (rf.c:6490-6494) */
  int32_t        ytry;
  int32_t        node_size;
  int32_t        node_depth;          /* < 0 = unlimited */
  int32_t        cr_weight_size;
  const double  *cr_weight;
  int32_t        ntree;
  int32_t        observation_size;    /* n */
  int32_t        y_size;              /* 1 */
  const char    *r_type;              /* y_size chars; 'C' / 'I' / 'B' */
  const int32_t *r_levels;
  const double  *r_data;              /* [y_size][n] response codes 1.0 / 2.0 */
  int32_t        x_size;              /* p */
  const char    *x_type;              /* p chars; 'R' real, 'C' unordered factor (ENOSUP) */
  const int32_t *x_levels;
  const double  *risk_time;           /* [n] */
  const int32_t *stratifier;          /* [n] >= 1 */
  int32_t        max_strat;
  const int32_t *var_categories;      /* [p] */
  const double  *category_weights;    /* [n_categories] */
  int32_t        n_categories;
  const int32_t *to_fix;
  int32_t        n_to_fix;
  double         k_for_alpha;
  int32_t        max_bootstrap_size;
  const int32_t *bootstrap_size;      /* [ntree] */
  const int32_t *bootstrap;           /* [ntree][n] in-bag counts when by.user, else unused */
  const double  *case_weight;         /* [n] must be uniform */
  const double  *x_split_stat_weight; /* [p] */
  const double  *y_weight;            /* [y_size] */
  const double  *x_weight;            /* [p] must be uniform */
  const double  *x_data;              /* [p][n]: p contiguous columns of n (rf.c:22190, rf.c:22469) */
  int32_t        time_interest_size;
  const double  *time_interest;
  int32_t        n_impute;
  int32_t        perf_block;
  int32_t        num_threads;         /* ignored: trees are scheduled on the GPU */

  /* --- build-specific knobs (no reference counterpart) --- */
  int32_t        device;              /* CUDA device ordinal; < 0 = current */
  int32_t        tree_first;          /* this process grows trees tree_first, tree_first+tree_step, ... */
  int32_t        tree_step;           /*   (1-based ids; 0/0 = all).  Seeds are derived for all ntree */
                                      /*   trees on every rank, so tree b is identical for any sharding. */
  int32_t        finalize_ensembles;  /* 1: divide by the denominators (single process); 0: return sums */
  int32_t        want_split_stat;     /* 1: also return splitStat (the winning delta per node) */
  int32_t        want_member_lists;   /* 1 (with RFSLAM_OPTH_MEMB_OUTG): also return rmbr/ambrMembership, 8 bytes x n x ntree */
  struct rfslam_comm *comm;           /* tree-sharded forest over several GPUs (rfslam_b200_comm_init): the ensemble sums are
                                         all-reduced ON THE DEVICE before they are finalized and sent home, so every rank returns the
                                         whole forest's oob/allEnsbCLS (and perfClas); NULL = this process is the whole job */
  int32_t        gather_forest;       /* with comm: 1 = gather the shards' forest records to rank 0 over NCCL and return the whole
                                         forest there in ascending treeID order (other ranks return no records); 0 = records stay sharded */
} rfslam_grow_args;

/* The reference's named output list for the class family (rf.c:9-79), sized by the
   ACTUAL node counts (R trims the reference's theoretical-maximum arrays to exactly this,
   R/rfsrc.R:1766-1861).  Trees appear in ascending treeID order, records in pre-order. */
typedef struct rfslam_forest_out {
  int32_t   ntree;             /* trees grown by this call (shard) */
  int32_t   n;
  int64_t   total_nodes;       /* sum over trees of 2*leafCount-1 */
  int64_t   total_leaves;
  int32_t  *tree_ids;          /* [ntree] which tree ids this shard holds */
  int32_t  *treeID;            /* [total_nodes] */
  int32_t  *nodeID;            /* [total_nodes] */
  int32_t  *parmID;            /* [total_nodes] 0 = terminal */
  double   *contPT;            /* [total_nodes] NA_REAL at terminals */
  int32_t  *mwcpSZ;            /* [total_nodes] words of the factor split of the node, 0 for a numeric split or a leaf (rf.c:10773) */
  double   *splitStat;         /* [total_nodes] winning delta (NaN at terminals) when want_split_stat, else NULL */
  int32_t  *leafCount;         /* [ntree] */
  int32_t  *seed;              /* [ntree] chain-A seed, as forest$seed */
  int32_t  *tnRCNT;            /* [total_leaves] in-bag count per leaf, by tree then leaf id */
  int32_t  *tnACNT;            /* [total_leaves] all-row count per leaf */
  int32_t  *tnCLAS;            /* [total_leaves][2] in-bag class counts (non-event, event) */
  int32_t  *nodeMembership;    /* [ntree][n] or NULL */
  int32_t  *bootMembership;    /* [ntree][n] or NULL */
  int32_t  *rmbrMembership;    /* [ntree][max_bootstrap_size] or NULL */
  int32_t  *ambrMembership;    /* [ntree][n] or NULL */
  double   *oobEnsbCLS;        /* [2][n] class-major: sums or ratios, see finalize_ensembles */
  double   *allEnsbCLS;        /* [2][n] */
  uint32_t *oobEnsbDen;        /* [n] */
  uint32_t *allEnsbDen;        /* [n] */
  /* measurement */
  int64_t   split_candidates;  /* (node, covariate, cutpoint) triples scored */
  int64_t   algorithmic_bytes; /* sum over attempted nodes of B_score + B_part + 4n per tree = the three below */
  double    grow_kernel_ms;    /* device time of the tree-growing kernel(s), CUDA events */
  double    total_device_ms;   /* device time of the whole call incl. bootstrap and ensembles */
  int64_t   kernel_launches;   /* kernels of this library launched by the call */
  int64_t   rmbr_stride;       /* row stride of rmbrMembership (= maxBootstrapSize, rf.c:18296) */
  /* perfClas (rf.c:25, rf.c:17454): [ntree][1 + 2] OOB misclassification rate of the forest and per class
     (getClassificationIndex rf.c:1067, getConditionalClassificationIndex rf.c:1025); with RFSLAM_OPT_PERF_CALB the
     normalised Brier score and its per-class terms (getBrierScore rf.c:974), with RFSLAM_OPT_PERF_GMN2 the g-mean
     index alone (getGMeanIndex rf.c:1096); RFSLAM_OPT_CLAS_RFQ switches the vote (getMaxVote rf.c:1175).  Present when opt_low has
     RFSLAM_OPT_PERF and the call finalizes the ensembles of the whole forest; only the last row is filled
     (perfBlock = ntree, rf.c:8155), the others hold NA_REAL as in the reference. */
  double   *perfClas;
  /* factor (MWCP) splits (rf.c:1416-1616, rf.c:10773-10788): the left-daughter level sets of all factor splits,
     32 levels per word, words of a node in ascending order, nodes in the forest's record order */
  int32_t  *mwcpPT;            /* [mwcp_total] or NULL when the forest has no factor split */
  int32_t  *mwcpCT;            /* [ntree] words per tree */
  int64_t   mwcp_total;
  /* algorithmic bytes and device time kernel by kernel: the tree kernel owns B_score and the in-bag
     half of B_part, the traversal kernel the all-row half (16 B per row per split node on its path), the
     bootstrap replay the 4 B per draw */
  int64_t   grow_kernel_algorithmic_bytes;
  int64_t   membership_algorithmic_bytes;
  int64_t   bootstrap_algorithmic_bytes;
  double    membership_kernel_ms;
  double    bootstrap_ms;
  double    collective_ms;     /* all-reduce (+ gather) through comm, host clock around the device work */
  /* vimpClas (rf.c:37, rf.c:17537): [xSize][1 + 2] permutation importance of every covariate (processDefaultGrow:
     intrPredictorSize = xSize, rf.c:1220) with RFSLAM_OPT_PERF | RFSLAM_OPT_VIMP | RFSLAM_OPT_VIMP_TYP1 on a whole forest
     (no tree sharding); per tree with RFSLAM_OPT_VIMP_LEOB, else from perturbed ensembles; NULL otherwise */
  double   *vimpClas;
  int32_t   vimp_count;
  double    vimp_ms;           /* device time of the importance pass (not part of total_device_ms) */
} rfslam_forest_out;

int  rfslam_b200_grow(const rfslam_grow_args *args, rfslam_forest_out *out);
void rfslam_b200_free_forest(rfslam_forest_out *out);
/* Output arrays are page-locked blocks recycled across calls (the R stub copies them into R vectors and
   frees the struct; a binding that adopts single arrays instead releases each with free_array, never
   with free()).  host_pool_trim returns the idle blocks to the OS (cap: RFSLAM_HOST_POOL_MB, default 8192) and
   frees the device scratch rfslam_b200_grow keeps between calls. */
void rfslam_b200_free_array(void *array);
void rfslam_b200_host_pool_trim(void);

/* ---------------------------------------------------------------- multi-GPU ------------------- */
/* One process per GPU; rank r grows trees r+1, r+1+world, ... (tree_first = rank + 1, tree_step = world) of the same
   ntree-tree forest with a full replica of the table -- nothing is exchanged while trees grow (rf.c:7082-7087).  The
   communicator (NCCL, bound at run time) carries the one exchange after growth: the all-reduce of the ensemble sums
   (rf.c:944-949) and, on request, the gather of the forest records to rank 0 (rf.c:20938-20946).  The 128-byte id
   made by rank 0 travels to the other ranks by whatever the host program has (torch.distributed, MPI, a file). */
typedef struct rfslam_comm rfslam_comm;
#define RFSLAM_COMM_ID_BYTES 128
int  rfslam_b200_comm_unique_id(void *id, int bytes);
int  rfslam_b200_comm_init(rfslam_comm **comm, int rank, int world, const void *id, int device);
void rfslam_b200_comm_destroy(rfslam_comm *comm);
int  rfslam_b200_comm_rank(const rfslam_comm *comm);
int  rfslam_b200_comm_world(const rfslam_comm *comm);
/* collectives issued so far through this communicator (returns their sum) */
int64_t rfslam_b200_comm_collectives(const rfslam_comm *comm, int64_t *allreduces, int64_t *gathers);
int  rfslam_b200_nccl_version(void);

/* Resident variant: upload + rank-code the CPIU table once, grow many forests from HBM.
   rfslam_b200_grow() == dataset_create + grow_resident + dataset_destroy. */
typedef struct rfslam_dataset rfslam_dataset;
int  rfslam_b200_dataset_create(const rfslam_grow_args *args, rfslam_dataset **ds);
int  rfslam_b200_grow_resident(rfslam_dataset *ds, const rfslam_grow_args *args, rfslam_forest_out *out);
void rfslam_b200_dataset_destroy(rfslam_dataset *ds);
/* bytes of HBM held by the resident table (codes + per-row arrays) */
int64_t rfslam_b200_dataset_bytes(const rfslam_dataset *ds);

/* ---------------------------------------------------------------- predict --------------------- */

/* Mirrors the 58 .Call arguments of rfsrcPredict in their order (rf.c:22219-22279, R side
   R/generic.predict.rfsrc.R:437-532); list arguments are flattened (hc_zero = parmID / contPT / mwcpSZ / mwcpPT;
   hc_one .. hc_mwcpPT exist only for htry > 0, which the path refuses; partial = its seven elements).
   Mode as in the reference (rf.c:22361): fobservation_size > 0 predicts the held-out rows fx_data, otherwise the
   call RESTORES the forest on its training rows x_data (OOB and full ensembles, perfClas).
   Leaf class counts: taken from tnCLAS when it is passed (OPT_TERM_INCG, per tree leafCount x 2 as R stores it);
   otherwise rebuilt the way restoreNodeMembership does (rf.c:3446-3900, rf.c:755-858): the bootstrap is replayed
   from forest$seed (rf.c:7046-7050) or taken from `bootstrap` (by.user) and the training rows are routed down
   every tree. */
typedef struct rfslam_predict_args {
  int32_t        trace_flag;
  int32_t        seed;
  uint32_t       opt_low;
  uint32_t       opt_high;
  int32_t        ntree;
  int32_t        observation_size;  /* n training rows */
  int32_t        y_size;
  const char    *r_type;
  const int32_t *r_levels;
  const double  *r_data;            /* [n] training response codes 1.0 / 2.0 */
  int32_t        x_size;            /* p */
  const char    *x_type;
  const int32_t *x_levels;
  const double  *x_data;            /* [p][n] training covariates */
  double         k_for_alpha;
  int32_t        max_bootstrap_size;
  const int32_t *bootstrap_size;    /* [ntree] */
  const int32_t *bootstrap;         /* [ntree][n] when by.user */
  const double  *case_weight;
  int32_t        time_interest_size;
  const double  *time_interest;
  int32_t        total_node_count;  /* records in the forest arrays */
  const int32_t *forest_seed;       /* [ntree] forest$seed: chain-A seeds (rf.c:7046-7050) */
  int32_t        htry;              /* must be 0 */
  const int32_t *treeID;            /* [total_node_count] records of one tree contiguous, pre-order */
  const int32_t *nodeID;
  const int32_t *parmID;            /* hc_zero[[1]] */
  const double  *contPT;            /* hc_zero[[2]] */
  const int32_t *mwcpSZ;            /* hc_zero[[3]] */
  const int32_t *mwcpPT;            /* hc_zero[[4]] or NULL */
  const int32_t *tnRMBR;            /* forest$nativeArrayTNDS, as returned by grow */
  const int32_t *tnAMBR;
  const int32_t *tnRCNT;            /* concatenated per tree id, leafCount[b] entries each */
  const int32_t *tnACNT;
  const double  *tnSURV, *tnMORT, *tnNLSN, *tnCSHZ, *tnCIFN, *tnREGR;   /* other families: unused */
  const int32_t *tnCLAS;            /* concatenated per tree id, leafCount[b] x 2 (non-event, event) */
  const int32_t *r_target;
  int32_t        r_target_count;
  int32_t        ptn_count;         /* must be 0 (pruning) */
  int32_t        intr_predictor_size;
  const int32_t *intr_predictor;
  int32_t        partial_type, partial_xvar, partial_length;   /* partial[[1..3]]; partial plots are outside the path */
  int32_t        sobservation_size;
  const int32_t *sobservation_indv;
  int32_t        fobservation_size; /* m held-out rows; 0 = restore mode */
  int32_t        fr_size;
  const double  *fr_data;
  const double  *fx_data;           /* [p][m] */
  int32_t        perf_block;
  int32_t        num_threads;       /* ignored */

  /* --- build-specific knobs (no reference counterpart) --- */
  int64_t        tnCLAS_len;        /* elements in tnCLAS (a SEXP carries its length, a pointer does not); 0 = absent */
  int64_t        mwcp_total;        /* elements in mwcpPT */
  int32_t        device;            /* CUDA device ordinal; < 0 = current */
  int32_t        finalize_ensembles;/* 1: divide by the denominators; 0: return sums (tree-sharded forests) */
  int32_t        frow_first;        /* row sharding: this process handles held-out rows [frow_first, frow_first + frow_count) */
  int32_t        frow_count;        /*   of fx_data (0 = all); outputs are sized by the shard.  Forest replicated, no collective. */
} rfslam_predict_args;

typedef struct rfslam_predict_out {
  int32_t   m;                 /* rows in the outputs: the shard's held-out rows, or n in restore mode */
  int32_t   ntree;
  double   *allEnsbCLS;        /* [2][m] */
  uint32_t *allEnsbDen;        /* [m] */
  double   *oobEnsbCLS;        /* [2][n] restore mode only */
  uint32_t *oobEnsbDen;        /* [n] restore mode only */
  double   *perfClas;          /* [ntree][3] with OPT_PERF, as rfslam_forest_out.perfClas: restore mode (OOB ensemble of the training
                                  rows) or held-out rows with responses (fr_size > 0: the full ensemble against fr_data, whole row set only) */
  int32_t  *leafCount;         /* [ntree] by tree id */
  int32_t  *nodeMembership;    /* [ntree][m] by tree id, with OPT_MEMB_USER, else NULL */
  int32_t  *bootMembership;    /* [ntree][n] in-bag counts of the training rows (replayed), with OPT_MEMB_USER */
  double    kernel_ms;         /* device time of the routing kernels (CUDA events) */
  double    total_device_ms;   /* device time of the whole call incl. uploads */
  int64_t   kernel_launches;
  int64_t   row_tree_visits;   /* (row, tree) pairs routed */
  int64_t   node_visits;       /* records read while routing (sum of path lengths) */
  /* vimpClas (rf.c:37, rf.c:17537): [vimp_count = intrPredictorSize][1 + 2] permutation importance of the covariates in
     intr_predictor, in restore mode (over the OOB rows) or on held-out rows that carry the outcome (fr_size > 0, whole row
     set), with RFSLAM_OPT_PERF | RFSLAM_OPT_VIMP | RFSLAM_OPT_VIMP_TYP1; per tree (Breiman-Cutler)
     with RFSLAM_OPT_VIMP_LEOB, else from perturbed ensembles; NULL otherwise */
  double   *vimpClas;
  int32_t   vimp_count;
  double    vimp_ms;           /* device time of the importance pass */
} rfslam_predict_out;

int  rfslam_b200_predict(const rfslam_predict_args *args, rfslam_predict_out *out);
void rfslam_b200_free_predict(rfslam_predict_out *out);

/* ---------------------------------------------------------------- split-rule registry --------- */

/* The reference selects a rule by family + slot: registerThis(func, family, slot) (rf.c:8856-8866),
   slot = ((optHigh & 0x0F00) >> 8) + 1 (rf.c:18457-18504); "poisson.splitN" is CLAS_FAM slot N+1
   (sc.c:47-62).  Here the six Poisson rules are device kernels; any other slot has no GPU
   implementation and is refused (no CPU fallback). */
#define RFSLAM_CLAS_FAM 0
#define RFSLAM_REGR_FAM 1
#define RFSLAM_SURV_FAM 2
#define RFSLAM_CRSK_FAM 3
/* rule 1..6 for a CLAS_FAM slot 2..7, 0 when the slot has no device rule */
int  rfslam_b200_rule_of_slot(int family, int slot);
/* re-map a slot to one of the six device rules (what registerThis does for function pointers) */
int  rfslam_b200_register_this(int rule, int family, int slot);
void rfslam_b200_register_defaults(void);

/* The reference's own plugin symbols, with its exact signatures (src/splitCustom.h:24-44, :246-330;
   src/splitCustom.c:23-67; src/randomForestSRC.c:8856-8866), so that code written against splitCustom.h
   links against this library unchanged.  Arrays are 1-BASED as in the reference (rf.c:13455-13482):
   membership[1..n] (LEFT = 1), response[1..n] in {1, 2}, feature[1][1..n] = stratum, feature[2][1..n] = risk
   time; time / event / eventTime are unused (NULL).  poissonSplitN evaluates the statistic of that ONE cut on
   the device.  registerThis maps a slot to the device rule when func is one of the six symbols below; any
   other pointer is a CPU rule with no device implementation: the slot is marked foreign and a grow call that
   selects it fails with RFSLAM_ENOSUP. */
typedef double (*rfslam_custom_function)(unsigned int n, double k_for_alpha, char *membership, double *time, double *event,
                                         unsigned int eventTypeSize, unsigned int eventTimeSize, double *eventTime,
                                         double *response, double mean, double variance, unsigned int maxLevel,
                                         double **feature, unsigned int featureCount);
#define RFSLAM_DECLARE_PLUGIN(name)                                                                              \
  double name(unsigned int n, double k_for_alpha, char *membership, double *time, double *event,                \
              unsigned int eventTypeSize, unsigned int eventTimeSize, double *eventTime, double *response,      \
              double mean, double variance, unsigned int maxLevel, double **feature, unsigned int featureCount)
RFSLAM_DECLARE_PLUGIN(poissonSplit1);   /* sc.c:639-783   Bayes, stratified */
RFSLAM_DECLARE_PLUGIN(poissonSplit2);   /* sc.c:814-943   Bayes, unstratified */
RFSLAM_DECLARE_PLUGIN(poissonSplit3);   /* sc.c:975-1120  Bayes, stratified, unit exposure */
RFSLAM_DECLARE_PLUGIN(poissonSplit4);   /* sc.c:1151-1314 raw rate, stratified */
RFSLAM_DECLARE_PLUGIN(poissonSplit5);   /* sc.c:1346-1491 raw rate, unstratified */
RFSLAM_DECLARE_PLUGIN(poissonSplit6);   /* sc.c:1524-1687 raw rate, stratified, unit exposure */
void registerThis(rfslam_custom_function func, unsigned int family, unsigned int slot);   /* rf.c:8856-8866 */
void registerCustomFunctions(void);                                                        /* sc.c:23-67 */

/* Score every cutpoint of ONE covariate in ONE node on the device: the work of the reference's
   per-cut loop rf.c:13404-13529 (virtuallySplitNode + gather + plugin call).
   In: m rows of the node: x (the covariate), response codes (1/2), stratum (>= 1), risk time.
   Out: the distinct values in ascending order and delta for the cut after each of them except the
   last (cut_value[k], delta[k], k < *n_cuts), exactly what the plugin returns for that membership. */
int  rfslam_b200_score_node(int rule, double k_for_alpha, int m, const double *x, const double *response,
                            const int32_t *stratum, const double *risk_time,
                            double *cut_value /* [m] */, double *delta /* [m] */, int32_t *n_cuts,
                            int device);

/* Chain-A replay only (rf.c:529-533): in-bag counts of tree `tree` (1-based) of an ntree forest. */
int  rfslam_b200_bootstrap_counts(int seed, int ntree, int n, int tree, int32_t *counts /* [n] */, int device);

/* Diagnostics: with RFSLAM_PROFILE set in the environment the tree kernel accumulates clock64()
   cycles per phase (slots 0..11), per node-size class log2(entries) (slots 16..47) and node counts
   per class pair (slots 48..63), summed over the trees of the last grow call. Returns slots written. */
int  rfslam_b200_last_profile(uint64_t *slots, int capacity);

const char *rfslam_b200_last_error(void);
const char *rfslam_b200_version(void);
int  rfslam_b200_device_count(void);

/* ---------------------------------------------------------------- CPIU construction ---------- */
/* The step before the forest path: `cpiu.fc` / `event` of utilities/cpiu.R:361-668 and the Rcpp helpers of
   utilities/cpiu.cpp:27-249, for all persons at once.  A person's follow-up is cut into ceil(t_outcome / (interval *
   unit_norm)) counting-process information units; rows of person 0 come first. */
#define RFSLAM_CPIU_INDICATOR   0   /* i.*  does an event fall in (t.start, t.end]?      assemble_indicator    cpiu.cpp:27-50   */
#define RFSLAM_CPIU_COUNT_PREV  1   /* n.*  events in PREVIOUS intervals                   assemble_accumulator  cpiu.cpp:75-113 flag 1 */
#define RFSLAM_CPIU_COUNT_CUR   2   /* ni.* events in THIS interval                        assemble_accumulator  flag 0 */
#define RFSLAM_CPIU_ELAPSED     3   /* t.*  time since the last event                      assemble_time_elapsed cpiu.cpp:129-165 */
#define RFSLAM_CPIU_CONTINUOUS  4   /* c.*  carried-forward measurement (k.to.persist)     assemble_continuous   cpiu.cpp:186-249 */

typedef struct {
  int32_t        persons;
  const double  *t_outcome;          /* [persons] time of the outcome / end of follow-up, in unit_norm units (cpiu.R:476) */
  const double  *t_death;            /* [persons] (cpiu.R:575) */
  const int32_t *death;              /* [persons] death indicator */
  double         interval;           /* interval length in years (default 0.5) */
  double         unit_norm;          /* time units per year (default 365): interval.mod = interval * unit_norm (cpiu.R:472) */
  int32_t        n_vars;             /* derived variables */
  const int32_t *var_kind;           /* [n_vars] RFSLAM_CPIU_* */
  const int32_t *var_width;          /* [n_vars] event / measurement slots per person (<= 64) */
  const double *const *var_times;    /* [n_vars] -> [persons][width]; NaN = missing slot (never in an interval, cpiu.R:583-584) */
  const double *const *var_values;   /* [n_vars] -> [persons][width] for CONTINUOUS (sorted by time first, cpiu.R:488-494), else NULL */
  int32_t        k_to_persist;       /* intervals a measurement is carried forward; -1 = for ever (cpiu.cpp:214-216) */
  int32_t        device;
} rfslam_cpiu_args;

typedef struct {
  int64_t  rows;                     /* CPIUs = sum over persons of their intervals */
  int32_t  n_vars;
  int32_t *person;                   /* [rows] 0-based person */
  int32_t *int_n;                    /* [rows] interval number 1.. = the stratifier of the Poisson rules (cpiu.R:567) */
  double  *t_start, *t_end;          /* [rows] (cpiu.R:568-571) */
  double  *rt;                       /* [rows] risk time in years (cpiu.R:572-574) */
  int32_t *i_death;                  /* [rows] (cpiu.R:575-577) */
  double  *vars;                     /* [n_vars][rows]; a value of -1 is NA_REAL (cpiu.R:642) */
  double   kernel_ms;
  int64_t  kernel_launches;
} rfslam_cpiu_out;

int  rfslam_b200_cpiu_build(const rfslam_cpiu_args *args, rfslam_cpiu_out *out);
void rfslam_b200_free_cpiu(rfslam_cpiu_out *out);

#ifdef __cplusplus
}
#endif
#endif
